// k_svr_fact.cu -- K-svr, factored form.
//
// The RBF kernel value of a candidate against support vector i factors over the feature blocks of
// the 192-vector (SVMipv4.cpp:60-113):
//     exp(-g ||x - s_i||^2) = exp(-g d_ext) * exp(-g d_lig) * exp(-g d_ins) * exp(-g d_lrc)
//       ext block  = features   1..22  + 191  (extension arm k-mers, length, log copy)
//       lig block  = features 153..190 + 192  (ligation arm k-mers, length, junction one-hot, log copy)
//       ins block  = features  67..152        (insert k-mers, scan size)
//       lrc block  = features  23..66         (long-range content: constant for a region)
// and the blocks of a candidate are shared with its neighbours: for one strand and capture size,
// a window of W scan starts holds W*n_pairs candidates but only
//     W*n_a  arms indexed by scan start   (+: extension arm [s-e, s-1]        -: ligation arm [s-l, s-1])
//   (W+d)*n_q arms indexed by q = s+cap-sum (+: ligation arm [q, q+l-1]       -: extension arm [q, q+e-1])
//     W*n_sums inserts                     ([s, s+cap-sum-1])
// distinct rows (defaults: 57*W candidates vs 12*W + 12*(W+5) + 6*W rows).  So per window and per
// chunk of 16 support vectors the kernel
//   1. contracts only the DISTINCT rows' blocks with the SV chunk on the FP64 tensor pipe
//      (DMMA.8x8x4, K = 24 / 24 / 88 instead of 192; the 16 junction one-hot columns of the ligation
//      block reduce to one table add), turns the distances into kernel factors with the fused exp
//      epilogue and parks them in three small shared-memory tables;
//   2. gives every candidate one thread that adds  sum_i  E_a[ra][i] * E_q[rq][i] * E_ins'[ri][i]
//      to its running score (E_ins' already carries alpha_i * exp(-g d_lrc(region, i))).
// Same FP64 arithmetic as the dense kernel per block (||x||^2 + ||s||^2 - 2 x.s, here scaled by -gamma at
// model upload so that the contraction ends on the exponent), ~1e-13 relative agreement with libsvm;
// work per (candidate, SV) drops from ~219 to ~35 FP64-pipe slots.
// The distinct rows' blocks come ready-made from K-feat's row-table mode (k_feat.cu: straight from its prefix-count tables, no
// 192-vector per candidate is ever materialised): one bulk copy per work item brings them into shared memory.
// SV blocks are re-tiled at model upload: one 20 KB bulk copy (cp.async.bulk / mbarrier) per chunk.
#include "mg_common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// tools/ablate_fact.py builds variants with parts of the kernel removed (wrong results, timing only):
// 1 no exp, 2 no insert-block DMMA, 4 no gather arithmetic, 8 no arm DMMA
#ifndef MG_FACT_ABLATE
#define MG_FACT_ABLATE 0
#endif

// exp of N exponents at once, written stage by stage (N independent chains hide the FP64 latency; same table and
// polynomial as exp_nonpos in k_svr.cu).  Exponents are <= 0 up to rounding (a few ulp above 0 is harmless);
// anything below -708, including the -inf of a parked row, is clamped to -708 on the high word (an integer min:
// magnitudes of negative doubles order like unsigned ints), i.e. contributes < 4e-308 instead of libsvm's 0 --
// far below one ulp of any score.
template <int N>
__device__ __forceinline__ void expn_nonpos(double (&t)[N], const double *__restrict__ tab64)
{
    const double kMagic = 6755399441055744.0;
    double kf0[N], r[N], q[N];
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = __hiloint2double((int)min((unsigned)__double2hiint(t[i]), 0xC0862000u), __double2loint(t[i]));
#pragma unroll
    for (int i = 0; i < N; i++) kf0[i] = fma(t[i], 92.332482616893657, kMagic);
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = kf0[i] - kMagic;
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(q[i], -0x1.62e42fee00000p-7, t[i]);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(q[i], -0x1.a39ef35793c76p-39, r[i]);
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = fma(r[i], 1.0 / 120.0, 1.0 / 24.0);
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = fma(q[i], r[i], 1.0 / 6.0);
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = fma(q[i], r[i], 0.5);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(r[i] * r[i], q[i], r[i]);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int k = __double2loint(kf0[i]);
        const double tj = tab64[k & 63];
        const double v = fma(tj, r[i], tj);
        t[i] = __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
    }
}

constexpr int kThreads = FACT_THREADS, kWarps = FACT_THREADS / 32;
constexpr int C = FACT_C, EST = FACT_C + 1;  // E-table row stride (odd: rows spread over the banks)

// w[r][i] = alpha_i * exp(-gamma * (sum_j (lrc_rj - s_i,22+j)^2 + tail_i))     (lrc block, features 23..66)
__global__ void __launch_bounds__(256) k_lrc_weights(const double *__restrict__ lrc_all, const double *__restrict__ sv,
                                                     const double *__restrict__ alpha, const double *__restrict__ tail, int n_sv_pad,
                                                     double gamma, double *__restrict__ w)
{
    __shared__ double l[MG_NLRC];
    const int r = blockIdx.x;
    if (threadIdx.x < MG_NLRC) l[threadIdx.x] = lrc_all ? lrc_all[(int64_t)r * MG_NLRC + threadIdx.x] : 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_sv_pad; i += blockDim.x) {
        const double *s = sv + (int64_t)i * MG_NFEAT + 22;
        double d = tail[i];
#pragma unroll 4
        for (int j = 0; j < MG_NLRC; j++) { const double t = l[j] - s[j]; d = fma(t, t, d); }
        w[(int64_t)r * n_sv_pad + i] = alpha[i] * exp(-gamma * d);
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp roles: warps [0, kMathWarps) build the factor tables of chunk c+1 on the FP64 pipe while warps
// [kMathWarps, kMathWarps + kGatherWarps) gather chunk c through the LSU; the first gather thread also
// fetches the SV blobs.  11 + 5 warps: the default configuration has 19 work units per chunk (3 insert,
// 16 arm), which balance over 11 warps, and the gather is co-critical.
constexpr int kMathWarps = FACT_MATH_WARPS, kGatherWarps = FACT_GATHER_WARPS, kGatherThreads = kGatherWarps * 32, kCpt = FACT_CPT;
constexpr int kUnitsPerWarp = 24;

__global__ void __launch_bounds__(kThreads, 1)
k_svr_fact(const DevFact *__restrict__ fc, const DevFTask *__restrict__ tasks, int task0, const double *__restrict__ rows,
           const uint8_t *__restrict__ cstate, const double *__restrict__ blob, const double *__restrict__ w_lrc,
           const double *__restrict__ exp2_tab, int n_sv_pad, double gamma, double rho, double zero_score, double *__restrict__ out,
           unsigned long long *__restrict__ work)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const DevFTask tk = tasks[task0 + blockIdx.x];
    const int strand = tk.strand;
    const int n_pairs = fc->n_pairs, n_cap = fc->n_cap, n_ext = fc->n_ext, n_lig = fc->n_lig, n_sums = fc->n_sums;
    const int dsum = fc->max_sum - fc->min_sum;
    // role of the two arm tables on this strand
    const int nA = strand ? n_lig : n_ext, nQ = strand ? n_ext : n_lig;
    const bool ligA = strand != 0, ligQ = strand == 0;  // which arm table plays the ligation role
    const int RA = (tk.nsi * nA + 15) & ~15, RQ = ((tk.nsi + dsum) * nQ + 15) & ~15, RI = (tk.nsi * n_sums + 15) & ~15;
    const int R = RA + RQ + RI;

    // ---- shared memory carve-up (capacities are for the larger of the two strands; host-checked) ----
    double *FA = reinterpret_cast<double *>(smem_raw);
    double *FQ = FA + fc->cap_FA;
    double *FI = FQ + fc->cap_FQ;
    double *xx = FI + fc->cap_FI;           // [cap_R]
    double *E = xx + fc->cap_R;             // [2][cap_R][EST]  factor tables, double buffered
    double *slab = E + 2 * fc->cap_R * EST; // [2][FACT_BLOB]
    double *wst = slab + 2 * FACT_BLOB;     // [2][C] alpha_i * exp(-g d_lrc)
    double *etab = wst + 2 * C;             // [64]
    uint64_t *bars = reinterpret_cast<uint64_t *>(etab + 64);
    uint64_t *slab_full = bars, *rows_full = bars + 2, *e_full = bars + 4, *e_empty = bars + 6;
    int *jc = reinterpret_cast<int *>(bars + 8);               // [cap_R] junction code of ligation-role rows (16: none)
    uint8_t *unit_list = reinterpret_cast<uint8_t *>(jc + fc->cap_R);  // [kMathWarps][kUnitsPerWarp] units of each math warp
    uint8_t *unit_cnt = unit_list + kMathWarps * kUnitsPerWarp;        // [kMathWarps]

    const int n_c = tk.nsi * n_pairs;  // candidates of this task (one strand, one capture size)
    const int n_chunks = n_sv_pad / C;
    const double *w_reg = w_lrc + (int64_t)tk.region * n_sv_pad;
    const int ESZ = fc->cap_R * EST;

    if (threadIdx.x < 64) etab[threadIdx.x] = exp2_tab[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++) {
            mbar_init(&slab_full[b], 1);
            mbar_init(&e_full[b], kMathWarps);
            mbar_init(&e_empty[b], kGatherWarps);
        }
        mbar_init(rows_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- phase 0 (all threads, one candidate each per round): every candidate names its three rows and reads its state
    //      (K-feat: 0 skipped, 1 invalid, 2 scored).  The result is handed to the gather thread that owns the candidate
    //      (gather thread t owns candidates t, t + kGatherThreads, ...) through the not yet used factor table. ----
    const bool gatherer = warp >= kMathWarps && warp < kMathWarps + kGatherWarps;
    const int gt = threadIdx.x - kMathWarps * 32;
    auto grid_index = [&](int t) {  // candidate t of the task -> index in the panel's grid
        const int s_rel = t / n_pairs, p = t - s_rel * n_pairs;
        return tk.g0 + ((((int64_t)(tk.si0 + s_rel) * n_cap + tk.ci) * n_pairs + p) * 2 + strand);
    };
    int4 *cinfo = reinterpret_cast<int4 *>(E);  // [n_c] {ra, rq, ri, state}; state: 0 skipped, 1 invalid (zero row), 2 scored
    int any_scored = 0;
    for (int t = threadIdx.x; t < n_c; t += kThreads) {
        const int s_rel = t / n_pairs, p = t - s_rel * n_pairs;
        const int e = fc->pair_e[p], l = fc->pair_l[p], sum = e + l;
        const int ie = fc->ext_idx[e], il = fc->lig_idx[l];
        int4 ci;
        ci.x = strand ? s_rel * n_lig + il : s_rel * n_ext + ie;
        ci.y = RA + (strand ? (s_rel + fc->max_sum - sum) * n_ext + ie : (s_rel + fc->max_sum - sum) * n_lig + il);
        ci.z = RA + RQ + s_rel * n_sums + fc->sum_idx[sum - fc->min_sum];
        ci.w = cstate[grid_index(t)];
        any_scored |= ci.w == 2;
        cinfo[t] = ci;
    }
    // a task whose candidates are all skipped or invalid (e.g. a capture size ruled out by mipgen.cpp:429)
    // has no factor tables to build
    const bool some = __syncthreads_or(any_scored);
    int ra[kCpt], rq[kCpt], ri[kCpt], state[kCpt];  // of the gather thread's candidates; state 3: scored with a non-finite feature
#pragma unroll
    for (int h = 0; h < kCpt; h++) {
        const int t = gt + h * kGatherThreads;
        int4 ci = make_int4(0, 0, 0, 0);
        if (gatherer && t < n_c) ci = cinfo[t];
        ra[h] = ci.x; rq[h] = ci.y; ri[h] = ci.z; state[h] = ci.w;
    }
    if (!some) {
        if (gatherer) {
#pragma unroll
            for (int h = 0; h < kCpt; h++) {
                const int t = gt + h * kGatherThreads;
                if (t < n_c) out[grid_index(t)] = state[h] == 1 ? zero_score : __longlong_as_double(0x7ff8000000000000LL);
            }
        }
        return;
    }
    // work units of the math warps: 16 rows x all C columns (8 accumulator chains and 8 independent exp chains per
    // lane: the epilogue is latency bound); longest-processing-time assignment
    const int uI = RI >> 4, uQ = RQ >> 4, uA = RA >> 4, n_units = uI + uQ + uA;
    if (threadIdx.x == 32 && work) {
        // work this task executes (tasks that returned above add nothing): every 8-row fragment meets C columns
        const unsigned long long chunks = (unsigned long long)(n_sv_pad / C);
        atomicAdd(&work[0], chunks * (unsigned long long)(((RA + RQ) / 8 * (FACT_K_ARM / 4) + RI / 8 * (FACT_K_INS / 4)) * (FACT_C / 8)));
        atomicAdd(&work[1], chunks * (unsigned long long)(R * FACT_C));
        atomicAdd(&work[2], chunks * (unsigned long long)(n_c * FACT_C));
    }
    if (threadIdx.x == 0) {
        int load[kMathWarps], cnt[kMathWarps];
        for (int w = 0; w < kMathWarps; w++) load[w] = cnt[w] = 0;
        for (int u = 0; u < n_units; u++) {  // units are ordered insert (heavy) first
            int best = -1;
            for (int w = 0; w < kMathWarps; w++)
                if (cnt[w] < kUnitsPerWarp && (best < 0 || load[w] < load[best])) best = w;
            unit_list[best * kUnitsPerWarp + cnt[best]++] = (uint8_t)u;
            load[best] += (u < uI ? FACT_K_INS : FACT_K_ARM) + 31;  // DMMAs + epilogue (measured: ~31 DMMA times)
        }
        for (int w = 0; w < kMathWarps; w++) unit_cnt[w] = (uint8_t)cnt[w];
    }
    __syncthreads();

    // ---- phase 1: the row tables of this work item (FA | FQ | FI | xx, then jc), as K-feat wrote them: bulk copies ----
    if (threadIdx.x == 0) {
        const double *src = rows + (int64_t)blockIdx.x * fc->blob_doubles;
        const uint32_t nFA = (uint32_t)fc->cap_FA * 8, nFQ = (uint32_t)fc->cap_FQ * 8, nFI = (uint32_t)fc->cap_FI * 8, nX = (uint32_t)fc->cap_R * 8,
                       nJ = (uint32_t)((fc->cap_R * 4 + 15) & ~15);
        mbar_arrive_expect_tx(rows_full, nFA + nFQ + nFI + nX + nJ);
        bulk_g2s(FA, src, nFA, rows_full);
        bulk_g2s(FQ, src + fc->cap_FA, nFQ, rows_full);
        bulk_g2s(FI, src + fc->cap_FA + fc->cap_FQ, nFI, rows_full);
        bulk_g2s(xx, src + fc->cap_FA + fc->cap_FQ + fc->cap_FI, nX, rows_full);
        bulk_g2s(jc, src + fc->cap_FA + fc->cap_FQ + fc->cap_FI + fc->cap_R, nJ, rows_full);
    }
    mbar_wait(rows_full, 0);

    // The SV blob (and the lrc weights) of chunk ch + 2 is fetched by one gather thread as soon as every math warp
    // has delivered chunk ch (e_full): the math warps arrive there after their last read of that blob buffer.
    auto fetch_blob = [&](int ch) {
        const int st = ch & 1;
        mbar_arrive_expect_tx(&slab_full[st], FACT_BLOB * 8 + C * 8);
        bulk_g2s(slab + st * FACT_BLOB, blob + (int64_t)ch * FACT_BLOB, FACT_BLOB * 8, &slab_full[st]);
        bulk_g2s(wst + st * C, w_reg + (int64_t)ch * C, C * 8, &slab_full[st]);
    };
    if (warp < kMathWarps) {
        // ======================= math warps: factor tables of the distinct rows =======================
        const int my_units = unit_cnt[warp];
        for (int ch = 0; ch < n_chunks; ch++) {
            const int st = ch & 1;
            mbar_wait(&slab_full[st], (ch >> 1) & 1);
            const double *sb = slab + st * FACT_BLOB;
            const double *ws = wst + st * C;
            double *Eb = E + st * ESZ;
            for (int ui = 0; ui < my_units; ui++) {
                const int u = unit_list[warp * kUnitsPerWarp + ui];
                const double *F, *S, *ssb;
                int ld, row0;
                bool is_ins = false, is_lig;
                if (u < uI) {
                    F = FI + (u * 16) * FACT_LD_INS; ld = FACT_LD_INS; S = sb + FACT_OFF_INS;
                    ssb = sb + FACT_OFF_SS + 2 * C; row0 = RA + RQ + u * 16; is_ins = true; is_lig = false;
                } else if (u < uI + uQ) {
                    const int m = u - uI;
                    F = FQ + (m * 16) * FACT_LD_ARM; ld = FACT_LD_ARM; row0 = RA + m * 16; is_lig = ligQ;
                    S = sb + (is_lig ? FACT_OFF_LIG : FACT_OFF_EXT); ssb = sb + FACT_OFF_SS + (is_lig ? C : 0);
                } else {
                    const int m = u - uI - uQ;
                    F = FA + (m * 16) * FACT_LD_ARM; ld = FACT_LD_ARM; row0 = m * 16; is_lig = ligA;
                    S = sb + (is_lig ? FACT_OFF_LIG : FACT_OFF_EXT); ssb = sb + FACT_OFF_SS + (is_lig ? C : 0);
                }
                // 4 accumulator chains per warp (2 row fragments x 2 column fragments; the other warps of the scheduler
                // keep the pipe fed).  They start from -gamma (||s||^2 [+ junction term]) (pre-scaled tables), -gamma ||row||^2
                // comes in through the spare column and the SV blocks carry the factor 2 gamma, so the contraction ends
                // on the exponent itself.  One 16-byte
                // load serves two k4 steps: thread-in-group t holds columns 8i+2t (even step) and 8i+2t+1 (odd step)
                double a[2][2][2];
#pragma unroll
                for (int mf = 0; mf < 2; mf++) {
                    const int row = row0 + mf * 8 + gid;
                    const double *tb = is_lig ? sb + FACT_OFF_JT + jc[row] * C : ssb;
#pragma unroll
                    for (int nf = 0; nf < 2; nf++) {
                        const double2 tv = *reinterpret_cast<const double2 *>(tb + nf * 8 + 2 * tig);
                        a[mf][nf][0] = tv.x; a[mf][nf][1] = tv.y;
                    }
                }
                const double2 *f0 = reinterpret_cast<const double2 *>(F + gid * ld + 2 * tig);
                const double2 *f1 = reinterpret_cast<const double2 *>(F + (8 + gid) * ld + 2 * tig);
                const double2 *s0 = reinterpret_cast<const double2 *>(S + gid * ld + 2 * tig);
                const double2 *s1 = reinterpret_cast<const double2 *>(S + (8 + gid) * ld + 2 * tig);
                const bool skip = ((MG_FACT_ABLATE & 2) && is_ins) || ((MG_FACT_ABLATE & 8) && !is_ins);
                auto dmma_iter = [&](int it) {
                    const double2 x0 = f0[it * 4], x1 = f1[it * 4], p0 = s0[it * 4], p1 = s1[it * 4];
                    dmma884(a[0][0][0], a[0][0][1], x0.x, p0.x);
                    dmma884(a[0][1][0], a[0][1][1], x0.x, p1.x);
                    dmma884(a[1][0][0], a[1][0][1], x1.x, p0.x);
                    dmma884(a[1][1][0], a[1][1][1], x1.x, p1.x);
                    dmma884(a[0][0][0], a[0][0][1], x0.y, p0.y);
                    dmma884(a[0][1][0], a[0][1][1], x0.y, p1.y);
                    dmma884(a[1][0][0], a[1][0][1], x1.y, p0.y);
                    dmma884(a[1][1][0], a[1][1][1], x1.y, p1.y);
                };
                // both trip counts are compile-time constants: fully unrolled, the fragment loads of later iterations
                // are issued under the DMMAs of earlier ones
                if (!skip) {
                    if (is_ins) {
#pragma unroll
                        for (int it = 0; it < FACT_LD_INS / 8; it++) dmma_iter(it);
                    } else {
#pragma unroll
                        for (int it = 0; it < FACT_LD_ARM / 8; it++) dmma_iter(it);
                    }
                }
                // epilogue, stage by stage over the lane's 8 elements so that 8 exp chains are in flight
                double t[8];
#pragma unroll
                for (int i = 0; i < 8; i++) t[i] = a[i >> 2][(i >> 1) & 1][i & 1];
                if (!(MG_FACT_ABLATE & 1)) expn_nonpos<8>(t, etab);
                if (is_ins) {
#pragma unroll
                    for (int i = 0; i < 8; i++) t[i] *= ws[((i >> 1) & 1) * 8 + 2 * tig + (i & 1)];
                }
                // the table of chunk ch - 2 must have been gathered before it is overwritten: waiting here, not at the
                // top of the chunk, lets the first unit's contraction and exp run beside the gather's tail
                if (ui == 0 && ch >= 2) mbar_wait(&e_empty[st], ((ch >> 1) & 1) ^ 1);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    Eb[(row0 + (i >> 2) * 8 + gid) * EST + ((i >> 1) & 1) * 8 + 2 * tig + (i & 1)] = t[i];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&e_full[st]);  // this warp's rows of chunk ch are in the table; its reads of the blob are done
        }
    } else {
        // ======================= gather warps: every candidate picks its three factors =======================
        double acc[kCpt];
#pragma unroll
        for (int h = 0; h < kCpt; h++) {
            acc[h] = 0.0;
            // a candidate with a parked (non-finite) row has every kernel value exactly 0, as in libsvm: state 3
            const double kNegInf = __longlong_as_double(0xfff0000000000000LL);
            if (state[h] == 2 && (xx[ra[h]] == kNegInf || xx[rq[h]] == kNegInf || xx[ri[h]] == kNegInf)) state[h] = 3;
        }
        if (gt == 0) {
            fetch_blob(0);
            if (n_chunks > 1) fetch_blob(1);
        }
        for (int ch = 0; ch < n_chunks; ch++) {
            const int st = ch & 1;
            mbar_wait(&e_full[st], (ch >> 1) & 1);
            if (gt == 0 && ch + 2 < n_chunks) fetch_blob(ch + 2);
            const double *Eb = E + st * ESZ;
#pragma unroll
            for (int h = 0; h < kCpt; h++) {
                if (state[h] != 2) continue;
                const double *ea = Eb + ra[h] * EST, *eq = Eb + rq[h] * EST, *ei = Eb + ri[h] * EST;
                double s = acc[h];
#pragma unroll
                for (int i = 0; i < C; i++) {
                    if (MG_FACT_ABLATE & 4) { if (i == 0) s += ea[0] + eq[0] + ei[0]; continue; }
                    s = fma(ea[i] * eq[i], ei[i], s);
                }
                acc[h] = s;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&e_empty[st]);
        }
#pragma unroll
        for (int h = 0; h < kCpt; h++) {
            const int t = gt + h * kGatherThreads;
            if (t < n_c)
                out[grid_index(t)] = state[h] >= 2 ? acc[h] - rho : (state[h] == 1 ? zero_score : __longlong_as_double(0x7ff8000000000000LL));
        }
    }
}

}  // namespace

int launch_fact_setup(mg_ctx *ctx)
{
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_svr_fact, cudaFuncAttributeMaxDynamicSharedMemorySize, FACT_SMEM_LIMIT));
    return MG_OK;
}

int launch_lrc_weights(mg_ctx *ctx, const mg_panel *p, double *d_w)
{
    mg_time_begin(ctx, TM_OTHER, p->n_regions);
    k_lrc_weights<<<p->n_regions, 256, 0, ctx->stream>>>(p->d_lrc, ctx->d_sv, ctx->d_alpha, ctx->d_tail, ctx->n_sv_pad, ctx->gamma, d_w);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}

int launch_svr_fact(mg_ctx *ctx, const mg_panel *p, int ftask0, int ftask1, const double *d_rows, int64_t n_cand,
                    const uint8_t *d_state, const double *d_w, double *d_out)
{
    if (ftask1 <= ftask0) return MG_OK;
    mg_time_begin(ctx, TM_SVR, n_cand);
    k_svr_fact<<<ftask1 - ftask0, kThreads, ctx->fact_smem, ctx->stream>>>(ctx->d_fact, p->d_ftasks, ftask0, d_rows, d_state,
                                                                         ctx->d_fact_blob, d_w, ctx->d_exp2tab, ctx->n_sv_pad, ctx->gamma,
                                                                         ctx->rho, ctx->zero_score, d_out, ctx->d_work);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}
