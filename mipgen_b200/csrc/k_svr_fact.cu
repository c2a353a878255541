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
//      (DMMA.8x8x4, K = 24 / 40 / 88 instead of 192), turns the distances into kernel factors with
//      the fused exp epilogue and parks them in three small shared-memory tables;
//   2. gives every candidate one thread that adds  sum_i  E_a[ra][i] * E_q[rq][i] * E_ins'[ri][i]
//      to its running score (E_ins' already carries alpha_i * exp(-g d_lrc(region, i))).
// Same FP64 arithmetic as the dense kernel per block (||x||^2 + ||s||^2 - 2 x.s), ~1e-13 relative
// agreement with libsvm; work per (candidate, SV) drops from ~219 to ~45 FP64-pipe slots.
// Row features are read from the feature rows K-feat wrote (a representative candidate per row).
// SV blocks are re-tiled at model upload: one 21 KB bulk copy (cp.async.bulk / mbarrier) per chunk.
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

// same routine as in k_svr.cu (kept local: both kernels inline it)
__device__ __forceinline__ double exp_nonpos(double t, const double *__restrict__ tab64)
{
    const double kMagic = 6755399441055744.0;
    const double kf0 = fma(t, 92.332482616893657, kMagic);
    const int k = __double2loint(kf0);
    const double kf = kf0 - kMagic;
    double r = fma(kf, -0x1.62e42fee00000p-7, t);
    r = fma(kf, -0x1.a39ef35793c76p-39, r);
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    const double tj = tab64[k & 63];
    const double v = fma(tj, p, tj);
    const double scaled = __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
    return t < -708.0 ? 0.0 : scaled;
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

__global__ void __launch_bounds__(kThreads, 1)
k_svr_fact(const DevFact *__restrict__ fc, const DevFTask *__restrict__ tasks, int task0, const double *__restrict__ x, int64_t g_base,
           const uint8_t *__restrict__ valid, const double *__restrict__ blob, const double *__restrict__ w_lrc,
           const double *__restrict__ exp2_tab, int n_sv_pad, double gamma, double rho, double zero_score, double *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const DevFTask tk = tasks[task0 + blockIdx.x];
    const int strand = tk.strand;
    const int n_pairs = fc->n_pairs, n_cap = fc->n_cap, n_ext = fc->n_ext, n_lig = fc->n_lig, n_sums = fc->n_sums;
    const int dsum = fc->max_sum - fc->min_sum;
    // role of the two arm tables on this strand
    const int nA = strand ? n_lig : n_ext, nQ = strand ? n_ext : n_lig;
    const int ldA = strand ? FACT_LD_LIG : FACT_LD_EXT, ldQ = strand ? FACT_LD_EXT : FACT_LD_LIG;
    const int kA = strand ? FACT_K_LIG : FACT_K_EXT, kQ = strand ? FACT_K_EXT : FACT_K_LIG;
    const int RA = (tk.nsi * nA + 7) & ~7, RQ = ((tk.nsi + dsum) * nQ + 7) & ~7, RI = (tk.nsi * n_sums + 7) & ~7;
    const int R = RA + RQ + RI;

    // ---- shared memory carve-up (capacities are for the larger of the two strands; host-checked) ----
    double *FA = reinterpret_cast<double *>(smem_raw);
    double *FQ = FA + fc->cap_FA;
    double *FI = FQ + fc->cap_FQ;
    double *xx = FI + fc->cap_FI;           // [cap_R]
    double *E = xx + fc->cap_R;             // [cap_R][EST]
    double *slab = E + fc->cap_R * EST;     // [2][FACT_BLOB]
    double *wst = slab + 2 * FACT_BLOB;     // [2][C] alpha_i * exp(-g d_lrc)
    double *etab = wst + 2 * C;             // [64]
    uint64_t *full = reinterpret_cast<uint64_t *>(etab + 64);  // [2]
    int *rep = reinterpret_cast<int *>(full + 2);              // [cap_R]

    const int n_c = tk.nsi * n_pairs;  // candidates of this task (one strand, one capture size)
    const int n_chunks = n_sv_pad / C;
    const double *w_reg = w_lrc + (int64_t)tk.region * n_sv_pad;

    if (threadIdx.x < 64) etab[threadIdx.x] = exp2_tab[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < R; i += kThreads) rep[i] = 0x7fffffff;
    __syncthreads();
    if (threadIdx.x == 0) {  // chunk 0 is in flight while the row tables are built
        mbar_arrive_expect_tx(&full[0], FACT_BLOB * 8 + C * 8);
        bulk_g2s(slab, blob, FACT_BLOB * 8, &full[0]);
        bulk_g2s(wst, w_reg, C * 8, &full[0]);
    }

    // ---- phase 0: every candidate names its three rows; the lowest candidate index represents a row ----
    int ra = 0, rq = 0, ri = 0, state = 0;  // state: 0 skipped, 1 invalid (zero row), 2 scored
    int64_t g = 0;
    double acc = 0.0;
    const bool mine = (int)threadIdx.x < n_c;
    if (mine) {
        const int s_rel = threadIdx.x / n_pairs, p = threadIdx.x - s_rel * n_pairs;
        g = tk.g0 + ((((int64_t)(tk.si0 + s_rel) * n_cap + tk.ci) * n_pairs + p) * 2 + strand);
        const int e = fc->pair_e[p], l = fc->pair_l[p], sum = e + l;
        const int ie = fc->ext_idx[e], il = fc->lig_idx[l];
        ra = strand ? s_rel * n_lig + il : s_rel * n_ext + ie;
        rq = RA + (strand ? (s_rel + fc->max_sum - sum) * n_ext + ie : (s_rel + fc->max_sum - sum) * n_lig + il);
        ri = RA + RQ + s_rel * n_sums + fc->sum_idx[sum - fc->min_sum];
        if (valid[g]) {
            // feature 22 (extension_arm_length) is zero only in the all-zero row of an invalid candidate
            state = x[(g - g_base) * MG_NFEAT + 21] == 0.0 ? 1 : 2;
            if (state == 2) {
                atomicMin(&rep[ra], (int)threadIdx.x);
                atomicMin(&rep[rq], (int)threadIdx.x);
                atomicMin(&rep[ri], (int)threadIdx.x);
            }
        }
    }
    __syncthreads();

    // ---- phase 1: copy each row's block out of its representative's feature row; ||row||^2 ----
    for (int row = warp; row < R; row += kWarps) {
        const int t = rep[row];
        int ld, kk, role;  // role 0 ext, 1 lig, 2 ins
        double *dst;
        if (row < RA) { ld = ldA; kk = kA; role = strand ? 1 : 0; dst = FA + row * ldA; }
        else if (row < RA + RQ) { ld = ldQ; kk = kQ; role = strand ? 0 : 1; dst = FQ + (row - RA) * ldQ; }
        else { ld = FACT_LD_INS; kk = FACT_K_INS; role = 2; dst = FI + (row - RA - RQ) * FACT_LD_INS; }
        const double *src = nullptr;
        if (t != 0x7fffffff) {
            const int s_rel = t / n_pairs, p = t - s_rel * n_pairs;
            const int64_t gr = tk.g0 + ((((int64_t)(tk.si0 + s_rel) * n_cap + tk.ci) * n_pairs + p) * 2 + strand);
            src = x + (gr - g_base) * MG_NFEAT;
        }
        double ssum = 0.0;
        for (int k0 = 0; k0 < ld; k0 += 32) {
            const int k = k0 + lane;
            double v = 0.0;
            if (src && k < kk) {
                if (role == 0) v = k < 22 ? src[k] : (k == 22 ? src[190] : 0.0);
                else if (role == 1) v = k < 38 ? src[152 + k] : (k == 38 ? src[191] : 0.0);
                else v = k < 86 ? src[66 + k] : 0.0;
            }
            if (k < ld) dst[k] = v;
            ssum = fma(v, v, ssum);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        if (lane == 0) xx[row] = ssum;
    }
    __syncthreads();

    // ---- main loop over chunks of C support vectors ----
    const double ngamma = -gamma;
    const int uI = RI >> 3, uQ = RQ >> 3, uA = RA >> 3, n_units = uI + uQ + uA;
    for (int ch = 0; ch < n_chunks; ch++) {
        const int st = ch & 1;
        if (threadIdx.x == 0 && ch + 1 < n_chunks) {  // prefetch the next chunk into the other stage
            mbar_arrive_expect_tx(&full[st ^ 1], FACT_BLOB * 8 + C * 8);
            bulk_g2s(slab + (st ^ 1) * FACT_BLOB, blob + (int64_t)(ch + 1) * FACT_BLOB, FACT_BLOB * 8, &full[st ^ 1]);
            bulk_g2s(wst + (st ^ 1) * C, w_reg + (int64_t)(ch + 1) * C, C * 8, &full[st ^ 1]);
        }
        mbar_wait(&full[st], (ch >> 1) & 1);
        const double *sb = slab + st * FACT_BLOB;
        const double *ws = wst + st * C;

        // (1) kernel factors of the distinct rows: DMMA over the row's block, fused exp epilogue.
        //     unit = one 8-row fragment x all C columns; largest units (insert, K=88) first.
        for (int u = warp; u < n_units; u += kWarps) {
            const double *F, *S, *ssb;
            int ld, ksteps, row0;
            bool is_ins = false;
            if (u < uI) { F = FI + (u * 8) * FACT_LD_INS; ld = FACT_LD_INS; ksteps = FACT_K_INS / 4; S = sb + FACT_OFF_INS; ssb = sb + FACT_OFF_SS + 2 * C; row0 = RA + RQ + u * 8; is_ins = true; }
            else if (u < uI + uQ) {
                const int m = u - uI;
                F = FQ + (m * 8) * ldQ; ld = ldQ; ksteps = kQ / 4; row0 = RA + m * 8;
                S = sb + (strand ? FACT_OFF_EXT : FACT_OFF_LIG); ssb = sb + FACT_OFF_SS + (strand ? 0 : C);
            } else {
                const int m = u - uI - uQ;
                F = FA + (m * 8) * ldA; ld = ldA; ksteps = kA / 4; row0 = m * 8;
                S = sb + (strand ? FACT_OFF_LIG : FACT_OFF_EXT); ssb = sb + FACT_OFF_SS + (strand ? C : 0);
            }
            double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0};
            const double *fa = F + gid * ld + tig, *s0 = S + gid * ld + tig, *s1 = S + (8 + gid) * ld + tig;
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ks++) {
                const double a = fa[ks * 4], b0 = s0[ks * 4], b1 = s1[ks * 4];
                dmma884(a0[0], a0[1], a, b0);
                dmma884(a1[0], a1[1], a, b1);
            }
            const double xr = xx[row0 + gid];
            const bool finite = fabs(xr) <= 1.7976931348623157e308;  // -inf copy feature: every factor is 0
            double *er = E + (row0 + gid) * EST;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c0 = 2 * tig + j, c1 = 8 + 2 * tig + j;
                double d0 = fmax(fma(-2.0, a0[j], xr + ssb[c0]), 0.0), d1 = fmax(fma(-2.0, a1[j], xr + ssb[c1]), 0.0);
                double e0 = finite ? exp_nonpos(ngamma * d0, etab) : 0.0, e1 = finite ? exp_nonpos(ngamma * d1, etab) : 0.0;
                if (is_ins) { e0 *= ws[c0]; e1 *= ws[c1]; }
                er[c0] = e0;
                er[c1] = e1;
            }
        }
        __syncthreads();

        // (2) every candidate gathers its three factors
        if (state == 2) {
            const double *ea = E + ra * EST, *eq = E + rq * EST, *ei = E + ri * EST;
#pragma unroll
            for (int i = 0; i < C; i++) acc = fma(ea[i] * eq[i], ei[i], acc);
        }
        __syncthreads();
    }

    if (mine) out[g] = state == 2 ? acc - rho : (state == 1 ? zero_score : __longlong_as_double(0x7ff8000000000000LL));
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

int launch_svr_fact(mg_ctx *ctx, const mg_panel *p, int ftask0, int ftask1, const double *d_x, int64_t g_base, int64_t n_cand,
                    const uint8_t *d_valid, const double *d_w, double *d_out)
{
    if (ftask1 <= ftask0) return MG_OK;
    mg_time_begin(ctx, TM_SVR, n_cand);
    k_svr_fact<<<ftask1 - ftask0, kThreads, ctx->fact_smem, ctx->stream>>>(ctx->d_fact, p->d_ftasks, ftask0, d_x, g_base, d_valid,
                                                                         ctx->d_fact_blob, d_w, ctx->d_exp2tab, ctx->n_sv_pad, ctx->gamma,
                                                                         ctx->rho, ctx->zero_score, d_out);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}
