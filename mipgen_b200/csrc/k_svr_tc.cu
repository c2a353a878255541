// k_svr_tc.cu -- K-svr on the 5th-generation tensor cores (tcgen05 / TMEM), split-precision form.
//
//   score(x) = sum_i alpha_i * exp(-gamma * ||x - s_i||^2) - rho        (svm.cpp:2504-2522, 328-368)
//
// tcgen05.mma has no FP64 kind, so the candidates x support-vectors contraction is recast so that the tensor cores
// only ever see numbers they represent EXACTLY and FP32 accumulation errors stay ~1e-8 absolute on a quantity that
// enters the exponent multiplied by 2*gamma (~1e-2):
//
//   * the 127 fractional feature columns (k-mer frequencies: ext 1..21, insert 67..151, lig 153..173) are centred on
//     the support vectors' column means (RBF kernels are translation invariant; |x'| ~ 0.05) and split into two FP16
//     terms  x' = hi + 2^-11 * lo  (22 significant bits; the lo term is pre-scaled so it stays a normal FP16 number):
//         x'.s' ~= hi.hi' + 2^-11 * (hi.lo' + lo.hi')       three K=128 kind::f16 MMAs, two FP32 TMEM accumulators
//   * the integer-valued columns (arm lengths 22 / 174, scan size 152, junction one-hot 175..190), centred on integers,
//     are exact in FP16 (|v| <= 2048) and their products sum exactly in FP32 (< 2^24): a third accumulator, K=32;
//   * the copy-number columns (191, 192) are zero unless BWA found extra copies: an FP64 correction per row that has them;
//   * the long-range block (23..66) is constant per region: it rides in the per-SV weight, as in the factored kernel;
//   * ||x'||^2, ||s'||^2 are FP64 (row constant R, per-SV weight), the exponent is assembled and exponentiated in FP64
//     (the 10-instruction exp of k_svr.cu) and the row sums are FP64.
// Measured deviation from libsvm's double arithmetic: see profiles/ (bench.py `tensor_core_kernel`); the FP64 kernels
// (k_svr.cu, k_svr_fact.cu) remain the default and the reference for it.
//
// CTA = one tile of 128 candidates of ONE region x all support vectors, 20 warps:
//   all      build the A operand once: rows read from K-feat's feature matrix, centred, split, written K-major with the
//            128-byte swizzle the tensor core expects (hand-applied: the operand is computed, not copied, so no TMA
//            tensor map), fence.proxy.async;
//   warp 16  one lane streams per-64-SV operand images (pre-swizzled at model upload: ONE 41 KB cp.async.bulk + the
//            region's weights) through a 3-stage mbarrier ring;
//   warp 17  one lane issues the 26 tcgen05.mma of a tile (M128 N64 K16; three accumulators -- hi.hi, cross terms, integer
//            block) into one of two TMEM accumulator stages and commits to the smem-empty / accumulator-full barriers.
//            The service warps have the highest warp ids: the issue arbiter prefers them;
//   warps 0-15  epilogue: tcgen05.ld (thread = candidate row, 8 columns at a time; four warps per TMEM lane quarter share the
//            64 columns), FP32 recombination, FP64 exponent, exp, weight, row sum; the four partial sums of a row are added
//            at the end.  The FP64 pipe is the busiest unit of this kernel and it is latency bound at low occupancy, hence
//            four epilogue warps per scheduler and a 7-instruction exp (256-entry 2^(j/256) table, degree-3 polynomial).
#include <cuda_fp16.h>

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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// tcgen05.commit: the mbarrier gets one arrival when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand, SWIZZLE_128B: 8-row x 128-byte atoms, 1024 bytes apart (tools/probe_tcgen05.cu checks this encoding)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;  // stride between 8-row groups
    d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the A operand in tensor memory (lane = row, two FP16 per 32-bit column: 8 columns per K = 16 step)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared memory -> tensor memory: 128 rows x 32 bytes (one K = 16 step of a K-major FP16 operand) into 8 columns; ordered with
// the tcgen05.mma instructions the same thread issues afterwards (tools/probe_tmem_a.cu checks both)
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// shared-space loads with 32-bit addresses (the dynamic shared-memory base is re-aligned by hand, so the compiler would
// otherwise fall back to generic loads with 64-bit address arithmetic)
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64_const(uint32_t a)  // data that never changes during the kernel: free to schedule
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

// exp of N exponents at once, stage by stage: t = (256 n + j) ln2/256 + r, exp(t) = 2^n * 2^(j/256) * (1 + r + r^2/2 + r^3/6)
// (|r| <= ln2/512: r^4/24 < 1.4e-13 -- two orders below the operand split's own error).  Exponents are clamped to >= -708 on
// the high word (magnitudes of negative doubles order like unsigned ints); positive ones (the per-SV weight carries the
// matching negative part) pass unchanged.
template <int N>
__device__ __forceinline__ void expn(double (&t)[N], uint32_t tab256)
{
    const double kMagic = 6755399441055744.0;
    double kf0[N], r[N], q[N];
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = __hiloint2double((int)min((unsigned)__double2hiint(t[i]), 0xC0862000u), __double2loint(t[i]));
#pragma unroll
    for (int i = 0; i < N; i++) kf0[i] = fma(t[i], 369.32993046757463, kMagic);   // 256 / ln2
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = kf0[i] - kMagic;
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(q[i], -0.0027076061740622863, t[i]);   // ln2 / 256 (|q| < 2^19: one constant is enough)
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = fma(r[i], 1.0 / 6.0, 0.5);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(r[i] * r[i], q[i], r[i]);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int k = __double2loint(kf0[i]);
        const double tj = lds_f64_const(tab256 + ((uint32_t)(k << 3) & 0x7f8u));
        const double v = fma(tj, r[i], tj);
        t[i] = __hiloint2double(__double2hiint(v) + ((k >> 8) << 20), __double2loint(v));
    }
}

// tools/ablate_tc.py builds variants with parts of the kernel removed (wrong results, timing only):
// 1 no exp arithmetic, 2 no float->double conversions, 4 no MMAs issued, 8 no epilogue arithmetic at all, 16 no operand build,
// 32 no operand-image copies
#ifndef MG_TC_ABLATE
#define MG_TC_ABLATE 0
#endif
// 1: the FP16 hi / lo tiles of the A operand are staged once per CTA from shared into tensor memory (columns 384..511, next to
// the 2 x 192 accumulator columns) and the 24 fractional MMAs of every SV tile read them there: the tensor core then takes
// 2 KB instead of 6 KB per MMA out of shared memory, which the FP64 epilogue's look-ups share
#ifndef MG_TC_A_TMEM
#define MG_TC_A_TMEM 1
#endif

// -DMG_TC_TRACE: CTA 1000 of a launch writes clock64 timestamps of its pipeline events to a global buffer (tools/trace_tc.py)
#ifdef MG_TC_TRACE
__device__ long long g_tc_trace[64 * 8];
#define TC_STAMP(j, k) do { if (blockIdx.x == 1000) g_tc_trace[(j) * 8 + (k)] = clock64(); } while (0)
#else
#define TC_STAMP(j, k) do { } while (0)
#endif

constexpr int kThreads = TC_THREADS;
constexpr int kEpiWarp0 = 0, kEpiWarps = 16;   // warps 16..19: producer, MMA issuer, TMEM allocator, spare
constexpr int kProducerWarp = 16, kMmaWarp = 17, kAllocWarp = 18;
constexpr uint32_t kTmemCols = 512;  // 2 accumulator stages x (hi.hi | cross | integer) x 64 columns = 384 -> next power of two

// shared-memory carve-up (bytes from the 1024-aligned base)
constexpr int kOffAhi = 0, kOffAlo = TC_A_F_BYTES, kOffAI = 2 * TC_A_F_BYTES, kOffB = 2 * TC_A_F_BYTES + TC_A_I_BYTES;
constexpr int kOffRow = kOffB + TC_STAGES * TC_STAGE_BYTES;             // R[128], xce[128], xcl[128], part[3][128] doubles
constexpr int kOffTab = kOffRow + 6 * TC_M * 8;                         // exp table [256] doubles
constexpr int kOffBar = kOffTab + 256 * 8;                              // 10 mbarriers
constexpr int kOffMisc = kOffBar + 16 * 8;                              // tmem base, row state[128] ints
constexpr int kSmemBytes = kOffMisc + 16 + TC_M * 4 + 1024;             // + slack for the 1024-byte alignment of the base

// byte offset of element (row, k) in a K-major FP16 tile of `rows` rows: K blocks of 64 columns (128 bytes) are separate
// [rows x 128 B] panels; inside a panel 8-row atoms of 1024 bytes, 16-byte chunks XOR-swizzled with the row
__host__ __device__ inline uint32_t tc_sw128(int rows, int row, int k)
{
    const int kb = k >> 6, kk = k & 63;
    const uint32_t chunk = (uint32_t)(kk >> 3) ^ (uint32_t)(row & 7);
    return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + chunk * 16u + (uint32_t)(kk & 7) * 2u;
}

// fractional column k (0..126) of the A/B operands -> feature index (0-based) of the 192-vector
__host__ __device__ inline int tc_frac_feature(int k) { return k < 21 ? k : (k < 106 ? 66 + (k - 21) : 152 + (k - 106)); }
// integer column k (0..18): ext_len, scan_size, lig_len, junction one-hot
__host__ __device__ inline int tc_int_feature(int k) { return k == 0 ? 21 : (k == 1 ? 151 : (k == 2 ? 173 : 174 + (k - 3))); }

__global__ void __launch_bounds__(kThreads, 1)
k_svr_tc(const DevRegion *__restrict__ regions, const int64_t *__restrict__ tile_off, const int *__restrict__ tile_region, int n_spans,
         const double *__restrict__ x, int64_t g_base, const uint8_t *__restrict__ valid, const uint8_t *__restrict__ b_img,
         const double *__restrict__ w_all, const double *__restrict__ centre, const double *__restrict__ exp2_tab, int n_sv_pad,
         double gamma, double rho, double zero_score, double *__restrict__ out)
{
    extern __shared__ uint8_t smem_unaligned[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) & ~(uintptr_t)1023);
    double *rowR = reinterpret_cast<double *>(smem + kOffRow), *rowCe = rowR + TC_M, *rowCl = rowCe + TC_M, *part = rowCl + TC_M;  // part[3][TC_M]
    double *etab = reinterpret_cast<double *>(smem + kOffTab);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint64_t *b_full = bars, *b_empty = bars + 3, *d_full = bars + 6, *d_empty = bars + 8;
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(smem + kOffMisc);
    int *rowState = reinterpret_cast<int *>(smem + kOffMisc + 16);  // 0 skipped, 1 invalid (zero row), 2 scored, 3 non-finite feature

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // which span (region x chunk) this tile belongs to: spans are few, binary search on their tile prefix sums
    int lo = 0, hi = n_spans - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_off[2 * mid] <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const int region = tile_region[lo];
    const int64_t g0 = tile_off[2 * lo + 1] + ((int64_t)blockIdx.x - tile_off[2 * lo]) * TC_M;   // first candidate of the tile
    const int64_t g_end = tile_off[2 * lo + 1] + tile_off[2 * n_spans + lo];                       // end of the span
    const int n_rows = (int)(g_end - g0 < (int64_t)TC_M ? g_end - g0 : (int64_t)TC_M);
    const double *w_reg = w_all + (int64_t)region * n_sv_pad;
    const int n_tiles = n_sv_pad / TC_N;

    if (tid < 256) etab[tid] = exp2_tab[tid];
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1 + kEpiWarps); }
        for (int s = 0; s < 2; s++) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kAllocWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- A operand: one warp per row; lanes over the columns.  Integer-block tile is cleared first (its K is padded). ----
    for (int i = tid; i < TC_A_I_BYTES / 16; i += kThreads) reinterpret_cast<int4 *>(smem + kOffAI)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
    // four rows per warp and step, so that their (latency-bound) global loads overlap
    constexpr int kRB = 4;
    for (int rb = warp * kRB; rb < ((MG_TC_ABLATE & 16) ? 0 : TC_M); rb += (kThreads / 32) * kRB) {
        double vf[kRB][4], vi[kRB], vce[kRB], vcl[kRB];
        uint8_t vd[kRB];
#pragma unroll
        for (int j = 0; j < kRB; j++) {
            const int row = rb + j;
            const int64_t g = g0 + row;
            vd[j] = row < n_rows ? valid[g] : 0;
            const double *xr = x + (g - g_base) * MG_NFEAT;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int k = lane + 32 * m;
                vf[j][m] = (vd[j] && k < TC_KF_USED) ? xr[tc_frac_feature(k)] : 0.0;
            }
            vi[j] = (vd[j] && lane < TC_KI_USED) ? xr[tc_int_feature(lane)] : 0.0;
            vce[j] = (vd[j] && lane == TC_KI_USED) ? xr[190] : 0.0;  // log10 copy numbers (0 for copy 1; -inf for copy 0)
            vcl[j] = (vd[j] && lane == TC_KI_USED) ? xr[191] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kRB; j++) {
            const int row = rb + j;
            double ssum = 0.0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int k = lane + 32 * m;
                const double v = (vd[j] && k < TC_KF_USED) ? vf[j][m] - centre[k] : 0.0;
                const __half h = __double2half(v);
                const __half l = __double2half((v - (double)__half2float(h)) * 2048.0);
                *reinterpret_cast<__half *>(smem + kOffAhi + tc_sw128(TC_M, row, k)) = h;
                *reinterpret_cast<__half *>(smem + kOffAlo + tc_sw128(TC_M, row, k)) = l;
                ssum = fma(v, v, ssum);
            }
            if (lane < TC_KI_USED) {
                const double v = vd[j] ? vi[j] - centre[TC_KF + lane] : 0.0;
                *reinterpret_cast<__half *>(smem + kOffAI + tc_sw128(TC_M, row, lane)) = __double2half(v);  // small integers: exact
                ssum = fma(v, v, ssum);
            }
            ssum = fma(vce[j], vce[j], fma(vcl[j], vcl[j], ssum));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
            const double f22 = __shfl_sync(0xffffffffu, vi[j], 0);   // extension_arm_length: zero only in the all-zero row of an invalid candidate
            const double ce = __shfl_sync(0xffffffffu, vce[j], TC_KI_USED), cl = __shfl_sync(0xffffffffu, vcl[j], TC_KI_USED);
            if (lane == 0) {
                const bool finite = fabs(ssum) <= 1.7976931348623157e308;
                rowState[row] = !vd[j] ? 0 : (f22 == 0.0 ? 1 : (finite ? 2 : 3));
                rowR[row] = finite ? -gamma * ssum : 0.0;
                rowCe[row] = finite ? 2.0 * gamma * ce : 0.0;
                rowCl[row] = finite ? 2.0 * gamma * cl : 0.0;
            }
        }
    }
    // generic-proxy writes of the operand -> visible to the tensor core (async proxy); TMEM address -> everyone
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_base_s;

    if (warp == kProducerWarp) {
        // ======================= producer: SV operand images + the region's weights =======================
        if (lane == 0) {
            for (int j = 0; j < n_tiles; j++) {
                const int s = j % TC_STAGES;
                if (j >= TC_STAGES) mbar_wait(&b_empty[s], ((j / TC_STAGES) - 1) & 1);
                uint8_t *dst = smem + kOffB + s * TC_STAGE_BYTES;
                if (MG_TC_ABLATE & 32) { mbar_arrive(&b_full[s]); continue; }
                TC_STAMP(j, 0);   // copy of stage j issued
                mbar_arrive_expect_tx(&b_full[s], TC_IMG_BYTES + TC_N * 8);
                bulk_g2s(dst, b_img + (size_t)j * TC_IMG_BYTES, TC_IMG_BYTES, &b_full[s]);
                bulk_g2s(dst + TC_IMG_BYTES, w_reg + (size_t)j * TC_N, TC_N * 8, &b_full[s]);
            }
        }
    } else if (warp == kMmaWarp) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            // D = F32, A = B = F16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            const uint32_t a_hi = smem_u32(smem + kOffAhi), a_lo = smem_u32(smem + kOffAlo), a_i = smem_u32(smem + kOffAI);
            const uint32_t ta_hi = tmem_d + 384, ta_lo = tmem_d + 448;
            if (MG_TC_A_TMEM) {
#pragma unroll
                for (int ks = 0; ks < TC_KF / 16; ks++) {
                    const uint32_t ao = (uint32_t)(ks >> 2) * (TC_M * 128) + (uint32_t)(ks & 3) * 32;
                    tmem_cp_128x256b(ta_hi + ks * 8, umma_desc(a_hi + ao));
                    tmem_cp_128x256b(ta_lo + ks * 8, umma_desc(a_lo + ao));
                }
            }
            for (int j = 0; j < n_tiles; j++) {
                const int s = j % TC_STAGES, ds = j & 1;
                mbar_wait(&b_full[s], (j / TC_STAGES) & 1);
                TC_STAMP(j, 1);   // operands of tile j have landed
                if (j >= 2) mbar_wait(&d_empty[ds], ((j >> 1) - 1) & 1);
                TC_STAMP(j, 2);   // accumulator stage free: MMAs of tile j issue now
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_hi = smem_u32(smem + kOffB + s * TC_STAGE_BYTES), b_lo = b_hi + TC_B_F_BYTES, b_i = b_hi + 2 * TC_B_F_BYTES;
                const uint32_t d_hh = tmem_d + (uint32_t)(ds * 3 * TC_N), d_x = d_hh + TC_N, d_i = d_hh + 2 * TC_N;
                if (!(MG_TC_ABLATE & 4)) {
                    // three accumulators (hi.hi | 2^11 * (hi.lo + lo.hi) | integer block), issued round-robin
#pragma unroll
                    for (int ks = 0; ks < TC_KF / 16; ks++) {  // K block of 64 columns = one 128-byte swizzle row; 4 steps of 32 bytes inside
                        const uint32_t ao = (uint32_t)(ks >> 2) * (TC_M * 128) + (uint32_t)(ks & 3) * 32, bo = (uint32_t)(ks >> 2) * (TC_N * 128) + (uint32_t)(ks & 3) * 32;
                        if (MG_TC_A_TMEM) {
                            umma_f16_ts(d_hh, ta_hi + ks * 8, umma_desc(b_hi + bo), idesc, ks > 0);
                            umma_f16_ts(d_x, ta_hi + ks * 8, umma_desc(b_lo + bo), idesc, ks > 0);
                            umma_f16_ts(d_x, ta_lo + ks * 8, umma_desc(b_hi + bo), idesc, 1);
                        } else {
                            umma_f16(d_hh, umma_desc(a_hi + ao), umma_desc(b_hi + bo), idesc, ks > 0);
                            umma_f16(d_x, umma_desc(a_hi + ao), umma_desc(b_lo + bo), idesc, ks > 0);
                            umma_f16(d_x, umma_desc(a_lo + ao), umma_desc(b_hi + bo), idesc, 1);
                        }
                        if (ks < 2) umma_f16(d_i, umma_desc(a_i + ks * 32), umma_desc(b_i + ks * 32), idesc, ks > 0);  // 19 integer columns, padded to 32
                    }
                }
                umma_commit(&b_empty[s]);   // the operand stage may be refilled once these MMAs have read it
                umma_commit(&d_full[ds]);   // ... and the accumulators are complete
            }
        }
    } else if (warp < kEpiWarp0 + kEpiWarps) {
        // ======================= epilogue: thread = candidate row (TMEM lane), half of the 64 columns =======================
        const int q = warp & 3, quarter = (warp - kEpiWarp0) >> 2, row = q * 32 + lane;   // quarter: which 16 of the 64 columns
        const double R = rowR[row], xce = rowCe[row], xcl = rowCl[row];
        const int state = rowState[row];
        const bool copies = xce != 0.0 || xcl != 0.0;
        const double g2 = 2.0 * gamma;
        const uint32_t tab_s = smem_u32(etab);
        double acc0 = 0.0, acc1 = 0.0;
        for (int j = 0; j < n_tiles; j++) {
            const int s = j % TC_STAGES, ds = j & 1;
            mbar_wait(&b_full[s], (j / TC_STAGES) & 1);   // completed long ago; orders our reads of the stage's constants
            mbar_wait(&d_full[ds], (j >> 1) & 1);
            if (warp == 0 && lane == 0) TC_STAMP(j, 3);   // accumulators of tile j complete (seen by epilogue warp 0)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t stage_s = smem_u32(smem + kOffB + s * TC_STAGE_BYTES);
            const uint32_t sce_s = stage_s + 2 * TC_B_F_BYTES + TC_B_I_BYTES + quarter * 128, scl_s = sce_s + TC_N * 8;
            const uint32_t w_s = stage_s + TC_IMG_BYTES + quarter * 128;
            const uint32_t t0 = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ds * 3 * TC_N + quarter * 16);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t vh[8], vx[8], vi[8];
                tmem_ld8(t0 + c * 8, vh);
                tmem_ld8(t0 + TC_N + c * 8, vx);
                tmem_ld8(t0 + 2 * TC_N + c * 8, vi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == 1) {
                    // every TMEM read of this accumulator stage is done: hand it back to the MMA issuer
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[ds]);
                    if (warp == 0 && lane == 0) TC_STAMP(j, 4);   // epilogue warp 0 has read its part of the accumulators
                }
                double ea[4], eb[4];
                if (MG_TC_ABLATE & 8) {
                    acc0 += __uint_as_float(vh[0] ^ vx[1] ^ vi[2] ^ vh[3] ^ vx[4] ^ vi[5] ^ vh[6] ^ vx[7]);
                    continue;
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float fa = fmaf(__uint_as_float(vx[i]), 1.0f / 2048.0f, __uint_as_float(vh[i]));
                    const float fb = fmaf(__uint_as_float(vx[4 + i]), 1.0f / 2048.0f, __uint_as_float(vh[4 + i]));
                    if (MG_TC_ABLATE & 2) {
                        ea[i] = fma(__hiloint2double(__float_as_int(fa), (int)vi[i]), g2, R);
                        eb[i] = fma(__hiloint2double(__float_as_int(fb), (int)vi[4 + i]), g2, R);
                    } else {
                        ea[i] = fma((double)fa + (double)__uint_as_float(vi[i]), g2, R);
                        eb[i] = fma((double)fb + (double)__uint_as_float(vi[4 + i]), g2, R);
                    }
                }
                if (copies) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        ea[i] = fma(xce, lds_f64(sce_s + (c * 8 + i) * 8), fma(xcl, lds_f64(scl_s + (c * 8 + i) * 8), ea[i]));
                        eb[i] = fma(xce, lds_f64(sce_s + (c * 8 + 4 + i) * 8), fma(xcl, lds_f64(scl_s + (c * 8 + 4 + i) * 8), eb[i]));
                    }
                }
                if (!(MG_TC_ABLATE & 1)) {
                    expn<4>(ea, tab_s);
                    expn<4>(eb, tab_s);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    acc0 = fma(ea[i], lds_f64(w_s + (c * 8 + i) * 8), acc0);
                    acc1 = fma(eb[i], lds_f64(w_s + (c * 8 + 4 + i) * 8), acc1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_empty[s]);  // constants of the stage are consumed
            if (warp == 0 && lane == 0) TC_STAMP(j, 5);   // epilogue warp 0 done with tile j
        }
        const double acc = acc0 + acc1;
        // the four column quarters of a row
        if (quarter > 0) part[(quarter - 1) * TC_M + row] = acc;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (quarter == 0 && row < n_rows) {
            const double total = ((acc + part[row]) + part[TC_M + row]) + part[2 * TC_M + row];
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            out[g0 + row] = state == 2 ? total - rho : (state == 3 ? -rho : (state == 1 ? zero_score : nan));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kAllocWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
}

// per region r and support vector i:  w'[r][i] = alpha_i * exp(-gamma * (||lrc_r - s_i[23..66]||^2 + tail_i)) * exp(-gamma ||s'_i||^2)
__global__ void __launch_bounds__(256) k_lrc_weights_tc(const double *__restrict__ lrc_all, const double *__restrict__ sv, const double *__restrict__ alpha,
                                                        const double *__restrict__ tail, const double *__restrict__ exp_c, int n_sv_pad,
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
        w[(int64_t)r * n_sv_pad + i] = alpha[i] * exp(-gamma * d) * exp_c[i];
    }
}

}  // namespace

#ifdef MG_TC_TRACE
extern "C" int mg_tc_trace_fetch(long long *out) { return cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 64 * 8) == cudaSuccess ? 0 : -1; }
#endif

int launch_tc_setup(mg_ctx *ctx)
{
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_svr_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return MG_OK;
}

// Host side of the model upload: centres, FP16 operand images, per-SV constants.  Returns false when the model cannot
// be represented (integer columns that are not small integers): the tensor-core mode is then unavailable.
bool mg_tc_prepare_model(const std::vector<double> &sv, int n_sv_pad, int n_sv, double gamma, std::vector<uint8_t> &img,
                         std::vector<double> &centre, std::vector<double> &exp_c)
{
    centre.assign(TC_KF + TC_KI, 0.0);
    if (n_sv <= 0) return false;
    for (int k = 0; k < TC_KF_USED; k++) {
        double m = 0;
        for (int i = 0; i < n_sv; i++) m += sv[(size_t)i * MG_NFEAT + tc_frac_feature(k)];
        centre[k] = m / n_sv;
    }
    for (int k = 0; k < 3; k++) {  // the three length columns: integer centres; the junction columns stay uncentred
        double m = 0;
        for (int i = 0; i < n_sv; i++) m += sv[(size_t)i * MG_NFEAT + tc_int_feature(k)];
        centre[TC_KF + k] = nearbyint(m / n_sv);
    }
    for (int i = 0; i < n_sv; i++)
        for (int k = 0; k < TC_KI_USED; k++) {
            const double v = sv[(size_t)i * MG_NFEAT + tc_int_feature(k)] - centre[TC_KF + k];
            if (v != nearbyint(v) || fabs(v) > 1024.0) return false;
        }
    const int n_tiles = n_sv_pad / TC_N;
    img.assign((size_t)n_tiles * TC_IMG_BYTES, 0);
    exp_c.assign((size_t)n_sv_pad, 0.0);
    for (int i = 0; i < n_sv_pad; i++) {
        uint8_t *t = &img[(size_t)(i / TC_N) * TC_IMG_BYTES];
        const int r = i % TC_N;
        const double *s = &sv[(size_t)i * MG_NFEAT];
        double ss = 0;
        for (int k = 0; k < TC_KF_USED; k++) {
            const double v = s[tc_frac_feature(k)] - centre[k];
            const __half h = __double2half(v);
            const __half l = __double2half((v - (double)__half2float(h)) * 2048.0);
            memcpy(t + tc_sw128(TC_N, r, k), &h, 2);
            memcpy(t + TC_B_F_BYTES + tc_sw128(TC_N, r, k), &l, 2);
            ss += v * v;
        }
        for (int k = 0; k < TC_KI_USED; k++) {
            const double v = s[tc_int_feature(k)] - centre[TC_KF + k];
            const __half h = __double2half(i < n_sv ? v : 0.0);
            memcpy(t + 2 * TC_B_F_BYTES + tc_sw128(TC_N, r, k), &h, 2);
            if (i < n_sv) ss += v * v;
        }
        double *sce = reinterpret_cast<double *>(t + 2 * TC_B_F_BYTES + TC_B_I_BYTES), *scl = sce + TC_N;
        sce[r] = s[190]; scl[r] = s[191];
        ss += s[190] * s[190] + s[191] * s[191];
        exp_c[i] = exp(-gamma * ss);
    }
    return true;
}

int launch_lrc_weights_tc(mg_ctx *ctx, const mg_panel *p, double *d_w)
{
    mg_time_begin(ctx, TM_OTHER, p->n_regions);
    k_lrc_weights_tc<<<p->n_regions, 256, 0, ctx->stream>>>(p->d_lrc, ctx->d_sv, ctx->d_alpha, ctx->d_tail, ctx->d_tc_expc, ctx->n_sv_pad, ctx->gamma, d_w);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}

// candidates [g0, g1) of the panel (feature rows at d_x, row g - g0), tiles of 128 that never straddle a region
int launch_svr_tc(mg_ctx *ctx, const mg_panel *p, const double *d_x, int64_t g0, int64_t g1, const uint8_t *d_valid, const double *d_w, double *d_out)
{
    if (g1 <= g0) return MG_OK;
    // spans = (region x this chunk); arrays: [2*s] first tile, [2*s+1] first candidate, then [2*n + s] length
    std::vector<int64_t> off;
    std::vector<int> reg;
    std::vector<int64_t> len;
    int64_t tiles = 0;
    for (int r = 0; r < p->n_regions; r++) {
        const int64_t a = std::max(g0, p->offsets[r]), b = std::min(g1, p->offsets[r + 1]);
        if (b <= a) continue;
        off.push_back(tiles); off.push_back(a);
        len.push_back(b - a);
        reg.push_back(r);
        tiles += (b - a + TC_M - 1) / TC_M;
    }
    const int n_spans = (int)reg.size();
    if (n_spans == 0 || tiles == 0) return MG_OK;
    if (tiles > 0x7fffffff) { ctx->err = "launch_svr_tc: too many tiles in one launch"; return MG_ERR_INVALID; }
    off.insert(off.end(), len.begin(), len.end());
    int64_t *d_off = nullptr;
    int *d_reg = nullptr;
    CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&d_off, off.size() * 8));
    CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&d_reg, reg.size() * 4));
    // pageable sources: the copies are staged by the runtime before the call returns
    CUDA_TRY(ctx, cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_reg, reg.data(), reg.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    mg_time_begin(ctx, TM_SVR, g1 - g0);
    k_svr_tc<<<(unsigned)tiles, kThreads, kSmemBytes, ctx->stream>>>(p->d_regions, d_off, d_reg, n_spans, d_x, g0, d_valid, ctx->d_tc_img, d_w,
                                                                    ctx->d_tc_centre, ctx->d_exp2tab256, ctx->n_sv_pad, ctx->gamma, ctx->rho,
                                                                    ctx->zero_score, d_out);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    mg_dev_free(ctx, d_off);
    mg_dev_free(ctx, d_reg);
    ctx->tm.svr_tc_mma += (double)tiles * (ctx->n_sv_pad / TC_N) * 26.0;
    return MG_OK;
}
