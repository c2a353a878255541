// k_svr.cu -- K-svr: the libsvm RBF epsilon-SVR decision function
//      score(x) = sum_i alpha_i * exp(-gamma * ||x - s_i||^2) - rho
// (svm.cpp:2504-2522 svm_predict_values, 328-368 Kernel::k_function) recast as a dense
// candidates x support-vectors contraction  x.s_i  on the FP64 tensor pipe
// (DMMA.8x8x4 -- every f64 mma.sync shape lowers to it on sm_100a; tcgen05 has no f64
// kind), with ||x-s||^2 = ||x||^2 + ||s||^2 - 2 x.s and the exp / alpha / row-sum fused
// into the epilogue, so the n_cand x n_sv kernel matrix never exists in memory.
//
// Precision: FP64 end to end.  The cancellation in the expanded square costs
// ~1e-16 * ||x||^2 (<= 2e-12 absolute on d, 1e-14 on gamma*d): measured max relative
// deviation from libsvm's sequential double arithmetic is ~1e-13 (tests/test_gpu_parity.py),
// seven orders inside the 1e-6 the north star allows.
//
// CTA = 8 consumer warps + 1 producer warp, one 64-candidate tile per CTA:
//   * the tile's 64 x 192 feature rows are bulk-copied (cp.async.bulk -> UBLKCP, the
//     TMA engine's non-tensor path) once into padded shared memory and stay resident;
//   * support vectors stream through a 4-stage ring of 64 SV x 32 k slabs.  The model is
//     re-tiled once at upload into exactly the padded slab layout shared memory wants, so a
//     slab is ONE 20 KB bulk copy completing on an mbarrier (full/empty pairs) -- the first
//     version issued 64 row copies of 256 B per slab and the consumers waited on the TMA
//     engine for a third of their cycles (profiles/r01_k_svr_v1_ncu.md);
//   * consumer warp (wm, wn) owns a 16 x 32 block of the 64 x 64 chunk: per 8-wide
//     k-block 2 + 4 conflict-free LDS.128 feed 16 DMMAs (k is interleaved even/odd so
//     one 16-byte load serves two k4 steps);
//   * after 6 slabs (k = 192) the epilogue turns 16 accumulators per lane into
//     alpha*exp(-gamma d) and adds them to the lane's running row sums; row sums are
//     reduced across the quad and the two wn halves at the end of the tile.
// The SV matrix (n_sv x 1536 B, 3 MB at 2048 SV) is L2-resident and shared by all CTAs.
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
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier
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

// exp(t) for t <= 0, FP64, ~1 ulp: t = (64 n + j) ln2/64 + r, |r| <= ln2/128;
//   exp(t) = 2^n * 2^(j/64) * (1 + r + r^2/2 + r^3/6 + r^4/24 + r^5/120)      (next term 3.5e-17)
// 10 FP64 instructions instead of ~21 for CUDA's exp(): the epilogue shares the FP64 pipe with
// the DMMAs (profiles/fp64_peaks.json), so every instruction saved here is contraction time.
// Results below 2^-1021 (t < -708) are flushed to 0: libsvm would add alpha * 1e-308.
__device__ __forceinline__ double exp_nonpos(double t, const double *__restrict__ tab64)
{
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: the low word of (x + kMagic) is rint(x)
    const double kf0 = fma(t, 92.332482616893657, kMagic);
    const int k = __double2loint(kf0);
    const double kf = kf0 - kMagic;
    double r = fma(kf, -0x1.62e42fee00000p-7, t);  // ln2/64 split: high part has 21 trailing zero bits
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

constexpr int kSlabs = MG_NFEAT / SVR_BK;                 // 6
constexpr int kXsDoubles = SVR_BM * SVR_LDX;              // 12800
constexpr int kBsDoubles = SVR_BN * SVR_LDB;              // 2560 per stage
constexpr size_t kSmemBytes = (size_t)(kXsDoubles + SVR_STAGES * kBsDoubles + 2 * SVR_BM + 64) * 8 + 16 * 8 + SVR_BM * 4;

__global__ void __launch_bounds__(SVR_THREADS, 1)
k_svr_dmma(const double *__restrict__ x, int64_t n, const double *__restrict__ sv_tiled, const double *__restrict__ ss,
           const double *__restrict__ alpha, const double *__restrict__ exp2_tab, int n_sv_pad, double gamma, double rho,
           const uint8_t *__restrict__ valid, double *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    double *Xs = reinterpret_cast<double *>(smem_raw);
    double *Bs = Xs + kXsDoubles;
    double *red = Bs + SVR_STAGES * kBsDoubles;  // [2][64]
    double *etab = red + 2 * SVR_BM;             // [64] 2^(j/64)
    uint64_t *bars = reinterpret_cast<uint64_t *>(etab + 64);
    uint64_t *full = bars, *empty = bars + SVR_STAGES, *xfull = bars + 2 * SVR_STAGES;
    int *rowflag = reinterpret_cast<int *>(bars + 16);  // [64] non-finite feature row

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * SVR_BM;
    const int n_chunks = n_sv_pad / SVR_BN;

    if (threadIdx.x >= 64 && threadIdx.x < 128) etab[threadIdx.x - 64] = exp2_tab[threadIdx.x - 64];
    if (threadIdx.x == 0) {
        for (int s = 0; s < SVR_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], SVR_CONSUMER_WARPS); }
        mbar_init(xfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == SVR_CONSUMER_WARPS) {
        // ============================ producer warp ============================
        if (lane == 0) mbar_arrive_expect_tx(xfull, SVR_BM * MG_NFEAT * 8);
        __syncwarp();
#pragma unroll
        for (int h = 0; h < SVR_BM / 32; h++) {
            int r = lane + 32 * h;
            bulk_g2s(Xs + r * SVR_LDX, x + (row0 + r) * MG_NFEAT, MG_NFEAT * 8, xfull);
        }
        int q = 0;
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            for (int slab = 0; slab < kSlabs; slab++, q++) {
                const int stage = q % SVR_STAGES;
                const uint32_t phase = (q / SVR_STAGES) & 1;
                if (q >= SVR_STAGES) mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[stage], kBsDoubles * 8);
                    bulk_g2s(Bs + stage * kBsDoubles, sv_tiled + ((int64_t)chunk * kSlabs + slab) * kBsDoubles, kBsDoubles * 8, &full[stage]);
                }
                __syncwarp();
            }
        }
    } else {
        // ============================ consumer warps ============================
        const int wm = warp & 3, wn = warp >> 2;
        const int gid = lane >> 2, tig = lane & 3;
        mbar_wait(xfull, 0);

        // ||x||^2 per row (fixed summation order: lane-strided partials, xor butterfly)
        double xx[2] = {0.0, 0.0};
        int bad[2] = {0, 0};
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const double *row = Xs + (wm * 16 + r) * SVR_LDX;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < MG_NFEAT / 32; k++) { double v = row[lane + 32 * k]; s = fma(v, v, s); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            int nf = !(fabs(s) <= 1.7976931348623157e308);  // inf or NaN anywhere in the row
            if ((r & 7) == gid) { xx[r >> 3] = s; bad[r >> 3] = nf; }
        }

        double part[2] = {0.0, 0.0};
        const double ngamma = -gamma;
        int q = 0;
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            // this lane's 8 columns of the chunk: ||s||^2 and alpha
            const int colb = chunk * SVR_BN + wn * 32 + 2 * tig;
            double2 ssv[4], alv[4];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                ssv[ni] = __ldg(reinterpret_cast<const double2 *>(ss + colb + ni * 8));
                alv[ni] = __ldg(reinterpret_cast<const double2 *>(alpha + colb + ni * 8));
            }
            double acc[2][4][2];
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

            for (int slab = 0; slab < kSlabs; slab++, q++) {
                const int stage = q % SVR_STAGES;
                const uint32_t phase = (q / SVR_STAGES) & 1;
                mbar_wait(&full[stage], phase);
                const double *As = Xs + (wm * 16 + gid) * SVR_LDX + slab * SVR_BK + 2 * tig;
                const double *Bp = Bs + stage * kBsDoubles + (wn * 32 + gid) * SVR_LDB + 2 * tig;
#pragma unroll
                for (int kb = 0; kb < SVR_BK / 8; kb++) {
                    double2 a[2], b[4];
#pragma unroll
                    for (int mi = 0; mi < 2; mi++) a[mi] = *reinterpret_cast<const double2 *>(As + mi * 8 * SVR_LDX + kb * 8);
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) b[ni] = *reinterpret_cast<const double2 *>(Bp + ni * 8 * SVR_LDB + kb * 8);
                    // two passes over the 8 accumulators so dependent DMMAs sit 8 instructions apart
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, b[ni].x);  // k = k0 + {0,2,4,6}
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, b[ni].y);  // k = k0 + {1,3,5,7}
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }

            // fused epilogue: d = ||x||^2 + ||s||^2 - 2 x.s ; k = exp(-gamma d) ; row += alpha k
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    double d0 = fma(-2.0, acc[mi][ni][0], xx[mi] + ssv[ni].x);
                    double d1 = fma(-2.0, acc[mi][ni][1], xx[mi] + ssv[ni].y);
                    d0 = fmax(d0, 0.0);
                    d1 = fmax(d1, 0.0);
                    part[mi] = fma(alv[ni].x, exp_nonpos(ngamma * d0, etab), part[mi]);
                    part[mi] = fma(alv[ni].y, exp_nonpos(ngamma * d1, etab), part[mi]);
                }
        }

        // row sums: quad reduce, then the two wn halves through shared memory
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            double s = part[mi];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (tig == 0) {
                red[wn * SVR_BM + wm * 16 + mi * 8 + gid] = s;
                if (wn == 0) rowflag[wm * 16 + mi * 8 + gid] = bad[mi];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < SVR_BM) {
        const int64_t g = row0 + threadIdx.x;
        if (g < n) {
            // a non-finite feature (log10(0) = -inf) makes every kernel value exp(-inf) = 0 in libsvm
            double s = rowflag[threadIdx.x] ? 0.0 : red[threadIdx.x] + red[SVR_BM + threadIdx.x];
            s -= rho;
            if (valid && !valid[g]) s = __longlong_as_double(0x7ff8000000000000LL);
            out[g] = s;
        }
    }
}

// Cross-check kernel: libsvm's own arithmetic order, one thread per row.
//   sum_k (x_k - s_k)^2 left to right with separately rounded mul/add (svm.cpp:330-365),
//   exp, then sum_i alpha_i k_i left to right, then - rho (svm.cpp:2511-2516).
__global__ void __launch_bounds__(64) k_svr_direct(const double *__restrict__ x, int64_t n, int64_t ld,
                                                   const double *__restrict__ sv, const double *__restrict__ alpha,
                                                   const double *__restrict__ tail, int n_sv, double gamma, double rho,
                                                   double *__restrict__ out)
{
    __shared__ double srow[MG_NFEAT];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double *xr = x + (i < n ? i : 0) * ld;
    double total = 0.0;
    for (int j = 0; j < n_sv; j++) {
        __syncthreads();
        for (int k = threadIdx.x; k < MG_NFEAT; k += blockDim.x) srow[k] = sv[(int64_t)j * MG_NFEAT + k];
        __syncthreads();
        double sum = 0.0;
        for (int k = 0; k < MG_NFEAT; k++) {
            double d = __dsub_rn(xr[k], srow[k]);
            sum = __dadd_rn(sum, __dmul_rn(d, d));
        }
        if (tail[j] != 0.0) sum = __dadd_rn(sum, tail[j]);  // SV features beyond index 192: + y*y (svm.cpp:360-364)
        double kv = exp(__dmul_rn(-gamma, sum));
        total = __dadd_rn(total, __dmul_rn(alpha[j], kv));
    }
    if (i < n) out[i] = __dsub_rn(total, rho);
}

}  // namespace

int launch_svr_setup(mg_ctx *ctx)
{
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_svr_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    return MG_OK;
}

int launch_svr(mg_ctx *ctx, const double *d_x, int64_t n, const uint8_t *d_valid, double *d_out)
{
    if (n <= 0) return MG_OK;
    const int64_t tiles = (n + SVR_BM - 1) / SVR_BM;
    mg_time_begin(ctx, TM_SVR, n);
    k_svr_dmma<<<(unsigned)tiles, SVR_THREADS, kSmemBytes, ctx->stream>>>(d_x, n, ctx->d_sv_tiled, ctx->d_ss, ctx->d_alpha, ctx->d_exp2tab, ctx->n_sv_pad,
                                                                       ctx->gamma, ctx->rho, d_valid, d_out);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    // executed work: every tile contracts 64 rows x n_sv_pad columns x 192
    ctx->tm.svr_dmma += (double)tiles * (ctx->n_sv_pad / SVR_BN) * (SVR_BM * SVR_BN * MG_NFEAT / 256.0);
    ctx->tm.svr_exp += (double)tiles * SVR_BM * ctx->n_sv_pad;
    return MG_OK;
}

int launch_svr_direct(mg_ctx *ctx, const double *d_x, int64_t n, int64_t ld, double *d_out)
{
    if (n <= 0) return MG_OK;
    mg_time_begin(ctx, TM_OTHER, n);
    k_svr_direct<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(d_x, n, ld, ctx->d_sv, ctx->d_alpha, ctx->d_tail, ctx->n_sv,
                                                                   ctx->gamma, ctx->rho, d_out);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}
