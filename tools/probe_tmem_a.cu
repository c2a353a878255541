// probe_tmem_a.cu -- stand-alone check of tcgen05.mma with the A operand in tensor memory (sm_100a), the building
// block of k_svr_tc's "A in TMEM" form: D[128 x N] = A[128 x K] . B[N x K]^T, kind::f16, FP32 accumulators.
//   variant 0  A from shared memory (control; same descriptors as k_svr_tc)
//   variant 1  A staged shared memory -> TMEM with tcgen05.cp.128x256b (one K = 16 slice = 8 columns per copy)
//   variant 2  A written registers -> TMEM with tcgen05.st.32x32b.x8 (thread = row, two FP16 per 32-bit column)
// plus the issue rate of the TMEM-A form on every SM at once.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/probe_tmem_a tools/probe_tmem_a.cu
//   timeout 60 tools/_bin/probe_tmem_a
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CHECK(x)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) in a K-major FP16 tile: K blocks of 64 columns (128 bytes) are separate [rows x 128 B]
// panels; inside a panel 8-row atoms of 1024 bytes, 16-byte chunks XOR-swizzled with the row (k_svr_tc.cu: tc_sw128)
__host__ __device__ inline uint32_t sw128_f16(int rows, int row, int k)
{
    const int kb = k >> 6, kk = k & 63;
    const uint32_t chunk = (uint32_t)(kk >> 3) ^ (uint32_t)(row & 7);
    return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + chunk * 16u + (uint32_t)(kk & 7) * 2u;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}

template <int N, int K, int VARIANT>
__global__ void __launch_bounds__(128, 1) k_probe(const __half *__restrict__ a_img, const __half *__restrict__ b_img, const __half *__restrict__ a_plain,
                                                  float *__restrict__ d_out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem;                         // [K/64][128 rows][128 B]
    uint8_t *sb = smem + (K / 64) * 128 * 128;  // [K/64][N rows][128 B]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t kCols = 512;

    for (int i = tid; i < (K / 64) * 128 * 64; i += 128) reinterpret_cast<__half *>(sa)[i] = a_img[i];
    for (int i = tid; i < (K / 64) * N * 64; i += 128) reinterpret_cast<__half *>(sb)[i] = b_img[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s, tmem_a = tmem_base_s + 256;   // accumulator at column 0, A operand from column 256

    if (VARIANT == 2) {
        // thread = row (TMEM lane): K / 2 packed columns, element k in the low half of column k / 2 when k is even
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const __half lo = a_plain[tid * K + 2 * (c0 + j)], hi = a_plain[tid * K + 2 * (c0 + j) + 1];
                v[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
            const uint32_t taddr = tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                         "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // F16 x F16 -> F32, K-major
        if (VARIANT == 1) {
            for (int ks = 0; ks < K / 16; ks++) {
                const uint64_t da = make_desc(smem_u32(sa + (ks >> 2) * 128 * 128 + (ks & 3) * 32));
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_a + (uint32_t)ks * 8), "l"(da) : "memory");
            }
        }
        for (int ks = 0; ks < K / 16; ks++) {
            const uint64_t da = make_desc(smem_u32(sa + (ks >> 2) * 128 * 128 + (ks & 3) * 32));
            const uint64_t db = make_desc(smem_u32(sb + (ks >> 2) * N * 128 + (ks & 3) * 32));
            if (VARIANT == 0) mma_ss(tmem_d, da, db, idesc, ks > 0);
            else mma_ts(tmem_d, tmem_a + (uint32_t)ks * 8, db, idesc, ks > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    wait_bar(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; j++) d_out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kCols) : "memory");
}

// ---- issue rate: one kind::f16 M128 x N x K16 MMA with A in TMEM (TS = 1) or in shared memory (TS = 0) ----
template <int N, int TS>
__global__ void __launch_bounds__(128, 1) k_rate(long long *cycles, int iters)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // FP16 ones
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 128 * 128));
        if (TS) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_d + 384), "l"(da) : "memory");
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const uint32_t d = tmem_d + (uint32_t)((i & 1) * N);
            if (TS) mma_ts(d, tmem_d + 384, db, idesc, i > 1 ? 1u : 0u);
            else mma_ss(d, da, db, idesc, i > 1 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
}

// ---- do the FP64 pipe and the tensor core get in each other's way?  Warp 0 issues MMAs (N = 64, A in TMEM) while warps 1..15
//      run 8 independent DFMA chains each; either side can be switched off ----
template <int N, typename T>
__global__ void __launch_bounds__(512, 1) k_contend(long long *cycles, double *sink, int mma_iters, int fma_iters)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 4; i += 512) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    if (warp == 0) {
        if (tid == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 128 * 128));
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_d + 504), "l"(da) : "memory");
            const long long t0 = clock64();
            for (int i = 0; i < mma_iters; i++) mma_ts(tmem_d + (uint32_t)((N <= 128 ? (i & 1) : 0) * N), tmem_d + 504, db, idesc, i > 1 ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            wait_bar(&bar, 0);
            cycles[blockIdx.x * 2] = clock64() - t0;
        }
    } else {
        T a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = (T)(1.0 + tid * 1e-6 + j);
        const T m = (T)1.0000001, c = (T)1e-9;
        const long long t0 = clock64();
        for (int i = 0; i < fma_iters; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = fma(a[j], m, c);
        }
        const long long t1 = clock64();
        double t = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) t += (double)a[j];
        sink[blockIdx.x * 512 + tid] = t;
        if (tid == 32) cycles[blockIdx.x * 2 + 1] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
}

template <int N, typename T>
static int run_contend(int mma_iters, int fma_iters)
{
    long long *d_c, h[296];
    double *d_s;
    CHECK(cudaMalloc(&d_c, sizeof h));
    CHECK(cudaMalloc(&d_s, 148 * 512 * 8));
    CHECK(cudaMemset(d_c, 0, sizeof h));
    const size_t smem = (128 + 256) * 128 + 1024;
    CHECK(cudaFuncSetAttribute(k_contend<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_contend<N, T><<<148, 512, smem>>>(d_c, d_s, mma_iters, fma_iters);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(h, d_c, sizeof h, cudaMemcpyDeviceToHost));
    long long mm = 0, mf = 0;
    for (int i = 0; i < 148; i++) { mm = h[2 * i] > mm ? h[2 * i] : mm; mf = h[2 * i + 1] > mf ? h[2 * i + 1] : mf; }
    printf("contention: %5d MMAs (M128 N%-3d K16, A in TMEM) + %5d x 8 %s per thread on 15 warps: MMA stream %8lld cycles (%.1f per MMA), FMA loop %8lld cycles (%.2f per warp-level FMA and scheduler)\n",
           mma_iters, N, fma_iters, sizeof(T) == 8 ? "DFMA" : "FFMA", mm, mma_iters ? (double)mm / mma_iters : 0.0, mf, fma_iters ? (double)mf / (fma_iters * 8.0 * 15.0 / 4.0) : 0.0);
    cudaFree(d_c); cudaFree(d_s);
    return 0;
}

template <int N, int TS>
static int run_rate()
{
    long long *d_c, h[148];
    CHECK(cudaMalloc(&d_c, sizeof h));
    const size_t smem = (128 + 256) * 128 + 1024;
    CHECK(cudaFuncSetAttribute(k_rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4096;
    k_rate<N, TS><<<148, 128, smem>>>(d_c, iters);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(h, d_c, sizeof h, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
    printf("M128 N%-3d K16 kind::f16, A in %s: %.1f cycles per MMA on every SM at once\n", N, TS ? "TMEM  " : "shared", (double)mx / iters);
    cudaFree(d_c);
    return 0;
}

template <int N, int K, int VARIANT>
static int run_case()
{
    std::vector<__half> A(128 * K), B(N * K), a_img((K / 64) * 128 * 64), b_img((K / 64) * N * 64);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((double)(s >> 11) / 9007199254740992.0 - 0.5); };
    for (auto &v : A) v = __float2half(rnd());
    for (auto &v : B) v = __float2half(rnd());
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < K; k++) a_img[sw128_f16(128, r, k) / 2] = A[r * K + k];
    for (int r = 0; r < N; r++)
        for (int k = 0; k < K; k++) b_img[sw128_f16(N, r, k) / 2] = B[r * K + k];
    __half *d_a, *d_b, *d_p;
    float *d_d;
    CHECK(cudaMalloc(&d_a, a_img.size() * 2));
    CHECK(cudaMalloc(&d_b, b_img.size() * 2));
    CHECK(cudaMalloc(&d_p, A.size() * 2));
    CHECK(cudaMalloc(&d_d, 128 * N * 4));
    CHECK(cudaMemcpy(d_a, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_b, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_p, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemset(d_d, 0xff, 128 * N * 4));
    const size_t smem = (size_t)(K / 64) * (128 + N) * 128 + 1024;
    CHECK(cudaFuncSetAttribute(k_probe<N, K, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<N, K, VARIANT><<<1, 128, smem>>>(d_a, d_b, d_p, d_d);
    CHECK(cudaGetLastError());
    CHECK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CHECK(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    for (int r = 0; r < 128; r++)
        for (int n = 0; n < N; n++) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)__half2float(A[r * K + k]) * (double)__half2float(B[n * K + k]);
            max_err = fmax(max_err, fabs(ref - (double)D[r * N + n]));
            max_ref = fmax(max_ref, fabs(ref));
        }
    static const char *names[] = {"A in shared memory", "A via tcgen05.cp.128x256b", "A via tcgen05.st.32x32b"};
    printf("M=128 N=%d K=%d, %-26s: max |err| %.3e (max |ref| %.3f) -> %s\n", N, K, names[VARIANT], max_err, max_ref, max_err < 1e-4 ? "OK" : "MISMATCH");
    fflush(stdout);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_p); cudaFree(d_d);
    return max_err < 1e-4 ? 0 : 1;
}

int main()
{
    int rc = 0;
    rc |= run_case<64, 128, 0>();
    rc |= run_case<64, 128, 2>() << 2;
    rc |= run_case<64, 128, 1>() << 1;
    rc |= run_case<64, 64, 1>() << 1;
    run_rate<64, 0>();
    run_rate<64, 1>();
    run_rate<128, 1>();
    run_contend<64, double>(0, 2048);      // FP64 alone
    run_contend<64, double>(8192, 0);      // tensor core alone
    run_contend<64, double>(8192, 2048);   // both (the MMA stream outlasts the DFMA loop)
    run_contend<64, double>(2048, 8192);   // both, the DFMA loop outlasts the MMAs
    run_contend<256, double>(4096, 2048);  // full-rate MMA shape
    run_contend<64, float>(0, 8192);       // FP32 alone
    run_contend<64, float>(16384, 8192);   // FP32 under MMAs
    printf("rc = %d (bit 0 control, bit 1 tcgen05.cp, bit 2 tcgen05.st)\n", rc);
    return 0;
}
