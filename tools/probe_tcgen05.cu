// probe_tcgen05.cu -- stand-alone check of the tcgen05 / TMEM building blocks K-svr's split-TF32 variant uses
// (sm_100a): hand-swizzled K-major shared-memory operands (SWIZZLE_128B), shared-memory matrix descriptors,
// the kind::tf32 instruction descriptor, TMEM allocation, tcgen05.mma issue by one thread, tcgen05.commit to an
// mbarrier, tcgen05.ld of the FP32 accumulator.  D[128 x N] = A[128 x K] . B[N x K]^T with K a multiple of 32
// (one 128-byte swizzle atom per 32 tf32 columns), compared with a double-precision host product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/probe_tcgen05 tools/probe_tcgen05.cu
//   timeout 60 tools/_bin/probe_tcgen05
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

// byte offset of element (row, k) of a K-major tile of 32-bit elements, 32 columns (= 128 bytes) per K block,
// rows grouped by 8 into 1024-byte swizzle atoms; K block kb is a separate [rows x 128 B] panel
__host__ __device__ inline uint32_t sw128_offset(int rows, int row, int k)
{
    const int kb = k >> 5, kk = k & 31;
    const uint32_t chunk = (uint32_t)(kk >> 2) ^ (uint32_t)(row & 7);  // 16-byte chunk index XOR row-in-atom
    return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + chunk * 16u + (uint32_t)(kk & 3) * 4u;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    // start address >> 4 [0,14) | LBO >> 4 [16,30) = 0 (single swizzle atom along K) | SBO >> 4 [32,46) = 1024 B
    // | version [46,48) = 1 | layout type [61,64) = 2 (SWIZZLE_128B)
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int N, int K>
__global__ void __launch_bounds__(128, 1) k_probe(const float *__restrict__ a_img, const float *__restrict__ b_img, float *__restrict__ d_out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem;                       // [K/32][128 rows][128 B]
    uint8_t *sb = smem + (K / 32) * 128 * 128; // [K/32][N rows][128 B]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // operands arrive as ready-made shared-memory images (the host applied the swizzle): plain copies
    for (int i = tid; i < (K / 32) * 128 * 32; i += 128) reinterpret_cast<float *>(sa)[i] = a_img[i];
    for (int i = tid; i < (K / 32) * N * 32; i += 128) reinterpret_cast<float *>(sb)[i] = b_img[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(N < 32 ? 32 : N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // generic-proxy writes of the operands -> visible to the async proxy (tensor core reads)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    if (tid == 0) {
        // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kb = 0; kb < K / 32; kb++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {  // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
                const uint64_t da = make_desc(smem_u32(sa + kb * 128 * 128 + ks * 32));
                const uint64_t db = make_desc(smem_u32(sb + kb * N * 128 + ks * 32));
                const uint32_t acc = (kb | ks) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                    "l"(da), "l"(db), "r"(idesc), "r"(acc)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the accumulator
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes [32w, 32w+32): thread = one row of D, 32 columns per load
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
            "%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j++) d_out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(N < 32 ? 32 : N) : "memory");
}

// ---- issue-rate probe: how long does one kind::f16 M128 x N x K16 MMA occupy the tensor pipe, per N? ----
template <int N>
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
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // F16 x F16 -> F32
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 128 * 128));
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const uint32_t d = tmem_d + (uint32_t)((i & 1) * N);  // two accumulators, alternating
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                "l"(da), "l"(db), "r"(idesc), "r"(i > 1 ? 1u : 0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
}

template <int N>
static int run_rate()
{
    long long *d_c, h[148];
    CHECK(cudaMalloc(&d_c, sizeof h));
    const size_t smem = (128 + 256) * 128 + 1024;
    CHECK(cudaFuncSetAttribute(k_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4096;
    k_rate<N><<<148, 128, smem>>>(d_c, iters);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(h, d_c, sizeof h, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
    printf("M128 N%-3d K16 kind::f16: %.1f cycles per MMA on every SM at once (=> %.0f MAC/clk/SM)\n", N, (double)mx / iters, 128.0 * N * 16 / ((double)mx / iters));
    cudaFree(d_c);
    return 0;
}

static float tf32_round(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;  // round to nearest on the 13 dropped bits
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

template <int N, int K>
static int run_case()
{
    std::vector<float> A(128 * K), B(N * K), a_img((K / 32) * 128 * 32), b_img((K / 32) * N * 32);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((double)(s >> 11) / 9007199254740992.0 - 0.5); };
    for (auto &v : A) v = tf32_round(rnd());
    for (auto &v : B) v = tf32_round(rnd());
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < K; k++) a_img[sw128_offset(128, r, k) / 4] = A[r * K + k];
    for (int r = 0; r < N; r++)
        for (int k = 0; k < K; k++) b_img[sw128_offset(N, r, k) / 4] = B[r * K + k];
    float *d_a, *d_b, *d_d;
    CHECK(cudaMalloc(&d_a, a_img.size() * 4));
    CHECK(cudaMalloc(&d_b, b_img.size() * 4));
    CHECK(cudaMalloc(&d_d, 128 * N * 4));
    CHECK(cudaMemcpy(d_a, a_img.data(), a_img.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_b, b_img.data(), b_img.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemset(d_d, 0xff, 128 * N * 4));
    const size_t smem = (size_t)(K / 32) * (128 + N) * 128 + 1024;
    CHECK(cudaFuncSetAttribute(k_probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<N, K><<<1, 128, smem>>>(d_a, d_b, d_d);
    CHECK(cudaGetLastError());
    CHECK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CHECK(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    for (int r = 0; r < 128; r++)
        for (int n = 0; n < N; n++) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)A[r * K + k] * (double)B[n * K + k];
            max_err = fmax(max_err, fabs(ref - (double)D[r * N + n]));
            max_ref = fmax(max_ref, fabs(ref));
        }
    printf("tcgen05 probe M=128 N=%d K=%d: max |err| %.3e (max |ref| %.3f) -> %s\n", N, K, max_err, max_ref, max_err < 1e-5 ? "OK" : "MISMATCH");
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_d);
    return max_err < 1e-5 ? 0 : 1;
}

int main()
{
    int rc = 0;
    rc |= run_case<64, 32>();
    rc |= run_case<64, 128>();
    rc |= run_case<128, 64>();
    run_rate<32>();
    run_rate<64>();
    run_rate<128>();
    run_rate<256>();
    printf(rc ? "PROBE FAILED\n" : "PROBE PASSED\n");
    return rc;
}
