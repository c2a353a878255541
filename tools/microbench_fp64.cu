// microbench_fp64.cu -- measures the FP64 roofline denominators MEASURED_PEAKS.json lacks:
// DMMA.8x8x4 (FP64 tensor pipe) peak, DFMA peak, whether they overlap, and exp() throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/microbench_fp64 tools/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double *out, int iters, double a0, double b0)
{
    double c[CHAINS][2];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int CHAINS>
__global__ void k_dfma(double *out, int iters, double a0, double b0)
{
    double c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c[i];
    if (s == 12345.678) out[0] = s;
}

// even warps DMMA, odd warps DFMA: do the two share one pipe?
__global__ void k_mixed(double *out, int iters, double a0, double b0)
{
    double a = a0 + threadIdx.x * 1e-9, b = b0, s = 0;
    if ((threadIdx.x >> 5) & 1) {
        double c[8];
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = i;
        for (int it = 0; it < iters * 16; it++) {  // 16 DFMA warp-instr ~ one DMMA's worth of pipe time at equal rates
#pragma unroll
            for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i];
    } else {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    }
    if (s == 12345.678) out[0] = s;
}

__global__ void k_exp(double *out, int iters, double x0)
{
    double x = x0 - threadIdx.x * 1e-3, s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s += exp(x - i * 0.01 - it * 1e-6);
    }
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out;
    cudaMalloc(&out, 64);
    const int iters = 4096;
    printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
    for (int threads : {128, 256, 512, 1024}) {
        for (int ctas_per_sm : {1, 2}) {
            if (threads * ctas_per_sm > 2048) continue;
            int grid = sms * ctas_per_sm;
            float ms = time_ms([&] { k_dmma<8><<<grid, threads>>>(out, iters, 1.0, 1.0); });
            double flops = 2.0 * 256 * 8.0 * iters * (threads / 32) * grid;
            printf("  {\"kernel\": \"dmma884_8chains\", \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f},\n", threads,
                   ctas_per_sm, ms, flops / ms / 1e9);
        }
    }
    for (int threads : {256, 512, 1024}) {
        int grid = sms * (2048 / threads);
        float ms = time_ms([&] { k_dfma<8><<<grid, threads>>>(out, iters * 8, 1.0000001, 1e-9); });
        double flops = 2.0 * 8.0 * iters * 8 * threads * (double)grid;
        printf("  {\"kernel\": \"dfma_8chains\", \"threads\": %d, \"grid\": %d, \"ms\": %.4f, \"tflops\": %.2f},\n", threads, grid, ms, flops / ms / 1e9);
    }
    {
        int threads = 512, grid = sms * 2;
        float ms = time_ms([&] { k_mixed<<<grid, threads>>>(out, iters, 1.0, 1.0); });
        double warps = (threads / 32) * (double)grid / 2;
        double f_dmma = 2.0 * 256 * 8.0 * iters * warps, f_dfma = 2.0 * 32 * 8.0 * iters * 16 * warps;
        printf("  {\"kernel\": \"mixed_dmma_dfma\", \"ms\": %.4f, \"tflops_dmma\": %.2f, \"tflops_dfma\": %.2f, \"tflops_sum\": %.2f},\n", ms,
               f_dmma / ms / 1e9, f_dfma / ms / 1e9, (f_dmma + f_dfma) / ms / 1e9);
    }
    {
        int threads = 512, grid = sms * 4;
        float ms = time_ms([&] { k_exp<<<grid, threads>>>(out, 512, -0.5); });
        double n = 8.0 * 512 * threads * (double)grid;
        printf("  {\"kernel\": \"exp_f64\", \"ms\": %.4f, \"gexp_per_s\": %.2f}\n", ms, n / ms / 1e6);
    }
    printf("]}\n");
    return 0;
}
