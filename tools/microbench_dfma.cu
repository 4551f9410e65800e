// microbench_dfma.cu -- DFMA issue rate of one SM on B200 (the roof of csrc/complement.cuh's gemm64_partial_kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_dfma tools/microbench_dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// outer-product form, as in a register-tiled GEMM: 16 accumulators, 4 + 4 operands that change every step (three distinct 64-bit
// register operands per DFMA; the scalar-times-constant loop above re-reads two of them from the operand reuse cache)
__global__ void dfma_outer_kernel(double *out, int iters, double a0, double b0) {
    double acc[4][4], a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = a0 + threadIdx.x + i; b[i] = b0 - i;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = i + j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = -a[i]; b[i] = -b[i]; }      // (8 cheap ops per 16 DFMAs keep the operands moving)
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// fp64 tensor path: mma.sync.m8n8k4.f64 (DMMA), NT independent 8x8 accumulator tiles per warp
template <int NT>
__global__ void dmma_kernel(double *out, int iters, double a0, double b0) {
    double c[NT][2], a = a0 + threadIdx.x, b = b0 - threadIdx.x;
#pragma unroll
    for (int t = 0; t < NT; ++t) c[t][0] = c[t][1] = t;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
        a = -a;
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) s += c[t][0] + c[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void ffma_kernel(float *out, int iters, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double *d; float *f;
    cudaMalloc(&d, sizeof(double) * sms * 1024 * 4); cudaMalloc(&f, sizeof(float) * sms * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); dfma_kernel<16><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        const double dfma = (double)sms * warps * 32 * 16 * iters;
        printf("DFMA, %2d warps/SM, 16 chains/thread: %.2f TDFMA/s = %.1f DFMA/clk/SM at %d MHz\n", warps, dfma / (ms * 1e-3) / 1e12, dfma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); dfma_outer_kernel<<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        printf("DFMA outer product 4x4, %2d warps/SM: %.2f TDFMA/s = %.1f DFMA/clk/SM\n", warps, dfma / (ms * 1e-3) / 1e12, dfma / (ms * 1e-3) / sms / (khz * 1e3));
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); dmma_kernel<8><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        { const double mac = (double)sms * warps * 8 * iters * 256; printf("DMMA m8n8k4, %2d warps/SM, 8 tiles/warp: %.2f TMAC/s = %.1f MAC/clk/SM\n", warps, mac / (ms * 1e-3) / 1e12, mac / (ms * 1e-3) / sms / (khz * 1e3)); }
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); ffma_kernel<16><<<sms, warps * 32>>>(f, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        printf("FFMA, %2d warps/SM, 16 chains/thread: %.2f TFFMA/s = %.1f FFMA/clk/SM\n", warps, dfma / (ms * 1e-3) / 1e12, dfma / (ms * 1e-3) / sms / (khz * 1e3));
    }
    return 0;
}
