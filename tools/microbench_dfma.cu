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
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); ffma_kernel<16><<<sms, warps * 32>>>(f, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        printf("FFMA, %2d warps/SM, 16 chains/thread: %.2f TFFMA/s = %.1f FFMA/clk/SM\n", warps, dfma / (ms * 1e-3) / 1e12, dfma / (ms * 1e-3) / sms / (khz * 1e3));
    }
    return 0;
}
