// microbench.cu -- B200 pipe measurements that size the F-update kernel's inner loop:
//   (1) FFMA issue rate of an 8x8 register outer product (operands in registers)
//   (2) the same with packed fma.rn.f32x2
//   (3) LDS.128 cost for the access patterns the kernel can choose from
//   (4) the real inner loop: 4 x LDS.128 + 64 FFMA per entry, per lane
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_ffma(float *out, int iters, long long *cyc) {
    float acc[8][8], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = threadIdx.x * 0.002f - i;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 1e-7f; }   // keep the loop from being hoisted
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void ffma2(float2 &d, float2 a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long *>(&d);
    unsigned long long aa = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long bb = *reinterpret_cast<unsigned long long *>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2 *>(&dd);
}

__global__ void __launch_bounds__(256) k_ffma2(float *out, int iters, long long *cyc) {
    float2 acc[8][4], b[4];
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = make_float2(threadIdx.x * 0.002f - j, threadIdx.x * 0.003f + j);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float2 aa = make_float2(a[i], a[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) ffma2(acc[i][j], aa, b[j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 1e-7f; }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// LDS.128 patterns.  mode 0: all lanes same address; 1: 8 distinct consecutive chunks per
// quarter-warp (conflict-free, no sharing); 2: two distinct rows per warp, lanes 0-15 / 16-31,
// each half one chunk (broadcast within half); 3: 5 distinct chunks of one row (lane % 5);
// 4: lanes read chunk (lane%8) of row (lane/8) with row stride 44 floats
__global__ void __launch_bounds__(256) k_lds(float *out, int iters, int mode, long long *cyc) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 0.5f;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int off;
    if (mode == 0) off = w * 64;
    else if (mode == 1) off = w * 256 + lane * 4;
    else if (mode == 2) off = w * 128 + (lane >> 4) * 44;
    else if (mode == 3) off = w * 64 + (lane % 5) * 4;
    else off = w * 400 + (lane >> 3) * 44 + (lane & 7) * 4;
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float4 v = *reinterpret_cast<const float4 *>(&sm[(off + u * 512 + it * 4) & 8188]);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the real inner loop: thread -> (group g, block b) with B blocks per group; per entry the
// lane loads a = row[chunk bi], row[chunk bi+NB]; b = row[chunk bj], row[chunk bj+NB] and
// does the 8x8 outer product.  rows = tile of E entries, stride RS floats.
template <int NB, int RS, int E, bool DOUBLE_BUF>
__global__ void __launch_bounds__(256) k_inner(float *out, int iters, long long *cyc) {
    constexpr int B = NB * (NB + 1) / 2, G = 256 / B;
    __shared__ __align__(16) float tile[E * RS];
    for (int i = threadIdx.x; i < E * RS; i += blockDim.x) tile[i] = (i % 97) * 0.01f;
    __syncthreads();
    const int t = threadIdx.x;
    int g = t / B, b = t % B;
    if (g >= G) { g = 0; b = 0; }
    int bi = 0, rem = b;
    while (rem >= NB - bi) { rem -= NB - bi; ++bi; }
    const int bj = bi + rem;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float *pa = tile + bi * 4, *pb = tile + bj * 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int e = g; e < E; e += G) {
            const float4 a0 = *reinterpret_cast<const float4 *>(pa + e * RS);
            const float4 a1 = *reinterpret_cast<const float4 *>(pa + e * RS + NB * 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(pb + e * RS);
            const float4 b1 = *reinterpret_cast<const float4 *>(pb + e * RS + NB * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    int dev = 0;
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, dev));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    float *out; long long *cyc, hc[4096];
    CHECK(cudaMalloc(&out, sizeof(float) * 4096 * 256));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * 4096));
    auto report = [&](const char *name, int blocks, double work_per_block, const char *unit) {
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(hc, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < blocks; ++i) if (hc[i] > mx) mx = (double)hc[i];
        printf("%-44s %10.0f cyc  -> %8.3f %s per clk per CTA\n", name, mx, work_per_block / mx, unit);
    };
    const int it = 2000;
    for (int occ = 1; occ <= 2; ++occ) {
        const int blocks = sms * occ;
        printf("--- %d CTA(s) of 256 threads per SM ---\n", occ);
        k_ffma<<<blocks, 256>>>(out, it, cyc);   report("FFMA 8x8 outer product (FMA lanes)", blocks, 256.0 * 64 * it, "FMA");
        k_ffma2<<<blocks, 256>>>(out, it, cyc);  report("FFMA2 8x4x2 outer product (FMA lanes)", blocks, 256.0 * 64 * it, "FMA");
        for (int m = 0; m < 5; ++m) {
            k_lds<<<blocks, 256>>>(out, it, m, cyc);
            char nm[64]; snprintf(nm, sizeof nm, "LDS.128 pattern %d (warp-instr)", m);
            report(nm, blocks, 8.0 * 8 * it, "LDS.128");
        }
        k_inner<5, 44, 136, false><<<blocks, 256>>>(out, it / 10, cyc);
        report("inner loop k=40 (NB=5,RS=44,E=136) FMA", blocks, 255.0 * 64 * 8 * (it / 10), "FMA");
        k_inner<5, 40, 136, false><<<blocks, 256>>>(out, it / 10, cyc);
        report("inner loop k=40 (NB=5,RS=40,E=136) FMA", blocks, 255.0 * 64 * 8 * (it / 10), "FMA");
        k_inner<8, 68, 56, false><<<blocks, 256>>>(out, it / 10, cyc);
        report("inner loop k=64 (NB=8,RS=68,E=56) FMA", blocks, 252.0 * 64 * 8 * (it / 10), "FMA");
        k_inner<8, 64, 56, false><<<blocks, 256>>>(out, it / 10, cyc);
        report("inner loop k=64 (NB=8,RS=64,E=56) FMA", blocks, 252.0 * 64 * 8 * (it / 10), "FMA");
    }
    return 0;
}
