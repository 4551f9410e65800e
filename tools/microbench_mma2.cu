// microbench_mma2.cu -- follow-up to microbench_mma.cu:
//   (1) does an HMMA block the issue port?  N HMMAs + x independent FFMAs per HMMA, time per HMMA
//   (2) rate of mma.sync.m16n8k16 f16 (fp32 accumulate)
//   (3) Gram prototype with a 2 x fp16 split (x*s = h1 + h2, s a power of two per column): 3 MMAs of K = 16 per
//       accumulator tile per 16 entries instead of 6 of K = 8 -- accuracy against fp64 and speed
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_mma2 tools/microbench_mma2.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                         const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                        const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16z(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}
__device__ __forceinline__ void mma_f16a(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_f16_k8(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
// KIND 0: tf32 k8, 1: f16 k16, 2: f16 k8.  XF independent FFMAs per HMMA.
template <int KIND, int XF>
__global__ void k_mix(float *out, int iters, long long *cyc) {
    float acc[9][4], f[8];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 0.001f + i;
    uint32_t a0 = 0x3c003c00u + threadIdx.x, a1 = 0x3c003800u, a2 = 0x38003c00u, a3 = 0x3c003c00u, b0 = 0x3c003c00u, b1 = 0x34003c00u;
    const float m = 1.0000001f, c = 1e-9f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (KIND == 0) mma_tf32(acc[i], a0, a1, a2, a3, b0, b1);
            else if (KIND == 1) mma_f16(acc[i], a0, a1, a2, a3, b0, b1);
            else mma_f16_k8(acc[i], a0, a1, b0);
#pragma unroll
            for (int x = 0; x < XF; ++x) f[(i * XF + x) & 7] = fmaf(f[(i * XF + x) & 7], m, c);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// fp16-split Gram prototype, k = 40
constexpr int K = 40, NC = 5, MT = 3, RS = 40, EV = 128;
__device__ __forceinline__ uint32_t pack_h2(const float lo, const float hi) {   // .lo = lo, .hi = hi
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(const uint32_t p) {
    const __half2 h = *reinterpret_cast<const __half2 *>(&p);
    return __half22float2(h);
}
__global__ void k_gram16(const float *__restrict__ X, float *__restrict__ G, int iters, long long *cyc, float scale) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = sm + warp * 16 * RS;
    const int ntile = (iters == 1) ? EV : 16;
    for (int i = lane; i < ntile * RS; i += 32) tile[i] = X[i];
    __syncwarp();
    const int g = lane >> 2, tig = lane & 3;
    float acc[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float sc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) sc[c] = scale;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c16 = 0; c16 < ntile / 16; ++c16) {
            const float *p = tile + (c16 * 16 + tig) * RS + g;
            uint32_t a1[MT][4], a2[MT][4], b1[NC][2], b2[NC][2];   // h1 / h2 parts, A- and B-arranged
#pragma unroll
            for (int c = 0; c < 2 * MT; ++c) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {                  // entries (tig + 8 hh, tig + 8 hh + 4)
                    uint32_t p1 = 0u, p2 = 0u;
                    if (c < NC) {
                        const float x0 = p[(8 * hh) * RS + 8 * c] * sc[c], x1 = p[(8 * hh + 4) * RS + 8 * c] * sc[c];
                        p1 = pack_h2(x0, x1);
                        const float2 f = unpack_h2(p1);
                        p2 = pack_h2(x0 - f.x, x1 - f.y);
                        b1[c][hh] = p1; b2[c][hh] = p2;
                    }
                    a1[c >> 1][(c & 1) + 2 * hh] = p1; a2[c >> 1][(c & 1) + 2 * hh] = p2;
                }
            }
            int t = 0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 2 * mt; nt < NC; ++nt) {
                    float d[4];
                    mma_f16z(d, a2[mt], b1[nt]);
                    mma_f16a(d, a1[mt], b2[nt]);
                    mma_f16a(d, a1[mt], b1[nt]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[t][j] += d[j];
                    ++t;
                }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0 && warp == 0) {
        const float inv = 1.f / (scale * scale);
        int t = 0;
        for (int mt = 0; mt < MT; ++mt)
            for (int nt = 2 * mt; nt < NC; ++nt) {
                const int r0 = 16 * mt + g, c0 = 8 * nt + 2 * tig;
                if (r0 < K) { G[r0 * K + c0] = acc[t][0] * inv; G[r0 * K + c0 + 1] = acc[t][1] * inv; }
                if (r0 + 8 < K) { G[(r0 + 8) * K + c0] = acc[t][2] * inv; G[(r0 + 8) * K + c0 + 1] = acc[t][3] * inv; }
                ++t;
            }
    }
}

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs\n", p.name, sms);
    float *out; long long *cyc; static long long hc[8192];
    CHECK(cudaMalloc(&out, sizeof(float) * 8192 * 64));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * 8192));
    auto maxcyc = [&](int blocks) {
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(hc, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < blocks; ++i) if (hc[i] > mx) mx = (double)hc[i];
        return mx;
    };
    const int it = 2000, warps = 16;
#define MIX(KIND, XF)                                                                                     \
    do {                                                                                                  \
        k_mix<KIND, XF><<<sms, warps * 32>>>(out, it, cyc);                                               \
        double c = maxcyc(sms);                                                                           \
        printf("%s + %d FFMA per HMMA, 16 warps/SM: %.2f clk per HMMA per SMSP\n", KIND == 2 ? "HMMA.1688.F16" : KIND ? "HMMA.16816.F16" : "HMMA.1688.TF32", XF, \
               c / (4.0 * 9 * it));                                                                       \
    } while (0)
    MIX(0, 0); MIX(0, 2); MIX(0, 4); MIX(0, 6); MIX(0, 8); MIX(0, 12);
    MIX(2, 0); MIX(2, 2); MIX(2, 4);
    MIX(1, 0); MIX(1, 2); MIX(1, 4); MIX(1, 6); MIX(1, 8); MIX(1, 12);
    // fp16-split Gram
    std::vector<float> hX(EV * RS);
    srand(1);
    for (auto &v : hX) v = (float)rand() / RAND_MAX;
    float *dX, *dG;
    CHECK(cudaMalloc(&dX, sizeof(float) * EV * RS));
    CHECK(cudaMalloc(&dG, sizeof(float) * K * K));
    CHECK(cudaMemcpy(dX, hX.data(), sizeof(float) * EV * RS, cudaMemcpyHostToDevice));
    CHECK(cudaFuncSetAttribute(k_gram16, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (float scale : {16384.f, 1024.f, 1.f}) {
        CHECK(cudaMemset(dG, 0, sizeof(float) * K * K));
        k_gram16<<<1, 32, EV * RS * sizeof(float)>>>(dX, dG, 1, cyc, scale);
        CHECK(cudaDeviceSynchronize());
        std::vector<float> hG(K * K);
        CHECK(cudaMemcpy(hG.data(), dG, sizeof(float) * K * K, cudaMemcpyDeviceToHost));
        double num = 0, den = 0, bias = 0; int cnt = 0;
        for (int r = 0; r < K; ++r)
            for (int c = r; c < K; ++c) {
                double ref = 0;
                for (int e = 0; e < EV; ++e) ref += (double)hX[e * RS + r] * (double)hX[e * RS + c];
                const double d = hG[r * K + c] - ref;
                num += d * d; den += ref * ref; bias += d / ref; ++cnt;
            }
        printf("Gram fp16-split, scale %g: %d entries, rel Frobenius error %.3e, mean signed rel error %.3e\n", scale, EV, sqrt(num / den), bias / cnt);
    }
    for (int w : {8, 12, 16}) {
        const int iters = 2000;
        k_gram16<<<sms, w * 32, w * 16 * RS * sizeof(float)>>>(dX, dG, iters, cyc, 16384.f);
        double c = maxcyc(sms);
        printf("Gram fp16-split, %2d warps/SM: %.3f clk per entry per SM\n", w, c / ((double)w * 16 * iters));
    }
    return 0;
}
