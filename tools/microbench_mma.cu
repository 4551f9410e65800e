// microbench_mma.cu -- can the legacy warp-level tensor path (mma.sync m16n8k8 TF32 -> SASS HMMA) carry the
// per-series Gram at fp32-class accuracy?  Measures on B200:
//   (1) raw issue rate of mma.sync.m16n8k8.tf32 with 9 independent accumulators per warp, 4/8/16 warps per SM
//   (2) a prototype of the Gram inner loop for k = 40: per 8 entries 10 x LDS.32, hi/lo TF32 split,
//       9 upper-triangle 16x8 tiles x 3 MMAs (hi*hi + hi*lo + lo*hi), fp32 FADD accumulation -- timing and the
//       error of the result against an fp64 Gram computed on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_mma tools/microbench_mma.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                         const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NACC>
__global__ void k_mma_rate(float *out, int iters, long long *cyc) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    uint32_t a0 = f2tf32(threadIdx.x * 0.001f), a1 = f2tf32(threadIdx.x * 0.002f), a2 = f2tf32(1.f), a3 = f2tf32(0.5f);
    uint32_t b0 = f2tf32(threadIdx.x * 0.003f), b1 = f2tf32(0.25f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma_tf32(acc[i], a0, a1, a2, a3, b0, b1);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Gram prototype, k = 40: each warp owns a tile of E entries (rows of RS floats) in shared memory and a
// 40 x 40 upper-triangle Gram in 9 x 4 fp32 registers per lane.
constexpr int K = 40, NC = 5, MT = 3, RS = 40, E = 32;
template <int SPLIT>   // 3: 3xTF32, 1: plain TF32
__global__ void k_gram(const float *__restrict__ X, float *__restrict__ G, int iters, long long *cyc) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = sm + warp * E * RS;
    for (int i = lane; i < E * RS; i += 32) tile[i] = X[i];
    __syncwarp();
    const int g = lane >> 2, tig = lane & 3;
    float acc[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int c8 = 0; c8 < E / 8; ++c8) {
            const float *p = tile + (c8 * 8 + tig) * RS + g;
            uint32_t hi[NC + 1][2], lo[NC + 1][2];
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float v = p[h * 4 * RS + 8 * c];
                    hi[c][h] = f2tf32(v);
                    lo[c][h] = f2tf32(v - __uint_as_float(hi[c][h]));
                }
            hi[NC][0] = hi[NC][1] = lo[NC][0] = lo[NC][1] = 0u;
            int t = 0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 2 * mt; nt < NC; ++nt) {
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
                    if (SPLIT == 3) {
                        mma_tf32(d, lo[2 * mt][0], lo[2 * mt + 1][0], lo[2 * mt][1], lo[2 * mt + 1][1], hi[nt][0], hi[nt][1]);
                        mma_tf32(d, hi[2 * mt][0], hi[2 * mt + 1][0], hi[2 * mt][1], hi[2 * mt + 1][1], lo[nt][0], lo[nt][1]);
                    }
                    mma_tf32(d, hi[2 * mt][0], hi[2 * mt + 1][0], hi[2 * mt][1], hi[2 * mt + 1][1], hi[nt][0], hi[nt][1]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[t][j] += d[j];
                    ++t;
                }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    // store: tile (mt, nt): c0 (16mt+g, 8nt+2tig), c1 (.., +1), c2 (16mt+g+8, 8nt+2tig), c3 (.., +1)
    if (blockIdx.x == 0 && warp == 0) {
        int t = 0;
        for (int mt = 0; mt < MT; ++mt)
            for (int nt = 2 * mt; nt < NC; ++nt) {
                const int r0 = 16 * mt + g, c0 = 8 * nt + 2 * tig;
                if (r0 < K) { G[r0 * K + c0] = acc[t][0]; G[r0 * K + c0 + 1] = acc[t][1]; }
                if (r0 + 8 < K) { G[(r0 + 8) * K + c0] = acc[t][2]; G[(r0 + 8) * K + c0 + 1] = acc[t][3]; }
                ++t;
            }
    }
}


__device__ __forceinline__ void mma_tf32_k4(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
template <int NACC>
__global__ void k_mma_rate_k4(float *out, int iters, long long *cyc) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    uint32_t a0 = f2tf32(threadIdx.x * 0.001f), a1 = f2tf32(threadIdx.x * 0.002f);
    uint32_t b0 = f2tf32(threadIdx.x * 0.003f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma_tf32_k4(acc[i], a0, a1, b0);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Variants of the Gram loop, EV entries per warp tile (all in shared memory):
//   MODE 0: cheap split (hi = (bits + 0x1000) & ~0x1fff, lo = x - hi passed raw), k8 MMAs, FADD accumulation per chunk
//   MODE 1: same split, k4 MMAs (no operand re-packing), FADD accumulation per chunk
//   MODE 2: same split, k8 MMAs accumulating directly in the tensor core (no FADD)
//   MODE 3: MODE 0 without the lo terms (plain TF32) -- speed bound of the non-split path
constexpr int EV = 128;
__device__ __forceinline__ void split_tf32(const float v, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
template <int MODE>
__global__ void k_gram2(const float *__restrict__ X, float *__restrict__ G, int iters, long long *cyc) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = sm + warp * 16 * RS;        // timing: 16-entry tile per warp, re-read; accuracy run: 1 warp, EV entries
    const int ntile = (iters == 1) ? EV : 16;
    for (int i = lane; i < ntile * RS; i += 32) tile[i] = X[i];
    __syncwarp();
    const int g = lane >> 2, tig = lane & 3;
    float acc[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int c8 = 0; c8 < ntile / 8; ++c8) {
            const float *p = tile + (c8 * 8 + tig) * RS + g;
            uint32_t hi[2][NC + 1], lo[2][NC + 1];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int c = 0; c < NC; ++c) split_tf32(p[h * 4 * RS + 8 * c], hi[h][c], lo[h][c]);
                hi[h][NC] = 0u; lo[h][NC] = 0u;
            }
            int t = 0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 2 * mt; nt < NC; ++nt) {
                    if (MODE == 2) {
                        mma_tf32(acc[t], lo[0][2 * mt], lo[0][2 * mt + 1], lo[1][2 * mt], lo[1][2 * mt + 1], hi[0][nt], hi[1][nt]);
                        mma_tf32(acc[t], hi[0][2 * mt], hi[0][2 * mt + 1], hi[1][2 * mt], hi[1][2 * mt + 1], lo[0][nt], lo[1][nt]);
                        mma_tf32(acc[t], hi[0][2 * mt], hi[0][2 * mt + 1], hi[1][2 * mt], hi[1][2 * mt + 1], hi[0][nt], hi[1][nt]);
                    } else {
                        float d[4] = {0.f, 0.f, 0.f, 0.f};
                        if (MODE == 0) {
                            mma_tf32(d, lo[0][2 * mt], lo[0][2 * mt + 1], lo[1][2 * mt], lo[1][2 * mt + 1], hi[0][nt], hi[1][nt]);
                            mma_tf32(d, hi[0][2 * mt], hi[0][2 * mt + 1], hi[1][2 * mt], hi[1][2 * mt + 1], lo[0][nt], lo[1][nt]);
                            mma_tf32(d, hi[0][2 * mt], hi[0][2 * mt + 1], hi[1][2 * mt], hi[1][2 * mt + 1], hi[0][nt], hi[1][nt]);
                        } else if (MODE == 3) {
                            mma_tf32(d, hi[0][2 * mt], hi[0][2 * mt + 1], hi[1][2 * mt], hi[1][2 * mt + 1], hi[0][nt], hi[1][nt]);
                        } else {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                mma_tf32_k4(d, lo[h][2 * mt], lo[h][2 * mt + 1], hi[h][nt]);
                                mma_tf32_k4(d, hi[h][2 * mt], hi[h][2 * mt + 1], lo[h][nt]);
                            }
#pragma unroll
                            for (int h = 0; h < 2; ++h) mma_tf32_k4(d, hi[h][2 * mt], hi[h][2 * mt + 1], hi[h][nt]);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[t][j] += d[j];
                    }
                    ++t;
                }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0 && warp == 0) {
        int t = 0;
        for (int mt = 0; mt < MT; ++mt)
            for (int nt = 2 * mt; nt < NC; ++nt) {
                const int r0 = 16 * mt + g, c0 = 8 * nt + 2 * tig;
                if (r0 < K) { G[r0 * K + c0] = acc[t][0]; G[r0 * K + c0 + 1] = acc[t][1]; }
                if (r0 + 8 < K) { G[(r0 + 8) * K + c0] = acc[t][2]; G[(r0 + 8) * K + c0 + 1] = acc[t][3]; }
                ++t;
            }
    }
}

template <int MODE>
static void run_gram2(int sms, const float *dX, float *dG, long long *cyc, const std::vector<float> &hX) {
    static long long hc[8192];
    const char *names[4] = {"k8 + FADD", "k4 + FADD", "k8 direct accumulate", "plain TF32 k8 + FADD"};
    CHECK(cudaFuncSetAttribute(k_gram2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CHECK(cudaMemset(dG, 0, sizeof(float) * K * K));
    k_gram2<MODE><<<1, 32, EV * RS * sizeof(float)>>>(dX, dG, 1, cyc);
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
    printf("Gram2 [%s]: %d entries, rel Frobenius error %.3e, mean signed rel error %.3e\n", names[MODE], EV, sqrt(num / den), bias / cnt);
    for (int warps : {8, 12, 16}) {
        const int iters = 1000;
        k_gram2<MODE><<<sms, warps * 32, warps * 16 * RS * sizeof(float)>>>(dX, dG, iters, cyc);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(hc, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < sms; ++i) if (hc[i] > mx) mx = (double)hc[i];
        printf("Gram2 [%s], %2d warps/SM: %.3f clk per entry per SM\n", names[MODE], warps, mx / ((double)warps * 16 * iters));
    }
}

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    float *out; long long *cyc; static long long hc[8192];
    CHECK(cudaMalloc(&out, sizeof(float) * 8192 * 64));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * 8192));
    auto maxcyc = [&](int blocks) {
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(hc, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < blocks; ++i) if (hc[i] > mx) mx = (double)hc[i];
        return mx;
    };
    const int it = 4000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        const int per_sm = warps * 32 / threads;
        k_mma_rate<9><<<sms * per_sm, threads>>>(out, it, cyc);
        double c = maxcyc(sms * per_sm);
        printf("mma.sync m16n8k8 tf32, %2d warps/SM, 9 chains: %10.0f cyc -> %.3f MMA/clk/SM = %.0f MAC/clk/SM\n", warps, c,
               (double)warps * 9 * it / c, (double)warps * 9 * it / c * 1024);
    }
    for (int warps : {8, 16}) {
        k_mma_rate<3><<<sms, warps * 32>>>(out, it, cyc);
        double c = maxcyc(sms);
        printf("mma.sync m16n8k8 tf32, %2d warps/SM, 3 chains: %10.0f cyc -> %.3f MMA/clk/SM\n", warps, c, (double)warps * 3 * it / c);
    }
    // Gram prototype
    std::vector<float> hX(E * RS);
    srand(1);
    for (auto &v : hX) v = (float)rand() / RAND_MAX;
    float *dX, *dG;
    CHECK(cudaMalloc(&dX, sizeof(float) * E * RS));
    CHECK(cudaMalloc(&dG, sizeof(float) * K * K));
    CHECK(cudaMemcpy(dX, hX.data(), sizeof(float) * E * RS, cudaMemcpyHostToDevice));
    std::vector<double> ref(K * K, 0.0);
    for (int e = 0; e < E; ++e)
        for (int r = 0; r < K; ++r)
            for (int c = 0; c < K; ++c) ref[r * K + c] += (double)hX[e * RS + r] * (double)hX[e * RS + c];
    for (int split : {3, 1}) {
        CHECK(cudaMemset(dG, 0, sizeof(float) * K * K));
        const size_t smem = 16 * E * RS * sizeof(float);
        if (split == 3) { CHECK(cudaFuncSetAttribute(k_gram<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_gram<3><<<1, 32, smem>>>(dX, dG, 1, cyc); }
        else { CHECK(cudaFuncSetAttribute(k_gram<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_gram<1><<<1, 32, smem>>>(dX, dG, 1, cyc); }
        CHECK(cudaDeviceSynchronize());
        std::vector<float> hG(K * K);
        CHECK(cudaMemcpy(hG.data(), dG, sizeof(float) * K * K, cudaMemcpyDeviceToHost));
        double num = 0, den = 0, mx = 0;
        for (int r = 0; r < K; ++r)
            for (int c = r; c < K; ++c) {
                const double d = hG[r * K + c] - ref[r * K + c];
                num += d * d; den += ref[r * K + c] * ref[r * K + c];
                mx = fmax(mx, fabs(d) / fabs(ref[r * K + c]));
            }
        printf("Gram prototype split=%d: upper-triangle rel Frobenius error %.3e, max rel %.3e (32 entries)\n", split, sqrt(num / den), mx);
        for (int warps : {8, 12, 16}) {
            const int iters = 500;
            if (split == 3) k_gram<3><<<sms, warps * 32, warps * E * RS * sizeof(float)>>>(dX, dG, iters, cyc);
            else k_gram<1><<<sms, warps * 32, warps * E * RS * sizeof(float)>>>(dX, dG, iters, cyc);
            double c = maxcyc(sms);
            printf("Gram prototype split=%d, %2d warps/SM: %10.0f cyc -> %.3f clk per entry per SM\n", split, warps, c,
                   c / ((double)warps * E * iters));
        }
    }

    {
        for (int warps : {8, 16}) {
            k_mma_rate_k4<9><<<sms, warps * 32>>>(out, it, cyc);
            double c = maxcyc(sms);
            printf("mma.sync m16n8k4 tf32, %2d warps/SM, 9 chains: %.3f MMA/clk/SM\n", warps, (double)warps * 9 * it / c);
        }
        std::vector<float> hX2(EV * RS);
        for (auto &v : hX2) v = (float)rand() / RAND_MAX;
        float *dX2;
        CHECK(cudaMalloc(&dX2, sizeof(float) * EV * RS));
        CHECK(cudaMemcpy(dX2, hX2.data(), sizeof(float) * EV * RS, cudaMemcpyHostToDevice));
        run_gram2<0>(sms, dX2, dG, cyc, hX2);
        run_gram2<1>(sms, dX2, dG, cyc, hX2);
        run_gram2<2>(sms, dX2, dG, cyc, hX2);
        run_gram2<3>(sms, dX2, dG, cyc, hX2);
    }
    return 0;
}
