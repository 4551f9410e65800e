// microbench_cg.cu -- prototype + measurement harness for the X-update's CG loop at small T*k (round-2 groundwork).
//
// Finding it follows up on ("Device-side CG control"; profiles/r01_rolling_phases.json): at the electricity /
// traffic shapes (T*k = 4-5e5, 24-26 lags) one CG step of the reference's trcg (rf_tron.h:412-505) is 5-7 tiny kernels
// whose launch-latency floors (4-7 us each) add up to 47-84 us, while the data they touch (a few MB, L2 resident) would
// take a few us.  Removing the per-step host round trip (chunked, gated launches) only bought 6-15 %.
//
// This file measures, on the dense-mode step (Hd = lI d + lAR A^T A d + d HTH; trmf.cpp:128-149, 209-214), three ways
// of driving the same arithmetic:
//   (A) one kernel per operation, host reads the scalar after every step       (round-1 first version)
//   (B) the same kernels, all steps enqueued, gated by a device flag           (what the library does now, chunk = all)
//   (C) ONE cooperative persistent kernel for the whole solve: 4 grid-wide barriers per step, reductions through a
//       per-block partial array summed redundantly by every block in a fixed order (deterministic, no atomics)
// and checks that (C) reproduces (A): same step count, iterates equal to rounding.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=false -o tools/microbench_cg tools/microbench_cg.cu
// Run:   tools/microbench_cg [T k L]      (defaults: 26304 20 24, then 10560 40 26)
#include <algorithm>
#include <cmath>
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

namespace cgx = cooperative_groups;
typedef float V;

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Lag { int L, mid; const uint32_t *lags; };

// ---------------------------------------------------------------------------------------------------------------
// device pieces shared by all three schemes (same arithmetic, element p = i*k + t)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rho_at(const V *S, const V *th, const Lag &ls, size_t p, size_t i, int t, int k) {
    if (i < (size_t)ls.mid) return 0.0;
    double r = (double)S[p];
    const V *tht = th + (size_t)ls.L * t;
    for (int l = 0; l < ls.L; ++l) r -= (double)tht[l] * (double)S[p - (size_t)ls.lags[l] * k];
    return r;
}
__device__ __forceinline__ double adj_at(const double *rho, const V *th, const Lag &ls, size_t p, size_t j, int t, int k, size_t T) {
    double a = rho[p];
    const V *tht = th + (size_t)ls.L * t;
    for (int l = 0; l < ls.L; ++l) {
        const size_t jj = j + ls.lags[l];
        if (jj < T) a -= (double)tht[l] * rho[jj * k + t];
    }
    return a;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum, result in every thread; red = shared double[33]
__device__ __forceinline__ double block_sum_all(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < (int)((blockDim.x + 31) >> 5) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// ---------------------------------------------------------------------------------------------------------------
// schemes (A) / (B): one kernel per operation; gate == nullptr -> always run
// ---------------------------------------------------------------------------------------------------------------
#define GATE(g) do { if ((g) != nullptr && *(g) == 0) return; } while (0)
#define GS for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x)

__global__ void k_rho(const V *S, const V *th, Lag ls, double *rho, size_t T, int k, const int *gate) {
    GATE(gate);
    const size_t total = T * k;
    GS { const size_t i = p / k; rho[p] = rho_at(S, th, ls, p, i, (int)(p - i * k), k); }
}
__global__ void k_apply(const V *S, const V *th, Lag ls, const double *rho, V *out, size_t T, int k, double lI, double lAR, const int *gate) {
    GATE(gate);
    const size_t total = T * k;
    GS { const size_t j = p / k; out[p] = (V)(lI * (double)S[p] + lAR * adj_at(rho, th, ls, p, j, (int)(p - j * k), k, T)); }
}
__global__ void k_hth(const V *S, const double *HTH, V *out, size_t T, int k, const int *gate) {   // out += S * HTH
    GATE(gate);
    extern __shared__ double sh[];
    for (int q = threadIdx.x; q < k * k; q += blockDim.x) sh[q] = HTH[q];
    __syncthreads();
    const size_t total = T * k;
    GS {
        const size_t i = p / k; const int t = (int)(p - i * k);
        double a = (double)out[p];
        for (int u = 0; u < k; ++u) a += (double)S[i * k + u] * sh[u * k + t];
        out[p] = (V)a;
    }
}
// two-level deterministic dot: partials, then the last block (ticket) sums them in index order
__global__ void k_dot(const V *a, const V *b, size_t total, double *part, unsigned *ticket, double *out, const int *gate) {
    GATE(gate);
    __shared__ double red[33];
    __shared__ bool last;
    double v = 0.0;
    GS v += (double)a[p] * (double)b[p];
    v = block_sum_all(v, red);
    if (threadIdx.x == 0) { part[blockIdx.x] = v; __threadfence(); last = atomicAdd(ticket, 1u) == gridDim.x - 1; }
    __syncthreads();
    if (last) {
        double t = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(part + i);
        t = block_sum_all(t, red);
        if (threadIdx.x == 0) { *out = t; *ticket = 0u; }
    }
}
__global__ void k_step1(V *s, V *r, const V *d, const V *Hd, size_t total, double *scal, double *part, unsigned *ticket, const int *gate) {
    GATE(gate);                                   // scal: [0] rTr, [1] dHd, [2] rnew, [3] cgtol, [4] steps, [5] rnorm
    __shared__ double red[33];
    __shared__ bool last;
    const V a = (V)(scal[0] / scal[1]);
    double v = 0.0;
    GS { s[p] = s[p] + a * d[p]; const V rv = r[p] - a * Hd[p]; r[p] = rv; v += (double)rv * (double)rv; }
    v = block_sum_all(v, red);
    if (threadIdx.x == 0) { part[blockIdx.x] = v; __threadfence(); last = atomicAdd(ticket, 1u) == gridDim.x - 1; }
    __syncthreads();
    if (last) {
        double t = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(part + i);
        t = block_sum_all(t, red);
        if (threadIdx.x == 0) { scal[2] = t; *ticket = 0u; }
    }
}
__global__ void k_step2(V *d, const V *r, size_t total, double *scal, int *ctl, int j) {
    if (ctl != nullptr && ctl[j] == 0) return;
    const V bm1 = (V)(scal[2] / scal[0]) - (V)1;
    GS { V dv = d[p]; dv = dv + bm1 * dv; d[p] = dv + r[p]; }
}
__global__ void k_advance(double *scal, int *ctl, int j) {   // rTr <- rnew, loop head for step j + 1
    if (ctl != nullptr && ctl[j] == 0) return;
    scal[0] = scal[2];
    scal[4] = (double)j;
    scal[5] = sqrt(scal[2]);
    if (ctl != nullptr) ctl[j + 1] = scal[5] <= scal[3] ? 0 : 1;
}
__global__ void k_init(const V *g, V *s, V *r, V *d, size_t total) { GS { s[p] = 0; r[p] = -g[p]; d[p] = -g[p]; } }

// ---------------------------------------------------------------------------------------------------------------
// scheme (C): the whole solve in one cooperative kernel
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double grid_sum(cgx::grid_group &grid, double v, double *part, double *red) {
    v = block_sum_all(v, red);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    grid.sync();
    double t = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(part + i);
    return block_sum_all(t, red);      // every block adds the same numbers in the same order: one value grid-wide
}

__global__ void __launch_bounds__(256)
k_cg_coop(V *s, V *r, V *d, V *Hd, double *rho, const V *__restrict__ g, const V *__restrict__ th, Lag ls,
          const double *__restrict__ HTH, size_t T, int k, double lI, double lAR, int max_cg, double eps_cg,
          double *part /* 2 * gridDim */, double *scal) {
    cgx::grid_group grid = cgx::this_grid();
    extern __shared__ double sh[];            // k*k of HTH
    __shared__ double red[33];
    for (int q = threadIdx.x; q < k * k; q += blockDim.x) sh[q] = HTH[q];
    const size_t total = T * (size_t)k;
    double *partA = part, *partB = part + gridDim.x;
    double v = 0.0;
    GS { const V gv = g[p]; s[p] = 0; r[p] = -gv; d[p] = -gv; v += (double)gv * (double)gv; }
    double rTr = grid_sum(grid, v, partA, red);
    const double cgtol = eps_cg * sqrt(rTr);
    int steps = 0;
    double rnorm = sqrt(rTr);
    while (rnorm > cgtol && steps < max_cg) {
        ++steps;
        // A: AR residual of the direction (reads d written by other blocks in phase D: ordered by the barrier below / at loop end)
        GS { const size_t i = p / k; rho[p] = rho_at(d, th, ls, p, i, (int)(p - i * k), k); }
        grid.sync();
        // B: Hd = lI d + lAR A^T rho + d HTH, and d'Hd
        v = 0.0;
        GS {
            const size_t i = p / k; const int t = (int)(p - i * k);
            const V base = (V)(lI * (double)d[p] + lAR * adj_at(rho, th, ls, p, i, t, k, T));
            double a = (double)base;
            for (int u = 0; u < k; ++u) a += (double)d[i * k + u] * sh[u * k + t];
            const V hv = (V)a;
            Hd[p] = hv;
            v += (double)d[p] * (double)hv;
        }
        const double dHd = grid_sum(grid, v, partB, red);
        // C: s += alpha d, r -= alpha Hd, <r, r>
        const V a = (V)(rTr / dHd);
        v = 0.0;
        GS { s[p] = s[p] + a * d[p]; const V rv = r[p] - a * Hd[p]; r[p] = rv; v += (double)rv * (double)rv; }
        const double rnew = grid_sum(grid, v, partA, red);
        // D: d <- d + (beta - 1) d + r      (rf_tron.h:495-501)
        const V bm1 = (V)(rnew / rTr) - (V)1;
        GS { V dv = d[p]; dv = dv + bm1 * dv; d[p] = dv + r[p]; }
        rTr = rnew;
        rnorm = sqrt(rTr);
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal[0] = rTr; scal[4] = (double)steps; scal[5] = rnorm; }
}

// ---------------------------------------------------------------------------------------------------------------
struct Problem {
    size_t T; int k, L;
    V *g, *s, *r, *d, *Hd, *th; double *rho, *HTH, *scal, *part; unsigned *ticket; int *ctl; uint32_t *lags; Lag ls;
    unsigned grid;
};

static void run_A(Problem &P, int max_cg, double eps, double lI, double lAR, bool host_sync, int *steps_out) {
    const size_t total = P.T * P.k;
    const size_t sm = sizeof(double) * P.k * P.k;
    k_init<<<P.grid, 256>>>(P.g, P.s, P.r, P.d, total);
    k_dot<<<P.grid, 256>>>(P.g, P.g, total, P.part, P.ticket, P.scal + 0, nullptr);
    double h[6];
    CHECK(cudaMemcpy(h, P.scal, sizeof h, cudaMemcpyDeviceToHost));
    const double cgtol = eps * std::sqrt(h[0]);
    h[3] = cgtol; h[4] = 0; h[5] = std::sqrt(h[0]);
    CHECK(cudaMemcpy(P.scal, h, sizeof h, cudaMemcpyHostToDevice));
    int ctl0[32] = {0};
    ctl0[1] = std::sqrt(h[0]) > cgtol ? 1 : 0;
    CHECK(cudaMemcpy(P.ctl, ctl0, sizeof ctl0, cudaMemcpyHostToDevice));
    int steps = 0;
    for (int j = 1; j <= max_cg; ++j) {
        const int *gate = host_sync ? nullptr : P.ctl + j;
        if (host_sync) {
            CHECK(cudaMemcpy(h, P.scal, sizeof h, cudaMemcpyDeviceToHost));   // the per-step round trip
            if (std::sqrt(h[0]) <= cgtol) break;
        }
        k_rho<<<P.grid, 256>>>(P.d, P.th, P.ls, P.rho, P.T, P.k, gate);
        k_apply<<<P.grid, 256>>>(P.d, P.th, P.ls, P.rho, P.Hd, P.T, P.k, lI, lAR, gate);
        k_hth<<<P.grid, 256, sm>>>(P.d, P.HTH, P.Hd, P.T, P.k, gate);
        k_dot<<<P.grid, 256>>>(P.d, P.Hd, total, P.part, P.ticket, P.scal + 1, gate);
        k_step1<<<P.grid, 256>>>(P.s, P.r, P.d, P.Hd, total, P.scal, P.part, P.ticket, gate);
        k_step2<<<P.grid, 256>>>(P.d, P.r, total, P.scal, host_sync ? nullptr : P.ctl, j);
        k_advance<<<1, 1>>>(P.scal, host_sync ? nullptr : P.ctl, j);
        ++steps;
    }
    CHECK(cudaMemcpy(h, P.scal, sizeof h, cudaMemcpyDeviceToHost));
    *steps_out = (int)h[4];
    (void)steps;
}

static void run_C(Problem &P, int max_cg, double eps, double lI, double lAR, unsigned cgrid, int *steps_out) {
    size_t T = P.T; int k = P.k;
    void *args[] = {&P.s, &P.r, &P.d, &P.Hd, &P.rho, &P.g, &P.th, &P.ls, &P.HTH, &T, &k, &lI, &lAR, &max_cg, &eps, &P.part, &P.scal};
    CHECK(cudaLaunchCooperativeKernel((void *)k_cg_coop, dim3(cgrid), dim3(256), args, sizeof(double) * k * k, 0));
    double h[6];
    CHECK(cudaMemcpy(h, P.scal, sizeof h, cudaMemcpyDeviceToHost));
    *steps_out = (int)h[4];
}

static void bench(size_t T, int k, int L) {
    Problem P; P.T = T; P.k = k; P.L = L;
    const size_t total = T * k;
    std::vector<V> g(total), th((size_t)L * k);
    std::vector<double> HTH((size_t)k * k);
    std::vector<uint32_t> lags(L);
    srand(7);
    auto rnd = []() { return (double)rand() / RAND_MAX - 0.5; };
    for (auto &x : g) x = (V)rnd();
    for (auto &x : th) x = (V)(0.2 * rnd());
    for (int l = 0; l < L; ++l) lags[l] = l + 1;
    if (L >= 26) { lags[L - 2] = 168; lags[L - 1] = 336; }
    // HTH = B^T B / scale: SPD, well conditioned enough for CG to take a good number of steps
    std::vector<double> B((size_t)2 * k * k);
    for (auto &x : B) x = rnd();
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
            double s = 0; for (int q = 0; q < 2 * k; ++q) s += B[(size_t)q * k + a] * B[(size_t)q * k + b];
            HTH[(size_t)a * k + b] = 40.0 * s;
        }
    CHECK(cudaMalloc(&P.g, total * sizeof(V))); CHECK(cudaMalloc(&P.s, total * sizeof(V))); CHECK(cudaMalloc(&P.r, total * sizeof(V)));
    CHECK(cudaMalloc(&P.d, total * sizeof(V))); CHECK(cudaMalloc(&P.Hd, total * sizeof(V))); CHECK(cudaMalloc(&P.rho, total * sizeof(double)));
    CHECK(cudaMalloc(&P.th, th.size() * sizeof(V))); CHECK(cudaMalloc(&P.HTH, HTH.size() * sizeof(double)));
    CHECK(cudaMalloc(&P.scal, 8 * sizeof(double))); CHECK(cudaMalloc(&P.part, 2 * 4096 * sizeof(double)));
    CHECK(cudaMalloc(&P.ticket, sizeof(unsigned))); CHECK(cudaMalloc(&P.ctl, 32 * sizeof(int))); CHECK(cudaMalloc(&P.lags, L * sizeof(uint32_t)));
    CHECK(cudaMemset(P.ticket, 0, sizeof(unsigned))); CHECK(cudaMemset(P.scal, 0, 8 * sizeof(double)));
    CHECK(cudaMemcpy(P.g, g.data(), total * sizeof(V), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(P.th, th.data(), th.size() * sizeof(V), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(P.HTH, HTH.data(), HTH.size() * sizeof(double), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(P.lags, lags.data(), L * sizeof(uint32_t), cudaMemcpyHostToDevice));
    P.ls.L = L; P.ls.mid = (int)lags.back(); P.ls.lags = P.lags;
    int dev = 0, sms = 0, occ = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    P.grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)sms * 8);
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cg_coop, 256, sizeof(double) * k * k));
    const unsigned cgrid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)sms * std::min(occ, 4));
    const double lI = 0.5, lAR = 50.0, eps = 1e-3;   // eps small so that the solve runs to the 20-step cap
    const int max_cg = 20;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    std::vector<V> sA(total), sC(total);
    int stA = 0, stB = 0, stC = 0;
    float msA = 0, msB = 0, msC = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CHECK(cudaEventRecord(e0)); run_A(P, max_cg, eps, lI, lAR, true, &stA); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        CHECK(cudaEventElapsedTime(&msA, e0, e1));
    }
    CHECK(cudaMemcpy(sA.data(), P.s, total * sizeof(V), cudaMemcpyDeviceToHost));
    for (int rep = 0; rep < 4; ++rep) {
        CHECK(cudaEventRecord(e0)); run_A(P, max_cg, eps, lI, lAR, false, &stB); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        CHECK(cudaEventElapsedTime(&msB, e0, e1));
    }
    for (int rep = 0; rep < 4; ++rep) {
        CHECK(cudaEventRecord(e0)); run_C(P, max_cg, eps, lI, lAR, cgrid, &stC); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        CHECK(cudaEventElapsedTime(&msC, e0, e1));
    }
    CHECK(cudaMemcpy(sC.data(), P.s, total * sizeof(V), cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (size_t p = 0; p < total; ++p) { const double df = (double)sA[p] - (double)sC[p]; num += df * df; den += (double)sA[p] * (double)sA[p]; }
    printf("T=%zu k=%d L=%d  grid %u / coop grid %u (occupancy %d/SM)\n", T, k, L, P.grid, cgrid, occ);
    printf("  (A) kernel per op, host round trip per step : %2d steps  %8.1f us  (%6.1f us/step)\n", stA, 1e3 * msA, 1e3 * msA / std::max(1, stA));
    printf("  (B) same kernels, gated, enqueued at once    : %2d steps  %8.1f us  (%6.1f us/step)\n", stB, 1e3 * msB, 1e3 * msB / std::max(1, stB));
    printf("  (C) one cooperative persistent kernel        : %2d steps  %8.1f us  (%6.1f us/step)\n", stC, 1e3 * msC, 1e3 * msC / std::max(1, stC));
    printf("  |s_A - s_C| / |s_A| = %.2e\n", std::sqrt(num / (den > 0 ? den : 1)));
    cudaFree(P.g); cudaFree(P.s); cudaFree(P.r); cudaFree(P.d); cudaFree(P.Hd); cudaFree(P.rho); cudaFree(P.th); cudaFree(P.HTH);
    cudaFree(P.scal); cudaFree(P.part); cudaFree(P.ticket); cudaFree(P.ctl); cudaFree(P.lags);
}

int main(int argc, char **argv) {
    if (argc == 4) { bench((size_t)atoll(argv[1]), atoi(argv[2]), atoi(argv[3])); return 0; }
    bench(26304, 20, 24);     // electricity shape (BASELINE configs[0])
    bench(10560, 40, 26);     // traffic shape (configs[2])
    bench(10000, 40, 3);      // C2's T*k with its 3 lags
    return 0;
}
