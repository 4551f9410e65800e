// f_update_tiled.cuh -- K1, the headline kernel: register-tiled symmetric Gram build +
// fp64 Cholesky per series (fp32 storage build only).
//
// Replaces the hot loop of l2r_ls_pY_IX_chol::solve (reference trmf.cpp:382-395): per
// observed entry the reference does k(k+1)/2 + k scalar multiply-adds into a k x k buffer.
// Here one CTA (8 "Gram" warps + 1 "rhs" warp) owns one series at a time:
//
//  * the observed rows x_i (k floats each, gathered through the by-series index list)
//    are staged into shared memory by cp.async (16 B per request, no register
//    staging), 3 stages deep, one __syncthreads per tile; the index stream is
//    prefetched into registers one tile ahead so the gather never waits on it;
//  * the k indices are split into NB = KP/8 sets R(t) of 8 (two 16-byte chunks, t and
//    t+NB, so that lanes of a group read distinct bank groups); the upper triangle of
//    the Gram is the B = NB(NB+1)/2 blocks R(bi) x R(bj), bi <= bj; a Gram thread is
//    (group g, block b): it keeps the 8x8 block in 64 fp32 registers and, per entry,
//    loads 4 x LDS.128 and issues 64 FFMA.  G = 256/B groups work on different entries
//    of the same tile, so a fp32 partial sum never covers more than 128 entries;
//  * every FL tiles the G partial blocks are reduced through shared memory into ONE
//    fp64 Gram per CTA (each matrix element has a single owner thread: no atomics,
//    fixed order => bitwise reproducible); the rhs warp does the same for sum Y_ij x_i;
//  * epilogue: + lambda I, CTA-wide fp64 Cholesky + substitutions (common.cuh), store
//    the k results.  Two CTAs per SM: one CTA's epilogue hides behind the other's FMAs.
//
// Tensor cores are deliberately not used: the operands would need M padded 40 -> 128
// and a 3 x TF32 split to stay inside the 1e-5 parity bar, which costs as many tensor
// cycles per entry (~7.5 clk/SM) as this FFMA formulation ("why no tcgen05").
#pragma once
#include "common.cuh"

#ifdef TRMF_F32

namespace ft {

constexpr int NT = 288;       // 8 Gram warps + 1 rhs warp
constexpr int NGRAM = 256;

template <int NB> struct Cfg {
    static constexpr int KP = 8 * NB;
    static constexpr int B = NB * (NB + 1) / 2;
    static constexpr int G = NGRAM / B;
    static constexpr int U = NB == 1 ? 1 : NB == 2 ? 2 : NB == 3 ? 3 : NB == 4 ? 5 : NB == 5 ? 6 : NB == 8 ? 6 : 8;
    static constexpr int ET = G * U;           // entries per tile
    static constexpr int RS = KP + 4;          // smem row stride (floats): odd number of 16-B chunks
    static constexpr int STAGES = 3;
    static constexpr int FL = (128 / U) > 0 ? (128 / U) : 1;   // tiles between fp32 -> fp64 flushes
    static constexpr int FBUF = G * B * 32;    // floats: half of every lane's 8x8 block
    static constexpr int RBUF = 32 * KP;       // floats: rhs partials of the 32 rhs lanes
    static constexpr int VALS = (STAGES * ET + 3) / 4 * 4;   // staged Y values, padded so fbuf stays 16-B aligned
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NB>
static size_t smem_bytes(int k) {
    typedef Cfg<NB> C;
    size_t dbl = (size_t)(k + 1) * (k + 1) + k;
    dbl = (dbl + 1) & ~(size_t)1;   // keep the float region 16-B aligned
    return dbl * sizeof(double) + sizeof(float) * ((size_t)C::STAGES * C::ET * C::RS + C::VALS + C::FBUF + C::RBUF);
}

// index of member m (0..7) of set R(t): chunk t then chunk t + NB
template <int NB> __device__ __forceinline__ int set_index(int t, int m) { return m < 4 ? 4 * t + m : 4 * (t + NB) + (m - 4); }

// SOLVE = true : F-update -- solve (Gram + lambda I) f = rhs and store the k results in F[j].
// SOLVE = false: "store" mode used by the X-update (rows = time stamps, X = the series factor):
//                the k x k Gram (full symmetric square, fp32) goes to Gout[j] and the rhs to
//                F[j]; rows without entries get zeros.  The Hessian-vector products of the CG
//                solve then read these Grams instead of re-walking Omega (x_update.cuh).
template <int NB, bool SOLVE>
__global__ void __launch_bounds__(NT, 2)
f_update_tiled_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                      const float *__restrict__ X, float *__restrict__ F, float *__restrict__ Gout, int k, double lambda,
                      uint32_t nseries, unsigned *__restrict__ queue) {
    typedef Cfg<NB> C;
    constexpr int KP = C::KP, B = C::B, G = C::G, U = C::U, ET = C::ET, RS = C::RS, STAGES = C::STAGES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = k + 1;
    double *A = reinterpret_cast<double *>(smem_raw);            // (k+1) x ld: lower triangle + rhs row
    double *dinv = A + (size_t)(k + 1) * ld;
    size_t dbl = (size_t)(k + 1) * ld + k;
    dbl = (dbl + 1) & ~(size_t)1;
    float *tiles = reinterpret_cast<float *>(A + dbl);           // STAGES x ET x RS
    float *vals = tiles + (size_t)STAGES * ET * RS;              // STAGES x ET
    float *fbuf = vals + C::VALS;                                // G x 8 x B float4
    float *rbuf = fbuf + C::FBUF;                                // KP/4 x 32 float4
    __shared__ unsigned next_series;

    const int tid = threadIdx.x;
    const int CH = k >> 2;                                       // 16-byte chunks per factor row (k % 4 == 0)
    const bool is_gram = tid < NGRAM;
    const int g = tid / B, b = tid - g * B;
    const bool active = is_gram && g < G;
    int bi = 0;
    {
        int rem = b;
        while (rem >= NB - bi) { rem -= NB - bi; ++bi; }
        // bj = bi + rem
    }
    int bj;
    {
        int rem = b, t = 0;
        while (rem >= NB - t) { rem -= NB - t; ++t; }
        bj = t + rem;
    }
    const int lane = tid & 31;

    // zero every staging row once: padding columns (k..KP+3) are never written by the copies
    for (int p = tid; p < STAGES * ET * RS; p += NT) tiles[p] = 0.f;

    constexpr int NCOPY = (ET * (KP / 4) + NT - 1) / NT;
    uint32_t nidx[NCOPY];

    if (tid == 0) next_series = atomicAdd(queue, 1u);
    __syncthreads();
    uint32_t j = next_series;

    while (j < nseries) {
        const uint64_t lo = ptr[j], hi = ptr[j + 1];
        const uint64_t nnz = hi - lo;
        if (nnz != 0) {
            const int ntiles = (int)((nnz + ET - 1) / ET);
            for (int p = tid; p < (k + 1) * ld; p += NT) A[p] = 0.0;

            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.f;
            // the rhs warp keeps its KP partial sums in the same registers: racc[t] == acc[t >> 3][t & 7]
#define RACC(t) acc[(t) >> 3][(t) & 7]

            auto tile_count = [&](int tt) -> int {
                const uint64_t base = lo + (uint64_t)tt * ET;
                return (int)((hi - base) < (uint64_t)ET ? (hi - base) : (uint64_t)ET);
            };
            auto load_idx = [&](int tt) {
                const uint64_t base = lo + (uint64_t)tt * ET;
                const int ncp = tile_count(tt) * CH;
#pragma unroll
                for (int q = 0; q < NCOPY; ++q) {
                    const int c = tid + q * NT;
                    nidx[q] = c < ncp ? __ldg(idx + base + c / CH) : 0u;
                }
            };
            auto issue = [&](int tt) {   // uses nidx loaded for tile tt
                const int stage = tt % STAGES;
                const uint64_t base = lo + (uint64_t)tt * ET;
                const int cnt = tile_count(tt);
                const int ncp = cnt * CH;
                float *dst = tiles + (size_t)stage * ET * RS;
#pragma unroll
                for (int q = 0; q < NCOPY; ++q) {
                    const int c = tid + q * NT;
                    if (c < ncp) {
                        const int e = c / CH, ch = c - e * CH;
                        cp_async16(dst + e * RS + ch * 4, X + (size_t)nidx[q] * k + ch * 4);
                    }
                }
                for (int e = tid; e < cnt; e += NT) cp_async4(vals + stage * ET + e, val + base + e);
                if (cnt < ET)   // tail tile: stale rows from an earlier tile must read as zero
                    for (int p = cnt * RS + tid; p < ET * RS; p += NT) dst[p] = 0.f;
            };
            auto flush = [&]() {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (active) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int jh = 0; jh < 2; ++jh) {
                                const int s4 = i * 2 + jh;
                                float4 v = make_float4(acc[h * 4 + i][jh * 4 + 0], acc[h * 4 + i][jh * 4 + 1],
                                                       acc[h * 4 + i][jh * 4 + 2], acc[h * 4 + i][jh * 4 + 3]);
                                *reinterpret_cast<float4 *>(fbuf + ((size_t)(g * 8 + s4) * B + b) * 4) = v;
                            }
                    }
                    if (h == 0 && !is_gram) {
#pragma unroll
                        for (int c4 = 0; c4 < KP / 4; ++c4)
                            *reinterpret_cast<float4 *>(rbuf + (c4 * 32 + lane) * 4) =
                                make_float4(RACC(4 * c4), RACC(4 * c4 + 1), RACC(4 * c4 + 2), RACC(4 * c4 + 3));
                    }
                    __syncthreads();
                    for (int u = tid; u < 32 * B; u += NT) {
                        const int c4 = u & 3, t = u >> 2;
                        const int bb = t % B, s4 = t / B;
                        int ti = 0, rem = bb;
                        while (rem >= NB - ti) { rem -= NB - ti; ++ti; }
                        const int tj = ti + rem;
                        const int r = set_index<NB>(ti, h * 4 + (s4 >> 1));
                        const int c = set_index<NB>(tj, (s4 & 1) * 4 + c4);
                        if (r < k && c < k && !(ti == tj && r > c)) {
                            double sum = 0.0;
#pragma unroll 4
                            for (int gg = 0; gg < G; ++gg) sum += (double)fbuf[(size_t)gg * 32 * B + u];
                            const int hi_ = r > c ? r : c, lo_ = r > c ? c : r;
                            A[hi_ * ld + lo_] += sum;
                        }
                    }
                    if (h == 0 && tid < k) {
                        double sum = 0.0;
#pragma unroll 8
                        for (int l = 0; l < 32; ++l) sum += (double)rbuf[((tid >> 2) * 32 + l) * 4 + (tid & 3)];
                        A[k * ld + tid] += sum;
                    }
                    __syncthreads();
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.f;
            };

            // ---- pipeline prologue: tiles 0 .. STAGES-2 in flight, indices of tile STAGES-1 in registers ----
#pragma unroll
            for (int s = 0; s < STAGES - 1; ++s) {
                if (s < ntiles) { load_idx(s); issue(s); }
                cp_async_commit();
            }
            if (STAGES - 1 < ntiles) load_idx(STAGES - 1);

            for (int t = 0; t < ntiles; ++t) {
                cp_async_wait<STAGES - 2>();
                __syncthreads();   // tile t landed for everyone; everyone is done with tile t-1
                if (t + STAGES - 1 < ntiles) issue(t + STAGES - 1);
                cp_async_commit();
                if (t + STAGES < ntiles) load_idx(t + STAGES);

                const float *tb = tiles + (size_t)(t % STAGES) * ET * RS;
                if (active) {
                    const float *pa = tb + 4 * bi, *pb = tb + 4 * bj;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int e = g + u * G;
                        const float4 a0 = *reinterpret_cast<const float4 *>(pa + e * RS);
                        const float4 a1 = *reinterpret_cast<const float4 *>(pa + e * RS + 4 * NB);
                        const float4 b0 = *reinterpret_cast<const float4 *>(pb + e * RS);
                        const float4 b1 = *reinterpret_cast<const float4 *>(pb + e * RS + 4 * NB);
                        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(a[i], bv[jj], acc[i][jj]);
                    }
                } else if (!is_gram) {
                    const int cnt = tile_count(t);
                    const float *tv = vals + (t % STAGES) * ET;
                    for (int e = lane; e < cnt; e += 32) {
                        const float y = tv[e];
                        const float *row = tb + e * RS;
#pragma unroll
                        for (int c4 = 0; c4 < KP / 4; ++c4) {
                            const float4 w = *reinterpret_cast<const float4 *>(row + 4 * c4);
                            RACC(4 * c4 + 0) = fmaf(y, w.x, RACC(4 * c4 + 0));
                            RACC(4 * c4 + 1) = fmaf(y, w.y, RACC(4 * c4 + 1));
                            RACC(4 * c4 + 2) = fmaf(y, w.z, RACC(4 * c4 + 2));
                            RACC(4 * c4 + 3) = fmaf(y, w.w, RACC(4 * c4 + 3));
                        }
                    }
                }
                if ((t + 1) % C::FL == 0 && t + 1 < ntiles) flush();
            }
            cp_async_wait<0>();
            flush();   // ends with __syncthreads
            if (SOLVE) {
                if (tid < k) A[tid * ld + tid] += lambda;      // trmf.cpp:393
                block_chol_solve(A, ld, dinv, k);              // starts and ends with __syncthreads
                if (tid < k) F[(size_t)j * k + tid] = (float)A[k * ld + tid];
            } else {
                float *Gj = Gout + (size_t)j * k * k;
                for (int p = tid; p < k * k; p += NT) {
                    const int r = p / k, c = p - r * k;
                    Gj[p] = (float)(r >= c ? A[r * ld + c] : A[c * ld + r]);
                }
                if (tid < k) F[(size_t)j * k + tid] = (float)A[k * ld + tid];
            }
#undef RACC
        } else if (!SOLVE) {
            float *Gj = Gout + (size_t)j * k * k;
            for (int p = tid; p < k * k; p += NT) Gj[p] = 0.f;
            if (tid < k) F[(size_t)j * k + tid] = 0.f;
        }
        __syncthreads();
        if (tid == 0) next_series = atomicAdd(queue, 1u);
        __syncthreads();
        j = next_series;
    }
}

}   // namespace ft

static inline bool f_update_tiled_supported(int k) { return k >= 4 && k <= 64 && (k % 4) == 0; }

// returns 0 on success
template <bool SOLVE>
static inline int f_update_tiled_launch(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *idx, const V *val,
                                        const V *X, V *F, V *Gout, int k, double lambda, uint32_t nseries, unsigned *queue,
                                        unsigned long long *launches) {
    const int NB = (k + 7) / 8;
    const unsigned grid = (unsigned)(nseries < (uint32_t)(2 * num_sms) ? nseries : (uint32_t)(2 * num_sms));
    if (cudaMemsetAsync(queue, 0, sizeof(unsigned), st) != cudaSuccess) return 1;
#define FT_CASE(N)                                                                                              \
    case N: {                                                                                                   \
        const size_t smem = ft::smem_bytes<N>(k);                                                               \
        auto kfn = ft::f_update_tiled_kernel<N, SOLVE>;                                                         \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        kfn<<<grid ? grid : 1, ft::NT, smem, st>>>(ptr, idx, val, X, F, Gout, k, lambda, nseries, queue);       \
        break;                                                                                                  \
    }
    switch (NB) {
        FT_CASE(1) FT_CASE(2) FT_CASE(3) FT_CASE(4) FT_CASE(5) FT_CASE(6) FT_CASE(7) FT_CASE(8)
        default: return 1;
    }
#undef FT_CASE
    ++*launches;
    return cudaGetLastError() != cudaSuccess;
}

#else   // float64 build: the generic kernel (fp64 FMAs) is the parity path

static inline bool f_update_tiled_supported(int) { return false; }
template <bool SOLVE>
static inline int f_update_tiled_launch(cudaStream_t, int, const uint64_t *, const uint32_t *, const V *, const V *, V *, V *, int,
                                        double, uint32_t, unsigned *, unsigned long long *) { return 1; }
#endif
