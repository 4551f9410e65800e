// f_update_mma.cuh -- K1 on the warp-level tensor path: per-series Gram by 3xTF32 mma.sync + fp64 Cholesky
// (fp32 storage build only).
//
// Replaces the hot loop of l2r_ls_pY_IX_chol::solve (reference trmf.cpp:382-395): per observed entry the
// reference does k(k+1)/2 + k scalar multiply-adds into a k x k buffer.  The Gram G = sum_e x_e x_e^T is a
// SYRK with M = N = k <= 64 and K = |Omega_j|.  tcgen05 cannot be fed by it (its tile is 128 rows of ONE
// accumulator; a series' Gram has k <= 64 and different series have different K ranges), but the warp-level
// m16n8k8 TF32 tile fits any k: a warp holds the whole upper triangle of the Gram as NT 16x8 accumulator tiles
// (9 at k = 40) and per 8 observed entries issues NT x 3 HMMAs.  Measured on B200
// (tools/microbench_mma.cu): 512 TF32 MAC/clk/SM from mma.sync, 9.5 clk per entry per SM for this loop at
// k = 40, against 20 clk per entry per SM for the FFMA formulation (f_update_tiled.cuh).
//
// Accuracy (the 1e-5 parity bar is on the factors; the fp32 reference build itself is 1e-5 off its fp64 twin):
//  * 3xTF32 split: x = hi + lo with hi = x rounded to TF32 (11 significant bits), lo = x - hi (exact, |lo| <=
//    2^-12 |x|, read by the tensor core through its own truncation to TF32); the product keeps hi*hi + hi*lo
//    + lo*hi and drops lo*lo (2^-24 relative);
//  * the tensor core adds with truncation, so it is only trusted with the 8 entries x 3 terms of one chunk
//    (small terms first); chunks are summed by FADD (round to nearest) for at most 128 entries and those
//    partial sums go into per-warp fp64 accumulators in shared memory.  Measured Gram error against fp64:
//    7.8e-8 relative Frobenius, -4.5e-8 mean (direct tensor-core accumulation over 128 entries: -1.1e-6 bias).
//
// Work decomposition: CTA = NW warps, one series at a time (atomic queue).  The series' entries are cut into
// tiles of 16; warp w takes tiles w, w + NW, ...  Each warp runs its own 3-stage cp.async (LDGSTS) pipeline
// (16 factor rows of k floats per stage, the Y values ride along with 4-byte copies) -- no CTA barrier inside
// a series.  Fragment loads are bank-conflict free because the staging row stride RS = 8 (mod 16) floats.
// At the end of a series the NW fp64 partials are added in warp order (bitwise reproducible), lambda goes on
// the diagonal and the CTA runs the blocked fp64 Cholesky of common.cuh.
#pragma once
#include "common.cuh"

#ifdef TRMF_F32

namespace fm {

constexpr int ET = 16;        // entries per tile (two m16n8k8 K-steps)
constexpr int FLUSH = 8;      // tiles between fp32 -> fp64 flushes (128 entries)

template <int K> struct Cfg {
    static constexpr int NC = (K + 7) / 8;          // 8-wide chunks of the factor index
    static constexpr int MT = (NC + 1) / 2;         // 16-row accumulator tiles
    static constexpr int ntiles_() { int n = 0; for (int mt = 0; mt < MT; ++mt) n += NC - 2 * mt; return n; }
    static constexpr int NT = ntiles_();            // upper-triangle 16x8 tiles (mt, nt), nt >= 2 mt
    static constexpr int CH = K / 4;                // 16-byte pieces of a factor row
    static constexpr int RS = (NC & 1) ? 8 * NC : 8 * NC + 8;   // staging row stride: >= 8 NC and = 8 (mod 16)
    static constexpr int STAGES = 3;
    static constexpr int NQ = (ET * CH + 31) / 32;  // warp-wide LDGSTS per tile
    static constexpr int STAGE_FLOATS = ET * RS;
    static constexpr int ld = K + 1;
    static constexpr size_t a_bytes = sizeof(double) * ((size_t)(K + 1) * (K + 1) + K);
    static constexpr size_t part_bytes(int nw) { return sizeof(double) * (size_t)nw * (NT * 128 + 8 * NC); }
    static constexpr size_t stage_bytes(int nw) {
        const size_t s = sizeof(float) * (size_t)nw * STAGES * (STAGE_FLOATS + ET);
        return ((s > a_bytes ? s : a_bytes) + 15) & ~(size_t)15;
    }
    static constexpr size_t smem(int nw) { return part_bytes(nw) + stage_bytes(nw); }
};

__device__ __forceinline__ void cp_async16(float *smem, const float *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(float *smem, const float *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x = hi + lo: hi = x rounded (half away) to 11 significant bits, lo = the exact remainder
__device__ __forceinline__ void split_tf32(const float v, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
// d (16x8, fp32) = a (16x8, tf32, row) * b (8x8, tf32, col) + c; fragment layout of PTX mma.m16n8k8:
// a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); b0 (t, g) b1 (t+4, g); c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// with g = lane / 4, t = lane % 4.
__device__ __forceinline__ void mma_zero(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}
__device__ __forceinline__ void mma_acc(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// SOLVE = true : F-update -- solve (Gram + lambda I) f = rhs and store the k results in F[j].
// SOLVE = false: "store" mode used by the X-update (rows = time stamps, X = the series factor): the k x k
//                Gram (full symmetric square, fp32) goes to Gout[j] and the rhs to F[j]; rows without entries
//                get zeros.
template <int K, int NW, int MINB, bool SOLVE>
__global__ void __launch_bounds__(NW * 32, MINB)
f_update_mma_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                    const float *__restrict__ X, float *__restrict__ F, float *__restrict__ Gout, double lambda,
                    uint32_t nseries, unsigned *__restrict__ queue) {
    typedef Cfg<K> C;
    constexpr int NC = C::NC, MT = C::MT, NT = C::NT, CH = C::CH, RS = C::RS, STAGES = C::STAGES, NQ = C::NQ;
    constexpr int SF = C::STAGE_FLOATS, ld = C::ld, NTH = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *P = reinterpret_cast<double *>(smem_raw);                 // [NW][NT*128]  per-warp Gram partials (fragment layout)
    double *R = P + (size_t)NW * NT * 128;                            // [NW][8*NC]    per-warp rhs partials
    constexpr size_t PART_BYTES = sizeof(double) * (size_t)NW * (NT * 128 + 8 * NC);
    float *stage = reinterpret_cast<float *>(smem_raw + PART_BYTES);   // [NW][STAGES][ET*RS]
    float *ystage = stage + (size_t)NW * STAGES * SF;                 // [NW][STAGES][ET]
    double *A = reinterpret_cast<double *>(stage);                    // epilogue only: (K+1) x ld lower triangle + rhs row
    double *dinv = A + (size_t)(K + 1) * ld;
    __shared__ unsigned next_series;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;

    for (int p = tid; p < NW * STAGES * SF; p += NTH) stage[p] = 0.f;   // padding columns stay zero
    if (tid == 0) next_series = atomicAdd(queue, 1u);
    __syncthreads();
    uint32_t j = next_series;

    float *st = stage + (size_t)warp * STAGES * SF;
    float *ys = ystage + (size_t)warp * STAGES * ET;
    double *Pw = P + (size_t)warp * NT * 128;
    double *Rw = R + (size_t)warp * 8 * NC;

    while (j < nseries) {
        const uint64_t lo_ = ptr[j];
        const uint32_t nnz = (uint32_t)(ptr[j + 1] - lo_);      // one series never holds 2^32 entries (T < 2^32)
        if (nnz != 0) {
            const uint32_t *sidx = idx + lo_;
            const float *sval = val + lo_;
            const int ntiles = (int)((nnz + ET - 1) / ET);
            const int nwa = ntiles < NW ? ntiles : NW;          // warps that own at least one tile
            if (warp < nwa) {
                const int my_tiles = (ntiles - warp + NW - 1) / NW;
                float acc[NT][4];
                float racc[NC];
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[t][q] = 0.f;
#pragma unroll
                for (int c = 0; c < NC; ++c) racc[c] = 0.f;
                bool first = true;

                uint32_t nidx;   // lane e < ET: row index of entry e of the next tile to issue
                auto load_idx = [&](int i) {
                    const uint32_t e = (uint32_t)(warp + i * NW) * ET + lane;
                    nidx = (lane < ET && i < my_tiles && e < nnz) ? __ldg(sidx + e) : 0u;
                };
                auto issue = [&](int i, int s) {   // gathers local tile i (indices in nidx) into stage s
                    const uint32_t base = (uint32_t)(warp + i * NW) * ET;
                    const uint32_t rem = nnz - base;
                    const int cnt = rem < (uint32_t)ET ? (int)rem : ET;
                    float *dst = st + s * SF;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int id = q * 32 + lane;
                        const int e = id / CH, cc = id - e * CH;
                        const uint32_t row = __shfl_sync(FULL_MASK, nidx, e & 15);
                        if (id < ET * CH && e < cnt) cp_async16(dst + e * RS + 4 * cc, X + (size_t)row * K + 4 * cc);
                    }
                    if (lane < ET) {
                        if (lane < cnt) cp_async4(ys + s * ET + lane, sval + base + lane);
                        else ys[s * ET + lane] = 0.f;
                    }
                    if (cnt < ET)   // tail tile: rows past the end must read as zero
                        for (int p = cnt * RS + lane; p < SF; p += 32) dst[p] = 0.f;
                };
                auto flush = [&]() {
#pragma unroll
                    for (int t = 0; t < NT; ++t)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            double *d = Pw + (t * 4 + q) * 32 + lane;
                            *d = first ? (double)acc[t][q] : *d + (double)acc[t][q];
                            acc[t][q] = 0.f;
                        }
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float r = racc[c];
                        r += __shfl_xor_sync(FULL_MASK, r, 1);
                        r += __shfl_xor_sync(FULL_MASK, r, 2);
                        if (tig == 0) Rw[8 * c + g] = first ? (double)r : Rw[8 * c + g] + (double)r;
                        racc[c] = 0.f;
                    }
                    first = false;
                };

                // ---- prologue: local tiles 0 .. STAGES-2 in flight ----
                load_idx(0);
#pragma unroll
                for (int s = 0; s < STAGES - 1; ++s) {
                    if (s < my_tiles) { issue(s, s); load_idx(s + 1); }
                    cp_async_commit();
                }
                int s_cur = 0;
                for (int i = 0; i < my_tiles; ++i) {
                    cp_async_wait<STAGES - 2>();
                    __syncwarp();                  // tile i visible to the whole warp; tile i-1 fully consumed
                    {
                        const int ni = i + STAGES - 1;
                        int ns = s_cur + STAGES - 1;
                        if (ns >= STAGES) ns -= STAGES;
                        if (ni < my_tiles) { issue(ni, ns); load_idx(ni + 1); }
                        cp_async_commit();
                    }
                    const float *tb = st + s_cur * SF;
                    const float *yb = ys + s_cur * ET;
#pragma unroll
                    for (int c8 = 0; c8 < ET / 8; ++c8) {
                        const float *p = tb + (c8 * 8 + tig) * RS + g;
                        const float y0 = yb[c8 * 8 + tig], y1 = yb[c8 * 8 + tig + 4];
                        uint32_t ah[MT][4], al[MT][4], bh[NC][2], bl[NC][2];
#pragma unroll
                        for (int c = 0; c < 2 * MT; ++c) {
                            if (c < NC) {
                                const float v0 = p[8 * c], v1 = p[4 * RS + 8 * c];
                                racc[c] = fmaf(y0, v0, racc[c]);
                                racc[c] = fmaf(y1, v1, racc[c]);
                                uint32_t h0, l0, h1, l1;
                                split_tf32(v0, h0, l0);
                                split_tf32(v1, h1, l1);
                                ah[c >> 1][(c & 1)] = h0; ah[c >> 1][(c & 1) + 2] = h1;
                                al[c >> 1][(c & 1)] = l0; al[c >> 1][(c & 1) + 2] = l1;
                                bh[c][0] = h0; bh[c][1] = h1;
                                bl[c][0] = l0; bl[c][1] = l1;
                            } else {   // virtual chunk past k (NC odd): zero rows of the last 16-row tile
                                ah[c >> 1][(c & 1)] = 0u; ah[c >> 1][(c & 1) + 2] = 0u;
                                al[c >> 1][(c & 1)] = 0u; al[c >> 1][(c & 1) + 2] = 0u;
                            }
                        }
                        int t = 0;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                            for (int nt = 2 * mt; nt < NC; ++nt) {
                                float d[4];
                                mma_zero(d, al[mt], bh[nt]);     // small terms first
                                mma_acc(d, ah[mt], bl[nt]);
                                mma_acc(d, ah[mt], bh[nt]);
#pragma unroll
                                for (int q = 0; q < 4; ++q) acc[t][q] += d[q];
                                ++t;
                            }
                    }
                    if (++s_cur == STAGES) s_cur = 0;
                    if ((i + 1) % FLUSH == 0) flush();
                }
                cp_async_wait<0>();
                if (first || my_tiles % FLUSH != 0) flush();
            }
            __syncthreads();   // all partials written, all staging reads done: the staging area becomes A
            // ---- reduce the per-warp partials (warp order) into the lower triangle of A + rhs row ----
            for (int u = tid; u < NT * 128; u += NTH) {
                int t = u >> 7, mt = 0;
                while (t >= NC - 2 * mt) { t -= NC - 2 * mt; ++mt; }
                const int nt = 2 * mt + t;
                const int q = (u >> 5) & 3, l = u & 31;
                const int r = 16 * mt + (l >> 2) + 8 * (q >> 1), c = 8 * nt + 2 * (l & 3) + (q & 1);
                if (r <= c && c < K) {
                    double s = P[u];
                    for (int w = 1; w < nwa; ++w) s += P[(size_t)w * NT * 128 + u];
                    A[c * ld + r] = s;
                }
            }
            for (int c = tid; c < K; c += NTH) {
                double s = R[c];
                for (int w = 1; w < nwa; ++w) s += R[w * 8 * NC + c];
                A[K * ld + c] = s;
            }
            __syncthreads();
            if (SOLVE) {
                if (tid < K) A[tid * ld + tid] += lambda;      // trmf.cpp:393
                block_chol_solve_blocked<(K + 32) / 32>(A, ld, dinv, K);   // starts and ends with __syncthreads
                if (tid < K) F[(size_t)j * K + tid] = (float)A[K * ld + tid];
            } else {
                float *Gj = Gout + (size_t)j * K * K;
                for (int p = tid; p < K * K; p += NTH) {
                    const int r = p / K, c = p - r * K;
                    Gj[p] = (float)(r >= c ? A[r * ld + c] : A[c * ld + r]);
                }
                if (tid < K) F[(size_t)j * K + tid] = (float)A[K * ld + tid];
            }
            if (RS > K) {   // A overwrote the staging area: padding columns must read as zero again
                __syncthreads();
                for (int p = tid; p < NW * STAGES * ET; p += NTH)
                    for (int c = K; c < RS; ++c) stage[(size_t)p * RS + c] = 0.f;
            }
        } else if (!SOLVE) {
            float *Gj = Gout + (size_t)j * K * K;
            for (int p = tid; p < K * K; p += NTH) Gj[p] = 0.f;
            if (tid < K) F[(size_t)j * K + tid] = 0.f;
        }
        __syncthreads();
        if (tid == 0) next_series = atomicAdd(queue, 1u);
        __syncthreads();
        j = next_series;
    }
}

}   // namespace fm

static inline bool f_update_mma_supported(int k) {
    switch (k) { case 8: case 16: case 20: case 24: case 32: case 40: case 48: return true; }
    return false;
}

// returns 0 on success.  `wide` selects one 12-warp CTA per SM (few series: finer load balance) instead of
// three 4-warp CTAs per SM.
template <bool SOLVE>
static inline int f_update_mma_launch(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *idx, const V *val,
                                      const V *X, V *F, V *Gout, int k, double lambda, uint32_t nseries, unsigned *queue,
                                      unsigned long long *launches) {
    if (cudaMemsetAsync(queue, 0, sizeof(unsigned), st) != cudaSuccess) return 1;
    const bool wide = nseries < (uint32_t)(24 * num_sms);
#define FM_LAUNCH(KK, NWW, MINBB)                                                                               \
    do {                                                                                                        \
        const size_t smem = fm::Cfg<KK>::smem(NWW);                                                             \
        auto kfn = fm::f_update_mma_kernel<KK, NWW, MINBB, SOLVE>;                                              \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        unsigned grid = (unsigned)(MINBB * num_sms);                                                            \
        if (grid > nseries) grid = nseries;                                                                     \
        kfn<<<grid ? grid : 1, NWW * 32, smem, st>>>(ptr, idx, val, X, F, Gout, lambda, nseries, queue);        \
    } while (0)
#define FM_CASE(KK)                                                                                             \
    case KK:                                                                                                    \
        if (wide) FM_LAUNCH(KK, 12, 1); else FM_LAUNCH(KK, 4, 3);                                               \
        break;
    switch (k) {
        FM_CASE(8) FM_CASE(16) FM_CASE(20) FM_CASE(24) FM_CASE(32) FM_CASE(40)
        case 48:   // 12 accumulator tiles: 23.5 KB of shared memory per warp -> 8 warps per SM
            if (wide) FM_LAUNCH(48, 8, 1); else FM_LAUNCH(48, 4, 2);
            break;
        default: return 1;
    }
#undef FM_CASE
#undef FM_LAUNCH
    ++*launches;
    return cudaGetLastError() != cudaSuccess;
}

#else   // float64 build: the generic kernel (fp64 FMAs) is the parity path

static inline bool f_update_mma_supported(int) { return false; }
template <bool SOLVE>
static inline int f_update_mma_launch(cudaStream_t, int, const uint64_t *, const uint32_t *, const V *, const V *, V *, V *, int,
                                      double, uint32_t, unsigned *, unsigned long long *) { return 1; }
#endif
