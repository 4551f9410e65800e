// f_update_mma2.cuh -- K1 (and the X-update's Gram build) on the warp-level tensor path, second generation: the factor is
// split into fp16 pairs ONCE per launch (presplit_kernel), gathered as fp16 rows and fed to mma.sync through ldmatrix.
//
// Replaces the hot loop of l2r_ls_pY_IX_chol::solve (reference trmf.cpp:382-395) and, in MODE_GRAD, arr_ls_pY_IX::fun / ::grad
// (trmf.cpp:231-267), like f_update_mma.cuh -- same decomposition (CTA = NW warps on one series, tiles of 16 entries dealt
// round-robin to the warps, per-warp cp.async double buffer, fp32 partial sums flushed to per-warp fp64 partials every 256
// entries, deterministic warp-order reduction, deferred fp64 Cholesky), same accuracy construction (exact per-column
// power-of-two scaling, x = h1 + h2 in fp16, products h2 h1 + h1 h2 + h1 h1 with the small terms first, the truncating
// tensor-core adder trusted with one tile only).  What changed is everything that was NOT an HMMA in the first kernel
// (profiles/r01_f_update_mma_k40_sass_mix.txt: 27 HMMA = 216 issue clocks of 620 per tile, ~320 other instructions):
//
//  * the fp32 -> (h1, h2) split leaves the hot loop: rows are gathered as [h1[0..8NC) | h2[0..8NC)] fp16 (same bytes per row as
//    fp32 when 8 | k), written by presplit_kernel together with the scaling -- no F2FP / HADD2.F32 / FADD per tile;
//  * fragments come from ldmatrix(.trans): 4 registers per instruction, already in the A-quad resp. B-pair register homes the
//    HMMA wants -- no LDS per value, no register moves between the two homes (both arrangements are loaded; shared memory
//    has the bandwidth, the issue port does not);
//  * the right-hand side sum_e y_e x_e rides on the tensor core too: one extra 8-wide B tile per 16-row group whose columns
//    0 / 1 are the fp16 pair (y1, y2) of the tile's weights.  F-update: pre-split once per Y with one exact global power-of-two
//    scale (ysplit_kernel).  MODE_GRAD / MODE_STORE: split per tile with an exact power of two taken from the tile's largest
//    magnitude and undone in the fp32 accumulate (an FFMA in place of an FADD);
//  * MODE_GONLY: no weights at all -- the Gram over the MISSING cells of a mostly observed Y (complement.cuh);
//  * MODE_GRAD: z_e = <w_j, x_e> is an HMMA chain over the same staged rows read un-transposed (entries as M, latent index as
//    the contraction), against a per-series B fragment of the point (w1 | w2), instead of 20 FFMA + 12 SHFL + 12 FADD.
//
// Per tile at k = 40: 33 HMMA + ~125 other instructions (MODE_DEFER) against 27 + ~320; 504 against 620 clk (profiles/r02_mma2_history.md).
#pragma once
#include <cuda_fp16.h>
#include <type_traits>
#include "f_update_mma.cuh"

#ifdef TRMF_F32

namespace fm {

template <int K, int ST = 2> struct Cfg2 {
    static constexpr int NC = Cfg<K>::NC, MT = Cfg<K>::MT, NT = Cfg<K>::NT, PW = Cfg<K>::PW, ld = K + 1;
    static constexpr bool ODD = (NC & 1) != 0;
    static constexpr int UR = 2 * NC;               // 16-byte units of a gathered row: h1 (NC units) | h2 (NC units)
    static constexpr int US = 2 * NC + 1;           // staging row stride in units: odd, so the 8 rows of an ldmatrix tile hit 8 different bank groups
    static constexpr int ROWB = 16 * UR;            // bytes of a pre-split factor row in global memory
    static constexpr int NQ = (ET * UR + 31) / 32;  // warp-wide LDGSTS per tile
    static constexpr int STAGES = ST;
    static constexpr int STAGE_B = ET * US * 16;    // bytes of one stage's rows
    static constexpr int FLUSH_TILES = 16;          // tiles between fp32 -> fp64 flushes (256 entries)
    static constexpr int RSTR = 16 * MT;            // per-warp rhs partial: one double per (padded) latent index
    static constexpr size_t a_bytes = sizeof(double) * ((size_t)(K + 1) * (K + 1) + K);
    static constexpr size_t part_bytes(int nw) { return sizeof(double) * (size_t)nw * (PW + RSTR); }
    static constexpr size_t stage_bytes(int nw) {
        const size_t s = (size_t)nw * STAGES * (STAGE_B + ET * sizeof(float));
        return ((s > a_bytes ? s : a_bytes) + 15) & ~(size_t)15;
    }
    static constexpr size_t smem(int nw) { return part_bytes(nw) + stage_bytes(nw); }
    static constexpr size_t xh_floats_per_row = 8 * NC;   // size of the pre-split copy in units of the fp32 factor's element
};

// Xh[i] = [h1 | h2] of X[i] * s (s = the exact per-column power-of-two scales of colscale_max_kernel: column maximum into
// [2^14, 2^15)), each half padded with zeros to kp = 8 * ceil(k / 8) columns.  invs[c] = 1 / s_c.
// Xr (optional, rows x k fp32) = the values the split actually carries, (h1 + h2) / s: exact in fp32, 22 of x's 24 significant bits.
// Whatever is combined with a Gram of the split factor (complement.cuh) reads THIS copy, so that Gram and right-hand side describe
// the same (slightly rounded) least-squares problem -- the normal equations forgive a consistent perturbation of the data, not an
// independent one of the Gram or the right-hand side.
__global__ void presplit_kernel(const float *__restrict__ X, size_t rows, int k, const unsigned *__restrict__ colmax,
                                __half *__restrict__ Xh, float *__restrict__ invs, float *__restrict__ Xr) {
    extern __shared__ float sc[];
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        const float m = __uint_as_float(colmax[c]);
        float s = 1.f;
        if (m > 0.f && m < 3.0e38f) {
            int e;
            frexpf(m, &e);
            e = 15 - e;
            e = e > 100 ? 100 : (e < -100 ? -100 : e);
            s = ldexpf(1.f, e);
        }
        sc[c] = s;
        if (blockIdx.x == 0) invs[c] = 1.f / s;
    }
    __syncthreads();
    const int kp = 8 * ((k + 7) / 8), kq = kp / 2;       // one thread: two adjacent columns -> one half2 store per split part
    const size_t total = rows * (size_t)kq;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = p / kq;
        const int c = 2 * (int)(p - i * kq);
        const float x0 = c < k ? X[i * k + c] * sc[c] : 0.f, x1 = c + 1 < k ? X[i * k + c + 1] * sc[c + 1] : 0.f;
        const __half2 h1 = __floats2half2_rn(x0, x1);
        const float2 f1 = __half22float2(h1);
        const __half2 h2 = __floats2half2_rn(x0 - f1.x, x1 - f1.y);
        __half2 *row = reinterpret_cast<__half2 *>(Xh + i * (size_t)(2 * kp));
        row[c / 2] = h1;
        row[(kp + c) / 2] = h2;
        if (Xr != nullptr) {
            const float2 f2 = __half22float2(h2);
            if (c < k) Xr[i * k + c] = (f1.x + f2.x) / sc[c];
            if (c + 1 < k) Xr[i * k + c + 1] = (f1.y + f2.y) / sc[c + 1];
        }
    }
}

__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");     // (.ca measured slower: profiles/r02_mma2_history.md)
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, unsigned a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t &r0, uint32_t &r1, unsigned a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, unsigned a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// exact power-of-two scale that brings a non-negative magnitude m into [2^14, 2^15), and its inverse.  m below 2^-113 (incl.
// 0) gets the largest scale and an inverse of 0: its contribution (below 2^-113 |x|) is dropped.
__device__ __forceinline__ void pow2_scale(const float m, float &s, float &inv) {
    int eb = (int)(__float_as_uint(m) >> 23);
    eb = eb < 14 ? 14 : eb;
    s = __uint_as_float((uint32_t)(268 - eb) << 23);
    inv = __uint_as_float((uint32_t)(eb - 14) << 23);
}
// (v0, v1) -> fp16 pairs: hi = (fp16(v0), fp16(v1)), lo = the fp16-rounded remainders
__device__ __forceinline__ void split_h2(const float v0, const float v1, uint32_t &hi, uint32_t &lo) {
    hi = pack_h2(v0, v1);
    const float2 f = unpack_h2(hi);
    lo = pack_h2(v0 - f.x, v1 - f.y);
}

// the last, partly filled tile of a series (at most one per warp and series): cnt < 16 rows are gathered, the rest of the stage
// is zeroed.  Out of line on purpose -- inlined, its address arithmetic is hoisted into every tile's path.
__device__ __noinline__ void issue_ragged_tile(const unsigned gdst_s, const unsigned char *gsrc, const uint32_t nidx, const int cnt,
                                               const bool g_on, const int ge, float *ysm, const float *yg, uint32_t *rows32,
                                               const int rpr, const int nq, const int us, const int rowb) {
    const int lane = threadIdx.x & 31;
    for (int q = 0; q < nq; ++q) {
        const uint32_t row = __shfl_sync(FULL_MASK, nidx, (ge + q * rpr) & 15);
        if (g_on && ge + q * rpr < cnt) cp_async16_s(gdst_s + (unsigned)(q * rpr * us * 16), gsrc + (size_t)row * rowb);
    }
    if (yg != nullptr) {
        if (lane < cnt) cp_async4(ysm + lane, yg + lane);
        else if (lane < ET) ysm[lane] = 0.f;
    }
    for (int p = cnt * us * 4 + lane; p < ET * us * 4; p += 32) rows32[p] = 0u;   // rows past the end read as zero
}

// weights of the right-hand side, pre-split: yh[e] = (fp16(y s) | fp16(y s - fp16(y s)) << 16) with s = ysc[0], the exact power of
// two that brings max |y| of the range into [2^14, 2^15); ysc[1] = 1 / s.  ymax = bit pattern of that maximum (atomicMax).
__global__ void yabsmax_kernel(const float *__restrict__ val, const uint64_t *__restrict__ ptr, uint32_t nseries, unsigned *__restrict__ ymax) {
    const uint64_t e0 = ptr[0], e1 = ptr[nseries];
    float m = 0.f;
    for (uint64_t e = e0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < e1; e += (uint64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(val[e]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(ymax, __float_as_uint(m));
}
__global__ void ysplit_kernel(const float *__restrict__ val, const uint64_t *__restrict__ ptr, uint32_t nseries, const unsigned *__restrict__ ymax,
                              uint32_t *__restrict__ yh, float *__restrict__ ysc) {
    const uint64_t e0 = ptr[0], e1 = ptr[nseries];
    float sy, inv;
    pow2_scale(__uint_as_float(*ymax), sy, inv);
    if (blockIdx.x == 0 && threadIdx.x == 0) { ysc[0] = sy; ysc[1] = inv; }
    for (uint64_t e = e0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < e1; e += (uint64_t)gridDim.x * blockDim.x) {
        const float y = val[e] * sy;
        const __half h1 = __float2half_rn(y);
        const __half h2 = __float2half_rn(y - __half2float(h1));
        yh[e] = (uint32_t)__half_as_ushort(h1) | ((uint32_t)__half_as_ushort(h2) << 16);
    }
}

// Modes as in f_update_mma.cuh.  Xh is the pre-split, column-scaled factor (presplit_kernel), invs its inverse scales.
// MODE_SOLVE / MODE_DEFER (the F-update): `val` is the PRE-SPLIT weight array of ysplit_kernel (uint32 per entry), ysc its scales.
// MODE_STORE / MODE_GRAD: `val` holds plain fp32 values (the residuals of MODE_GRAD are split tile by tile on the fly).
template <int K, int NW, int MINB, int MODE, int ST = 2>
__global__ void __launch_bounds__(NW * 32, MINB)
f_update_mma2_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                     const unsigned char *__restrict__ Xh, const float *__restrict__ invs, float *__restrict__ F,
                     float *__restrict__ Gout, double lambda, uint32_t nseries, unsigned *__restrict__ queue,
                     const float *__restrict__ Wv, int gaccum, double *__restrict__ frow, const float *__restrict__ ysc) {
    typedef Cfg2<K, ST> C;
    constexpr bool GONLY = MODE == MODE_GONLY;      // Gram only: no weights, no right-hand side
    constexpr bool SOLVE = MODE == MODE_SOLVE || MODE == MODE_DEFER || GONLY, GRAD = MODE == MODE_GRAD, DEFER = MODE == MODE_DEFER || GONLY;
    constexpr bool YPRE = SOLVE && !GONLY;          // weights arrive pre-split, one global scale
    constexpr int NC = C::NC, MT = C::MT, NT = C::NT, UR = C::UR, US = C::US, STAGES = C::STAGES;
    constexpr int SB = C::STAGE_B, ld = C::ld, NTH = NW * 32, PW = C::PW, RSTR = C::RSTR, ROWB = C::ROWB;
    constexpr bool ODD = C::ODD;
    constexpr int RPR = 32 / UR < ET ? 32 / UR : ET;      // rows per warp-wide gather request: lane = (row within the request, unit)
    constexpr int NQ = (ET + RPR - 1) / RPR;              // requests per tile
    constexpr int FL = C::FLUSH_TILES;
    static_assert(STAGES == 2, "the tile loop is unrolled over two stages");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *P = reinterpret_cast<double *>(smem_raw);                 // [NW][PW]      per-warp Gram partials (fragment layout)
    double *R = P + (size_t)NW * PW;                                  // [NW][RSTR]    per-warp rhs partials
    constexpr size_t PART_BYTES = sizeof(double) * (size_t)NW * (PW + RSTR);
    unsigned char *stage = smem_raw + PART_BYTES;                     // [NW][STAGES][ET rows of US units]
    float *ystage = reinterpret_cast<float *>(stage + (size_t)NW * STAGES * SB);   // [NW][STAGES][ET]
    double *A = reinterpret_cast<double *>(stage);                    // epilogue only: (K+1) x ld lower triangle + rhs row
    double *dinv = A + (size_t)(K + 1) * ld;
    __shared__ unsigned next_series;
    __shared__ double fwarp[NW];                                      // MODE_GRAD: per-warp sum of squared residuals
    __shared__ __align__(16) uint32_t zeros[STAGES * ET + ET];        // what the lanes outside the two weight columns read

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;

    if (tid < STAGES * ET + ET) zeros[tid] = 0u;
    if (tid == 0) next_series = atomicAdd(queue, 1u);
    __syncthreads();
    uint32_t j = next_series;

    unsigned char *st = stage + (size_t)warp * STAGES * SB;
    float *ys = ystage + (size_t)warp * STAGES * ET;
    double *Pw = P + (size_t)warp * PW;
    double *Rw = R + (size_t)warp * RSTR;
    const unsigned st_s = (unsigned)__cvta_generic_to_shared(st);
    const unsigned ys_s = (unsigned)__cvta_generic_to_shared(ys);
    // ldmatrix lane addressing (lane = 8 m + r supplies row r of 8x8 matrix m; rows = entries, 16 bytes = 8 latent columns):
    //  "pair" order : m = 0 (entries 0-7, unit u) 1 (entries 8-15, u) 2 (entries 0-7, u+1) 3 (entries 8-15, u+1)
    //                 -> .trans: B pairs of n-tiles u, u+1;  plain: the A quad of the z pass (16 entries x 16 latent columns)
    //  "quad" order : m = 0 (entries 0-7, u) 1 (entries 0-7, u+1) 2 (entries 8-15, u) 3 (entries 8-15, u+1)
    //                 -> .trans: the A quad of the Gram (16 latent rows x 16 entries)
    // A unit past a half (odd NC, last 16-row group) only feeds Gram rows >= 8 NC >= K, which the epilogue never reads.
    const int lm = lane >> 3, lr = lane & 7;
    const unsigned pair_s = st_s + (unsigned)(((lr + 8 * (lm & 1)) * US + (lm >> 1)) * 16);
    const unsigned quad_s = st_s + (unsigned)(((lr + 8 * (lm >> 1)) * US + (lm & 1)) * 16);
    // gather: lane = (row ge of the request, 16-byte unit gc); request q brings rows q RPR + ge
    const int ge = lane / UR, gc = lane - ge * UR;
    const bool g_on = lane < RPR * UR;
    const unsigned char *gsrc = Xh + 16 * gc;
    asm volatile("" : "+l"(gsrc));     // keep the lane's source pointer in registers: one IMAD.WIDE per request instead of base + offset + row
    const unsigned gdst_s = st_s + (unsigned)((ge * US + gc) * 16);
    // pre-split weights: lanes g < 2 read entries (2 tig, 2 tig + 1) and (+8) and keep half g of each; the others read zeros
    const unsigned yrd_s = g < 2 ? ys_s + 8u * tig : (unsigned)__cvta_generic_to_shared(zeros);
    const uint32_t ysel_lo = g == 0 ? 0x5410u : 0x7632u;
    float inv_sy_pre = 0.f;
    if (YPRE) inv_sy_pre = __ldg(ysc + 1);

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
                float racc[MT][2];     // rhs of 16-row group mt, both weight columns added: lanes tig == 0 hold rows g and g + 8
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[t][q] = 0.f;
#pragma unroll
                for (int m = 0; m < MT; ++m) racc[m][0] = racc[m][1] = 0.f;
                bool first = true;
                uint32_t bw[MT][2];    // MODE_GRAD: B fragments of the point, columns 0 / 1 = (w1 | w2), 16 latent rows per step
                float inv_sw = 0.f;
                double fsum = 0.0;
                if (GRAD) {
                    // w^ = w / (column scales), scaled by a power of two so that its largest magnitude is in [2^14, 2^15)
                    float wv[MT][4], m = 0.f;
#pragma unroll
                    for (int kk = 0; kk < MT; ++kk)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c = 16 * kk + 8 * (q >> 1) + 2 * tig + (q & 1);
                            wv[kk][q] = c < K ? __ldg(Wv + (size_t)j * K + c) * __ldg(invs + c) : 0.f;
                            m = fmaxf(m, fabsf(wv[kk][q]));
                        }
                    m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 2));
                    float sw;
                    pow2_scale(m, sw, inv_sw);
#pragma unroll
                    for (int kk = 0; kk < MT; ++kk) {
                        uint32_t h0, l0, h1, l1;
                        split_h2(wv[kk][0] * sw, wv[kk][1] * sw, h0, l0);
                        split_h2(wv[kk][2] * sw, wv[kk][3] * sw, h1, l1);
                        bw[kk][0] = g == 0 ? h0 : (g == 1 ? l0 : 0u);
                        bw[kk][1] = g == 0 ? h1 : (g == 1 ? l1 : 0u);
                    }
                }

                uint32_t nidx = 0;                                  // lane e (and e + 16): row index of entry e of the next tile to issue
                uint32_t ib = (uint32_t)warp * ET;                  // first entry of the next tile to issue (a tile exists iff ib < nnz)
                uint32_t lb = (uint32_t)warp * ET + (lane & 15);    // this lane's entry of the next tile whose indices get loaded
                auto load_idx = [&]() {
                    if (lb < nnz) nidx = __ldg(sidx + lb);
                    lb += NW * ET;
                };
                auto issue = [&](auto sc) {   // gathers the tile at ib (indices in nidx) into stage s
                    constexpr int s = decltype(sc)::value;
                    if (ib + ET <= nnz) {          // full tile (warp-uniform)
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const uint32_t row = __shfl_sync(FULL_MASK, nidx, (ge + q * RPR) & 15);
                            if (g_on && ((q + 1) * RPR <= ET || ge + q * RPR < ET))
                                cp_async16_s(gdst_s + (unsigned)(s * SB + q * RPR * US * 16), gsrc + (size_t)row * ROWB);
                        }
                        if (!GONLY && lane < ET) cp_async4(ys + s * ET + lane, sval + ib + lane);
                    } else {
                        issue_ragged_tile(gdst_s + (unsigned)(s * SB), gsrc, nidx, (int)(nnz - ib), g_on, ge, ys + s * ET, GONLY ? nullptr : sval + ib,
                                          reinterpret_cast<uint32_t *>(st + s * SB), RPR, NQ, US, ROWB);
                    }
                    ib += NW * ET;
                };
                auto flush = [&]() {
                    {
                        int t = 0;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                            for (int nt = 2 * mt; nt < NC; ++nt) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (nt > 2 * mt || q < 2) {
                                        double *d = Pw + Cfg<K>::poff(mt, nt - 2 * mt) + q * 32 + lane;
                                        *d = first ? (double)acc[t][q] : *d + (double)acc[t][q];
                                    }
                                    acc[t][q] = 0.f;
                                }
                                ++t;
                            }
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        if (tig == 0) {
                            const double r0 = (double)racc[mt][0], r1 = (double)racc[mt][1];
                            Rw[16 * mt + g] = first ? r0 : Rw[16 * mt + g] + r0;
                            Rw[16 * mt + g + 8] = first ? r1 : Rw[16 * mt + g + 8] + r1;
                        }
                        racc[mt][0] = racc[mt][1] = 0.f;
                    }
                    first = false;
                };
                // one tile out of stage s (compile-time): every shared-memory address below is a lane constant + an immediate
                auto tile = [&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    uint32_t by[2];      // B fragment of the weights: column 0 = first fp16 part, column 1 = second, rows = the tile's entries
                    float inv_sy = 0.f;
                    if (GONLY) {
                        by[0] = by[1] = 0u;
                    } else if (YPRE) {
                        uint32_t w0, w1, w2, w3;
                        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(yrd_s + (unsigned)(s * ET * 4)) : "memory");
                        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w2), "=r"(w3) : "r"(yrd_s + (unsigned)(s * ET * 4 + 32)) : "memory");
                        by[0] = prmt(w0, w1, ysel_lo);
                        by[1] = prmt(w2, w3, ysel_lo);
                    } else if (!GRAD) {
                        const float *yb = ys + s * ET;
                        const float2 ya = *reinterpret_cast<const float2 *>(yb + 2 * tig), yc = *reinterpret_cast<const float2 *>(yb + 8 + 2 * tig);
                        float m = fmaxf(fmaxf(fabsf(ya.x), fabsf(ya.y)), fmaxf(fabsf(yc.x), fabsf(yc.y)));
                        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 1));
                        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 2));
                        float sy;
                        pow2_scale(m, sy, inv_sy);
                        uint32_t h0, l0, h1, l1;
                        split_h2(ya.x * sy, ya.y * sy, h0, l0);
                        split_h2(yc.x * sy, yc.y * sy, h1, l1);
                        by[0] = g == 0 ? h0 : (g == 1 ? l0 : 0u);
                        by[1] = g == 0 ? h1 : (g == 1 ? l1 : 0u);
                    } else {
                        // z = X w^ on the tensor core: rows = the tile's 16 entries, contraction over the latent index, h2 terms first
                        const float *yb = ys + s * ET;
                        float zl[4], zh[4];     // two independent HMMA chains: the h2 products and the h1 products
#pragma unroll
                        for (int kk = 0; kk < MT; ++kk)
#pragma unroll
                            for (int part = 1; part >= 0; --part) {
                                uint32_t a[4];
                                const unsigned ad = pair_s + (unsigned)(s * SB + (part * NC + 2 * kk) * 16);
                                if (ODD && kk == MT - 1) {      // latent columns past 8 NC: the point is zero there, so is this half of A
                                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(a[0]), "=r"(a[1]) : "r"(ad) : "memory");
                                    a[2] = 0u; a[3] = 0u;
                                } else ldsm_x4(a[0], a[1], a[2], a[3], ad);
                                if (part == 1) { if (kk == 0) mma_zero(zl, a, bw[kk]); else mma_acc(zl, a, bw[kk]); }
                                else           { if (kk == 0) mma_zero(zh, a, bw[kk]); else mma_acc(zh, a, bw[kk]); }
                            }
                        // lanes tig == 0: entry g in columns (0 | 1) = registers (0 | 1), entry g + 8 in registers (2 | 3); small terms first;
                        // residual r = z - y
                        const float r0 = fmaf((zl[0] + zl[1]) + (zh[1] + zh[0]), inv_sw, -yb[g]);
                        const float r1 = fmaf((zl[2] + zl[3]) + (zh[3] + zh[2]), inv_sw, -yb[g + 8]);
                        if (tig == 0) fsum += (double)r0 * (double)r0 + (double)r1 * (double)r1;   // rows past the end: x = 0, y = 0
                        float m = fmaxf(fabsf(r0), fabsf(r1));
                        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 4));
                        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 8));
                        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 16));
                        m = __shfl_sync(FULL_MASK, m, 0);
                        float sy;
                        pow2_scale(m, sy, inv_sy);
                        uint32_t ph, pl;           // (entry g, entry g + 8) on the lanes tig == 0
                        split_h2(r0 * sy, r1 * sy, ph, pl);
                        // lane (n = g, t = tig) needs entries 2t, 2t+1 (b0) and 2t+8, 2t+9 (b1) of part n: lanes 8t and 8t+4 hold them
                        const uint32_t h_a = __shfl_sync(FULL_MASK, ph, 8 * tig), h_b = __shfl_sync(FULL_MASK, ph, 8 * tig + 4);
                        const uint32_t l_a = __shfl_sync(FULL_MASK, pl, 8 * tig), l_b = __shfl_sync(FULL_MASK, pl, 8 * tig + 4);
                        const uint32_t va = g == 0 ? h_a : l_a, vb = g == 0 ? h_b : l_b;
                        by[0] = g < 2 ? prmt(va, vb, 0x5410u) : 0u;
                        by[1] = g < 2 ? prmt(va, vb, 0x7632u) : 0u;
                    }
                    uint32_t b1[NC][2], b2[NC][2];     // B pairs of every 8-wide column tile, h1 and h2 parts
#pragma unroll
                    for (int nt = 0; nt + 1 < NC; nt += 2) {
                        ldsm_x4_t(b1[nt][0], b1[nt][1], b1[nt + 1][0], b1[nt + 1][1], pair_s + (unsigned)(s * SB + nt * 16));
                        ldsm_x4_t(b2[nt][0], b2[nt][1], b2[nt + 1][0], b2[nt + 1][1], pair_s + (unsigned)(s * SB + (NC + nt) * 16));
                    }
                    if (ODD) {   // lanes 16-31 address unit NC resp. 2 NC of their row (in bounds; .x2 ignores them)
                        ldsm_x2_t(b1[NC - 1][0], b1[NC - 1][1], pair_s + (unsigned)(s * SB + (NC - 1) * 16));
                        ldsm_x2_t(b2[NC - 1][0], b2[NC - 1][1], pair_s + (unsigned)(s * SB + (2 * NC - 1) * 16));
                    }
                    int t = 0;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        uint32_t a1[4], a2[4];         // A quads of the 16-row group, h1 and h2 parts
                        ldsm_x4_t(a1[0], a1[1], a1[2], a1[3], quad_s + (unsigned)(s * SB + 2 * mt * 16));
                        ldsm_x4_t(a2[0], a2[1], a2[2], a2[3], quad_s + (unsigned)(s * SB + (NC + 2 * mt) * 16));
#pragma unroll
                        for (int nt = 2 * mt; nt < NC; ++nt) {
                            float d[4];
                            mma_zero(d, a2, b1[nt]);     // small terms first
                            mma_acc(d, a1, b2[nt]);
                            mma_acc(d, a1, b1[nt]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[t][q] += d[q];
                            ++t;
                        }
                        if (GONLY) continue;
                        float d[4];
                        mma_zero(d, a2, by);
                        mma_acc(d, a1, by);
                        if (YPRE) {
                            racc[mt][0] += d[0] + d[1];
                            racc[mt][1] += d[2] + d[3];
                        } else {
                            asm("fma.rn.f32 %0, %1, %2, %0;" : "+f"(racc[mt][0]) : "f"(d[0] + d[1]), "f"(inv_sy));
                            asm("fma.rn.f32 %0, %1, %2, %0;" : "+f"(racc[mt][1]) : "f"(d[2] + d[3]), "f"(inv_sy));
                        }
                    }
                };

                // ---- prologue: local tile 0 in flight ----
                load_idx();
                issue(std::integral_constant<int, 0>());
                load_idx();
                cp_async_commit();
                for (int i = 0; i < my_tiles; i += 2) {
                    cp_async_wait<0>();
                    __syncwarp();                  // tile i visible to the whole warp; tile i-1 fully consumed
                    if (ib < nnz) { issue(std::integral_constant<int, 1>()); load_idx(); }
                    cp_async_commit();
                    tile(std::integral_constant<int, 0>());
                    if (i + 1 < my_tiles) {
                        cp_async_wait<0>();
                        __syncwarp();
                        if (ib < nnz) { issue(std::integral_constant<int, 0>()); load_idx(); }
                        cp_async_commit();
                        tile(std::integral_constant<int, 1>());
                    }
                    if ((i + 2) % FL == 0) flush();       // FL is even
                }
                cp_async_wait<0>();
                if (first || ((my_tiles + 1) & ~1) % FL != 0) flush();
                if (GRAD) {
                    fsum += __shfl_xor_sync(FULL_MASK, fsum, 4);
                    fsum += __shfl_xor_sync(FULL_MASK, fsum, 8);
                    fsum += __shfl_xor_sync(FULL_MASK, fsum, 16);
                    if (lane == 0) fwarp[warp] = fsum;
                }
            }
            __syncthreads();   // all partials written, all staging reads done: the staging area becomes A
            // ---- reduce the per-warp partials (warp order) into the lower triangle of A + rhs row ----
            for (int u = tid; u < NT * 128; u += NTH) {
                int t = u >> 7, mt = 0;
                while (t >= NC - 2 * mt) { t -= NC - 2 * mt; ++mt; }
                const int nt = 2 * mt + t;
                const int q = (u >> 5) & 3, l = u & 31;
                const int r = 16 * mt + (l >> 2) + 8 * (q >> 1), c = 8 * nt + 2 * (l & 3) + (q & 1);
                if (r <= c && c < K) {     // (r <= c excludes the slots the partials do not store)
                    const int pu = 128 * (u >> 7) - 64 * (mt + (t > 0 ? 1 : 0)) + (q * 32 + l);
                    double s = P[pu];
                    for (int w = 1; w < nwa; ++w) s += P[(size_t)w * PW + pu];
                    A[c * ld + r] = s * ((double)invs[r] * (double)invs[c]);     // undo the column scaling (exact)
                }
            }
            for (int c = tid; c < K; c += NTH) {
                double s = R[c];
                for (int w = 1; w < nwa; ++w) s += R[w * RSTR + c];
                A[K * ld + c] = s * (YPRE ? (double)invs[c] * (double)inv_sy_pre : (double)invs[c]);
            }
            __syncthreads();
            if (DEFER) {   // frow carries the scratch: one (K+1) x ld fp64 system per series
                double *dst = frow + (size_t)j * ((K + 1) * ld);
                for (int p = tid; p < (K + 1) * ld; p += NTH) dst[p] = A[p];
            } else if (SOLVE) {
                if (tid < K) A[tid * ld + tid] += lambda;      // trmf.cpp:393
                block_chol_solve_blocked<(K + 32) / 32>(A, ld, dinv, K);   // starts and ends with __syncthreads
                if (tid < K) F[(size_t)j * K + tid] = (float)A[K * ld + tid];
            } else {
                float *Gj = Gout + (size_t)j * K * K;
                for (int p = tid; p < K * K; p += NTH) {
                    const int r = p / K, c = p - r * K;
                    Gj[p] = (float)(r >= c ? A[r * ld + c] : A[c * ld + r]);
                }
                if (GRAD) {
                    if (tid < K) {
                        float *o = F + (size_t)j * K + tid;
                        *o = gaccum ? (float)((double)*o + A[K * ld + tid]) : (float)A[K * ld + tid];
                    }
                    if (tid == 0) {
                        double fs = fwarp[0];
                        for (int w = 1; w < nwa; ++w) fs += fwarp[w];
                        frow[j] = fs;
                    }
                } else if (tid < K) F[(size_t)j * K + tid] = (float)A[K * ld + tid];
            }
            // (the staging area needs no clean-up after A: every unit a tile reads is rewritten by that tile's gather)
        } else if (!SOLVE) {
            float *Gj = Gout + (size_t)j * K * K;
            for (int p = tid; p < K * K; p += NTH) Gj[p] = 0.f;
            if (GRAD) { if (tid == 0) frow[j] = 0.0; if (!gaccum && tid < K) F[(size_t)j * K + tid] = 0.f; }
            else if (tid < K) F[(size_t)j * K + tid] = 0.f;
        }
        __syncthreads();
        if (tid == 0) next_series = atomicAdd(queue, 1u);
        __syncthreads();
        j = next_series;
    }
}

}   // namespace fm

// Pre-split weights of the series [0, nseries) behind `ptr` (entries ptr[0] .. ptr[nseries]) for the F-update modes: yh (one
// uint32 per entry, indexed like val) and ysc[0..1].  `ymax` = one scratch word.  Two streaming passes over val.
static inline int f_update_mma2_split_y(cudaStream_t st, int num_sms, const V *val, const uint64_t *ptr, uint32_t nseries, uint32_t *yh,
                                        float *ysc, unsigned *ymax, unsigned long long *launches) {
    if (cudaMemsetAsync(ymax, 0, sizeof(unsigned), st) != cudaSuccess) return 1;
    fm::yabsmax_kernel<<<(unsigned)(8 * num_sms), 256, 0, st>>>(val, ptr, nseries, ymax);
    fm::ysplit_kernel<<<(unsigned)(8 * num_sms), 256, 0, st>>>(val, ptr, nseries, ymax, yh, ysc);
    *launches += 2;
    return cudaGetLastError() != cudaSuccess;
}

// Same contract as f_update_mma_launch, except: Xs is reused as the pre-split copy of the factor (8 * ceil(k / 8) floats per
// row), and in MODE_SOLVE / MODE_DEFER `val` is the pre-split weight array of f_update_mma2_split_y with its scales in `ysc`.
template <int MODE>
static inline int f_update_mma2_launch(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *idx, const V *val,
                                       const V *X, size_t xrows, V *Xs, float *invs, V *F, V *Gout, int k, double lambda,
                                       uint32_t nseries, unsigned *queue, unsigned long long *launches,
                                       const V *Wv = nullptr, int gaccum = 0, double *frow = nullptr, bool rescale = true,
                                       const float *ysc = nullptr, V *Xr = nullptr, uint32_t rows_total = 0) {
    if (cudaMemsetAsync(queue, 0, sizeof(unsigned) * (rescale ? 8 + 128 : 1), st) != cudaSuccess) return 1;
    if (rescale) {
        const size_t total = xrows * (size_t)k;
        unsigned g1 = (unsigned)((total + 255) / 256);
        if (g1 > (unsigned)(4 * num_sms)) g1 = (unsigned)(4 * num_sms);
        if (g1 == 0) g1 = 1;
        if ((size_t)g1 * 256 < (size_t)k) g1 = (unsigned)((k + 255) / 256);
        fm::colscale_max_kernel<<<g1, 256, 0, st>>>(X, xrows, k, queue + 8);
        fm::presplit_kernel<<<g1, 256, sizeof(float) * k, st>>>(X, xrows, k, queue + 8, reinterpret_cast<__half *>(Xs), invs, Xr);
        *launches += 2;
    }
    // (a launch over one range of a longer row list -- rows_total rows in all -- takes the CTA shape the whole list would get: a row's
    //  sums then do not depend on how the list was cut)
    const bool wide = (rows_total ? rows_total : nseries) < (uint32_t)(24 * num_sms);
#define FM2_LAUNCH(KK, NWW, MINBB)                                                                              \
    do {                                                                                                        \
        const size_t smem = fm::Cfg2<KK>::smem(NWW);                                                            \
        auto kfn = fm::f_update_mma2_kernel<KK, NWW, MINBB, MODE>;                                             \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        unsigned grid = (unsigned)(MINBB * num_sms);                                                            \
        if (grid > nseries) grid = nseries;                                                                     \
        kfn<<<grid ? grid : 1, NWW * 32, smem, st>>>(ptr, idx, val, reinterpret_cast<const unsigned char *>(Xs), invs, F, Gout, lambda, \
                                                     nseries, queue, Wv, gaccum, frow, ysc);                    \
    } while (0)
#define FM2_CASE(KK, NWIDE, MINBB)                                                                              \
    case KK:                                                                                                    \
        if (wide) FM2_LAUNCH(KK, NWIDE, 1);                                                                     \
        else if constexpr (MODE == fm::MODE_GONLY) {                                                            \
            if (gonly_nw == 2) FM2_LAUNCH(KK, 2, 2 * MINBB); else FM2_LAUNCH(KK, 4, MINBB);                     \
        } else FM2_LAUNCH(KK, 4, MINBB);                                                                        \
        break;
    // rows of the complement lists are short (~1000 cells at C2 against ~9000 observed ones): a row goes to two warps instead of
    // four (twice the rows in flight per SM, half the partials to add per row) -- 0.368 -> 0.351 ms per pass at C2, C3's X-update
    // 0.947 -> 0.908 ms; TRMF_B200_GONLY_WARPS=4 restores the four-warp CTAs
    int gonly_nw = 2;
    if (const char *e = getenv("TRMF_B200_GONLY_WARPS")) gonly_nw = atoi(e);
    (void)gonly_nw;
    switch (k) {
        FM2_CASE(8, 16, 4) FM2_CASE(12, 16, 4) FM2_CASE(16, 16, 4) FM2_CASE(20, 16, 4) FM2_CASE(24, 16, 4) FM2_CASE(28, 16, 4)
        FM2_CASE(32, 16, 4) FM2_CASE(36, 16, 4) FM2_CASE(40, 16, 4)
        FM2_CASE(44, 12, 3) FM2_CASE(48, 12, 3)
        FM2_CASE(52, 8, 2) FM2_CASE(56, 8, 2) FM2_CASE(60, 8, 2) FM2_CASE(64, 8, 2)
        default: return 1;
    }
#undef FM2_CASE
#undef FM2_LAUNCH
    ++*launches;
    return cudaGetLastError() != cudaSuccess;
}

#else   // float64 build

static inline int f_update_mma2_split_y(cudaStream_t, int, const V *, const uint64_t *, uint32_t, uint32_t *, float *, unsigned *,
                                        unsigned long long *) { return 1; }
template <int MODE>
static inline int f_update_mma2_launch(cudaStream_t, int, const uint64_t *, const uint32_t *, const V *, const V *, size_t, V *, float *,
                                       V *, V *, int, double, uint32_t, unsigned *, unsigned long long *, const V * = nullptr, int = 0,
                                       double * = nullptr, bool = true, const float * = nullptr, V * = nullptr, uint32_t = 0) { return 1; }
#endif
