// f_update_mma.cuh -- K1 on the warp-level tensor path: per-series Gram by split-fp16 mma.sync + fp64 Cholesky
// (fp32 storage build only).
//
// Replaces the hot loop of l2r_ls_pY_IX_chol::solve (reference trmf.cpp:382-395): per observed entry the
// reference does k(k+1)/2 + k scalar multiply-adds into a k x k buffer.  The Gram G = sum_e x_e x_e^T is a
// SYRK with M = N = k <= 64 and K = |Omega_j|.  tcgen05 cannot be fed by it (its tile is 128 rows of ONE
// accumulator; a series' Gram has k <= 64 rows and every series its own K range), but the warp-level
// m16n8k16 tile fits any k: a warp holds the whole upper triangle of the Gram as NT 16x8 accumulator tiles
// (9 at k = 40) and per 16 observed entries issues NT x 3 HMMAs.
//
// Measured on B200 (tools/microbench_mma.cu, microbench_mma2.cu; profiles/): a legacy HMMA (TF32 m16n8k8 or
// fp16 m16n8k16 alike) holds the SM sub-partition's issue port for 8 clk and every other instruction adds 1 clk
// on top (strictly additive), so the cost of this loop is 8 x #HMMA + #other instructions.  That is why the
// operands are fp16 pairs (2048 MACs per HMMA) rather than TF32 (1024): 6.0 against 9.5 clk per entry per SM
// for the bare loop at k = 40, and 20 clk per entry per SM for the FFMA formulation (f_update_tiled.cuh).
//
// Accuracy (the 1e-5 parity bar is on the factors; the fp32 reference build itself is 1e-5 off its fp64 twin):
//  * the factor is first scaled per latent dimension by a power of two so that its largest magnitude lands in
//    [2^14, 2^15) (colscale kernels below; exact, undone exactly in the epilogue) -- fp16's narrow exponent
//    range then covers 2^-17 of every column's maximum with full precision and everything below with an
//    absolute error of 2^-39 of that maximum;
//  * split: x = h1 + h2, h1 = fp16(x), h2 = fp16(x - h1): 22 significant bits; the product keeps h1*h1 + h1*h2
//    + h2*h1 and drops h2*h2 (2^-22 relative to the dropped-term-free product's own rounding);
//  * the tensor core adds with truncation, so it is only trusted with the 16 entries x 3 terms of one tile
//    (small terms first); tiles are summed by FADD (round to nearest) for at most 128 entries and those partial
//    sums go into per-warp fp64 accumulators in shared memory.  Measured Gram error against fp64 over 128
//    entries: 7.0e-8 relative Frobenius, -4.5e-8 mean (tensor-core accumulation over 128 entries instead:
//    -1.1e-6 bias).
//
// Work decomposition: CTA = NW warps, one series at a time (atomic queue).  The series' entries are cut into
// tiles of 16; warp w takes tiles w, w + NW, ...  Each warp runs its own 2-stage cp.async (LDGSTS) pipeline
// (16 factor rows of k floats per stage, the Y values ride along with 4-byte copies) -- no CTA barrier inside
// a series.  Fragment loads are bank-conflict free because the staging row stride RS = 8 (mod 16) floats.
// At the end of a series the NW fp64 partials are added in warp order (bitwise reproducible) and the scaling is
// undone; the assembled system is solved by chol_solve_kernel (MODE_DEFER, default) or in place (MODE_SOLVE) with
// the blocked fp64 Cholesky of common.cuh, or stored as the X-update's Gram (MODE_STORE / MODE_GRAD).
#pragma once
#include "common.cuh"

#ifdef TRMF_F32

namespace fm {

constexpr int ET = 16;        // entries per tile (one m16n8k16 K-step)
constexpr int FLUSH = 8;      // tiles between fp32 -> fp64 flushes (128 entries)

template <int K> struct Cfg {
    static constexpr int NC = (K + 7) / 8;          // 8-wide chunks of the factor index
    static constexpr int MT = (NC + 1) / 2;         // 16-row accumulator tiles
    static constexpr int ntiles_() { int n = 0; for (int mt = 0; mt < MT; ++mt) n += NC - 2 * mt; return n; }
    static constexpr int NT = ntiles_();            // upper-triangle 16x8 tiles (mt, nt), nt >= 2 mt
    static constexpr int CH = K / 4;                // 16-byte pieces of a factor row
    static constexpr int RS = (NC & 1) ? 8 * NC : 8 * NC + 8;   // staging row stride: >= 8 NC and = 8 (mod 16)
    static constexpr int STAGES = 2;               // a warp spends ~1000 clk on a tile: two stages already cover L2 latency  // a warp spends >1000 clk on a tile: two stages already cover L2 latency
    static constexpr int NQ = (ET * CH + 31) / 32;  // warp-wide LDGSTS per tile
    static constexpr int STAGE_FLOATS = ET * RS;
    static constexpr int ld = K + 1;
    static constexpr size_t a_bytes = sizeof(double) * ((size_t)(K + 1) * (K + 1) + K);
    // per-warp fp64 partial in fragment layout: 4 x 32 values per tile, but the first tile of every 16-row group
    // (nt = 2 mt) only keeps its upper 8 rows -- rows g+8 there lie below the diagonal
    static constexpr int PW = NT * 128 - MT * 64;
    __host__ __device__ static constexpr int poff(int mt, int tin) { int t = tin; for (int m = 0; m < mt; ++m) t += NC - 2 * m; return 128 * t - 64 * (mt + (tin > 0 ? 1 : 0)); }
    static constexpr size_t part_bytes(int nw) { return sizeof(double) * (size_t)nw * (PW + 8 * NC); }
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

// two fp32 -> one register of two fp16 (round to nearest): .lo = a, .hi = b
__device__ __forceinline__ uint32_t pack_h2(const float a, const float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// opaque register copy: a value that sits in an A quad AND a B pair needs two homes; hiding the copy from the
// optimiser gives it exactly one MOV per value instead of a re-pack in front of every HMMA
__device__ __forceinline__ uint32_t reg_copy(const uint32_t v) {
    uint32_t r;
    asm("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(const uint32_t p) {
    float2 f;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(f.x), "=f"(f.y) : "r"(p));
    return f;
}
// d (16x8, fp32) = a (16x16, fp16, row) * b (16x8, fp16, col) + c; fragment layout of PTX mma.m16n8k16, g = lane / 4,
// t = lane % 4: a0 (g, 2t..2t+1) a1 (g+8, 2t..2t+1) a2 (g, 2t+8..2t+9) a3 (g+8, 2t+8..2t+9); b0 (2t..2t+1, g)
// b1 (2t+8..2t+9, g); c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).  The contraction index is free to
// permute: slots (2t, 2t+1, 2t+8, 2t+9) carry the tile's entries (t, t+4, t+8, t+12).
__device__ __forceinline__ void mma_zero(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}
__device__ __forceinline__ void mma_acc(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---- per-column power-of-two scaling of the factor (exact) ----
// colmax[c] = max_i |X[i][c]| as the bit pattern of a non-negative float (atomicMax on the bits is order-free)
__global__ void colscale_max_kernel(const float *__restrict__ X, size_t rows, int k, unsigned *__restrict__ colmax) {
    const size_t total = rows * (size_t)k;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // a thread's column is fixed when the stride is a multiple of k: round the stride down to one
    const size_t step = stride - stride % k;
    const size_t p0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p0 >= step) return;
    float m = 0.f;
    for (size_t p = p0; p < total; p += step) m = fmaxf(m, fabsf(X[p]));
    if (m > 0.f) atomicMax(colmax + (p0 % k), __float_as_uint(m));
}
// scale[c] = 2^(15 - e) with max = f * 2^e, f in [0.5, 1): the scaled column maximum lies in [2^14, 2^15).
// Xs = X * scale (exact), invs = 1 / scale.  All-zero / non-finite columns keep scale 1.
__global__ void colscale_apply_kernel(const float *__restrict__ X, size_t rows, int k, const unsigned *__restrict__ colmax,
                                      float *__restrict__ Xs, float *__restrict__ invs) {
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
    const size_t total = rows * (size_t)k;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x)
        Xs[p] = X[p] * sc[p % k];
}

// MODE_SOLVE : F-update -- solve (Gram + lambda I) f = rhs and store the k results in F[j].
// MODE_STORE : "store" mode used by the X-update (rows = time stamps, X = the series factor): the k x k
//              Gram (full symmetric square, fp32) goes to Gout[j] and the rhs to F[j]; rows without entries
//              get zeros.
// MODE_GRAD  : MODE_STORE, and the same gather also evaluates the sparse loss at the point Wv (arr_ls_pY_IX::fun
//              and ::grad, trmf.cpp:231-267): per entry z = <Wv_j, x_e> - Y_je (fp32 dot), F[j] (+)= sum_e z x_e
//              (fp32 FMAs, fp64 every 128 entries) and frow[j] = sum_e z^2 (fp64) -- the X-update then needs no
//              separate walk over Omega for fun(w) / grad(w).
// MODE_DEFER : MODE_SOLVE without the solve: the assembled fp64 system (lower triangle + rhs row, (K+1) x (K+1)) goes
//              to sys[j] and chol_solve_kernel factors it afterwards.  A CTA that solves in place keeps its 4 warps
//              off the tensor pipe for the length of a 128-thread Cholesky; with few entries per series (C5: 2 000
//              at k = 64) that is half of the F-update, while the separate kernel runs 6-16 solves per SM at once.
// X is the column-scaled factor (colscale_apply_kernel), invs its inverse scales.
// MODE_GONLY : (f_update_mma2.cuh only) MODE_DEFER without a right-hand side: the Gram alone goes to sys[j]; used over the COMPLEMENT of
//              the observed set (complement.cuh), where there are no Y values.
enum { MODE_SOLVE = 0, MODE_STORE = 1, MODE_GRAD = 2, MODE_DEFER = 3, MODE_GONLY = 4 };
template <int K, int NW, int MINB, int MODE>
__global__ void __launch_bounds__(NW * 32, MINB)
f_update_mma_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                    const float *__restrict__ X, const float *__restrict__ invs, float *__restrict__ F,
                    float *__restrict__ Gout, double lambda, uint32_t nseries, unsigned *__restrict__ queue,
                    const float *__restrict__ Wv, int gaccum, double *__restrict__ frow) {
    typedef Cfg<K> C;
    constexpr bool SOLVE = MODE == MODE_SOLVE || MODE == MODE_DEFER, GRAD = MODE == MODE_GRAD, DEFER = MODE == MODE_DEFER;
    constexpr int NC = C::NC, MT = C::MT, NT = C::NT, CH = C::CH, RS = C::RS, STAGES = C::STAGES, NQ = C::NQ;
    constexpr int SF = C::STAGE_FLOATS, ld = C::ld, NTH = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PW = C::PW;
    double *P = reinterpret_cast<double *>(smem_raw);                 // [NW][PW]      per-warp Gram partials (fragment layout)
    double *R = P + (size_t)NW * PW;                                  // [NW][8*NC]    per-warp rhs partials
    constexpr size_t PART_BYTES = sizeof(double) * (size_t)NW * (PW + 8 * NC);
    float *stage = reinterpret_cast<float *>(smem_raw + PART_BYTES);   // [NW][STAGES][ET*RS]
    float *ystage = stage + (size_t)NW * STAGES * SF;                 // [NW][STAGES][ET]
    double *A = reinterpret_cast<double *>(stage);                    // epilogue only: (K+1) x ld lower triangle + rhs row
    double *dinv = A + (size_t)(K + 1) * ld;
    __shared__ unsigned next_series;
    __shared__ double fwarp[NW];                                      // MODE_GRAD: per-warp sum of squared residuals

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;

    for (int p = tid; p < NW * STAGES * SF; p += NTH) stage[p] = 0.f;   // padding columns stay zero
    if (tid == 0) next_series = atomicAdd(queue, 1u);
    __syncthreads();
    uint32_t j = next_series;

    float *st = stage + (size_t)warp * STAGES * SF;
    float *ys = ystage + (size_t)warp * STAGES * ET;
    double *Pw = P + (size_t)warp * PW;
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
                float wl[NC];          // MODE_GRAD: this lane's columns of the point, pre-divided by the column scales
                double fsum = 0.0;
                if (GRAD) {
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        wl[c] = (g + 8 * c < K) ? __ldg(Wv + (size_t)j * K + g + 8 * c) * __ldg(invs + g + 8 * c) : 0.f;
                }

                uint32_t nidx;   // lane e (and e + 16): row index of entry e of the next tile to issue
                auto load_idx = [&](int i) {   // clamped: always a valid address, also one tile past the end
                    uint32_t e = (uint32_t)(warp + i * NW) * ET + (lane & 15);
                    e = e < nnz ? e : nnz - 1;
                    nidx = __ldg(sidx + e);
                };
                auto issue = [&](int i, int s) {   // gathers local tile i (indices in nidx) into stage s
                    const uint32_t base = (uint32_t)(warp + i * NW) * ET;
                    float *dst = st + s * SF;
                    if (base + ET <= nnz) {        // full tile (warp-uniform): no per-request predicates
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const int id = q * 32 + lane;
                            const int e = id / CH, cc = id - e * CH;
                            const uint32_t row = __shfl_sync(FULL_MASK, nidx, e & 15);
                            if ((ET * CH) % 32 == 0 || id < ET * CH) cp_async16(dst + e * RS + 4 * cc, X + (size_t)row * K + 4 * cc);
                        }
                        if (lane < ET) cp_async4(ys + s * ET + lane, sval + base + lane);
                    } else {
                        const int cnt = (int)(nnz - base);
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const int id = q * 32 + lane;
                            const int e = id / CH, cc = id - e * CH;
                            const uint32_t row = __shfl_sync(FULL_MASK, nidx, e & 15);
                            if (id < ET * CH && e < cnt) cp_async16(dst + e * RS + 4 * cc, X + (size_t)row * K + 4 * cc);
                        }
                        if (lane < cnt) cp_async4(ys + s * ET + lane, sval + base + lane);
                        else if (lane < ET) ys[s * ET + lane] = 0.f;
                        for (int p = cnt * RS + lane; p < SF; p += 32) dst[p] = 0.f;   // rows past the end read as zero
                    }
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
                                        double *d = Pw + C::poff(mt, nt - 2 * mt) + q * 32 + lane;
                                        *d = first ? (double)acc[t][q] : *d + (double)acc[t][q];
                                    }
                                    acc[t][q] = 0.f;
                                }
                                ++t;
                            }
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
                    {
                        const float *p = tb + tig * RS + g;
                        float y[4];   // weights of the tile's entries (t, t+4, t+8, t+12) in the rhs sum: Y, or the residual
#pragma unroll
                        for (int q = 0; q < 4; ++q) y[q] = yb[tig + 4 * q];
                        if (GRAD) {
                            float z[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                z[q] = 0.f;
#pragma unroll
                                for (int c = 0; c < NC; ++c) z[q] = fmaf(wl[c], p[(4 * q) * RS + 8 * c], z[q]);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {   // fold the 8 lanes that share entry t + 4q (same bits on every lane)
                                z[q] += __shfl_xor_sync(FULL_MASK, z[q], 4);
                                z[q] += __shfl_xor_sync(FULL_MASK, z[q], 8);
                                z[q] += __shfl_xor_sync(FULL_MASK, z[q], 16);
                            }
                            if (g == 0) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) { const double rr = (double)y[q] - (double)z[q]; fsum += rr * rr; }
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) y[q] = z[q] - y[q];   // rows past the end: x = 0, Y = 0 -> 0
                        }
                        uint32_t a1[MT][4], a2[MT][4], b1[NC][2], b2[NC][2];   // h1 / h2 parts, A- and B-arranged
#pragma unroll
                        for (int c = 0; c < 2 * MT; ++c) {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {   // contraction slots 2t..2t+1 (+8): entries (t + 8 hh, t + 8 hh + 4)
                                uint32_t p1 = 0u, p2 = 0u;     // virtual chunk past k (NC odd): zero rows of the last 16-row tile
                                if (c < NC) {
                                    const float x0 = p[(8 * hh) * RS + 8 * c], x1 = p[(8 * hh + 4) * RS + 8 * c];
                                    racc[c] = fmaf(y[2 * hh], x0, racc[c]);
                                    racc[c] = fmaf(y[2 * hh + 1], x1, racc[c]);
                                    p1 = pack_h2(x0, x1);
                                    const float2 f = unpack_h2(p1);
                                    p2 = pack_h2(x0 - f.x, x1 - f.y);
                                    b1[c][hh] = reg_copy(p1); b2[c][hh] = reg_copy(p2);
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
                                mma_zero(d, a2[mt], b1[nt]);     // small terms first
                                mma_acc(d, a1[mt], b2[nt]);
                                mma_acc(d, a1[mt], b1[nt]);
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
                if (GRAD) {
                    fsum += __shfl_xor_sync(FULL_MASK, fsum, 1);
                    fsum += __shfl_xor_sync(FULL_MASK, fsum, 2);
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
                for (int w = 1; w < nwa; ++w) s += R[w * 8 * NC + c];
                A[K * ld + c] = s * (double)invs[c];
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
            if (RS > K) {   // A overwrote the staging area: padding columns must read as zero again
                __syncthreads();
                for (int p = tid; p < NW * STAGES * ET; p += NTH)
                    for (int c = K; c < RS; ++c) stage[(size_t)p * RS + c] = 0.f;
            }
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

// The solve of MODE_DEFER: one CTA per system at a time, many CTAs per SM.  Same arithmetic as the in-kernel solve
// (block_chol_solve_blocked does not depend on the block size), so the factors are bit-identical either way.
// NTH threads per system (32 / 64 / 128, chosen by the launcher): see cm::solve_kernel -- the factorisation is a chain of short
// dependent steps, more systems in flight per SM beat more threads per system.
template <int K, int NTH>
__global__ void __launch_bounds__(NTH)
chol_solve_kernel(const uint64_t *__restrict__ ptr, const double *__restrict__ sys, float *__restrict__ F, double lambda,
                  uint32_t nseries) {
    constexpr int ld = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *A = reinterpret_cast<double *>(smem_raw);
    double *dinv = A + (size_t)(K + 1) * ld;
    const int tid = threadIdx.x;
    for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
        if (ptr[j + 1] == ptr[j]) continue;             // no observation: the row keeps its value (trmf.cpp:374)
        const double *src = sys + (size_t)j * ((K + 1) * ld);
        for (int p = tid; p < (K + 1) * ld; p += NTH) A[p] = src[p];
        __syncthreads();
        for (int c = tid; c < K; c += NTH) A[c * ld + c] += lambda;       // trmf.cpp:393
        block_chol_solve_blocked<(K + 32) / 32>(A, ld, dinv, K);   // starts and ends with __syncthreads
        for (int c = tid; c < K; c += NTH) F[(size_t)j * K + c] = (float)A[K * ld + c];
        __syncthreads();
    }
}

// One warp per system on packed triangular storage (warp_chol_solve_packed): the default of the deferred solve.
template <int K>
__global__ void __launch_bounds__(32)
chol_solve_warp_kernel(const uint64_t *__restrict__ ptr, const double *__restrict__ sys, float *__restrict__ F, double lambda,
                       uint32_t nseries) {
    constexpr int ld = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *A = reinterpret_cast<double *>(smem_raw);
    double *dinv = A + (K * (K + 1) / 2 + K);
    const int lane = threadIdx.x;
    for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
        if (ptr[j + 1] == ptr[j]) continue;             // no observation: the row keeps its value (trmf.cpp:374)
        const double *src = sys + (size_t)j * ((K + 1) * ld);
#pragma unroll 4
        for (int p = lane; p < (K + 1) * ld; p += 32) {      // (flat and unrolled: several loads in flight)
            const int r = p / ld, c = p - r * ld;
            if (c <= r && c < K) A[tri_off(r) + c] = src[p] + (c == r ? lambda : 0.0);   // trmf.cpp:393
        }
        warp_chol_solve_packed<(K + 32) / 32>(A, dinv, K);   // starts and ends with __syncwarp
        for (int c = lane; c < K; c += 32) F[(size_t)j * K + c] = (float)A[tri_off(K) + c];
        __syncwarp();
    }
}

// out = scale * sum_i v[i], deterministic two-level fp64 sum (MODE_GRAD's per-row losses -> the objective's loss part)
__global__ void sum_rows_kernel(const double *__restrict__ v, size_t n, double scale, double *part, unsigned *ticket, double *out) {
    __shared__ double red[32];
    double a = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) a += v[p];
    a = block_sum(a, red);
    grid_sum_commit(a, part, ticket, out, scale, red);
}

}   // namespace fm

static inline size_t f_update_mma_sys_doubles(int k) { return (size_t)(k + 1) * (k + 1); }

// solve the nseries systems a MODE_DEFER launch left in `sys`
static inline int f_update_mma_solve(cudaStream_t st, int num_sms, const uint64_t *ptr, const double *sys, V *F, int k, double lambda,
                                     uint32_t nseries, unsigned long long *launches) {
#define FS_LAUNCH(KK, NTH)                                                                                      \
    do {                                                                                                        \
        const size_t smem = sizeof(double) * ((size_t)(KK + 1) * (KK + 1) + KK);                                \
        auto kfn = fm::chol_solve_kernel<KK, NTH>;                                                              \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        int per_sm = 0;                                                                                         \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, NTH, smem) != cudaSuccess) return 1;    \
        unsigned grid = (unsigned)(per_sm > 0 ? per_sm : 1) * (unsigned)num_sms;                                \
        if (grid > nseries) grid = nseries;                                                                     \
        kfn<<<grid ? grid : 1, NTH, smem, st>>>(ptr, sys, F, lambda, nseries);                                  \
    } while (0)
#define FS_WARP(KK)                                                                                             \
    do {                                                                                                        \
        const size_t smem = sizeof(double) * ((size_t)KK * (KK + 1) / 2 + 2 * KK);                              \
        auto kfn = fm::chol_solve_warp_kernel<KK>;                                                              \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        int per_sm = 0;                                                                                         \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, 32, smem) != cudaSuccess) return 1;     \
        unsigned grid = (unsigned)(per_sm > 0 ? per_sm : 1) * (unsigned)num_sms;                                \
        if (grid > nseries) grid = nseries;                                                                     \
        kfn<<<grid ? grid : 1, 32, smem, st>>>(ptr, sys, F, lambda, nseries);                                   \
    } while (0)
#define FS_CASE(KK)                                                                                             \
    case KK:                                                                                                    \
        if (nth == 32) FS_WARP(KK); else if (nth == 64) FS_LAUNCH(KK, 64); else FS_LAUNCH(KK, 128);             \
        break;
    // 128 / 64 = one CTA per system, 32 = one warp per system on packed storage.  Measured (F-update, ms): C4 32.33 / 32.48, C5
    // 186.2 / 184.6 for 128 / 32 -- the solve is not what bounds those F-updates; identical factors either way.
    int nth = 128;
    if (const char *e = getenv("TRMF_B200_SOLVE_THREADS")) nth = atoi(e);
    switch (k) {
        FS_CASE(8) FS_CASE(12) FS_CASE(16) FS_CASE(20) FS_CASE(24) FS_CASE(28) FS_CASE(32) FS_CASE(36) FS_CASE(40) FS_CASE(44) FS_CASE(48) FS_CASE(52)
        FS_CASE(56) FS_CASE(60) FS_CASE(64)
        default: return 1;
    }
#undef FS_CASE
#undef FS_LAUNCH
#undef FS_WARP
    ++*launches;
    return cudaGetLastError() != cudaSuccess;
}

static inline bool f_update_mma_supported(int k) {
    // every multiple of 4 up to 64
    return k >= 8 && k <= 64 && k % 4 == 0;
}

// returns 0 on success.  `wide` selects one 16-warp CTA per SM (few series: finer load balance) instead of
// four 4-warp CTAs per SM (16 warps at 128 registers; measured 3.20 ms against 3.50 ms for 3 x 4 warps at 168
// registers at C2 -- with the issue port held 8 clk per HMMA, more resident warps is what hides the rest).
template <int MODE>
static inline int f_update_mma_launch(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *idx, const V *val,
                                      const V *X, size_t xrows, V *Xs, float *invs, V *F, V *Gout, int k, double lambda,
                                      uint32_t nseries, unsigned *queue, unsigned long long *launches,
                                      const V *Wv = nullptr, int gaccum = 0, double *frow = nullptr, bool rescale = true) {
    // queue[0] = series counter, queue[8 .. 8+k) = per-column max |x| (bit patterns).  rescale = false: Xs / invs are
    // still valid from the previous launch over the same factor (series-slab launches of one F-update).
    if (cudaMemsetAsync(queue, 0, sizeof(unsigned) * (rescale ? 8 + 128 : 1), st) != cudaSuccess) return 1;
    if (rescale) {
        const size_t total = xrows * (size_t)k;
        unsigned g1 = (unsigned)((total + 255) / 256);
        if (g1 > (unsigned)(4 * num_sms)) g1 = (unsigned)(4 * num_sms);
        if (g1 == 0) g1 = 1;
        if ((size_t)g1 * 256 < (size_t)k) g1 = (unsigned)((k + 255) / 256);
        fm::colscale_max_kernel<<<g1, 256, 0, st>>>(X, xrows, k, queue + 8);
        fm::colscale_apply_kernel<<<g1, 256, sizeof(float) * k, st>>>(X, xrows, k, queue + 8, Xs, invs);
        *launches += 2;
    }
    const bool wide = nseries < (uint32_t)(24 * num_sms);
#define FM_LAUNCH(KK, NWW, MINBB)                                                                               \
    do {                                                                                                        \
        const size_t smem = fm::Cfg<KK>::smem(NWW);                                                             \
        auto kfn = fm::f_update_mma_kernel<KK, NWW, MINBB, MODE>;                                              \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        unsigned grid = (unsigned)(MINBB * num_sms);                                                            \
        if (grid > nseries) grid = nseries;                                                                     \
        kfn<<<grid ? grid : 1, NWW * 32, smem, st>>>(ptr, idx, val, Xs, invs, F, Gout, lambda, nseries, queue, Wv, gaccum, frow);        \
    } while (0)
#define FM_CASE(KK)                                                                                             \
    case KK:                                                                                                    \
        if (wide) FM_LAUNCH(KK, 16, 1); else FM_LAUNCH(KK, 4, 4);                                               \
        break;
    switch (k) {
        FM_CASE(8) FM_CASE(16) FM_CASE(20) FM_CASE(24) FM_CASE(32) FM_CASE(40)
        case 48:   // 12 accumulator tiles: 18 KB of shared memory per warp -> 12 warps per SM
            if (wide) FM_LAUNCH(48, 12, 1); else FM_LAUNCH(48, 4, 3);
            break;
        case 56:   // 16 / 20 accumulator tiles: 8 warps per SM
            if (wide) FM_LAUNCH(56, 8, 1); else FM_LAUNCH(56, 4, 2);
            break;
        case 60:
            if (wide) FM_LAUNCH(60, 8, 1); else FM_LAUNCH(60, 4, 2);
            break;
        case 64:
            if (wide) FM_LAUNCH(64, 8, 1); else FM_LAUNCH(64, 4, 2);   // (4 CTAs of 2 warps: 230 vs 214 ms at C5 -- measured, rejected)
            break;
        // the in-between ranks (added in round 2 so that no k % 4 == 0 falls back to the FFMA / generic kernels): the narrow
        // shape of their next larger neighbour only
        case 12: FM_LAUNCH(12, 4, 4); break;
        case 28: FM_LAUNCH(28, 4, 4); break;
        case 36: FM_LAUNCH(36, 4, 4); break;
        case 44: FM_LAUNCH(44, 4, 3); break;
        case 52: FM_LAUNCH(52, 4, 2); break;
        default: return 1;
    }
#undef FM_CASE
#undef FM_LAUNCH
    ++*launches;
    return cudaGetLastError() != cudaSuccess;
}

#else   // float64 build: the generic kernel (fp64 FMAs) is the parity path

static inline bool f_update_mma_supported(int) { return false; }
static inline size_t f_update_mma_sys_doubles(int) { return 0; }
static inline int f_update_mma_solve(cudaStream_t, int, const uint64_t *, const double *, V *, int, double, uint32_t, unsigned long long *) { return 1; }
namespace fm {
enum { MODE_SOLVE = 0, MODE_STORE = 1, MODE_GRAD = 2, MODE_DEFER = 3, MODE_GONLY = 4 };
__global__ void sum_rows_kernel(const double *, size_t, double, double *, unsigned *, double *) {}
}
template <int MODE>
static inline int f_update_mma_launch(cudaStream_t, int, const uint64_t *, const uint32_t *, const V *, const V *, size_t, V *, float *,
                                      V *, V *, int, double, uint32_t, unsigned *, unsigned long long *, const V * = nullptr, int = 0,
                                      double * = nullptr, bool = true) { return 1; }
#endif
