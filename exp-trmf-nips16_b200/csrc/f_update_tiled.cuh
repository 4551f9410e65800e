// f_update_tiled.cuh -- K1, the headline kernel: register-tiled symmetric Gram build +
// fp64 Cholesky per series (fp32 storage build only).
//
// Replaces the hot loop of l2r_ls_pY_IX_chol::solve (reference trmf.cpp:382-395): per
// observed entry the reference does k(k+1)/2 + k scalar multiply-adds into a k x k buffer.
// Here one CTA (7 "Gram" warps + 1 "rhs" warp) owns one series at a time:
//
//  * the observed rows x_i (k floats each, gathered through the by-series index list)
//    are staged into shared memory by cp.async (LDGSTS, 16 B per request, no register
//    staging): a lane owns a fixed (row-in-instruction, chunk) slot and a warp every 8th
//    warp-wide request of a tile, the row indices are prefetched into registers one tile
//    ahead, so a request costs ~4 instructions and a warp ~25 per tile; 3 stages deep, one
//    __syncthreads per tile.  (Measured alternatives that lost: all copies from the light
//    warp -- it becomes the straggler at the barrier; per-stage mbarriers instead of the
//    barrier -- the spin loops eat the issue slots; TMA row copies -- see below.)
//  * the k indices are split into NB = KP/8 sets R(t) of 8 (two 16-byte chunks, t and
//    t+NB, so that lanes of a group read distinct bank groups); the upper triangle of
//    the Gram is the B = NB(NB+1)/2 blocks R(bi) x R(bj), bi <= bj; a Gram thread is
//    (group g, block b): it keeps the 8x8 block in 64 fp32 registers and, per entry,
//    loads 4 x LDS.128 and issues 64 FFMA.  G = 224/B groups work on different entries
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

// One CTA = 8 warps x 128 registers, two CTAs per SM (the register file is handed out in pairs of
// warps: a 9-warp CTA is billed as 10 and drops to one CTA/SM -- measured).  Roles: 6 "Gram" warps,
// one producer warp (all cp.async traffic) and one rhs warp.  Which warp ids play which role depends
// on the CTA's "flavor" (0 or 1 = arrival order on its SM) so that the two co-resident CTAs together
// put exactly 3 Gram warps + 1 light warp on each of the SM's 4 schedulers (warp id % 4).
constexpr int NT = 256;
constexpr int NGRAM = 192;
#ifndef TRMF_FT_MBAR
#define TRMF_FT_MBAR 0   // 1: per-stage full/empty mbarriers (warps drift freely); 0: one __syncthreads per tile
#endif                   // (measured equal at C2: 6.77 vs 6.79 ms; the barrier version issues 13 % fewer instructions)
#ifndef TRMF_FT_FFMA2
#define TRMF_FT_FFMA2 1  // 1: packed fma.rn.f32x2 (sm_100+: two IEEE fp32 FMAs per issue slot); 0: scalar FFMA
#endif

template <int K> struct Cfg {
    static constexpr int NB = (K + 7) / 8;
    static constexpr int KP = 8 * NB;
    static constexpr int CH = K / 4;           // 16-byte chunks per factor row
    static constexpr int B = NB * (NB + 1) / 2;
    static constexpr int G = NGRAM / B;
    static constexpr int U = NB == 1 ? 1 : NB == 2 ? 2 : NB == 3 ? 3 : NB == 4 ? 5 : NB == 5 ? 8 : NB == 6 ? 10 : NB == 7 ? 12 : 10;
    static constexpr int ET = G * U;           // entries per tile
    static constexpr int RS = KP + 4;          // smem row stride (floats): odd number of 16-B chunks
    static constexpr int STAGES = 3;
    static constexpr int FL = (128 / U) > 0 ? (128 / U) : 1;   // tiles between fp32 -> fp64 flushes
    static constexpr int FBUF = G * B * 32;    // floats: half of every lane's 8x8 block
    static constexpr int RBUF = 32 * KP;       // floats: rhs partials of the 32 rhs lanes
    static constexpr int NQ = (ET + 31) / 32;  // tile rows per lane of a light warp
    static constexpr int RPI = 32 / CH;        // factor rows moved by one warp-wide LDGSTS
    static constexpr int NCI = (ET + RPI - 1) / RPI;   // warp-wide LDGSTS instructions per tile
    // group g works on tile rows u*G + (g*D mod G): neighbouring groups sit D rows apart, which
    // keeps quarter-warps that straddle two groups off the same shared-memory banks
    static constexpr int gcd_(int a, int b) { return b == 0 ? a : gcd_(b, a % b); }
    static constexpr int pick_d_() {
        for (int d = 1; d < G; ++d) {
            const int shift = (d * (RS / 4)) % 8;   // bank-group offset between neighbouring groups' rows
            if (gcd_(d, G) == 1 && shift >= 5) return d;
        }
        return 1;
    }
    static constexpr int D = pick_d_();
};

// d.{x,y} += a.{x,y} * b.{x,y}: one FFMA2 (bitwise the same results as two fmaf)
__device__ __forceinline__ void ffma2(float2 &d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long *>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long *>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long *>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2 *>(&dd);
}

// ---- cp.async (LDGSTS) primitives ----
// (A TMA variant -- one cp.async.bulk per gathered 160-byte row, mbarrier completion -- was
//  measured 3x slower end to end: per-row bulk requests are far too small for the copy engine.)
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(s), "l"(gmem), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- mbarrier primitives: per-stage "full" (data landed) / "empty" (all warps done) barriers ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive once every cp.async this thread has issued so far has landed (counted in the barrier's expected arrivals)
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int K>
static size_t smem_bytes() {
    typedef Cfg<K> C;
    const int k = K;
    size_t dbl = (size_t)(k + 1) * (k + 1) + k;
    dbl = (dbl + 1) & ~(size_t)1;   // keep the float region 16-B aligned
    return dbl * sizeof(double) + sizeof(float) * ((size_t)C::STAGES * C::ET * C::RS + C::FBUF + C::RBUF);
}

// index of member m (0..7) of set R(t): chunk t then chunk t + NB
template <int NB> __device__ __forceinline__ int set_index(int t, int m) { return m < 4 ? 4 * t + m : 4 * (t + NB) + (m - 4); }

// SOLVE = true : F-update -- solve (Gram + lambda I) f = rhs and store the k results in F[j].
// SOLVE = false: "store" mode used by the X-update (rows = time stamps, X = the series factor):
//                the k x k Gram (full symmetric square, fp32) goes to Gout[j] and the rhs to
//                F[j]; rows without entries get zeros.  The Hessian-vector products of the CG
//                solve then read these Grams instead of re-walking Omega (x_update.cuh).
template <int K, bool SOLVE>
__global__ void __launch_bounds__(NT, 2)
f_update_tiled_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                      const float *__restrict__ X, float *__restrict__ F, float *__restrict__ Gout, double lambda,
                      uint32_t nseries, unsigned *__restrict__ queue, unsigned *__restrict__ sm_slots) {
    typedef Cfg<K> C;
    constexpr int k = K, NB = C::NB, CH = C::CH;
    constexpr int KP = C::KP, B = C::B, G = C::G, U = C::U, ET = C::ET, RS = C::RS, STAGES = C::STAGES, NQ = C::NQ;
    constexpr int RPI = C::RPI, NCI = C::NCI;
    constexpr int STAGE_FLOATS = ET * RS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int ld = k + 1;
    double *A = reinterpret_cast<double *>(smem_raw);            // (k+1) x ld: lower triangle + rhs row
    double *dinv = A + (size_t)(k + 1) * ld;
    constexpr size_t dbl = (((size_t)(k + 1) * ld + k) + 1) & ~(size_t)1;
    float *tiles = reinterpret_cast<float *>(A + dbl);           // STAGES x ET x RS
    float *fbuf = tiles + (size_t)STAGES * STAGE_FLOATS;         // G x 8 x B float4
    float *rbuf = fbuf + C::FBUF;                                // KP/4 x 32 float4
    __shared__ unsigned next_series;
    __shared__ unsigned flavor_s;
#if TRMF_FT_MBAR
    // full[s]: the producer's copies into stage s have landed (32 cp.async-completion arrivals) and its tail
    // zero fill is published (1 release arrival); empty[s]: the 7 consumer warps finished reading stage s
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
#endif

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    // zero every staging row once: padding columns (k..KP+3) are never written by the copies
    for (int p = tid; p < STAGES * STAGE_FLOATS; p += NT) tiles[p] = 0.f;
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        flavor_s = atomicAdd(sm_slots + smid, 1u) & 1u;
        next_series = atomicAdd(queue, 1u);
#if TRMF_FT_MBAR
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 33); mbar_init(&empty_bar[s], NT / 32 - 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncthreads();
    uint32_t j = next_series;
#if TRMF_FT_MBAR
    int r_stage = 0;          // ring position, running over the CTA's lifetime: consumers: next tile to read;
    unsigned r_phase = 0;     // producer: next tile to issue (it runs STAGES-1 tiles ahead of the consumers)
#endif

    // ---- roles ----
    // flavor 0: Gram warps 0,1,2,3,4,5   producer 6   rhs 7     (Gram per scheduler 2,2,1,1)
    // flavor 1: Gram warps 0,1,2,3,6,7   producer 4   rhs 5     (Gram per scheduler 1,1,2,2)
    const unsigned flavor = flavor_s;
    const int w_prod = flavor ? 4 : 6, w_rhs = flavor ? 5 : 7;
    const bool is_prod = warp == w_prod, is_rhs = warp == w_rhs;
    const bool is_gram = !is_prod && !is_rhs;
    const int gw = warp < 4 ? warp : (flavor ? warp - 2 : warp);          // ordinal among the Gram warps (0..5)
    const int gtid = gw * 32 + lane;
    const int g = gtid / B, b = gtid - g * B;
    const bool active = is_gram && g < G;
    int bi = 0, bj = 0;
    {
        int rem = b;
        while (rem >= NB - bi) { rem -= NB - bi; ++bi; }
        bj = bi + rem;
    }
    const int grow = (g * C::D) % G;                 // this group's row inside every G-row slice of a tile
    const int off_a = grow * RS + 4 * bi, off_b = grow * RS + 4 * bj;
    // producer lane -> (row within a warp-wide LDGSTS, 16-byte chunk of that row)
    const int cp_r = lane / CH, cp_c = lane - cp_r * CH;
    const bool cp_on = lane < RPI * CH;
    const float *cp_src = X + cp_c * 4;
    const int cp_dst = cp_r * RS + cp_c * 4;

    while (j < nseries) {
        const uint64_t lo = ptr[j];
        const uint32_t nnz = (uint32_t)(ptr[j + 1] - lo);       // one series never holds 2^32 entries (T < 2^32)
        if (nnz != 0) {
            const uint32_t *sidx = idx + lo;
            const float *sval = val + lo;
            const int ntiles = (int)((nnz + ET - 1) / ET);
            for (int p = tid; p < (k + 1) * ld; p += NT) A[p] = 0.0;

            __align__(8) float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.f;
            // the rhs warp keeps its KP partial sums in the same registers: racc[t] == acc[t >> 3][t & 7]
#define RACC(t) acc[(t) >> 3][(t) & 7]

            // light-warp state (lane l <-> tile rows l, l+32, ...): producer: row indices of the next tile to
            // issue; rhs warp: Y values of the tile being consumed and of the next one
            uint32_t lw[2 * NQ];          // producer: lw[0..NQ) = nidx;  rhs warp: lw[0..NQ) = vcur, lw[NQ..2NQ) = vnext (as bits)
#define nidx lw
#define VCUR(q) __uint_as_float(lw[q])
#define VNEXT_SET(q, v) lw[NQ + (q)] = __float_as_uint(v)
            auto tile_count = [&](int tt) -> int {
                const uint32_t rem = nnz - (uint32_t)tt * ET;
                return (int)(rem < (uint32_t)ET ? rem : (uint32_t)ET);
            };
            auto load_idx = [&](int tt) {
                const uint32_t base = (uint32_t)tt * ET;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const uint32_t e = base + lane + 32 * q;
                    nidx[q] = (lane + 32 * q < ET && e < nnz) ? __ldg(sidx + e) : 0u;
                }
            };
            auto load_vals = [&](int tt) {
                const uint32_t base = (uint32_t)tt * ET;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const uint32_t e = base + lane + 32 * q;
                    VNEXT_SET(q, (lane + 32 * q < ET && e < nnz) ? __ldg(sval + e) : 0.f);
                }
            };
            auto issue = [&](int tt, int stage) {   // producer warp: gathers tile tt (indices in nidx) into `stage`
#if TRMF_FT_MBAR
                stage = r_stage;
                if (r_phase > 0) mbar_wait(&empty_bar[stage], (r_phase - 1) & 1u);   // previous tenant consumed by all
#endif
                float *dst = tiles + stage * STAGE_FLOATS + cp_dst;
                const int cnt = tile_count(tt);
#pragma unroll
                for (int q = 0; q < NCI; ++q) {
                    // rows RPI*q .. RPI*q + RPI-1 of the tile; their indices sit in lane e & 31, register e >> 5
                    const int e = RPI * q + cp_r;
                    constexpr int QMAX = NQ - 1;
                    const int m_lo = (RPI * q) >> 5, m_hi = (RPI * q + RPI - 1) >> 5;
                    uint32_t row = __shfl_sync(FULL_MASK, nidx[m_lo < QMAX ? m_lo : QMAX], e & 31);
                    if (m_hi != m_lo && m_hi <= QMAX) {
                        const uint32_t row2 = __shfl_sync(FULL_MASK, nidx[m_hi <= QMAX ? m_hi : QMAX], e & 31);
                        row = (e >> 5) == m_hi ? row2 : row;
                    }
                    cp_async16(dst + q * RPI * RS, cp_src + (size_t)row * k, cp_on && e < cnt);
                }
#if TRMF_FT_MBAR
                mbar_arrive_on_cp_async(&full_bar[stage]);
#endif
                if (cnt < ET)   // tail tile: stale rows from an earlier tile must read as zero
                    for (int p = cnt * RS + lane; p < STAGE_FLOATS; p += 32) tiles[stage * STAGE_FLOATS + p] = 0.f;
#if TRMF_FT_MBAR
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);   // release: publishes the zero fill
                if (++r_stage == STAGES) { r_stage = 0; ++r_phase; }
#endif
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
                    if (h == 0 && is_rhs) {
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
                            // four independent fp64 chains (fixed association: still bitwise reproducible)
                            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                            const float *fp = fbuf + u;
#pragma unroll
                            for (int gg = 0; gg + 3 < G; gg += 4) {
                                s0 += (double)fp[(size_t)(gg + 0) * 32 * B];
                                s1 += (double)fp[(size_t)(gg + 1) * 32 * B];
                                s2 += (double)fp[(size_t)(gg + 2) * 32 * B];
                                s3 += (double)fp[(size_t)(gg + 3) * 32 * B];
                            }
#pragma unroll
                            for (int gg = G & ~3; gg < G; ++gg) s0 += (double)fp[(size_t)gg * 32 * B];
                            const int hi_ = r > c ? r : c, lo_ = r > c ? c : r;
                            A[hi_ * ld + lo_] += (s0 + s1) + (s2 + s3);
                        }
                    }
                    if (h == 0 && tid < k) {
                        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                        const float *rp = rbuf + (tid >> 2) * 128 + (tid & 3);
#pragma unroll
                        for (int l = 0; l < 32; l += 4) {
                            s0 += (double)rp[(l + 0) * 4];
                            s1 += (double)rp[(l + 1) * 4];
                            s2 += (double)rp[(l + 2) * 4];
                            s3 += (double)rp[(l + 3) * 4];
                        }
                        A[k * ld + tid] += (s0 + s1) + (s2 + s3);
                    }
                    __syncthreads();
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.f;
            };

            // ---- pipeline prologue: tiles 0 .. STAGES-2 in flight, indices of tile STAGES-1 loaded ----
            if (is_prod) {
#pragma unroll
                for (int s = 0; s < STAGES - 1; ++s) {
                    if (s < ntiles) { load_idx(s); issue(s, s); }
                    cp_async_commit();
                }
                if (STAGES - 1 < ntiles) load_idx(STAGES - 1);
            }
            if (is_rhs) load_vals(0);

            int stage = 0;                         // stage holding tile t
            for (int t = 0; t < ntiles; ++t) {
#if TRMF_FT_MBAR
                if (is_prod) {
                    // the producer runs ahead on its own: refill the ring as soon as every consumer warp has
                    // released the stage (empty barrier), no CTA-wide barrier involved
                    const int nt = t + STAGES - 1;
                    if (nt < ntiles) issue(nt, 0);
                    if (nt + 1 < ntiles) load_idx(nt + 1);
                } else {
                    stage = r_stage;
                    mbar_wait(&full_bar[stage], r_phase & 1u);   // tile t has landed
#else
                if (is_prod) cp_async_wait<STAGES - 2>();   // the producer issued every copy of tile t
                __syncthreads();                   // tile t is visible to everyone; tile t-1 is fully consumed
                if (is_prod) {
                    const int nt = t + STAGES - 1;
                    int ns = stage + STAGES - 1;
                    if (ns >= STAGES) ns -= STAGES;
                    if (nt < ntiles) issue(nt, ns);
                    cp_async_commit();
                    if (nt + 1 < ntiles) load_idx(nt + 1);
                } else {
#endif
                    if (is_gram) {
                        if (active) {
                            const float *pa = tiles + stage * STAGE_FLOATS + off_a, *pb = tiles + stage * STAGE_FLOATS + off_b;
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const float4 a0 = *reinterpret_cast<const float4 *>(pa + u * G * RS);
                                const float4 a1 = *reinterpret_cast<const float4 *>(pa + u * G * RS + 4 * NB);
                                const float4 b0 = *reinterpret_cast<const float4 *>(pb + u * G * RS);
                                const float4 b1 = *reinterpret_cast<const float4 *>(pb + u * G * RS + 4 * NB);
                                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#if TRMF_FT_FFMA2
                                const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                                                      make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float2 ap = make_float2(a[i], a[i]);
#pragma unroll
                                    for (int jp = 0; jp < 4; ++jp)
                                        ffma2(*reinterpret_cast<float2 *>(&acc[i][2 * jp]), ap, bp[jp]);
                                }
#else
                                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i)
#pragma unroll
                                    for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(a[i], bv[jj], acc[i][jj]);
#endif
                            }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < NQ; ++q) lw[q] = lw[NQ + q];
                        if (t + 1 < ntiles) load_vals(t + 1);
                        const float *tb = tiles + stage * STAGE_FLOATS;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const int e = lane + 32 * q;     // rows past the tile's count are zero and carry y = 0
                            if (e < ET) {
                                const float y = VCUR(q);
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
                    }
#if TRMF_FT_MBAR
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);   // this warp is done with tile t's stage
                    if (++r_stage == STAGES) { r_stage = 0; ++r_phase; }
#endif
                }
#if !TRMF_FT_MBAR
                if (++stage == STAGES) stage = 0;
#endif
                if ((t + 1) % C::FL == 0 && t + 1 < ntiles) flush();
            }
#if TRMF_FT_MBAR
            // bring the producer's ring position in line with the consumers' for the next series: it has issued
            // exactly ntiles tiles, like they have consumed (prologue + loop), so both counters already agree
#endif
            if (is_prod) cp_async_wait<0>();
            flush();   // ends with __syncthreads
            if (SOLVE) {
                if (tid < k) A[tid * ld + tid] += lambda;      // trmf.cpp:393
                block_chol_solve_blocked<(K + 32) / 32>(A, ld, dinv, k);   // starts and ends with __syncthreads
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
#undef nidx
#undef VCUR
#undef VNEXT_SET
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

static inline bool f_update_tiled_supported(int k) {
    switch (k) { case 8: case 16: case 20: case 24: case 32: case 40: case 48: case 56: case 60: case 64: return true; }
    return false;
}

// returns 0 on success
template <bool SOLVE>
static inline int f_update_tiled_launch(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *idx, const V *val,
                                        const V *X, V *F, V *Gout, int k, double lambda, uint32_t nseries, unsigned *queue,
                                        unsigned long long *launches) {
    // queue[0] = series counter, queue[1 .. 1+num_sms) = per-SM CTA arrival counters (role flavor)
    const unsigned grid = (unsigned)(nseries < (uint32_t)(2 * num_sms) ? nseries : (uint32_t)(2 * num_sms));
    if (cudaMemsetAsync(queue, 0, sizeof(unsigned) * (size_t)(1 + num_sms), st) != cudaSuccess) return 1;
#define FT_CASE(KK)                                                                                             \
    case KK: {                                                                                                  \
        const size_t smem = ft::smem_bytes<KK>();                                                               \
        auto kfn = ft::f_update_tiled_kernel<KK, SOLVE>;                                                        \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        kfn<<<grid ? grid : 1, ft::NT, smem, st>>>(ptr, idx, val, X, F, Gout, lambda, nseries, queue, queue + 1); \
        break;                                                                                                  \
    }
    switch (k) {
        FT_CASE(8) FT_CASE(16) FT_CASE(20) FT_CASE(24) FT_CASE(32) FT_CASE(40) FT_CASE(48) FT_CASE(56) FT_CASE(60) FT_CASE(64)
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
