// x_pass_fast.cuh -- walks over Omega for the X-update, fp32 build, k % 4 == 0, k <= 64.
//
// Same contract as sparse_pass_kernel (x_update.cuh; reference trmf.cpp:231-288) for
// MODE_FUN / MODE_GRAD / MODE_GRADFUN / MODE_HV, rebuilt around the gather so that the
// pass is bound by L2 bandwidth (4k + 8 bytes per entry) instead of by instruction issue:
//
//  * one warp per time stamp (dynamic row queue), 32 entries per tile, two tiles in flight;
//    the observed series' factor rows go global -> shared with cp.async (16 B requests; the
//    row indices sit one per lane and are handed out by shuffle), no register staging;
//  * phase A, lane e: z_e = <s_i, h_e> with s_i held in registers and the row read as k/4
//    conflict-free LDS.128 (row stride is an odd number of 16-byte chunks);
//  * phase B, lane (entry group g, chunk c): acc[4] += z_e * h_e[4c..4c+3] for e = g, g+EG, ...
//    (one LDS.128 + one LDS.32 per 4 FMA, no shuffles); the EG groups are folded at row end;
//  * fp32 FMAs, partial sums moved to fp64 every 8 tiles (<= 88 terms each) and for the
//    objective after every tile; deterministic two-level fp64 sum for the scalar.
// ~5 warp instructions per entry against ~17 for the generic kernel.
#pragma once
#include "common.cuh"
#include "x_update.cuh"

#ifdef TRMF_F32
namespace xp {

__device__ __forceinline__ void cp16(void *smem, const void *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(s), "l"(gmem), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int WARPS = 4;

template <int K> struct Cfg {
    static constexpr int CH = K / 4;                          // 16-byte chunks per factor row
    static constexpr int RSC = (CH % 2 == 0) ? CH + 1 : CH;   // row stride in chunks: odd => lane-per-row LDS.128 is conflict free
    static constexpr int RS = RSC * 4;
    static constexpr int EG = 32 / CH;                        // entry groups in phase B
    static constexpr int RPI = 32 / CH;                       // rows per warp-wide LDGSTS
    static constexpr int NCI = (32 + RPI - 1) / RPI;          // LDGSTS instructions per 32-entry tile
    static constexpr int TILE = 32 * RS;                      // floats per stage
    static constexpr int WARP_FLOATS = 2 * TILE + 32;         // two stages + z buffer
    static constexpr size_t SMEM = sizeof(float) * WARPS * WARP_FLOATS;
};

template <int K, int MODE>
__global__ void __launch_bounds__(WARPS * 32)
pass_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ col, const float *__restrict__ val,
            const float *__restrict__ H, const float *__restrict__ S, float *__restrict__ out, uint32_t T, bool accum,
            unsigned *__restrict__ queue, double *part, unsigned *ticket, double *fout) {
    typedef Cfg<K> C;
    constexpr int CH = C::CH, RS = C::RS, EG = C::EG, RPI = C::RPI, NCI = C::NCI, TILE = C::TILE;
    constexpr bool WANT_F = MODE == MODE_FUN || MODE == MODE_GRADFUN;
    constexpr bool WANT_V = MODE != MODE_FUN;
    constexpr bool USE_Y = MODE != MODE_HV;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *wbase = reinterpret_cast<float *>(smem_raw) + (size_t)wid * C::WARP_FLOATS;
    float *zbuf = wbase + 2 * TILE;
    // copy slots: lane -> (row within an LDGSTS instruction, chunk)
    const int cp_r = lane / CH, cp_c = lane - cp_r * CH;
    const bool cp_on = lane < RPI * CH;
    const float *cp_src = H + cp_c * 4;
    const int cp_dst = cp_r * RS + cp_c * 4;
    // phase B slots: lane -> (entry group, chunk)
    const int bg = lane / CH, bc = lane - bg * CH;
    const bool b_on = lane < EG * CH;
    double fsum = 0.0;

    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(queue, 1u);
        i = __shfl_sync(FULL_MASK, i, 0);
        if (i >= T) break;
        const uint64_t lo = ptr[i];
        const uint32_t nnz = (uint32_t)(ptr[i + 1] - lo);
        const uint32_t *rcol = col + lo;
        const float *rval = val + lo;
        // s_i in registers (every lane holds the whole row)
        float4 sv[CH];
        {
            const float4 *sp = reinterpret_cast<const float4 *>(S + (size_t)i * K);
#pragma unroll
            for (int c = 0; c < CH; ++c) sv[c] = __ldg(sp + c);
        }
        float accf[4] = {0.f, 0.f, 0.f, 0.f};
        double accd[4] = {0.0, 0.0, 0.0, 0.0};
        const int ntiles = (int)((nnz + 31) / 32);

        auto issue = [&](int tt, uint32_t myidx) {   // gather tile tt (lane e holds the column index of entry e)
            float *dst = wbase + (tt & 1) * TILE + cp_dst;
            const int cnt = (int)((nnz - (uint32_t)tt * 32) < 32u ? (nnz - (uint32_t)tt * 32) : 32u);
#pragma unroll
            for (int q = 0; q < NCI; ++q) {
                const int e = RPI * q + cp_r;
                const uint32_t row = __shfl_sync(FULL_MASK, myidx, e & 31);
                cp16(dst + q * RPI * RS, cp_src + (size_t)row * K, cp_on && e < cnt);
            }
            cp_commit();
        };
        auto fetch_idx = [&](int tt) -> uint32_t {
            const uint32_t e = (uint32_t)tt * 32 + lane;
            return e < nnz ? __ldg(rcol + e) : 0u;
        };
        auto fetch_val = [&](int tt) -> float {
            const uint32_t e = (uint32_t)tt * 32 + lane;
            return (USE_Y && e < nnz) ? __ldg(rval + e) : 0.f;
        };

        // prologue: tile 0 in flight, indices/values of tile 1 in registers
        uint32_t nidx = 0;
        float ycur = 0.f, ynext = 0.f;
        if (ntiles > 0) {
            nidx = fetch_idx(0);
            ynext = fetch_val(0);
            issue(0, nidx);
            if (ntiles > 1) nidx = fetch_idx(1);
        }
        for (int t = 0; t < ntiles; ++t) {
            ycur = ynext;
            if (t + 1 < ntiles) { issue(t + 1, nidx); ynext = fetch_val(t + 1); } else cp_commit();
            if (t + 2 < ntiles) nidx = fetch_idx(t + 2);
            cp_wait<1>();          // tile t has landed (this lane's requests) ...
            __syncwarp();          // ... and every other lane's
            const float *tb = wbase + (t & 1) * TILE;
            const int cnt = (int)((nnz - (uint32_t)t * 32) < 32u ? (nnz - (uint32_t)t * 32) : 32u);
            // phase A: lane e forms <s_i, h_e>
            float z = 0.f;
            if (lane < cnt) {
                const float4 *hr = reinterpret_cast<const float4 *>(tb + lane * RS);
                float z0 = 0.f, z1 = 0.f;
#pragma unroll
                for (int c = 0; c < CH; c += 2) {
                    const float4 h = hr[c];
                    z0 = fmaf(sv[c].x, h.x, z0); z0 = fmaf(sv[c].y, h.y, z0); z0 = fmaf(sv[c].z, h.z, z0); z0 = fmaf(sv[c].w, h.w, z0);
                    if (c + 1 < CH) {
                        const float4 g = hr[c + 1];
                        z1 = fmaf(sv[c + 1].x, g.x, z1); z1 = fmaf(sv[c + 1].y, g.y, z1); z1 = fmaf(sv[c + 1].z, g.z, z1); z1 = fmaf(sv[c + 1].w, g.w, z1);
                    }
                }
                z = z0 + z1;
                if (WANT_F) { const double rr = (double)ycur - (double)z; fsum += rr * rr; }
                if (USE_Y) z -= ycur;
            }
            if (WANT_V) {
                zbuf[lane] = z;     // 0 for lanes past the tile's count
                __syncwarp();
                // phase B: lane (bg, bc): acc += z_e * h_e[4 bc .. 4 bc + 3], e = bg, bg + EG, ...
                if (b_on) {
#pragma unroll
                    for (int e = 0; e < 32; e += EG) {
                        const int ee = e + bg;
                        if (ee < cnt) {   // rows past the tile's count hold stale bytes
                            const float ze = zbuf[ee];
                            const float4 h = *reinterpret_cast<const float4 *>(tb + ee * RS + 4 * bc);
                            accf[0] = fmaf(ze, h.x, accf[0]); accf[1] = fmaf(ze, h.y, accf[1]);
                            accf[2] = fmaf(ze, h.z, accf[2]); accf[3] = fmaf(ze, h.w, accf[3]);
                        }
                    }
                }
                if ((t & 7) == 7) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { accd[q] += (double)accf[q]; accf[q] = 0.f; }
                }
            }
            __syncwarp();   // the stage (and zbuf) may be overwritten by the next iteration's copies
        }
        cp_wait<0>();
        if (WANT_V) {
#pragma unroll
            for (int q = 0; q < 4; ++q) accd[q] += (double)accf[q];
            // fold the EG entry groups: lanes bc, bc + CH, bc + 2 CH, ... hold partials of the same 4 outputs
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double v = b_on ? accd[q] : 0.0, tot = 0.0;
#pragma unroll
                for (int gI = 0; gI < EG; ++gI) tot += __shfl_sync(FULL_MASK, v, (bc + gI * CH) & 31);
                accd[q] = tot;
            }
            if (lane < CH) {
                float4 *op = reinterpret_cast<float4 *>(out + (size_t)i * K) + lane;
                float4 o = accum ? *op : make_float4(0.f, 0.f, 0.f, 0.f);
                o.x = (float)((double)o.x + accd[0]); o.y = (float)((double)o.y + accd[1]);
                o.z = (float)((double)o.z + accd[2]); o.w = (float)((double)o.w + accd[3]);
                *op = o;
            }
        }
    }
    if (WANT_F) {
        double v = block_sum(fsum, red);
        grid_sum_commit(v, part, ticket, fout, 0.5, red);
    }
}

}   // namespace xp

static inline bool x_pass_fast_supported(int k) {
    return k >= 8 && k <= 64 && k % 4 == 0;
}

template <int MODE>
static inline int x_pass_fast_launch(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *col, const V *val,
                                     const V *H, const V *S, V *out, int k, uint32_t T, bool accum, unsigned *queue,
                                     double *part, unsigned *ticket, double *fout) {
    if (cudaMemsetAsync(queue, 0, sizeof(unsigned), st) != cudaSuccess) return 1;
    const unsigned rows_ctas = (T + xp::WARPS - 1) / xp::WARPS;
#define XP_CASE(KK)                                                                                          \
    case KK: {                                                                                               \
        auto kfn = xp::pass_kernel<KK, MODE>;                                                                \
        const size_t smem = xp::Cfg<KK>::SMEM;                                                               \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
        int per_sm = 0;                                                                                      \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, xp::WARPS * 32, smem) != cudaSuccess) return 1; \
        unsigned grid = (unsigned)(per_sm > 0 ? per_sm : 1) * (unsigned)num_sms;                             \
        if (grid > rows_ctas) grid = rows_ctas ? rows_ctas : 1;                                              \
        if (grid > 4096) grid = 4096;   /* `part` holds 8192 partials */                                     \
        kfn<<<grid, xp::WARPS * 32, smem, st>>>(ptr, col, val, H, S, out, T, accum, queue, part, ticket, fout); \
        break;                                                                                               \
    }
    switch (k) {
        XP_CASE(8) XP_CASE(12) XP_CASE(16) XP_CASE(20) XP_CASE(24) XP_CASE(28) XP_CASE(32) XP_CASE(36) XP_CASE(40) XP_CASE(44) XP_CASE(48) XP_CASE(52)
        XP_CASE(56) XP_CASE(60) XP_CASE(64)
        default: return 1;
    }
#undef XP_CASE
    return cudaGetLastError() != cudaSuccess;
}
#else
static inline bool x_pass_fast_supported(int) { return false; }
template <int MODE>
static inline int x_pass_fast_launch(cudaStream_t, int, const uint64_t *, const uint32_t *, const V *, const V *, const V *, V *,
                                     int, uint32_t, bool, unsigned *, double *, unsigned *, double *) { return 1; }
#endif
