// f_update_tc.cuh -- K1 (and the X-update's Gram build) on the 5th-generation tensor core: per-series Gram + right-hand
// side by tcgen05.mma with the accumulator in TMEM (fp32 storage build, k % 4 == 0, k <= 64).
//
// Replaces the hot loop of l2r_ls_pY_IX_chol::solve (reference trmf.cpp:382-395) and, over the by-time CSR, the Omega walks
// of arr_ls_pY_IX::fun / ::grad / ::Hv (trmf.cpp:231-288) exactly like f_update_mma.cuh does -- same inputs, same outputs,
// same modes -- but the k x k SYRK  G = sum_e x_e x_e^T  runs on tcgen05 instead of warp-level mma.sync:
//
//   * split  x = h1 + h2,  h1 = fp16(x), h2 = fp16(x - h1)  (22 significant bits; the factor is column-scaled by exact
//     powers of two first, colscale_*_kernel of f_update_mma.cuh);
//   * ONE operand tile per 16 entries serves as A and B: A = [h1 ; h2] STACKED IN M (rows 0..k-1 and 64..64+k-1 of a
//     128 x 16 operand, one entry per contraction step), B = the h1 part (N = 16 ceil(k/16)).  D_top = h1^T h1 and
//     D_bot = h2^T h1 =: S come out of one M128 instruction and G = D_top + S + S^T is assembled once per series: the
//     third split product h1^T h2 is S^T, and h2^T h2 (2^-22 relative) is dropped as in the mma.sync kernel;
//   * the tile's weights (Y values, or the residuals z = <w, x> - y in MODE_GRAD), scaled by a power of two and split the
//     same way, ride in two spare B columns (k, k+1 when 16 does not divide k, else a 16-column side tile and a second
//     N = 16 MMA): rhs_c = D[c][y1] + D[c][y2] + D[64+c][y1] is a by-product of the same instruction;
//   * both operands are MN-major, unswizzled core matrices (8 entries x 16 bytes): for one entry, 8 consecutive factor
//     columns are one 16-byte row -- exactly how a gathered factor row arrives, so the converter writes 8-byte halves of
//     such rows with no transpose (descriptor strides LBO = 128 B, SBO = 256 B, verified by tools/microbench_tcgen05_gram.cu).
//
// Warp-specialised, one CTA of 13/17 warps per SM, persistent over a static round-robin of the series:
//   converter warps (8)  gather: LDG.128 of the fp32 rows straight into registers (two tiles in flight per warp), split,
//                        st.shared.v2 into the operand tile, fence.proxy.async, mbarrier arrive.  (TMA was measured and
//                        rejected for this gather: one request per row / per 4 rows runs at 46-130 clk per row per SM,
//                        profiles/r02_microbench_gather.txt; the register path needs ~5.)  In MODE_GRAD the same
//                        registers give z = <w, x> - y per entry (fp32 FMAs) and sum z^2 (fp64).
//   MMA warp (1 thread)  per chunk of 8 tiles (128 entries of one series): waits for the chunk's operand slot and a free
//                        TMEM accumulator, issues the 8 (or 16) tcgen05.mma, commits to the slot's and the accumulator's
//                        mbarriers.
//   drain warps (4 / 8)  tcgen05.ld the finished accumulator (thread = accumulator row), release it, add into fp64
//                        registers; at the end of a series assemble G and the rhs in shared memory, undo the scaling and
//                        write the fp64 system (MODE_DEFER, solved by fm::chol_solve_kernel) or the fp32 Gram + gradient
//                        row (MODE_STORE / MODE_GRAD).
// The tensor core adds with truncation; 128 entries per TMEM run measured 4.7e-7 relative on the Gram
// (profiles/r02_microbench_tcgen05_gram.txt) and the truncation acts on the right-hand side alike.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

#ifdef TRMF_F32

namespace tc {

constexpr int ET = 16;                 // entries per tile = one MMA contraction step
constexpr int CT = 8;                  // tiles per chunk = one accumulation run in TMEM (128 entries)
constexpr int NCW = 4;                 // producer warps (CT / NCW tiles of a chunk each)
constexpr int DEPTH = 2;               // chunks a producer warp keeps in flight (cp.async groups) before it publishes the oldest
// Operand tile of 16 entries: MN-major, SWIZZLE_128B.  One 128-byte line per entry and half (64 fp16 columns; the 16-byte chunk c
// of entry e sits at chunk c ^ (e & 7) of its line), the 16 h1 lines followed by the 16 h2 lines.  A gathered row half therefore
// lands in ONE line: a quarter warp of LDGSTS writes one shared-memory wavefront.  (The unswizzled core-matrix layout puts the
// chunks of a row 256 bytes apart -- every 16-byte piece its own wavefront, 32 per LDGSTS: measured 10 clk per entry in the LSU
// alone, profiles/r02_tc_k40_v2_ncu.txt.)
constexpr int LINE = 128;
constexpr int HALF_BYTES = ET * LINE;      // 2048: the h1 (or h2) lines of a tile = one 64-column MN block of the descriptor (LBO)
constexpr int TILE_BYTES = 2 * HALF_BYTES;
constexpr int YT_TILE = 512;               // side tile of the weights when k fills its MN block: unswizzled, 2 groups x 256 B

template <int K> struct Cfg {
    static_assert(K % 4 == 0 && K >= 4 && K <= 64, "rank");
    static constexpr int NG = (K + 7) / 8;                      // 16-byte chunks (8 columns) per half of a pre-split row
    static constexpr int ROWB = 32 * NG;                        // bytes of a pre-split row: [h1 | h2], each padded to 8 NG halfs
    static constexpr int NB = ((K + 15) / 16) * 16;             // N of the Gram MMA
    static constexpr bool YIN = (K % 8 == 0) && (NB - K) >= 8;  // weights ride in columns K, K+1 of the h1 part (a group no copy touches)
    static constexpr int YC = YIN ? K : NB;                     // accumulator column of the y1 products (y2: YC + 1)
    static constexpr int DCOLS = YIN ? NB : NB + 16;            // accumulator columns
    static constexpr int ACC_STRIDE = DCOLS <= 32 ? 32 : (DCOLS <= 64 ? 64 : 128);
    static constexpr int NACC = 512 / ACC_STRIDE;               // TMEM accumulator ring
    static constexpr int YT_BYTES = YIN ? 0 : YT_TILE;          // side tile of the weights (two unswizzled groups)
    static constexpr int SLOT_BYTES = CT * (TILE_BYTES + YT_BYTES);
    static constexpr int NSLOT = K <= 48 ? 6 : 5;               // operand ring, in chunks (shared memory: 192 / 180 KB)
    static constexpr int IL = NSLOT - DEPTH - 1;                // chunks whose MMAs are interleaved (their slots are all held at once)
    static constexpr int NDG = DCOLS > 48 ? 2 : 1;              // drain warpgroups (each owns DCOLS / NDG columns)
    static constexpr int DPER = DCOLS / NDG;                    // columns per drain thread
    static_assert(DPER % 8 == 0, "drain split");
    static constexpr int NRW = K > 48 ? 4 : 8;                  // MODE_GRAD: residual warps (CT / NRW tiles of a chunk each)
    // Warps: 4 NDG drain | NCW producers | NRW residual (MODE_GRAD) | 4 in the MMA group (one issues, three idle so that the
    // count is a multiple of 4: ptxas sizes the launch allocation for whole groups of four warps).  MODE_GRAD redistributes
    // registers with setmaxnreg -- the pool is the CTA's launch allocation nwarps x 32 x launch_regs, launch_regs being ptxas'
    // cap for this block size (checked on the host against cudaFuncGetAttributes before the first launch).
    __host__ __device__ static constexpr int nwarps(bool grad) { return 4 * NDG + NCW + (grad ? NRW : 0) + 4; }
    __host__ __device__ static constexpr int launch_regs(bool grad) { int r = 65536 / (32 * nwarps(grad)); r = r > 255 ? 255 : r; return r & ~7; }
    static constexpr int MMA_REGS = 48;
    __host__ __device__ static constexpr int prod_regs(bool grad) { return launch_regs(grad) >= 96 ? 88 : 72; }
    __host__ __device__ static constexpr int res_regs() { return launch_regs(true) >= 88 ? 80 : 72; }
    __host__ __device__ static constexpr int drain_regs(bool grad) {     // what is left of the pool for the drain warps
        int x = (32 * nwarps(grad) * launch_regs(grad) - 32 * (NCW * prod_regs(grad) + (grad ? NRW * res_regs() : 0) + 4 * MMA_REGS)) / (32 * 4 * NDG);
        x = x > 232 ? 232 : x;
        return x & ~7;
    }
    static constexpr int ld = K + 1;
    static constexpr size_t g_bytes = sizeof(double) * (size_t)(K + 1) * (K + 1);
    static constexpr size_t smem = 1024 + (size_t)NSLOT * SLOT_BYTES + g_bytes + sizeof(double) * 16 * 8;
    __host__ __device__ static constexpr uint32_t idesc(int n) { return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(r[q]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// MN-major SWIZZLE_128B: LBO = bytes between 64-column MN blocks (h1 lines -> h2 lines), SBO = bytes between groups of 8 entries
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(HALF_BYTES >> 4) << 16) | ((uint64_t)(8 * LINE >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// the unswizzled side tile of the weights: LBO (between the two 8-entry halves) = 128 B, SBO (between 8-column groups) = 256 B
__device__ __forceinline__ uint64_t make_desc_y(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t pack_h2(const float a, const float b) {   // .lo = a, .hi = b, round to nearest
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(const uint32_t p) {
    float2 f;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(f.x), "=f"(f.y) : "r"(p));
    return f;
}
// 16 bytes global -> shared without a register stop (LDGSTS); nbytes = 0 writes zeros
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// Pre-split copy of the gathered factor: row i = [h1 (8 NG halfs) | h2 (8 NG halfs)], h1 = fp16(x s_c), h2 = fp16(x s_c - h1), with
// the exact per-column power-of-two scales s_c of fm::colscale_max_kernel (column maximum into [2^14, 2^15)); columns past k are 0.
// Same bytes per row as the fp32 factor when 8 divides k.  invs[c] = 1 / s_c.
__global__ void presplit_kernel(const float *__restrict__ X, size_t rows, int k, int ng, const unsigned *__restrict__ colmax,
                                __half *__restrict__ Xh, float *__restrict__ invs) {
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
    const int kp = 8 * ng;
    const size_t total = rows * (size_t)kp;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = p / kp;
        const int c = (int)(p - i * kp);
        __half h1 = __float2half_rn(0.f), h2 = h1;
        if (c < k) {
            const float x = X[i * k + c] * sc[c];
            h1 = __float2half_rn(x);
            h2 = __float2half_rn(x - __half2float(h1));
        }
        Xh[i * (size_t)(2 * kp) + c] = h1;
        Xh[i * (size_t)(2 * kp) + kp + c] = h2;
    }
}

// a CTA's walk over its series, chunk by chunk (every role runs the same walk)
struct Walk {
    const uint64_t *ptr;
    uint32_t nseries, j, nnz, q;     // q = running chunk number of this CTA
    uint64_t lo;
    int c, nch, ntiles;
    bool ok;
    __device__ __forceinline__ void seek() {   // first series at or after j with entries
        ok = false;
        while (j < nseries) {
            lo = ptr[j];
            nnz = (uint32_t)(ptr[j + 1] - lo);
            if (nnz != 0) { ntiles = (int)((nnz + ET - 1) / ET); nch = (ntiles + CT - 1) / CT; c = 0; ok = true; return; }
            j += gridDim.x;
        }
    }
    __device__ __forceinline__ void init(const uint64_t *p, uint32_t n) { ptr = p; nseries = n; j = blockIdx.x; q = 0; seek(); }
    __device__ __forceinline__ void next() {
        ++q;
        if (++c == nch) { j += gridDim.x; seek(); }
    }
};

enum { MODE_SOLVE = 0, MODE_STORE = 1, MODE_GRAD = 2, MODE_DEFER = 3 };   // = fm::MODE_*

#ifdef TC_DEBUG
__device__ volatile unsigned *g_tc_dbg = nullptr;    // host-mapped progress words (tools/test_f_update_tc.cu): [warp][4]
#define TC_DBG(slot, v) do { if (g_tc_dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { g_tc_dbg[(threadIdx.x >> 5) * 4 + (slot)] = (v); __threadfence_system(); } } while (0)
__device__ unsigned long long g_tc_clk[32 * 8];     // CTA 0: per warp, cycles spent in [0..5] the waits below, [7] the whole role
#define TC_T0() const long long tc_t0_ = clock64()
#define TC_T1(slot) do { tc_acc_[slot] += clock64() - tc_t0_; } while (0)
#define TC_TDECL() long long tc_acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long tc_role0_ = clock64()
#define TC_TFLUSH() do { tc_acc_[7] = clock64() - tc_role0_; if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) for (int q_ = 0; q_ < 8; ++q_) g_tc_clk[(threadIdx.x >> 5) * 8 + q_] = (unsigned long long)tc_acc_[q_]; } while (0)
#else
#define TC_DBG(slot, v) do { } while (0)
#define TC_T0() do { } while (0)
#define TC_T1(slot) do { } while (0)
#define TC_TDECL() do { } while (0)
#define TC_TFLUSH() do { } while (0)
#endif

// Xh = the pre-split factor (presplit_kernel); ysc[0] = power-of-two scale of the weights (Y, or the residual bound in MODE_GRAD),
// ysc[1] = its inverse
template <int K, int MODE>
__global__ void __launch_bounds__(32 * Cfg<K>::nwarps(MODE == MODE_GRAD), 1)
f_update_tc_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                   const unsigned char *__restrict__ Xh, const float *__restrict__ invs, float *__restrict__ F, float *__restrict__ Gout,
                   uint32_t nseries, const float *__restrict__ Wv, int gaccum, double *__restrict__ frow, const float *__restrict__ ysc) {
    typedef Cfg<K> C;
    constexpr bool GRAD = MODE == MODE_GRAD, DEFER = MODE == MODE_DEFER;
    constexpr int NG = C::NG, NB = C::NB, NACC = C::NACC, ld = C::ld, NDG = C::NDG, DPER = C::DPER;
    constexpr int NTH = 32 * C::nwarps(GRAD);
    constexpr int NSLOT = C::NSLOT;
    constexpr int NRW = C::NRW, W_PROD = 4 * NDG, W_RES = W_PROD + NCW, W_MMA = W_RES + (GRAD ? NRW : 0);
    extern __shared__ unsigned char smem_raw[];
    unsigned char *slots = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    double *Gs = reinterpret_cast<double *>(slots + (size_t)NSLOT * C::SLOT_BYTES);   // (K+1) x ld: Gram rows, row K = rhs
    double *fpart = Gs + (size_t)(K + 1) * ld;                                       // [16][NRW] MODE_GRAD: sum z^2 per residual warp
    __shared__ __align__(8) unsigned long long bar_full[NSLOT], bar_wfull[NSLOT], bar_empty[NSLOT], bar_afull[16], bar_aempty[16];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t slots_s = smem_u32(slots);

    for (int p = tid; p < NSLOT * C::SLOT_BYTES / 16; p += NTH) reinterpret_cast<uint4 *>(slots)[p] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(smem_u32(&bar_full[s]), NCW); mbar_init(smem_u32(&bar_wfull[s]), NRW); mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < NACC; ++b) { mbar_init(smem_u32(&bar_afull[b]), 1); mbar_init(smem_u32(&bar_aempty[b]), 4 * NDG); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < W_PROD) {
        // =============================== drain warps ===============================
        if constexpr (C::drain_regs(GRAD) > C::launch_regs(GRAD)) reg_inc<C::drain_regs(GRAD)>();
        const int dg = warp >> 2;                        // drain group: columns [dg * DPER, dg * DPER + DPER)
        const int r = (warp & 3) * 32 + lane;            // accumulator row = TMEM lane
        const bool top = r < K, bot = r >= 64 && r < 64 + K;
        const int rr = r - 64;
        const int dtid = tid;                            // 0 .. 128 NDG - 1 within the drain warps
        constexpr int NDT = 128 * NDG;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(dg * DPER);
        double accr[DPER];          // top rows: fp64 totals of the series
        float accf[DPER];           // fp32 partial sums (top: since the last flush; bot: of the whole series)
        constexpr int FLUSH_CH = 8;
        int nflush = 0;
        bool flushed = false;
        Walk w;
        w.init(ptr, nseries);
        TC_TDECL();
        // series without entries that precede / sit between this CTA's non-empty ones: the store modes zero their outputs
        uint32_t jz = blockIdx.x;
        auto zero_until = [&](uint32_t jend) {
            if (!DEFER) {
                for (; jz < jend && jz < nseries; jz += gridDim.x) {
                    if (ptr[jz + 1] != ptr[jz]) continue;
                    float *Gj = Gout + (size_t)jz * K * K;
                    for (int p = dtid; p < K * K; p += NDT) Gj[p] = 0.f;
                    if (GRAD) { if (dtid == 0) frow[jz] = 0.0; if (!gaccum && dtid < K) F[(size_t)jz * K + dtid] = 0.f; }
                    else if (dtid < K) F[(size_t)jz * K + dtid] = 0.f;
                }
            }
        };
        while (w.ok) {
            const uint32_t b = w.q % NACC;
            { TC_T0(); mbar_wait(smem_u32(&bar_afull[b]), (w.q / NACC) & 1); TC_T1(0); }
            tc_fence_after();
            float v[DPER];
#pragma unroll
            for (int c8 = 0; c8 < DPER / 8; ++c8) tmem_ld8(lane_addr + b * C::ACC_STRIDE + c8 * 8, *reinterpret_cast<float(*)[8]>(&v[c8 * 8]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_aempty[b]));
            // fp32 first: the chunk sums of up to FLUSH_CH chunks (1024 entries) are added by FADD; the h1^T h1 rows and the rhs
            // (top) then go into fp64 registers.  (F2F.F64.F32 + DADD per value and chunk kept the drain warps busy 1250 clk per
            // chunk -- the bottleneck of the first versions.)  The h2^T h1 rows (bot) are 2^-11 of the Gram: fp32 all the way.
            if (top) {
                if (nflush == 0) {
#pragma unroll
                    for (int c = 0; c < DPER; ++c) accf[c] = v[c];
                } else {
#pragma unroll
                    for (int c = 0; c < DPER; ++c) accf[c] += v[c];
                }
                if (++nflush == FLUSH_CH || w.c == w.nch - 1) {
                    if (!flushed) {
#pragma unroll
                        for (int c = 0; c < DPER; ++c) accr[c] = (double)accf[c];
                    } else {
#pragma unroll
                        for (int c = 0; c < DPER; ++c) accr[c] += (double)accf[c];
                    }
                    flushed = true;
                    nflush = 0;
                }
            } else if (bot) {
                if (w.c == 0) {
#pragma unroll
                    for (int c = 0; c < DPER; ++c) accf[c] = v[c];
                } else {
#pragma unroll
                    for (int c = 0; c < DPER; ++c) accf[c] += v[c];
                }
            }
            if (w.c == w.nch - 1) {
                TC_T0();
                // ---- series epilogue: G = D_top + S + S^T, rhs = D[c][y1] + D[c][y2] + D[64+c][y1] ----
                const uint32_t j = w.j;
                const int c0 = dg * DPER;
                if (top) {
#pragma unroll
                    for (int c = 0; c < DPER; ++c) if (c0 + c < K) Gs[r * ld + c0 + c] = accr[c];
                }
                constexpr int YG = C::YC / DPER, YI = C::YC % DPER;     // drain group and local column of the y1 products
                if (top && dg == YG) Gs[K * ld + r] = accr[YI] + accr[YI + 1];
                named_bar(1, NDT);
                if (bot) {
#pragma unroll
                    for (int c = 0; c < DPER; ++c) if (c0 + c < K) Gs[rr * ld + c0 + c] += (double)accf[c];
                    if (dg == YG) Gs[K * ld + rr] += (double)accf[YI];
                }
                named_bar(1, NDT);
                if (bot) {
#pragma unroll
                    for (int c = 0; c < DPER; ++c) if (c0 + c < K) Gs[(c0 + c) * ld + rr] += (double)accf[c];
                }
                named_bar(1, NDT);
                const float yinv = ysc[1];
                if (DEFER) {    // frow carries the scratch: one (K+1) x ld fp64 system per series (lower triangle + rhs row)
                    double *dst = frow + (size_t)j * ((K + 1) * ld);
                    for (int p = dtid; p < (K + 1) * ld; p += NDT) {
                        const int row = p / ld, col = p - row * ld;
                        double o = 0.0;
                        if (col < K) o = row < K ? Gs[p] * ((double)invs[row] * (double)invs[col]) : Gs[p] * ((double)invs[col] * (double)yinv);
                        dst[p] = o;
                    }
                } else {
                    float *Gj = Gout + (size_t)j * K * K;
                    for (int p = dtid; p < K * K; p += NDT) {
                        const int row = p / K, col = p - row * K;
                        Gj[p] = (float)(Gs[row * ld + col] * ((double)invs[row] * (double)invs[col]));
                    }
                    if (dtid < K) {
                        const double rhs = Gs[K * ld + dtid] * ((double)invs[dtid] * (double)yinv);
                        float *o = F + (size_t)j * K + dtid;
                        if (GRAD) *o = gaccum ? (float)((double)*o + rhs) : (float)rhs;
                        else *o = (float)rhs;
                    }
                    if (GRAD && dtid == 0) {
                        __threadfence_block();
                        const double *fp = fpart + (size_t)(w.q & 15) * NRW;
                        double fs = fp[0];
                        for (int cw = 1; cw < NRW; ++cw) fs += fp[cw];
                        frow[j] = fs;
                    }
                }
                named_bar(1, NDT);     // Gs is free again
                zero_until(j);
                if (jz == j) jz += gridDim.x;
                flushed = false;
                TC_T1(1);
            }
            w.next();
        }
        zero_until(nseries);
        TC_TFLUSH();
    } else if (warp < W_RES) {
        // =============================== producer warps ===============================
        // Tile cw of every chunk: NG warp-wide LDGSTS lay the 16 pre-split rows straight into the operand tile (lane l, step i:
        // piece id = l + 32 i of the tile, entry id / (2 NG), 16-byte chunk id % (2 NG) of the row -- a quarter warp reads 128
        // contiguous bytes and writes 8 different bank groups); the Y values follow by st.shared.  DEPTH tiles stay in flight.
        if constexpr (C::prod_regs(GRAD) < C::launch_regs(GRAD)) reg_dec<C::prod_regs(GRAD)>();
        const int cw = warp - W_PROD;
        const float yscale = ysc[0];
        Walk w0, w1, w2;            // the chunk being issued, the next, the one after
        w0.init(ptr, nseries);
        w1 = w0; if (w1.ok) w1.next();
        w2 = w1; if (w2.ok) w2.next();
        constexpr int TPP = CT / NCW;                    // tiles of a chunk per producer warp: cw, cw + NCW, ..
        // LDGSTS step i (of 8) of a tile: quarter warp q = lane / 8 copies half (4 i + q) % 2 of entry (4 i + q) / 2; lane % 8 = the
        // 16-byte chunk (lanes past the row's NG chunks stay idle: those columns were zeroed once and are never written)
        const int pc = lane & 7;
        const bool pact = pc < NG;
        int pe[8];
        uint32_t psrc[8], pdst[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rh = 4 * i + (lane >> 3), e = rh >> 1, h = rh & 1;
            pe[i] = e;
            psrc[i] = (uint32_t)((h * NG + pc) * 16);
            pdst[i] = (uint32_t)(h * HALF_BYTES + e * LINE + ((pc ^ (e & 7)) << 4));
        }
        auto ld_row = [&](const Walk &w, int u) -> uint32_t {   // row index of entry (lane & 15) of tile cw + u NCW (clamped)
            if (!w.ok) return 0u;
            uint32_t pos = (uint32_t)(w.c * CT + cw + u * NCW) * ET + (uint32_t)(lane & 15);
            pos = pos < w.nnz ? pos : w.nnz - 1;
            return __ldg(idx + w.lo + pos);
        };
        auto ld_yv = [&](const Walk &w, int u) -> float {        // lanes 0..15: the Y value of entry `lane` of tile cw + u NCW
            if (GRAD || !w.ok || lane >= ET) return 0.f;
            const uint32_t pos = (uint32_t)(w.c * CT + cw + u * NCW) * ET + (uint32_t)lane;
            return pos < w.nnz ? __ldg(val + w.lo + pos) : 0.f;
        };
        uint32_t row0[TPP], row1[TPP], row2[TPP];
        float yv0[TPP], yv1[TPP];
#pragma unroll
        for (int u = 0; u < TPP; ++u) { row0[u] = ld_row(w0, u); row1[u] = ld_row(w1, u); yv0[u] = ld_yv(w0, u); }
        TC_TDECL();
        uint32_t pend[DEPTH] = {};    // slots of the chunks in flight, oldest first (constant indices only: registers)
        int npend = 0;
        auto retire = [&]() {       // this warp's tiles of the oldest chunk in flight have landed: publish them
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_full[pend[0]]));
#pragma unroll
            for (int d = 0; d + 1 < DEPTH; ++d) pend[d] = pend[d + 1];
            --npend;
        };
        while (w0.ok) {
#pragma unroll
            for (int u = 0; u < TPP; ++u) { row2[u] = ld_row(w2, u); yv1[u] = ld_yv(w1, u); }
            const uint32_t s = w0.q % NSLOT;
            { TC_T0(); mbar_wait(smem_u32(&bar_empty[s]), ((w0.q / NSLOT) & 1) ^ 1); TC_T1(0); }
            TC_T0();
#pragma unroll
            for (int u = 0; u < TPP; ++u) {
                const int tl = cw + u * NCW, t = w0.c * CT + tl;
                if (t < w0.ntiles) {
                    const uint32_t tile = slots_s + s * C::SLOT_BYTES + tl * TILE_BYTES;
                    const uint32_t cnt = min((uint32_t)ET, w0.nnz - (uint32_t)t * ET);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t row = __shfl_sync(FULL_MASK, row0[u], pe[i]);
                        if (pact) cp_async16(tile + pdst[i], Xh + ((size_t)row * C::ROWB + psrc[i]), (uint32_t)pe[i] < cnt ? 16u : 0u);
                    }
                    if (!GRAD && lane < ET) {   // the entry's weight, scaled and split, in columns YC, YC + 1 of the B operand
                        const float ys = yv0[u] * yscale;     // (entries past the end were loaded as 0)
                        const uint32_t py = pack_h2(ys, 0.f);
                        const float2 fy = unpack_h2(py);
                        const uint32_t pw = (py & 0xffffu) | (pack_h2(ys - fy.x, 0.f) << 16);
                        sts32(C::YIN ? tile + lane * LINE + ((((K / 8) ^ (lane & 7))) << 4) + (K % 8) * 2
                                     : slots_s + s * C::SLOT_BYTES + CT * TILE_BYTES + tl * YT_TILE + (lane >> 3) * 128 + (lane & 7) * 16, pw);
                    }
                }
            }
            cp_async_commit();
            TC_T1(2);
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) if (d == npend) pend[d] = s;
            ++npend;
            if (npend == DEPTH) { TC_T0(); cp_async_wait<DEPTH - 1>(); TC_T1(1); retire(); }
            w0 = w1; w1 = w2; if (w2.ok) w2.next();
#pragma unroll
            for (int u = 0; u < TPP; ++u) { row0[u] = row1[u]; row1[u] = row2[u]; yv0[u] = yv1[u]; }
        }
        cp_async_wait<0>();
        while (npend > 0) retire();
        TC_TFLUSH();
    } else if (GRAD && warp < W_MMA) {
        // =============================== residual warps (MODE_GRAD) ===============================
        // Tile rw of every landed chunk: z = <w, x> - y per entry from the operand tile itself (x = h1 + h2, fp32 FMAs), sum z^2 in
        // fp64, and z -- scaled and split like Y -- into the weight columns; then the chunk goes to the MMA warp.
        if constexpr (C::res_regs() < C::launch_regs(true)) reg_dec<C::res_regs()>();
        const int rw = warp - W_RES;
        const int e = lane & 15, par = lane >> 4;        // entry; parity of the 8-column groups this lane sums
        const uint32_t eline = (uint32_t)e * LINE, esw = (uint32_t)(e & 7);
        const float yscale = ysc[0];
        constexpr int NGL = (NG + 1) / 2;
        float wl[NGL][8];
        double fsum = 0.0;
        uint32_t wl_j = 0xffffffffu;
        Walk w, wn;
        w.init(ptr, nseries);
        wn = w; if (wn.ok) wn.next();
        constexpr int TPW = CT / NRW;                    // tiles of a chunk per residual warp: rw, rw + NRW, ..
        auto ld_y = [&](const Walk &x, int u) -> float {
            if (!x.ok) return 0.f;
            const uint32_t pos = (uint32_t)(x.c * CT + rw + u * NRW) * ET + (uint32_t)e;
            return pos < x.nnz ? __ldg(val + x.lo + pos) : 0.f;
        };
        TC_TDECL();
        float y0[TPW], y1[TPW];
#pragma unroll
        for (int u = 0; u < TPW; ++u) y0[u] = ld_y(w, u);
        while (w.ok) {
#pragma unroll
            for (int u = 0; u < TPW; ++u) y1[u] = ld_y(wn, u);
            const uint32_t s = w.q % NSLOT;
            if (wl_j != w.j) {
                wl_j = w.j;
#pragma unroll
                for (int i = 0; i < NGL; ++i)
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int c = 8 * (par + 2 * i) + q;
                        wl[i][q] = c < K ? __ldg(Wv + (size_t)w.j * K + c) * __ldg(invs + c) : 0.f;
                    }
            }
            { TC_T0(); mbar_wait(smem_u32(&bar_full[s]), (w.q / NSLOT) & 1); TC_T1(0); }
            TC_T0();
#pragma unroll
            for (int u = 0; u < TPW; ++u) {
                const int tl = rw + u * NRW, t = w.c * CT + tl;
                if (t < w.ntiles) {
                    const uint32_t tile = slots_s + s * C::SLOT_BYTES + tl * TILE_BYTES;
                    float z = 0.f, zq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < NGL; ++i) {
                        const int g = par + 2 * i;
                        if (g < NG) {
                            const uint32_t ca = tile + eline + (((uint32_t)g ^ esw) << 4);
                            const uint4 a = lds128(ca), b2 = lds128(ca + HALF_BYTES);
                            const uint32_t ha[4] = {a.x, a.y, a.z, a.w}, hb[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 f1 = unpack_h2(ha[q]), f2 = unpack_h2(hb[q]);
                                zq[q] = fmaf(wl[i][2 * q], f1.x + f2.x, zq[q]);
                                zq[q] = fmaf(wl[i][2 * q + 1], f1.y + f2.y, zq[q]);
                            }
                        }
                    }
                    z = (zq[0] + zq[1]) + (zq[2] + zq[3]);
                    z += __shfl_xor_sync(FULL_MASK, z, 16);          // the partner lane holds the other groups of the same entry
                    const uint32_t cnt = min((uint32_t)ET, w.nnz - (uint32_t)t * ET);
                    const bool live = (uint32_t)e < cnt;
                    if (par == 0) {
                        if (live) { const double rr = (double)y0[u] - (double)z; fsum += rr * rr; }
                        const float ys = live ? (z - y0[u]) * yscale : 0.f;
                        const uint32_t py = pack_h2(ys, 0.f);
                        const float2 fy = unpack_h2(py);
                        const uint32_t pw = (py & 0xffffu) | (pack_h2(ys - fy.x, 0.f) << 16);
                        sts32(C::YIN ? tile + eline + ((((uint32_t)(K / 8)) ^ esw) << 4) + (K % 8) * 2
                                     : slots_s + s * C::SLOT_BYTES + CT * TILE_BYTES + tl * YT_TILE + (e >> 3) * 128 + (e & 7) * 16, pw);
                    }
                }
            }
            fence_proxy_async();
            TC_T1(1);
            if (w.c == w.nch - 1) {      // series finished: publish this warp's sum of squared residuals
                const double fs = warp_sum(fsum);
                if (lane == 0) fpart[(size_t)(w.q & 15) * NRW + rw] = fs;
                fsum = 0.0;
                __threadfence_block();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_wfull[s]));
            w = wn; if (wn.ok) wn.next();
#pragma unroll
            for (int u = 0; u < TPW; ++u) y0[u] = y1[u];
        }
        TC_TFLUSH();
    } else {
        // =============================== MMA warp (+ three idle warps of its group) ===============================
        if constexpr (C::MMA_REGS < C::launch_regs(GRAD)) reg_dec<C::MMA_REGS>();
        if (warp == W_MMA) {
        // Chunks are issued IL at a time, their tiles interleaved: MMAs into one accumulator run back to back at ~120 clk
        // each (measured; a fixed cost per instruction, far above the 24 clk of math of an M128 N48 K16 step), MMAs into
        // different accumulators overlap.
        constexpr int IL = C::IL;
        Walk w;
        w.init(ptr, nseries);
        TC_TDECL();
        while (w.ok) {
            Walk wq[IL];
            int nq = 0;
#pragma unroll
            for (int u = 0; u < IL; ++u) {
                if (w.ok) { wq[u] = w; ++nq; w.next(); }
            }
#pragma unroll
            for (int u = 0; u < IL; ++u) {
                if (u < nq) {
                    const uint32_t s = wq[u].q % NSLOT, b = wq[u].q % NACC;
                    { TC_T0(); mbar_wait(smem_u32(GRAD ? &bar_wfull[s] : &bar_full[s]), (wq[u].q / NSLOT) & 1); TC_T1(0); }
                    { TC_T0(); mbar_wait(smem_u32(&bar_aempty[b]), ((wq[u].q / NACC) & 1) ^ 1); TC_T1(1); }
                }
            }
            TC_T0();
            tc_fence_after();
            if (lane == 0) {
                for (int t = 0; t < CT; ++t) {
#pragma unroll
                    for (int u = 0; u < IL; ++u) {
                        if (u < nq && t < min(CT, wq[u].ntiles - wq[u].c * CT)) {
                            const uint32_t sbase = slots_s + (wq[u].q % NSLOT) * C::SLOT_BYTES;
                            const uint32_t dacc = tmem_base + (wq[u].q % NACC) * C::ACC_STRIDE;
                            const uint64_t da = make_desc(sbase + t * TILE_BYTES);
#ifdef TC_EXP_M      // timing experiment only (results are garbage): another instruction shape on the same tile
                            umma_f16(dacc, da, da, (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(TC_EXP_N >> 3) << 17) | ((uint32_t)(TC_EXP_M >> 4) << 24), t > 0 ? 1u : 0u);
#else
                            umma_f16(dacc, da, da, C::idesc(NB), t > 0 ? 1u : 0u);
#endif
                            if (!C::YIN) umma_f16(dacc + NB, da, make_desc_y(sbase + CT * TILE_BYTES + t * YT_TILE), C::idesc(16), t > 0 ? 1u : 0u);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < IL; ++u) {
                    if (u < nq) {
                        umma_commit(smem_u32(&bar_empty[wq[u].q % NSLOT]));
                        umma_commit(smem_u32(&bar_afull[wq[u].q % NACC]));
                    }
                }
            }
            __syncwarp();
            TC_T1(2);
        }
        TC_TFLUSH();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// max |v[lo .. hi)| as float bits (atomicMax on the bits of non-negative floats is order-free); lo = ptr[0], hi = ptr[n]
__global__ void absmax_range_kernel(const float *__restrict__ v, const uint64_t *__restrict__ ptr, uint32_t n, unsigned *__restrict__ out) {
    const uint64_t lo = ptr[0], hi = ptr[n];
    float m = 0.f;
    for (uint64_t p = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < hi; p += (uint64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(v[p]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
// ysc[0] = 2^s with bound * 2^s in [2^13, 2^14), ysc[1] = 2^-s.  bound = max |y|, plus (MODE_GRAD) sum_c max|w_c| max|x_c| so
// that no residual <w, x> - y can overflow fp16 after scaling; a zero / non-finite bound keeps scale 1.
__global__ void weight_scale_kernel(const unsigned *__restrict__ ymax, const unsigned *__restrict__ xcolmax, const unsigned *__restrict__ wcolmax,
                                    int k, float *__restrict__ ysc) {
    float bound = __uint_as_float(*ymax);
    if (wcolmax != nullptr)
        for (int c = 0; c < k; ++c) bound += __uint_as_float(wcolmax[c]) * __uint_as_float(xcolmax[c]);
    float s = 1.f;
    if (bound > 0.f && bound < 3.0e38f) {
        int e;
        frexpf(bound, &e);
        e = 14 - e;
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
        s = ldexpf(1.f, e);
    }
    ysc[0] = s;
    ysc[1] = 1.f / s;
}

}   // namespace tc

static inline bool f_update_tc_supported(int k) {
    switch (k) { case 40: case 64: return true; }
    return false;
}

// Launch of the tcgen05 Gram pipeline: same contract as f_update_mma_launch (f_update_mma.cuh) for MODE_DEFER / MODE_GRAD.
// `Xh` receives the pre-split fp16 copy of the gathered factor (max rows x 32 ceil(k/8) bytes: the session's Xs buffer serves when
// 8 divides k); queue[0] is unused here, queue[8 .. 8+k) = per-column max |x|, queue[256 .. 256+k) = per-column max |w| (MODE_GRAD),
// queue[400] = max |y|; ysc = 2 floats.  The setmaxnreg arithmetic of the kernel assumes ptxas' launch allocation: it is checked
// against cudaFuncGetAttributes once per instantiation, and the launch is refused (return 2) if it does not hold.
template <int K, int MODE>
static inline int f_update_tc_launch_k(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *idx, const float *val, const float *X,
                                       size_t xrows, void *Xh, float *invs, float *F, float *Gout, uint32_t nseries, unsigned *queue,
                                       float *ysc, unsigned long long *launches, const float *Wv, int gaccum, double *frow, bool rescale) {
    typedef tc::Cfg<K> C;
    constexpr bool GRAD = MODE == tc::MODE_GRAD;
    auto kfn = tc::f_update_tc_kernel<K, MODE>;
    static int checked = 0;     // 0 = not yet, 1 = ok, -1 = refused
    if (checked == 0) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kfn) != cudaSuccess) return 1;
        checked = fa.numRegs == C::launch_regs(GRAD) ? 1 : -1;
        if (checked == 1 && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem) != cudaSuccess) return 1;
    }
    if (checked < 0) return 2;
    if (rescale) {
        if (cudaMemsetAsync(queue, 0, sizeof(unsigned) * 512, st) != cudaSuccess) return 1;
        const size_t total = xrows * (size_t)K;
        unsigned g1 = (unsigned)((total + 255) / 256);
        if (g1 > (unsigned)(4 * num_sms)) g1 = (unsigned)(4 * num_sms);
        if (g1 == 0) g1 = 1;
        fm::colscale_max_kernel<<<g1, 256, 0, st>>>(X, xrows, K, queue + 8);
        tc::presplit_kernel<<<g1, 256, sizeof(float) * K, st>>>(X, xrows, K, C::NG, queue + 8, reinterpret_cast<__half *>(Xh), invs);
        *launches += 2;
    } else if (cudaMemsetAsync(queue + 400, 0, sizeof(unsigned), st) != cudaSuccess) return 1;
    tc::absmax_range_kernel<<<2 * num_sms, 256, 0, st>>>(val, ptr, nseries, queue + 400);
    if (GRAD) {
        const size_t total = (size_t)nseries * K;
        unsigned g1 = (unsigned)((total + 255) / 256);
        if (g1 > (unsigned)(4 * num_sms)) g1 = (unsigned)(4 * num_sms);
        if (g1 == 0) g1 = 1;
        fm::colscale_max_kernel<<<g1, 256, 0, st>>>(Wv, nseries, K, queue + 256);
        ++*launches;
    }
    tc::weight_scale_kernel<<<1, 1, 0, st>>>(queue + 400, GRAD ? queue + 8 : nullptr, GRAD ? queue + 256 : nullptr, K, ysc);
    unsigned grid = (unsigned)num_sms;
    if (grid > nseries) grid = nseries;
    kfn<<<grid ? grid : 1, 32 * C::nwarps(GRAD), C::smem, st>>>(ptr, idx, val, reinterpret_cast<const unsigned char *>(Xh), invs, F, Gout, nseries, Wv,
                                                                gaccum, frow, ysc);
    *launches += 3;
    return cudaGetLastError() != cudaSuccess;
}
template <int MODE>
static inline int f_update_tc_launch(cudaStream_t st, int num_sms, const uint64_t *ptr, const uint32_t *idx, const float *val, const float *X,
                                     size_t xrows, void *Xh, float *invs, float *F, float *Gout, int k, uint32_t nseries, unsigned *queue,
                                     float *ysc, unsigned long long *launches, const float *Wv = nullptr, int gaccum = 0, double *frow = nullptr,
                                     bool rescale = true) {
    switch (k) {
        case 40: return f_update_tc_launch_k<40, MODE>(st, num_sms, ptr, idx, val, X, xrows, Xh, invs, F, Gout, nseries, queue, ysc, launches, Wv, gaccum, frow, rescale);
        case 64: return f_update_tc_launch_k<64, MODE>(st, num_sms, ptr, idx, val, X, xrows, Xh, invs, F, Gout, nseries, queue, ysc, launches, Wv, gaccum, frow, rescale);
    }
    return 1;
}

#else   // float64 build

static inline bool f_update_tc_supported(int) { return false; }
template <int MODE>
static inline int f_update_tc_launch(cudaStream_t, int, const uint64_t *, const uint32_t *, const V *, const V *, size_t, void *, float *, V *, V *, int,
                                     uint32_t, unsigned *, float *, unsigned long long *, const V * = nullptr, int = 0, double * = nullptr, bool = true) { return 1; }
#endif   // TRMF_F32
