// microbench_tcgen05_gram.cu -- can tcgen05 build the per-series Gram after all?  (round-2 groundwork; compiled for
// sm_100a here, NOT YET RUN: round 1's GPU budget was spent when the idea came up.)
//
// DESIGN.md ("Why no tcgen05 for the Gram") dismissed the 5th-generation tensor core because a series' Gram has only
// k <= 64 rows while a tcgen05 tile wants M = 128, and counted 7.5 clk per entry per SM for a padded formulation against
// 6.0 for the bare mma.sync loop.  Two things that comparison missed:
//   1. `mma.sync` HOLDS the issue port (8 clk per HMMA, measured) -- the F-update kernel spends 216 of its 620 clk per
//      16-entry tile per SM sub-partition on that and the rest on CUDA-core work the HMMAs cannot overlap with.
//      `tcgen05.mma` is asynchronous: one thread issues it, the tensor core reads shared memory on its own.
//   2. The split-fp16 Gram  G ~ h1'h1 + h2'h1 + (h2'h1)'  (x = h1 + h2, h2'h2 dropped) fits ONE M = 128 tile when h1 and
//      h2 are STACKED in the M dimension:   A = [h1 ; h2]  (128 x 16 entries),  B = h1  (16 entries x 48)
//          D[0..39][c']    = sum_e h1[e][c] h1[e][c']          (rows   0.. 39)
//          D[64..103][c']  = sum_e h2[e][c] h1[e][c'] =: S     (rows  64..103)
//      and G = D_top + S + S'.  One `tcgen05.mma` (M 128, N 48, K 16) per 16 entries: floor 128*48/256 = 24 clk
//      (B300_MICROARCH.md, "tcgen05 floor") = 1.5 clk per entry per SM, against 9.7 measured for the whole mma.sync
//      kernel.  Both operands are MN-major (for one entry, 8 consecutive factor columns are 16 contiguous bytes), which
//      is how gathered factor rows arrive: no transpose, the converter writes one 16-byte row per (entry, column group)
//      and B is simply a second descriptor onto the h1 part of A.
// What is left per tile is the CUDA-core split (fp32 -> h1, h2), ~150 warp-instructions per 16 entries, and the L2
// gather itself -- the estimate in DESIGN.md section 7 is an F-update of ~1.6 ms at C2 (L2-gather bound) instead of 3.0.
//
// This prototype answers the questions that estimate rests on, on dense synthetic rows (no gather):
//   (a) does the descriptor / layout construction below produce the right numbers (G against fp64)?
//   (b) how does the error grow with the number of entries accumulated in TMEM (the tensor core adds with truncation:
//       -1.1e-6 over 128 entries for mma.sync) -- i.e. how often must the accumulator be flushed to registers?
//   (c) the issue floor of the M 128 / N 48 / K 16 instruction, and the cycles per tile with a 4-warp converter.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_tcgen05_gram tools/microbench_tcgen05_gram.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int K = 40;                 // rank (factor columns)
constexpr int NCG = K / 8;            // 8-column groups that carry data (5)
constexpr int ET = 16;                // entries per MMA = UMMA K for fp16
constexpr int UM = 128, UN = 48;      // MMA shape (N must be a multiple of 16 at M = 128)
constexpr int NST = 4;                // operand stages
constexpr int GROUP_BYTES = 256;      // one 8-column group of a tile: 2 K-groups x (8 entries x 16 B)
constexpr int TILE_BYTES = 16 * GROUP_BYTES;   // 16 groups of 8 rows of A: h1 in groups 0..4, h2 in groups 8..12, rest zero
constexpr int TMEM_COLS = 64;

// ---- PTX wrappers (forms as in CUTLASS's cute/arch/{mma_sm100_umma,copy_sm100,tmem_allocator_sm100}.hpp, cutlass/arch/barrier.h)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(r[q]);
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 in [0,14), LBO >> 4 in [16,30),
// SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64).  MN-major canonical layout, in 16-byte units:
// ((1,n),(8,k)) : ((x,SBO),(1,LBO)) -- 8 consecutive K (entries) 16 B apart form a core matrix whose 16-byte rows hold 8
// consecutive MN (factor columns); K groups LBO apart, MN groups SBO apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6) = 1, A/B fp16 [7,10) = [10,13) = 0, A and B MN-major
// (bits 15, 16), N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);

// One CTA (4 warps) builds the Gram of `ntiles` x 16 dense rows of X (row stride K floats).
//   flush_tiles > 0: every flush_tiles tiles the TMEM accumulator is read into fp32 registers and restarted.
//   mode 0: convert + MMA;  mode 1: MMAs only, re-using the first NST converted tiles (issue floor).
// out: 128 x 48 floats (D, summed over flushes); cyc: clock64 ticks of the main loop.
__global__ void __launch_bounds__(128, 1)
gram_tc_kernel(const float *__restrict__ X, int ntiles, int flush_tiles, int mode, float *__restrict__ out, long long *__restrict__ cyc,
               uint32_t lbo_bytes, uint32_t sbo_bytes) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[NST + 1];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const float *Xc = X + (size_t)blockIdx.x * ntiles * ET * K;

    for (int p = tid; p < NST * TILE_BYTES / 16; p += 128) reinterpret_cast<uint4 *>(smem)[p] = make_uint4(0, 0, 0, 0);   // padding groups stay zero
    if (tid == 0) {
        for (int s = 0; s <= NST; ++s) mbar_init(smem_u32(&bars[s]), 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t bar_acc = smem_u32(&bars[NST]);

    float accr[UN];
#pragma unroll
    for (int c = 0; c < UN; ++c) accr[c] = 0.f;
    uint32_t acc_phase = 0;
    int since_flush = 0;

    auto drain = [&]() {      // all MMAs issued so far have completed -> add the accumulator to the registers
        if (tid == 0) umma_commit(bar_acc);
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes (= rows of D)
#pragma unroll
        for (int c8 = 0; c8 < UN / 8; ++c8) {
            float v[8];
            tmem_ld8(taddr + c8 * 8, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) accr[c8 * 8 + q] += v[q];
        }
        tc_fence_before();
        __syncthreads();      // nobody restarts the accumulator while someone still reads it
        since_flush = 0;
    };

    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % NST;
        unsigned char *tile = smem + s * TILE_BYTES;
        if (mode == 0 || t < NST) {
            if (t >= NST) mbar_wait(smem_u32(&bars[s]), (uint32_t)((t / NST - 1) & 1));   // the MMA that read this stage is done
            if (tid < ET * NCG) {
                const int e = tid / NCG, cg = tid - e * NCG;
                const float4 *src = reinterpret_cast<const float4 *>(Xc + ((size_t)t * ET + e) * K + 8 * cg);
                const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
                const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                __half h1[8], h2[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    h1[q] = __float2half_rn(xs[q]);
                    h2[q] = __float2half_rn(xs[q] - __half2float(h1[q]));
                }
                const uint32_t off = (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16;
                *reinterpret_cast<uint4 *>(tile + cg * GROUP_BYTES + off) = *reinterpret_cast<const uint4 *>(h1);
                *reinterpret_cast<uint4 *>(tile + (8 + cg) * GROUP_BYTES + off) = *reinterpret_cast<const uint4 *>(h2);
            }
            fence_proxy_async();      // generic-proxy stores -> visible to the tensor core's async proxy
            __syncthreads();
        }
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(tile);
            const uint64_t da = make_desc(a_addr, lbo_bytes, sbo_bytes);   // A: 16 groups (128 rows) x 16 entries
            const uint64_t db = make_desc(a_addr, lbo_bytes, sbo_bytes);   // B: the first 6 groups of the same tile (h1, N = 48)
            umma_f16(tmem_d, da, db, IDESC, since_flush > 0 ? 1u : 0u);
            if (mode == 0) umma_commit(smem_u32(&bars[s]));
        }
        ++since_flush;
        if (flush_tiles > 0 && since_flush == flush_tiles && t + 1 < ntiles) drain();
    }
    drain();
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    float *o = out + ((size_t)blockIdx.x * UM + tid) * UN;
#pragma unroll
    for (int c = 0; c < UN; ++c) o[c] = accr[c];
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

static double check(const std::vector<float> &X, int rows, const float *D) {
    std::vector<double> G((size_t)K * K, 0.0);
    for (int e = 0; e < rows; ++e)
        for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b) G[(size_t)a * K + b] += (double)X[(size_t)e * K + a] * (double)X[(size_t)e * K + b];
    double num = 0, den = 0;
    for (int a = 0; a < K; ++a)
        for (int b = 0; b < K; ++b) {
            const double g = (double)D[(size_t)a * UN + b] + (double)D[(size_t)(64 + a) * UN + b] + (double)D[(size_t)(64 + b) * UN + a];
            const double df = g - G[(size_t)a * K + b];
            num += df * df; den += G[(size_t)a * K + b] * G[(size_t)a * K + b];
        }
    return std::sqrt(num / den);
}

int main() {
    int dev = 0, sms = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int max_tiles = 1024;                       // 16 384 entries per CTA
    const size_t rows = (size_t)max_tiles * ET;
    std::vector<float> X(rows * K * sms);
    srand(3);
    for (auto &x : X) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dX, *dD; long long *dC;
    CHECK(cudaMalloc(&dX, X.size() * sizeof(float)));
    CHECK(cudaMalloc(&dD, (size_t)sms * UM * UN * sizeof(float)));
    CHECK(cudaMalloc(&dC, sms * sizeof(long long)));
    CHECK(cudaMemcpy(dX, X.data(), X.size() * sizeof(float), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)NST * TILE_BYTES + 1024;
    CHECK(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> D((size_t)UM * UN);
    long long cyc = 0;
    // (a0) which descriptor field carries which stride?  Reading of cute::make_umma_desc<Major::MN> for SWIZZLE_NONE:
    // LBO = distance between the two 8-entry K groups (128 B), SBO = distance between 8-column MN groups (256 B).
    // Try that and the swapped assignment; continue with whichever reproduces the fp64 Gram.
    uint32_t lbo = 128, sbo = GROUP_BYTES;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const uint32_t l = attempt == 0 ? 128u : (uint32_t)GROUP_BYTES, sb = attempt == 0 ? (uint32_t)GROUP_BYTES : 128u;
        gram_tc_kernel<<<1, 128, smem>>>(dX, 1, 0, 0, dD, dC, l, sb);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(D.data(), dD, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
        const double err = check(X, ET, D.data());
        printf("(a0) descriptor LBO = %u B, SBO = %u B: one tile, rel. error %.2e%s\n", l, sb, err, err < 1e-3 ? "  <- correct" : "");
        if (err < 1e-3) { lbo = l; sbo = sb; break; }
    }
    printf("(a)/(b) accuracy of G = D_top + S + S' against fp64, one CTA, by entries accumulated in TMEM between flushes\n");
    const int tiles_list[] = {1, 8, 64, 512};
    const int flush_list[] = {0, 1, 8, 32};
    for (int nt : tiles_list)
        for (int fl : flush_list) {
            if (fl >= nt && fl != 0) continue;
            // one CTA reads the first nt tiles of its slice; slice stride = nt tiles, so CTA 0 sees X[0 .. nt*16)
            gram_tc_kernel<<<1, 128, smem>>>(dX, nt, fl, 0, dD, dC, lbo, sbo);
            CHECK(cudaDeviceSynchronize());
            CHECK(cudaMemcpy(D.data(), dD, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
            CHECK(cudaMemcpy(&cyc, dC, sizeof cyc, cudaMemcpyDeviceToHost));
            printf("  entries %6d  flush every %4d entries : rel. Frobenius error %.2e   (%lld clk, %.0f per tile)\n", nt * ET,
                   fl == 0 ? nt * ET : fl * ET, check(X, nt * ET, D.data()), cyc, (double)cyc / nt);
        }
    printf("(c) all %d SMs, %d tiles each\n", sms, max_tiles);
    for (int mode = 0; mode < 2; ++mode) {
        gram_tc_kernel<<<sms, 128, smem>>>(dX, max_tiles, 8, mode, dD, dC, lbo, sbo);
        CHECK(cudaDeviceSynchronize());
        std::vector<long long> c(sms);
        CHECK(cudaMemcpy(c.data(), dC, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0; for (long long v : c) avg += (double)v; avg /= sms;
        printf("  mode %d (%s): %.0f clk per 16-entry tile per SM (flush every 128 entries)\n", mode,
               mode == 0 ? "4-warp converter + MMA" : "MMA issue only", avg / max_tiles);
    }
    printf("reference points: mma.sync F-update kernel 155 clk per tile per SM (620 per sub-partition); tcgen05 floor 24\n");
    cudaFree(dX); cudaFree(dD); cudaFree(dC);
    return 0;
}
