// proto_f_update_tc.cu -- stage 2 of the tcgen05 experiment (see tools/microbench_tcgen05_gram.cu and DESIGN.md,
// "Why no tcgen05 for the Gram", re-examined): a complete sparse F-update Gram pass on the 5th-generation tensor core,
// standalone (own synthetic problem, own fp64 check), shaped like the product's MODE_DEFER launch:
//
//   per series j:  G_j = sum_{e in Omega_j} x_e x_e^T   (k x k, fp64 out),   rhs_j = sum_e y_e x_e
//
//   * CTA = 4 warps, one series at a time from an atomic queue (as f_update_mma.cuh);
//   * 16 entries per tile: the gathered factor rows (cp.async, 16-byte pieces, fp32, entry-major) are split by the same
//     4 warps into h1 = fp16(x), h2 = fp16(x - h1) and written as 16-byte rows of MN-major, unswizzled core matrices:
//     A = [h1 ; h2] stacked in M (groups 0..4 and 8..12 of a 128 x 16 operand), B = the h1 part (N = 48);
//   * the tile's Y values ride in the spare columns 40, 41 of the h1 part (scaled by a power of two, split like x), so
//     rhs_c = D[c][40] + D[c][41] + D[64+c][40] falls out of the same instruction;
//   * ONE tcgen05.mma (M 128, N 48, K 16) per tile, issued by thread 0, accumulating in TMEM; its commit frees the
//     operand stage;  every FLUSH tiles (128 entries) the accumulator is drained into fp64 registers (tcgen05.ld);
//   * series end: G = D_top + S + S^T assembled in shared memory, column scaling undone, written out.
//
// NOT YET RUN (round 1's GPU budget was spent); compiled for sm_100a.  Expected first results: correctness against
// fp64, then entries/s against the mma.sync kernel's 28.3 G entries/s at k = 40 (3.18 ms at C2).  The descriptor
// strides (LBO / SBO) are taken from stage 1's finding: pass "swap" as argv[1] if stage 1 reports the swapped assignment.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/proto_f_update_tc tools/proto_f_update_tc.cu
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int K = 40, NCG = K / 8, ET = 16, UM = 128, UN = 48;
constexpr int NST = 4;                         // operand stages
constexpr int GST = 4;                         // gather stages
constexpr int FLUSH = 8;                       // tiles between accumulator drains
constexpr int GROUP_BYTES = 256, TILE_BYTES = 16 * GROUP_BYTES;
constexpr int ROW_BYTES = K * 4;               // one gathered factor row
constexpr int GTILE_BYTES = ET * ROW_BYTES + ET * 4;   // 16 rows + 16 y values
constexpr int PIECES = ET * (ROW_BYTES / 16);  // 160 16-byte pieces per tile
constexpr int TMEM_COLS = 64;
constexpr int LD = K + 1;

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
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory"); }
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);

// out[j] = K x K Gram (row-major, fp64) followed by K rhs values: (K + 1) * K doubles per series
__global__ void __launch_bounds__(128)
f_update_tc_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                   const float *__restrict__ Xs, const float *__restrict__ invs, float yscale, float yinv, uint32_t nseries,
                   unsigned *__restrict__ queue, double *__restrict__ out, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *optile = smem;                                   // [NST][TILE_BYTES]   operand stages
    unsigned char *gtile = smem + NST * TILE_BYTES;                 // [GST][GTILE_BYTES]  gathered rows + y
    double *A = reinterpret_cast<double *>(gtile + GST * GTILE_BYTES + 64);   // K x LD: assembled Gram, column K = rhs
    __shared__ __align__(8) unsigned long long bars[NST + 1];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned next_series;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int p = tid; p < NST * TILE_BYTES / 16; p += 128) reinterpret_cast<uint4 *>(optile)[p] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s <= NST; ++s) mbar_init(smem_u32(&bars[s]), 1);
        fence_barrier_init();
        next_series = atomicAdd(queue, 1u);
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t bar_acc = smem_u32(&bars[NST]);
    uint32_t acc_phase = 0;
    uint32_t issued = 0;                // MMAs issued so far by this CTA (same on every thread): operand stage = issued % NST,
                                        // and that stage has been committed issued / NST times before

    // this thread's role in a tile
    const bool xunit = tid < ET * NCG, yunit = tid >= ET * NCG && tid < ET * NCG + ET;
    const int ue = xunit ? tid / NCG : tid - ET * NCG, ucg = xunit ? tid - (tid / NCG) * NCG : NCG;
    const uint32_t uoff = (uint32_t)(ue >> 3) * 128 + (uint32_t)(ue & 7) * 16;

    uint32_t j = next_series;
    while (j < nseries) {
        const uint64_t lo = ptr[j];
        const uint32_t nnz = (uint32_t)(ptr[j + 1] - lo);
        double accr[UN];
#pragma unroll
        for (int c = 0; c < UN; ++c) accr[c] = 0.0;
        if (nnz != 0) {
            const uint32_t *sidx = idx + lo;
            const float *sval = val + lo;
            const int ntiles = (int)((nnz + ET - 1) / ET);
            // gather of tile t into gather stage t % GST: thread handles pieces tid and tid + 128 (< 160) and, for
            // tid < 16, the y value of entry tid.  Rows past the end of the series are not fetched (the converter
            // writes zeros for them).
            auto issue = [&](int t) {
                if (t < ntiles) {
                    unsigned char *g = gtile + (t % GST) * GTILE_BYTES;
                    const uint32_t base = (uint32_t)t * ET;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int p = tid + 128 * h;
                        if (p < PIECES) {
                            const int e = p / (ROW_BYTES / 16), pc = p - e * (ROW_BYTES / 16);
                            if (base + e < nnz) {
                                const uint32_t row = __ldg(sidx + base + e);
                                cp_async16(g + e * ROW_BYTES + pc * 16, Xs + (size_t)row * K + pc * 4);
                            }
                        }
                    }
                    if (tid < ET && base + tid < nnz) cp_async4(g + ET * ROW_BYTES + tid * 4, sval + base + tid);
                }
                cp_async_commit();
            };
            auto drain = [&]() {
                if (tid == 0) umma_commit(bar_acc);
                mbar_wait(bar_acc, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll
                for (int c8 = 0; c8 < UN / 8; ++c8) {
                    float v[8];
                    tmem_ld8(taddr + c8 * 8, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) accr[c8 * 8 + q] += (double)v[q];
                }
                tc_fence_before();
                __syncthreads();
            };

#pragma unroll
            for (int t = 0; t < GST - 1; ++t) issue(t);
            int since_flush = 0;
            for (int t = 0; t < ntiles; ++t) {
                issue(t + GST - 1);
                cp_async_wait<GST - 1>();                      // this thread's pieces of tile t have landed ...
                __syncthreads();                               // ... and so have everybody else's
                const uint32_t ost = issued % NST;
                if (issued >= NST) mbar_wait(smem_u32(&bars[ost]), (issued / NST - 1) & 1);   // the MMA that last read this stage is done
                unsigned char *tile = optile + ost * TILE_BYTES;
                const unsigned char *g = gtile + (t % GST) * GTILE_BYTES;
                const int cnt = (int)min((uint32_t)ET, nnz - (uint32_t)t * ET);
                if (xunit) {
                    uint4 w1 = make_uint4(0, 0, 0, 0), w2 = make_uint4(0, 0, 0, 0);
                    if (ue < cnt) {
                        const float4 x0 = *reinterpret_cast<const float4 *>(g + ue * ROW_BYTES + ucg * 32);
                        const float4 x1 = *reinterpret_cast<const float4 *>(g + ue * ROW_BYTES + ucg * 32 + 16);
                        const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                        __half h1[8], h2[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            h1[q] = __float2half_rn(xs[q]);
                            h2[q] = __float2half_rn(xs[q] - __half2float(h1[q]));
                        }
                        w1 = *reinterpret_cast<const uint4 *>(h1);
                        w2 = *reinterpret_cast<const uint4 *>(h2);
                    }
                    *reinterpret_cast<uint4 *>(tile + ucg * GROUP_BYTES + uoff) = w1;
                    *reinterpret_cast<uint4 *>(tile + (8 + ucg) * GROUP_BYTES + uoff) = w2;
                } else if (yunit) {     // columns 40 (y1) and 41 (y2) of the h1 part; its h2 counterpart stays zero
                    uint4 w = make_uint4(0, 0, 0, 0);
                    if (ue < cnt) {
                        const float y = *reinterpret_cast<const float *>(g + ET * ROW_BYTES + ue * 4) * yscale;
                        const __half y1 = __float2half_rn(y), y2 = __float2half_rn(y - __half2float(y1));
                        w.x = (uint32_t)__half_as_ushort(y1) | ((uint32_t)__half_as_ushort(y2) << 16);
                    }
                    *reinterpret_cast<uint4 *>(tile + NCG * GROUP_BYTES + uoff) = w;
                }
                fence_proxy_async();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tile);
                    umma_f16(tmem_d, make_desc(a_addr, lbo_bytes, sbo_bytes), make_desc(a_addr, lbo_bytes, sbo_bytes), IDESC,
                             since_flush > 0 ? 1u : 0u);
                    umma_commit(smem_u32(&bars[ost]));
                }
                ++issued;
                ++since_flush;
                if (since_flush == FLUSH && t + 1 < ntiles) { drain(); since_flush = 0; }
            }
            cp_async_wait<0>();
            drain();
        }
        // ---- series epilogue: G = D_top + S + S^T, rhs = D[c][40] + D[c][41] + D[64+c][40]; undo the scaling ----
        const int r = tid < 64 ? tid : tid - 64;
        const bool top = tid < K, bot = tid >= 64 && tid < 64 + K;
        if (top) {
#pragma unroll
            for (int c = 0; c < K; ++c) A[r * LD + c] = accr[c];
            A[r * LD + K] = accr[K] + accr[K + 1];
        }
        __syncthreads();
        if (bot) {
#pragma unroll
            for (int c = 0; c < K; ++c) A[r * LD + c] += accr[c];
            A[r * LD + K] += accr[K];
        }
        __syncthreads();
        if (bot) {
#pragma unroll
            for (int c = 0; c < K; ++c) A[c * LD + r] += accr[c];
        }
        __syncthreads();
        double *o = out + (size_t)j * ((K + 1) * K);
        for (int p = tid; p < K * K; p += 128) {
            const int rr = p / K, cc = p - rr * K;
            o[p] = A[rr * LD + cc] * ((double)invs[rr] * (double)invs[cc]);
        }
        if (tid < K) o[K * K + tid] = A[tid * LD + K] * ((double)invs[tid] * (double)yinv);
        if (tid == 0) next_series = atomicAdd(queue, 1u);
        __syncthreads();
        j = next_series;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

int main(int argc, char **argv) {
    const bool swap = argc > 1 && !strcmp(argv[1], "swap");
    const uint32_t lbo = swap ? GROUP_BYTES : 128, sbo = swap ? 128 : GROUP_BYTES;
    const size_t T = 10000;
    const uint32_t n = 4096;
    srand(11);
    auto rnd = []() { return (double)rand() / RAND_MAX; };
    std::vector<float> X(T * K);
    for (size_t i = 0; i < T; ++i)
        for (int c = 0; c < K; ++c) X[i * K + c] = (float)((rnd() - 0.3) * (c % 7 == 0 ? 40.0 : 1.0) * (c == 3 ? 1e-3 : 1.0));   // badly scaled columns
    std::vector<uint64_t> ptr(n + 1, 0);
    std::vector<uint32_t> idx;
    std::vector<float> val;
    for (uint32_t j = 0; j < n; ++j) {
        const double dens = j == 5 ? 0.0 : (j % 97 == 0 ? 0.001 : 0.5 + 0.4 * rnd());   // an empty series, some tiny ones
        for (size_t i = 0; i < T; ++i)
            if (rnd() < dens) { idx.push_back((uint32_t)i); val.push_back((float)(rnd() * 6 - 2)); }
        ptr[j + 1] = idx.size();
    }
    const size_t nnz = idx.size();
    // per-column power-of-two scales: column maximum into [2^14, 2^15)   (colscale_*_kernel of the product)
    std::vector<float> invs(K), Xs(T * K);
    for (int c = 0; c < K; ++c) {
        float m = 0; for (size_t i = 0; i < T; ++i) m = std::max(m, std::fabs(X[i * K + c]));
        int e; std::frexp(m, &e);
        const float s = std::ldexp(1.f, 15 - e);
        invs[c] = 1.f / s;
        for (size_t i = 0; i < T; ++i) Xs[i * K + c] = X[i * K + c] * s;
    }
    float ym = 0; for (float v : val) ym = std::max(ym, std::fabs(v));
    int ye; std::frexp(ym, &ye);
    const float yscale = std::ldexp(1.f, 15 - ye), yinv = 1.f / yscale;

    uint64_t *dptr; uint32_t *didx; float *dval, *dXs, *dinvs; unsigned *dq; double *dout;
    CHECK(cudaMalloc(&dptr, (n + 1) * sizeof(uint64_t))); CHECK(cudaMalloc(&didx, nnz * sizeof(uint32_t)));
    CHECK(cudaMalloc(&dval, nnz * sizeof(float))); CHECK(cudaMalloc(&dXs, Xs.size() * sizeof(float)));
    CHECK(cudaMalloc(&dinvs, K * sizeof(float))); CHECK(cudaMalloc(&dq, sizeof(unsigned)));
    CHECK(cudaMalloc(&dout, (size_t)n * (K + 1) * K * sizeof(double)));
    CHECK(cudaMemcpy(dptr, ptr.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(didx, idx.data(), nnz * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dval, val.data(), nnz * sizeof(float), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dXs, Xs.data(), Xs.size() * sizeof(float), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dinvs, invs.data(), K * sizeof(float), cudaMemcpyHostToDevice));
    CHECK(cudaMemset(dout, 0, (size_t)n * (K + 1) * K * sizeof(double)));
    int dev = 0, sms = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = (size_t)NST * TILE_BYTES + (size_t)GST * GTILE_BYTES + 64 + sizeof(double) * K * LD + 1024;
    CHECK(cudaFuncSetAttribute(f_update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, f_update_tc_kernel, 128, smem));
    const int ctas = sms * std::max(1, std::min(occ, 4));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CHECK(cudaMemset(dq, 0, sizeof(unsigned)));
        CHECK(cudaEventRecord(e0));
        f_update_tc_kernel<<<ctas, 128, smem>>>(dptr, didx, dval, dXs, dinvs, yscale, yinv, n, dq, dout, lbo, sbo);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("tcgen05 F-update Gram pass: n = %u series, T = %zu, k = %d, %zu entries, %d CTAs (%d per SM): %.3f ms = %.2f G entries/s"
           "   (mma.sync kernel of the product at C2: 28.3 G entries/s)\n", n, T, K, nnz, ctas, std::max(1, std::min(occ, 4)), ms, nnz / (ms * 1e-3) / 1e9);
    // fp64 check of a few series
    std::vector<double> o((size_t)(K + 1) * K);
    double worstG = 0, worstR = 0;
    for (uint32_t j : {0u, 1u, 5u, 97u, 1000u, n - 1}) {
        CHECK(cudaMemcpy(o.data(), dout + (size_t)j * (K + 1) * K, o.size() * sizeof(double), cudaMemcpyDeviceToHost));
        std::vector<double> G((size_t)K * K, 0.0), R(K, 0.0);
        for (uint64_t e = ptr[j]; e < ptr[j + 1]; ++e) {
            const float *x = &X[(size_t)idx[e] * K];
            for (int a = 0; a < K; ++a) {
                R[a] += (double)val[e] * (double)x[a];
                for (int b = 0; b < K; ++b) G[(size_t)a * K + b] += (double)x[a] * (double)x[b];
            }
        }
        double ng = 0, dg = 0, nr = 0, dr = 0;
        for (int p = 0; p < K * K; ++p) { const double d = o[p] - G[p]; ng += d * d; dg += G[p] * G[p]; }
        for (int a = 0; a < K; ++a) { const double d = o[K * K + a] - R[a]; nr += d * d; dr += R[a] * R[a]; }
        const double eg = dg > 0 ? std::sqrt(ng / dg) : std::sqrt(ng), er = dr > 0 ? std::sqrt(nr / dr) : std::sqrt(nr);
        printf("  series %5u (%6llu entries): Gram rel. error %.2e, rhs rel. error %.2e\n", j, (unsigned long long)(ptr[j + 1] - ptr[j]), eg, er);
        worstG = std::max(worstG, eg); worstR = std::max(worstR, er);
    }
    printf("worst: Gram %.2e, rhs %.2e  (mma.sync kernel: 7e-8 on the Gram)\n", worstG, worstR);
    return 0;
}
