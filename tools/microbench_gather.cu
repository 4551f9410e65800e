// microbench_gather.cu -- how fast can one SM stage gathered factor rows in shared memory on B200?
//
// The Gram kernels gather one factor row (k values) per observed entry out of L2.  Round 1 used LDGSTS (cp.async,
// 16-byte pieces) and claimed "TMA was 3x slower" without a committed measurement.  This harness measures the pure
// staging rate -- producer warp(s) fill a ring of shared-memory stages, a consumer warp releases them at once -- for
//
//   L  cp.async 16-byte pieces (LDGSTS), all 4 warps, 2-stage wait_group pipeline per warp   (the round-1 scheme)
//   B  cp.async.bulk (UBLKCP), one contiguous row per request, mbarrier complete_tx
//   T  cp.async.bulk.tensor.2d tile, box {64 halfs, 1 row}, SWIZZLE_128B: two requests per entry (h1 | h2 halves)
//   G  cp.async.bulk.tensor.2d tile::gather4: four rows per request, SWIZZLE_128B: two requests per 4 entries
//
// over a table of `rows` pre-split fp16 factor rows [h1 (k) | h2 (k)] = 4k bytes, with sorted random row indices of
// density p per series (C2: T = 10 000 rows, p = 0.9, k = 40; C5: T = 100 000, p = 0.02, k = 64).  `verify` copies one
// stage back and checks the layout the tensor-map modes produce (row r of the stage = 128-byte line r, 16-byte chunk
// c stored at chunk c ^ (r % 8)).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench_gather tools/microbench_gather.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int SE = 64;          // entries per stage
constexpr int NST = 6;          // stages
constexpr int LINE = 128;       // bytes of one swizzle row (64 halfs)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_tile2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *gmem) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

enum { MODE_L = 0, MODE_B = 1, MODE_T = 2, MODE_G = 3 };

// every CTA walks `nseries` index lists round-robin; clk[bid] = cycles spent, dump = one stage copied out (verify)
template <int MODE>
__global__ void __launch_bounds__(160) gather_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                                                     const unsigned char *__restrict__ table, int rowb, int k,
                                                     const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, uint32_t nseries,
                                                     unsigned long long *__restrict__ clk, unsigned long long *__restrict__ total_entries,
                                                     unsigned char *__restrict__ dump, int dump_stage_no) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[NST], empty[NST];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_bytes = (MODE == MODE_B) ? SE * rowb : 2 * SE * LINE;     // tensor modes: h1 lines then h2 lines
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    unsigned long long entries = 0;
    if (MODE == MODE_L) {
        // round-1 scheme: each of 4 warps stages tiles of 16 rows with LDGSTS pieces, 2-stage wait_group pipeline
        if (warp < 4) {
            const int pieces = rowb / 16;
            unsigned char *st = smem + warp * 2 * (16 * 272);
            for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
                const uint64_t lo = ptr[j];
                const uint32_t nnz = (uint32_t)(ptr[j + 1] - lo);
                const int ntiles = (int)((nnz + 15) / 16);
                int s_cur = 0;
                for (int t = warp; t < ntiles + 4; t += 4) {
                    if (t < ntiles) {
                        const uint32_t base = (uint32_t)t * 16;
                        for (int id = lane; id < 16 * pieces; id += 32) {
                            const int e = id / pieces, c = id - e * pieces;
                            uint32_t ee = base + e; ee = ee < nnz ? ee : nnz - 1;
                            const uint32_t row = __ldg(idx + lo + ee);
                            cp_async16(smem_u32(st + s_cur * (16 * 272) + e * 272 + c * 16), table + (size_t)row * rowb + c * 16);
                        }
                        entries += (lane == 0) ? min(16u, nnz - base) : 0;
                    }
                    cp_async_commit();
                    cp_async_wait<1>();
                    __syncwarp();
                    s_cur ^= 1;
                }
            }
            cp_async_wait<0>();
        }
    } else if (warp == 0) {
        // ---- producer warp ----
        uint32_t it = 0;
        for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
            const uint64_t lo = ptr[j];
            const uint32_t nnz = (uint32_t)(ptr[j + 1] - lo);
            for (uint32_t base = 0; base < nnz; base += SE, ++it) {
                const uint32_t s = it % NST, ph = (it / NST) & 1;
                mbar_wait(smem_u32(&empty[s]), ph ^ 1);
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes), bar = smem_u32(&full[s]);
                const uint32_t cnt = min((uint32_t)SE, nnz - base);
                if (MODE == MODE_B) {
                    if (lane == 0) mbar_expect_tx(bar, cnt * rowb);
                    __syncwarp();
                    for (uint32_t e = lane; e < cnt; e += 32) {
                        const uint32_t row = __ldg(idx + lo + base + e);
                        bulk_copy(dst + e * rowb, table + (size_t)row * rowb, rowb, bar);
                    }
                } else if (MODE == MODE_T) {
                    if (lane == 0) mbar_expect_tx(bar, cnt * 2 * LINE);
                    __syncwarp();
                    for (uint32_t e = lane; e < cnt; e += 32) {
                        const uint32_t row = __ldg(idx + lo + base + e);
                        tma_tile2d(dst + e * LINE, &map1, 0, (int)row, bar);
                        tma_tile2d(dst + SE * LINE + e * LINE, &map2, 0, (int)row, bar);
                    }
                } else {   // MODE_G: lane q handles entries 4q .. 4q+3 (q < 16); rows past the end re-read the last row
                    if (lane == 0) mbar_expect_tx(bar, SE * 2 * LINE);
                    __syncwarp();
                    if (lane < SE / 4) {
                        uint32_t r[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t e = base + 4 * lane + q;
                            e = e < nnz ? e : nnz - 1;
                            r[q] = __ldg(idx + lo + e);
                        }
                        tma_gather4(dst + 4 * lane * LINE, &map1, 0, r[0], r[1], r[2], r[3], bar);
                        tma_gather4(dst + SE * LINE + 4 * lane * LINE, &map2, 0, r[0], r[1], r[2], r[3], bar);
                    }
                }
                entries += cnt;
            }
        }
    } else if (warp == 1) {
        // ---- consumer warp: release every stage as soon as it has landed ----
        uint32_t it = 0;
        for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
            const uint32_t nnz = (uint32_t)(ptr[j + 1] - ptr[j]);
            for (uint32_t base = 0; base < nnz; base += SE, ++it) {
                const uint32_t s = it % NST, ph = (it / NST) & 1;
                mbar_wait(smem_u32(&full[s]), ph);
                if (dump != nullptr && blockIdx.x == 0 && (int)it == dump_stage_no) {
                    for (int p = lane; p < stage_bytes / 4; p += 32)
                        reinterpret_cast<uint32_t *>(dump)[p] = reinterpret_cast<const uint32_t *>(smem + (size_t)s * stage_bytes)[p];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
            }
        }
    }
    __syncthreads();
    if (tid == 0) clk[blockIdx.x] = (unsigned long long)(clock64() - t0);
    if (lane == 0 && entries) atomicAdd(total_entries, entries);
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void *base, uint64_t rows, int k, int rowb, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)k, rows};
    cuuint64_t strides[1] = {(cuuint64_t)rowb};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box rows %d)\n", (int)r, box_rows); exit(1); }
    return m;
}

template <int MODE>
static void run(const char *name, const CUtensorMap &m1, const CUtensorMap &m2, const unsigned char *dtab, int rowb, int k, const uint64_t *dptr,
                const uint32_t *didx, uint32_t nseries, size_t nnz, int sms, const std::vector<__half> &tab, const std::vector<uint64_t> &ptr,
                const std::vector<uint32_t> &idx, bool verify) {
    unsigned long long *dclk, *dtot;
    unsigned char *ddump;
    const int stage_bytes = (MODE == MODE_B) ? SE * rowb : 2 * SE * LINE;
    const size_t smem = MODE == MODE_L ? 4 * 2 * 16 * 272 : (size_t)NST * stage_bytes;
    CHECK(cudaMalloc(&dclk, sms * sizeof(unsigned long long)));
    CHECK(cudaMalloc(&dtot, sizeof(unsigned long long)));
    CHECK(cudaMalloc(&ddump, stage_bytes));
    CHECK(cudaMemset(ddump, 0xff, stage_bytes));
    CHECK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CHECK(cudaMemset(dtot, 0, sizeof(unsigned long long)));
        CHECK(cudaEventRecord(e0));
        gather_kernel<MODE><<<sms, 160, smem + 1024>>>(m1, m2, dtab, rowb, k, dptr, didx, nseries, dclk, dtot, verify ? ddump : nullptr, 1);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
    }
    std::vector<unsigned long long> clk(sms);
    unsigned long long tot = 0;
    CHECK(cudaMemcpy(clk.data(), dclk, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(&tot, dtot, sizeof tot, cudaMemcpyDeviceToHost));
    unsigned long long cmax = 0;
    for (auto c : clk) cmax = std::max(cmax, c);
    printf("  %-44s %8.3f ms  %7.2f G rows/s  %6.2f TB/s of row bytes  %5.2f clk per row per SM  (%llu rows)\n", name, ms, nnz / (ms * 1e-3) / 1e9,
           (double)nnz * rowb / (ms * 1e-3) / 1e12, (double)cmax * sms / (double)nnz, tot);
    if (verify && MODE != MODE_L) {
        std::vector<unsigned char> d(stage_bytes);
        CHECK(cudaMemcpy(d.data(), ddump, stage_bytes, cudaMemcpyDeviceToHost));
        // CTA 0, stage number 1 = entries 64..127 of series 0
        int bad = 0;
        for (int e = 0; e < SE && bad < 4; ++e) {
            const uint32_t row = idx[ptr[0] + SE + e];
            const __half *src = &tab[(size_t)row * (rowb / 2)];
            if (MODE == MODE_B) {
                if (memcmp(&d[(size_t)e * rowb], src, rowb)) { ++bad; printf("    verify: row %d differs\n", e); }
            } else {
                for (int half = 0; half < 2 && bad < 4; ++half)
                    for (int c = 0; c < 8 && bad < 4; ++c) {
                        const unsigned char *got = &d[(size_t)half * SE * LINE + (size_t)e * LINE + ((c ^ (e & 7)) * 16)];
                        unsigned char want[16];
                        for (int q = 0; q < 8; ++q) {
                            const int col = 8 * c + q;
                            const __half v = col < k ? src[half * k + col] : __float2half(0.f);
                            memcpy(&want[2 * q], &v, 2);
                        }
                        if (memcmp(got, want, 16)) { ++bad; printf("    verify: entry %d half %d chunk %d differs\n", e, half, c); }
                    }
            }
        }
        printf("    verify: %s\n", bad ? "MISMATCH" : "layout ok (line e, chunk c at c ^ (e %% 8), columns >= k zero-filled)");
    }
    cudaFree(dclk); cudaFree(dtot); cudaFree(ddump);
}

int main(int argc, char **argv) {
    const bool verify = !(argc > 1 && !strcmp(argv[1], "noverify"));
    int dev = 0, sms = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qres));
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    struct Case { const char *name; size_t rows; int k; double p; uint32_t nseries; } cases[] = {
        {"C2-like: 10 000 rows, k = 40 (160 B rows), p = 0.9", 10000, 40, 0.9, 2960},
        {"C5-like: 100 000 rows, k = 64 (256 B rows), p = 0.02", 100000, 64, 0.02, 29600},
    };
    for (const Case &c : cases) {
        const int rowb = 4 * c.k;
        srand(7);
        std::vector<__half> tab((c.rows + 1) * (size_t)(2 * c.k));
        for (auto &v : tab) v = __float2half((float)(rand() % 2001 - 1000) / 64.f);
        for (int q = 0; q < 2 * c.k; ++q) tab[c.rows * (size_t)(2 * c.k) + q] = __float2half(0.f);   // the all-zero row
        std::vector<uint64_t> ptr(c.nseries + 1, 0);
        std::vector<uint32_t> idx;
        for (uint32_t j = 0; j < c.nseries; ++j) {
            for (size_t i = 0; i < c.rows; ++i)
                if ((double)rand() / RAND_MAX < c.p) idx.push_back((uint32_t)i);
            ptr[j + 1] = idx.size();
        }
        const size_t nnz = idx.size();
        unsigned char *dtab; uint64_t *dptr; uint32_t *didx;
        CHECK(cudaMalloc(&dtab, tab.size() * 2)); CHECK(cudaMalloc(&dptr, ptr.size() * 8)); CHECK(cudaMalloc(&didx, nnz * 4));
        CHECK(cudaMemcpy(dtab, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(dptr, ptr.data(), ptr.size() * 8, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(didx, idx.data(), nnz * 4, cudaMemcpyHostToDevice));
        printf("%s: %zu rows gathered by %d CTAs (1 per SM), stages of %d rows x %d in flight\n", c.name, nnz, sms, SE, NST);
        CUtensorMap m1 = make_map(enc, dtab, c.rows + 1, c.k, rowb, 1), m2 = make_map(enc, dtab + 2 * c.k, c.rows + 1, c.k, rowb, 1);
        run<MODE_L>("L  LDGSTS 16 B pieces, 4 warps x 2 stages x 16 rows", m1, m2, dtab, rowb, c.k, dptr, didx, c.nseries, nnz, sms, tab, ptr, idx, verify);
        run<MODE_B>("B  cp.async.bulk, one row per request", m1, m2, dtab, rowb, c.k, dptr, didx, c.nseries, nnz, sms, tab, ptr, idx, verify);
        run<MODE_T>("T  tensor tile {64,1} SW128, 2 requests per row", m1, m2, dtab, rowb, c.k, dptr, didx, c.nseries, nnz, sms, tab, ptr, idx, verify);
        run<MODE_G>("G  tensor tile::gather4 SW128, box {64,1}", m1, m2, dtab, rowb, c.k, dptr, didx, c.nseries, nnz, sms, tab, ptr, idx, verify);
        if (argc > 2 && !strcmp(argv[2], "box4")) {
            CUtensorMap g1 = make_map(enc, dtab, c.rows + 1, c.k, rowb, 4), g2 = make_map(enc, dtab + 2 * c.k, c.rows + 1, c.k, rowb, 4);
            run<MODE_G>("G' tensor tile::gather4 SW128, box {64,4}", g1, g2, dtab, rowb, c.k, dptr, didx, c.nseries, nnz, sms, tab, ptr, idx, verify);
        }
        cudaFree(dtab); cudaFree(dptr); cudaFree(didx);
    }
    return 0;
}
