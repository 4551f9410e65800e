// complement.cuh -- the sparse F-update and the X-update's Gram build for a MOSTLY OBSERVED Y (fp32 build): work over the missing
// cells instead of the observed ones.
//
// Same results as l2r_ls_pY_IX_chol::solve (reference trmf.cpp:369-397) and arr_ls_pY_IX::fun / ::grad (trmf.cpp:231-267), which walk
// the observed set Omega_r of every row r (r = a series for the F-update, a time stamp for the X-update; X = the other factor):
//
//      Gram_r  = sum_{e in Omega_r} x_e x_e^T          = X^T X  -  sum_{e NOT in Omega_r} x_e x_e^T
//      rhs_r   = sum_{e in Omega_r} y_re x_e            = (Y0 X)_r          with Y0 = Y zero-filled at the missing cells
//
// BASELINE configs[1] (C2) and the traffic shape (C3) observe 90 % of their cells: the complement has a NINTH of the entries, so the
// gather-bound Gram kernel (f_update_mma2.cuh, MODE_GONLY) runs over 1e7 instead of 9e7 entries, and what needs the Y values collapses
// into one tall-skinny dense product that reads Y0 (T x n floats) once.  It is also the more accurate formulation: X^T X is summed in
// fp64 from fp32 products and the split-fp16 tensor-core part (7e-8 relative) only carries the tenth of the Gram that is subtracted.
//
// For the X-update the same Gram (stored fp32 for the CG mat-vecs) also gives the loss gradient and value at the current point w_r
// while it is still in fp64:   sum_{e in Omega_r} z_e x_e = Gram_r w_r - rhs_r,   sum z_e^2 = w_r^T (Gram_r w_r - 2 rhs_r) + sum y^2.
//
// Used when nnz >= 0.6 T n (TRMF_B200_COMPLEMENT=0 / 1 pins it); rows without any observation are handled like in the walks (F row
// untouched, zero Gram).  Everything is deterministic (fixed summation orders).
#pragma once
#include "common.cuh"
#include "dense.cuh"
#include "ingest.cuh"
// (-DValueType=... is a macro; CUB uses that word as a template parameter name)
#pragma push_macro("ValueType")
#undef ValueType
#include <cub/device/device_scan.cuh>
#pragma pop_macro("ValueType")

#ifdef TRMF_F32
namespace cm {

// ---- the missing cells of every row as an index list (complement of a CSR/CSC half with ascending indices) ----
__global__ void count_missing_kernel(const uint64_t *__restrict__ ptr, uint64_t rows, uint64_t dim, uint64_t *__restrict__ cnt) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r <= rows; r += (uint64_t)gridDim.x * blockDim.x)
        cnt[r] = r < rows ? dim - (ptr[r + 1] - ptr[r]) : 0;
}
// bitmap[r][w] = ~(bits of the observed indices): one warp per row
__global__ void missing_bitmap_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx, uint64_t rows, uint32_t words,
                                      uint32_t *__restrict__ bitmap) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < rows; r += nwarps) {
        uint32_t *bm = bitmap + r * (uint64_t)words;
        for (uint32_t w = lane; w < words; w += 32) bm[w] = 0u;
        __syncwarp();
        const uint64_t e0 = ptr[r], e1 = ptr[r + 1];
        for (uint64_t e = e0; e < e1; e += 32) {     // ascending indices: the lanes of one word are neighbours; one atomic per word and step
            const bool on = e + lane < e1;
            const uint32_t i = on ? idx[e + lane] : 0xffffffffu;
            const unsigned grp = __match_any_sync(FULL_MASK, i >> 5);
            const uint32_t bits = __reduce_or_sync(grp, on ? 1u << (i & 31) : 0u);
            if (on && lane == __ffs(grp) - 1) atomicOr(bm + (i >> 5), bits);
        }
        __syncwarp();
        for (uint32_t w = lane; w < words; w += 32) bm[w] = ~bm[w];     // (bits past `dim` in the last word are masked by the expansion)
    }
}

// Y0[t * n + j] = Y_tj at the observed cells of the series [0, nseries) behind `col_ptr` (the buffer is zeroed first); from the
// by-series CSC, so that a series slab can be placed the moment it has landed.  A thread walks one of SC_CHUNKS pieces of one series
// (its reads run along the series and stay in L1); the 32 lanes of a warp hold the same piece of 32 neighbouring series and, Y being
// mostly observed, sit on nearly the same time stamp, so a warp's stores fall into a few 128-byte lines of Y0 and the partial
// sectors merge in L2.  Block = (32, 8), grid = (ceil(nseries / 32), SC_CHUNKS / 8).
constexpr int SC_CHUNKS = 32;
__global__ void scatter_dense_csc_kernel(const uint64_t *__restrict__ col_ptr, const uint32_t *__restrict__ row_idx, const float *__restrict__ val,
                                         uint64_t nseries, uint64_t n, float *__restrict__ Y0col) {
    const uint64_t j = blockIdx.x * 32ull + threadIdx.x;
    if (j >= nseries) return;
    const uint64_t c = blockIdx.y * (uint64_t)blockDim.y + threadIdx.y;
    const uint64_t p0 = col_ptr[j], len = col_ptr[j + 1] - p0;
    const uint64_t e0 = p0 + len * c / SC_CHUNKS, e1 = p0 + len * (c + 1) / SC_CHUNKS;
    float *col = Y0col + j;
#pragma unroll 4
    for (uint64_t e = e0; e < e1; ++e) col[(uint64_t)row_idx[e] * n] = val[e];
}
// yy[t] = sum of squares of row t of Y0 (= of the observed values of time stamp t: the zero-filled cells add nothing); fp64, one warp
// per row, fixed order
__global__ void row_sumsq_dense_kernel(const float *__restrict__ Y0, uint64_t T, uint64_t n, double *__restrict__ yy) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t t = warp; t < T; t += nwarps) {
        const float *row = Y0 + t * n;
        double a = 0.0;
        for (uint64_t j = lane; j < n; j += 32) { const double v = (double)row[j]; a += v * v; }
        a = warp_sum(a);
        if (lane == 0) yy[t] = a;
    }
}
// the by-series list of missing cells, transposed into one bitmap per time stamp: bit j of bmT[t] is set iff (t, j) is missing.
// atomicOr: the result does not depend on the order, and series slabs can add their bits as they arrive.  `j0` = index of the
// first series behind cptr.
__global__ void transpose_missing_kernel(const uint64_t *__restrict__ cptr, const uint32_t *__restrict__ cidx, uint64_t nseries, uint64_t j0,
                                         uint32_t words, uint32_t *__restrict__ bmT) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t jj = warp; jj < nseries; jj += nwarps) {
        const uint64_t j = j0 + jj, e1 = cptr[jj + 1];
        const uint32_t bit = 1u << (j & 31);
        for (uint64_t e = cptr[jj] + lane; e < e1; e += 32) atomicOr(bmT + (uint64_t)cidx[e] * words + (j >> 5), bit);
    }
}
// cnt[r] = set bits of row r's bitmap among the first `dim` (cnt[rows] = 0: the scan's total lands there)
__global__ void count_bits_kernel(const uint32_t *__restrict__ bm, uint64_t rows, uint64_t dim, uint32_t words, uint64_t *__restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r <= rows; r += nwarps) {
        uint32_t c = 0;
        if (r < rows)
            for (uint32_t w = lane; w < words; w += 32) {
                uint32_t bits = bm[r * (uint64_t)words + w];
                if (w == words - 1 && (dim & 31)) bits &= (1u << (dim & 31)) - 1u;
                c += __popc(bits);
            }
        c = __reduce_add_sync(FULL_MASK, c);
        if (lane == 0) cnt[r] = c;
    }
}

// ---- tall-skinny product in fp64 on the DMMA path ----
// Cpart[split][m][c] = sum_{kappa in the split's range} A(m, kappa) * B[kappa][c],  A(m, kappa) = A[m * sm + kappa * sk] (one of sm, sk is 1),
// B row-major (kappa x N), N <= 64; fp32 inputs, fp64 products and sums (the right-hand sides need it: on an ill-conditioned system a
// 1e-7 relative error of the right-hand side alone moves the solution by 3e-5, tests/test_complement_gpu.py).
//
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) runs at the rate of the DFMA pipe on B200 -- 63.5 against 60.6 MAC per clock and SM,
// tools/microbench_dfma.cu -- but takes its operands spread over the warp: 9 doubles per lane from shared memory for 20 DMMAs
// (160 MACs per lane) at k = 40, where a register-tiled DFMA loop needs 14 for 40.  Two DFMA versions (8 column groups of 5; warp-owned
// 32-row tiles of 4 x 10 per lane) were bound by exactly that: an LDS.128 costs 4 shared-memory wavefronts whether its lanes share an
// address or not, 28 wavefronts per kappa and warp against 20 DFMA issue cycles per SM sub-partition -- 7 of the 17.6 TDFMA/s the
// pipe delivers (profiles/r02_gemm64_history.md).
//
// CTA = 4 warps on a tile of 128 rows x all N columns, 32 kappa per step staged in shared memory as doubles while the next step's values
// are already in flight in registers; warp w owns rows 32 w .. 32 w + 31 as 4 x NB accumulator tiles of 8 x 8.  Row strides of both
// tiles are 4 mod 16 doubles, so the 16 lanes of a half warp (4 kappa x 4 rows / columns) hit 16 different 8-byte banks.
// Same split-K / finish scheme as dense.cuh (gemm_finish_kernel sums the splits in order).
constexpr int GM = 128, GK = 32, GAS = GM + 4;
__device__ __forceinline__ void dmma884(double (&c)[2], const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__host__ __device__ inline int gemm64_bstride(int N) { return (N % 8 == 4) ? N : N + 4; }      // N % 4 == 0
template <int NB>
__global__ void __launch_bounds__(128, NB <= 5 ? 3 : 2)
gemm64_partial_kernel(const float *__restrict__ A, size_t sm, size_t sk, const float *__restrict__ B, size_t M, int N, size_t K,
                      size_t kchunk, double *__restrict__ Cpart) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double (*As)[GAS] = reinterpret_cast<double (*)[GAS]>(smem_raw);            // [GK][GAS]
    double *Bs = reinterpret_cast<double *>(smem_raw + sizeof(double) * GK * GAS);   // [GK][nbs] (+ slack: the last 8-column block may reach past N)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;      // fragment coordinates: row (A) / column (B) within the 8-block, kappa within the step of 4
    const int nbs = gemm64_bstride(N);
    const size_t m0 = (size_t)blockIdx.x * GM;
    const size_t k0 = (size_t)blockIdx.y * kchunk, k1 = (k0 + kchunk < K) ? k0 + kchunk : K;
    double acc[4][NB][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < NB; ++q) acc[i][q][0] = acc[i][q][1] = 0.0;
    // this thread's share of a step's tiles: A tile element u = either (row tid, kappa u) -- rows contiguous in memory -- or
    // (row tid / 32 + 4 u, kappa tid % 32) -- kappa contiguous; B tile = GK * N consecutive floats, element tid + 128 u
    constexpr int NA = GM * GK / 128, NBL = (GK * 64 + 127) / 128;
    const bool kfast = sk == 1;
    const int a_row0 = kfast ? tid >> 5 : tid, a_k0 = kfast ? tid & 31 : 0;
    const int a_drow = kfast ? 4 : 0, a_dk = kfast ? 0 : 1;
    const size_t a_step = kfast ? 4 * sm : sk;
    const int nb = GK * N;
    float ra[NA], rb[NBL];
    auto fetch = [&](size_t kk) {
        const float *pa = A + (m0 + a_row0) * sm + (kk + a_k0) * sk;
        if (m0 + GM <= M && kk + GK <= k1) {         // interior tile: no bounds to check
#pragma unroll
            for (int u = 0; u < NA; ++u) ra[u] = __ldg(pa + u * a_step);
        } else {
#pragma unroll
            for (int u = 0; u < NA; ++u)
                ra[u] = (m0 + a_row0 + a_drow * u < M && kk + a_k0 + a_dk * u < k1) ? __ldg(pa + u * a_step) : 0.f;
        }
        const float *pb = B + kk * N + tid;
        const int lim = (kk + GK <= k1) ? nb : (int)(k1 - kk) * N;
#pragma unroll
        for (int u = 0; u < NBL; ++u) rb[u] = (tid + 128 * u < lim) ? __ldg(pb + 128 * u) : 0.f;
    };
    // B tile element p = tid + 128 u sits at row p / N, column p % N: walk (row, column) instead of dividing
    const int b_r0 = tid / N, b_c0 = tid - b_r0 * N, b_dr = 128 / N, b_dc = 128 - b_dr * N;
    for (int p = tid; p < GK * nbs + 16; p += 128) Bs[p] = 0.0;     // (padding columns and the slack stay finite)
    fetch(k0);
    for (size_t kk = k0; kk < k1; kk += GK) {
        __syncthreads();          // the previous step's tiles are consumed
#pragma unroll
        for (int u = 0; u < NA; ++u) As[a_k0 + a_dk * u][a_row0 + a_drow * u] = (double)ra[u];
        {
            int r = b_r0, c = b_c0;
#pragma unroll
            for (int u = 0; u < NBL; ++u) {
                if (tid + 128 * u < nb) Bs[r * nbs + c] = (double)rb[u];
                r += b_dr; c += b_dc;
                if (c >= N) { c -= N; ++r; }
            }
        }
        __syncthreads();
        if (kk + GK < k1) fetch(kk + GK);
        const double *ap = &As[fk][32 * warp + fr], *bp = Bs + fk * nbs + fr;
#pragma unroll 2
        for (int kq = 0; kq < GK; kq += 4) {
            double a[4], b[NB];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ap[kq * GAS + 8 * i];
#pragma unroll
            for (int q = 0; q < NB; ++q) b[q] = bp[kq * nbs + 8 * q];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < NB; ++q) dmma884(acc[i][q], a[i], b[q]);
        }
    }
    // accumulator tile (i, q): lane holds row fr, columns 2 fk and 2 fk + 1
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t m = m0 + 32 * warp + 8 * i + fr;
        if (m < M) {
            double *dst = Cpart + ((size_t)blockIdx.y * M + m) * N;
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int c = 8 * q + 2 * fk;
                if (c < N) dst[c] = acc[i][q][0];
                if (c + 1 < N) dst[c + 1] = acc[i][q][1];
            }
        }
    }
}
template <int NB> constexpr size_t gemm64_smem() { return sizeof(double) * ((size_t)GK * GAS + (size_t)GK * 68 + 16); }

// ---- F-update: (X^T X - Gmiss_j + lambda I) f = rhs_j, one CTA per system at a time (fp64 Cholesky of common.cuh) ----
// sys[j] = the (K+1) x (K+1) lower-triangle layout MODE_GONLY leaves (only read when the series has missing cells), XtX = K x K fp64
// (full), rhs = n x K fp64.  A series without any observation keeps its row (trmf.cpp:374).
// NT threads per system: the factorisation is a chain of short dependent steps (panel by warp 0, barrier, trailing update, barrier),
// so a CTA of 128 spends half its time at barriers (ncu: 50 % barrier stalls, 43k clk per 40 x 40 system); fewer threads per
// system and more systems in flight per SM is the better trade -- NT is chosen by the launcher (TRMF_B200_SOLVE_THREADS pins it).
template <int K, int NT>
__global__ void __launch_bounds__(NT)
solve_kernel(const uint64_t *__restrict__ ptr, const uint64_t *__restrict__ cptr, const double *__restrict__ sys, const double *__restrict__ XtX,
             const double *__restrict__ rhs, float *__restrict__ F, double lambda, uint32_t nseries) {
    constexpr int ld = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *A = reinterpret_cast<double *>(smem_raw);
    double *dinv = A + (size_t)(K + 1) * ld;
    const int tid = threadIdx.x;
    for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
        if (ptr[j + 1] == ptr[j]) continue;
        const bool miss = cptr[j + 1] != cptr[j];
        const double *src = sys + (size_t)j * ((K + 1) * ld);
        for (int p = tid; p < K * K; p += NT) {
            const int r = p / K, c = p - r * K;
            if (c <= r) A[r * ld + c] = XtX[p] - (miss ? src[r * ld + c] : 0.0);
        }
        for (int c = tid; c < K; c += NT) A[K * ld + c] = rhs[(size_t)j * K + c];
        __syncthreads();
        for (int c = tid; c < K; c += NT) A[c * ld + c] += lambda;       // trmf.cpp:393
        block_chol_solve_blocked<(K + 32) / 32>(A, ld, dinv, K);   // starts and ends with __syncthreads
        for (int c = tid; c < K; c += NT) F[(size_t)j * K + c] = (float)A[K * ld + c];
        __syncthreads();
    }
}

// One warp per system on packed triangular storage (warp_chol_solve_packed, common.cuh): the default.  CTA = 32 threads.
template <int K> constexpr size_t solve_warp_smem() { return sizeof(double) * ((size_t)K * (K + 1) / 2 + K + K); }
template <int K>
__global__ void __launch_bounds__(32)
solve_warp_kernel(const uint64_t *__restrict__ ptr, const uint64_t *__restrict__ cptr, const double *__restrict__ sys, const double *__restrict__ XtX,
                  const double *__restrict__ rhs, float *__restrict__ F, double lambda, uint32_t nseries) {
    constexpr int ld = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *A = reinterpret_cast<double *>(smem_raw);
    double *dinv = A + (K * (K + 1) / 2 + K);
    const int lane = threadIdx.x;
    for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
        if (ptr[j + 1] == ptr[j]) continue;
        const bool miss = cptr[j + 1] != cptr[j];
        const double *src = sys + (size_t)j * ((K + 1) * ld);
        // (one flat, unrolled loop over the square: the loads of several steps are in flight together; a loop per row would pay a
        //  global-memory round trip per row)
#pragma unroll 4
        for (int p = lane; p < K * K; p += 32) {
            const int r = p / K, c = p - r * K;
            if (c <= r) {
                double v = XtX[p] - (miss ? src[r * ld + c] : 0.0);
                if (c == r) v += lambda;       // trmf.cpp:393
                A[tri_off(r) + c] = v;
            }
        }
        for (int c = lane; c < K; c += 32) A[tri_off(K) + c] = rhs[(size_t)j * K + c];
        warp_chol_solve_packed<(K + 32) / 32>(A, dinv, K);   // starts and ends with __syncwarp
        for (int c = lane; c < K; c += 32) F[(size_t)j * K + c] = (float)A[tri_off(K) + c];
        __syncwarp();
    }
}

// ---- X-update: Gram_t = H^T H - Gmiss_t -> Gout[t] (fp32, full square); loss gradient and value at the point Wv from the fp64 Gram ----
// grad row: F[t] (+)= Gram_t w_t - rhs_t;  frow[t] = w_t^T (Gram_t w_t - 2 rhs_t) + yy[t]  (= sum of squared residuals of row t).
// A time stamp without observations: zero Gram, frow 0, gradient row untouched (gaccum) or zero -- as the walk leaves it.
template <int K>
__global__ void __launch_bounds__(128)
xgram_kernel(uint64_t dim, const uint64_t *__restrict__ cptr, const double *__restrict__ sys, const double *__restrict__ HtH,
             const double *__restrict__ rhs, const double *__restrict__ yy, const float *__restrict__ Wv, float *__restrict__ F,
             float *__restrict__ Gout, int gaccum, double *__restrict__ frow, uint32_t nrows) {
    constexpr int ld = K + 1;
    __shared__ double G[K * ld];
    __shared__ double w[K];
    __shared__ double red[4];
    const int tid = threadIdx.x;
    for (uint32_t t = blockIdx.x; t < nrows; t += gridDim.x) {
        float *Gt = Gout + (size_t)t * K * K;
        if (cptr[t + 1] - cptr[t] == dim) {      // no observation at this time stamp
            for (int p = tid; p < K * K; p += 128) Gt[p] = 0.f;
            if (tid == 0) frow[t] = 0.0;
            if (!gaccum && tid < K) F[(size_t)t * K + tid] = 0.f;
            continue;
        }
        const bool miss = cptr[t + 1] != cptr[t];
        const double *src = sys + (size_t)t * ((K + 1) * ld);
        for (int p = tid; p < K * K; p += 128) {
            const int r = p / K, c = p - r * K;
            const double g = HtH[p] - (miss ? (c <= r ? src[r * ld + c] : src[c * ld + r]) : 0.0);
            G[r * ld + c] = g;
            Gt[p] = (float)g;
        }
        if (tid < K) w[tid] = (double)Wv[(size_t)t * K + tid];
        __syncthreads();
        double part = 0.0;
        if (tid < K) {
            double a = 0.0;
#pragma unroll 8
            for (int c = 0; c < K; ++c) a += G[tid * ld + c] * w[c];
            const double r = rhs[(size_t)t * K + tid];
            float *o = F + (size_t)t * K + tid;
            *o = gaccum ? (float)((double)*o + (a - r)) : (float)(a - r);
            part = w[tid] * (a - 2.0 * r);
        }
        part = warp_sum(part);
        if ((tid & 31) == 0) red[tid >> 5] = part;
        __syncthreads();
        if (tid == 0) frow[t] = ((red[0] + red[1]) + (red[2] + red[3])) + yy[t];
        __syncthreads();
    }
}

}   // namespace cm
#endif
