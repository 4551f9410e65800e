// f_update.cuh -- K1: the per-series ridge solve ("F rows"), sparse mode.
//
// Replaces l2r_ls_pY_IX_chol::solve (reference python/trmf/corelib/trmf.cpp:369-397):
// for every series j with at least one observation
//     ( sum_{i in Omega_j} x_i x_i^T + lambda I ) f_j = sum_{i in Omega_j} Y_ij x_i
// where x_i = row i of the T x k temporal factor (the reference's W) and f_j =
// row j of the n x k series factor (the reference's H).  Series without
// observations keep their previous row (trmf.cpp:374).
//
// Input is the by-series orientation of Y, i.e. the CSC arrays of the T x n
// matrix (PyMatrix col_ptr / row_idx / val), which is what the reference reads
// after its CSR<->CSC pointer swap (rf_matrix.h:1633-1640).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// Generic kernel: any k <= 128.  One CTA per series (grid-stride), the Gram's
// upper triangle distributed over the threads' registers, factor rows staged
// through shared memory ENT entries at a time.  fp32 build: products are
// accumulated with fp32 FMAs inside one staged tile and flushed into fp64
// partners after every tile (<= ENT terms per fp32 partial sum), so the Gram
// that reaches the fp64 Cholesky carries ~1e-7 relative error instead of the
// ~1e-5 of a plain fp32 running sum over thousands of entries.
// ---------------------------------------------------------------------------
template <int MAXP, int THREADS, int ENT>
__global__ void __launch_bounds__(THREADS)
f_update_generic_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ idx,
                        const V *__restrict__ val, const V *__restrict__ X, V *__restrict__ F,
                        int k, double lambda, uint32_t nseries) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = k + 1;                                  // odd-ish stride for the fp64 Gram
    double *A = reinterpret_cast<double *>(smem_raw);      // (k+1) x ld : Gram rows + rhs row
    double *dinv = A + (size_t)(k + 1) * ld;               // k
    V *tile = reinterpret_cast<V *>(dinv + k);             // ENT x k
    V *tval = tile + (size_t)ENT * k;                      // ENT
    const int tid = threadIdx.x;
    const int npairs = k * (k + 1) / 2;

    // thread-owned pairs (s <= t) of the upper triangle, fixed for the whole launch
    unsigned char ps[MAXP], pt[MAXP];
#pragma unroll
    for (int q = 0; q < MAXP; ++q) {
        int p = tid + q * THREADS;
        int s = 0, t = 0;
        if (p < npairs) {
            // row-major walk of the upper triangle: row s has k - s entries
            int rem = p;
            s = 0;
            while (rem >= k - s) { rem -= k - s; ++s; }
            t = s + rem;
        }
        ps[q] = (unsigned char)s;
        pt[q] = (unsigned char)t;
    }

    for (uint32_t j = blockIdx.x; j < nseries; j += gridDim.x) {
        const uint64_t lo = ptr[j], hi = ptr[j + 1];
        if (lo == hi) continue;   // uniform across the CTA
        double accd[MAXP];
        double rhsd = 0.0;
#pragma unroll
        for (int q = 0; q < MAXP; ++q) accd[q] = 0.0;

        for (uint64_t base = lo; base < hi; base += ENT) {
            const int cnt = (int)((hi - base) < (uint64_t)ENT ? (hi - base) : (uint64_t)ENT);
            __syncthreads();   // previous tile fully consumed
            // stage: cnt factor rows, coalesced along k
            for (int e = tid / 32; e < cnt; e += THREADS / 32) {
                const uint32_t row = idx[base + e];
                for (int t = tid & 31; t < k; t += 32) tile[e * k + t] = X[(size_t)row * k + t];
            }
            for (int e = tid; e < cnt; e += THREADS) tval[e] = val[base + e];
            __syncthreads();
            V accf[MAXP];
#pragma unroll
            for (int q = 0; q < MAXP; ++q) accf[q] = (V)0;
            V rhsf = (V)0;
            for (int e = 0; e < cnt; ++e) {
                const V *row = tile + e * k;
#pragma unroll
                for (int q = 0; q < MAXP; ++q) accf[q] += row[ps[q]] * row[pt[q]];
                if (tid < k) rhsf += tval[e] * row[tid];
            }
#pragma unroll
            for (int q = 0; q < MAXP; ++q) accd[q] += (double)accf[q];
            rhsd += (double)rhsf;
        }
        __syncthreads();
        // scatter into the LOWER triangle of A (A[t][s], t >= s), lambda on the diagonal
#pragma unroll
        for (int q = 0; q < MAXP; ++q) {
            int p = tid + q * THREADS;
            if (p < npairs) A[pt[q] * ld + ps[q]] = accd[q] + (ps[q] == pt[q] ? lambda : 0.0);
        }
        if (tid < k) A[k * ld + tid] = rhsd;
        block_chol_solve(A, ld, dinv, k);   // syncs internally
        if (tid < k) F[(size_t)j * k + tid] = (V)A[k * ld + tid];
    }
}

static inline size_t f_update_generic_smem(int k, int ent) {
    return sizeof(double) * ((size_t)(k + 1) * (k + 1) + k) + sizeof(V) * ((size_t)ent * k + ent);
}
