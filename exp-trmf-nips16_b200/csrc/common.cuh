// common.cuh -- shared device helpers for the B200 TRMF solver (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef ValueType
#define ValueType float
#endif
typedef ValueType V;

#define TRMF_WARP 32
#define FULL_MASK 0xffffffffu

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Deterministic block reduction (fixed tree): every thread passes a value,
// thread 0 returns the block sum.  `red` = shared double[32].
__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();               // protect `red` against a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (wid == 0) {
        v = lane < nw ? red[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// Grid-wide deterministic sum into out[slot]: each block writes its partial to
// part[blockIdx.x]; the last block to finish (atomic ticket) adds the partials
// in index order.  The result does not depend on block scheduling.
__device__ __forceinline__ void grid_sum_commit(double block_val, double *part, unsigned *ticket,
                                                double *out, double scale, double *red) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = block_val;
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double v = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(part + i);
        v = block_sum(v, red);
        if (threadIdx.x == 0) {
            *out = v * scale;
            *ticket = 0u;   // re-arm for the next launch on this stream
        }
    }
}

// Cholesky solve of an SPD system held in shared memory, by the whole CTA.
//
// Layout: `A` has n+1 rows of leading dimension ld (row-major, fp64).  Rows
// 0..n-1 hold the LOWER triangle of the matrix (entries j <= i are read), row n
// holds the right-hand side b.  On return row n holds the solution x.
// `dinv` = shared double[n] scratch.
//
// Replaces LAPACK ?posv('U', n, 1) as called by ls_solve_chol
// (rf_matrix.h:3008-3014): same factorisation (U = L^T), always in fp64.  The
// forward substitution is folded into the factorisation by treating b as an
// extra row of L (y[c] = L[n][c]); only the backward substitution remains and
// is done by warp 0 with the running vector in registers.  n <= 128.
// As in the reference, a non-positive pivot is not reported (posv's `info` is
// ignored by trmf.cpp:395,482); it yields NaN/Inf in the affected row only.
__device__ __forceinline__ void block_chol_solve(double *A, int ld, double *dinv, int n) {
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4, ny = blockDim.x >> 4;
    for (int c = 0; c < n; ++c) {
        __syncthreads();   // trailing update of column c-1 finished
        const double inv = 1.0 / sqrt(A[c * ld + c]);
        if (tid == 0) dinv[c] = inv;
        for (int i = c + 1 + tid; i <= n; i += blockDim.x) A[i * ld + c] *= inv;
        __syncthreads();
        for (int i = c + 1 + ty; i <= n; i += ny) {
            const double lic = A[i * ld + c];
            const int jmax = i < n ? i : n - 1;
            for (int j = c + 1 + tx; j <= jmax; j += 16) A[i * ld + j] -= lic * A[j * ld + c];
        }
    }
    __syncthreads();
    // backward substitution L^T x = y, warp 0; lane owns x[lane + 32 q]
    if (tid < 32) {
        double y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int i = tid + 32 * q; y[q] = i < n ? A[n * ld + i] : 0.0; }
        for (int c = n - 1; c >= 0; --c) {
            const int q = c >> 5, owner = c & 31;
            double v = q == 0 ? y[0] : q == 1 ? y[1] : q == 2 ? y[2] : y[3];
            v *= dinv[c];
            v = __shfl_sync(FULL_MASK, v, owner);
            if (tid == owner) { if (q == 0) y[0] = v; else if (q == 1) y[1] = v; else if (q == 2) y[2] = v; else y[3] = v; }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) { const int i = tid + 32 * qq; if (i < c) y[qq] -= A[c * ld + i] * v; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int i = tid + 32 * q; if (i < n) A[n * ld + i] = y[q]; }
    }
    __syncthreads();
}
