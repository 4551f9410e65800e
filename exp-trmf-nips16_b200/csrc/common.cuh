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

// Blocked variant of block_chol_solve (same contract, same layout): panels of 8 columns.
// The panel (rows c0..n, 8 columns) is factored by warp 0 entirely in registers -- lane l owns
// rows c0+l, c0+l+32, ... (SLOTS of them, so n+1 <= 32*SLOTS) and pivots travel by shuffle --
// and the trailing matrix gets one rank-8 update by all threads: 2 CTA barriers per 8 columns
// instead of 2 per column, no shared-memory round trip inside a panel, one rsqrt per pivot.
template <int SLOTS>
__device__ __forceinline__ void block_chol_solve_blocked(double *A, int ld, double *dinv, int n) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = tid & 15, ty = tid >> 4, ny = blockDim.x >> 4;
    for (int c0 = 0; c0 < n; c0 += 8) {
        const int w = n - c0 < 8 ? n - c0 : 8;
        __syncthreads();   // trailing update of the previous panel finished
        if (tid < 32) {
            double p[SLOTS][8];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int row = c0 + lane + 32 * s;
#pragma unroll
                for (int q = 0; q < 8; ++q) p[s][q] = (row <= n && q < w) ? A[row * ld + c0 + q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q < w) {
                    const double d = __shfl_sync(FULL_MASK, p[0][q], q);      // pivot A[c][c], c = c0 + q
                    const double inv = rsqrt(d);
                    if (lane == 0) dinv[c0 + q] = inv;
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s)
                        if (lane + 32 * s > q) p[s][q] *= inv;                 // rows below the pivot: L[i][c]
#pragma unroll
                    for (int q2 = q + 1; q2 < 8; ++q2) {
                        const double lj = __shfl_sync(FULL_MASK, p[0][q], q2);  // L[c0+q2][c]
#pragma unroll
                        for (int s = 0; s < SLOTS; ++s)
                            if (lane + 32 * s >= q2) p[s][q2] -= p[s][q] * lj; // lower part of panel column q2
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int row = c0 + lane + 32 * s;
                if (row <= n) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (q < w && lane + 32 * s > q) A[row * ld + c0 + q] = p[s][q];
                }
            }
        }
        __syncthreads();
        // trailing update: A[i][j] -= sum_q L[i][c0+q] L[j][c0+q]   (i in (c0+w .. n], j in [c0+w .. min(i, n-1)])
        const int t0 = c0 + w;
        for (int i = t0 + ty; i <= n; i += ny) {
            double li[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) li[q] = q < w ? A[i * ld + c0 + q] : 0.0;
            const int jmax = i < n ? i : n - 1;
            for (int j = t0 + tx; j <= jmax; j += 16) {
                double acc = A[i * ld + j];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc -= li[q] * (q < w ? A[j * ld + c0 + q] : 0.0);
                A[i * ld + j] = acc;
            }
        }
    }
    __syncthreads();
    // backward substitution L^T x = y, warp 0; lane owns x[lane + 32 q]
    if (tid < 32) {
        double y[SLOTS];
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) { const int i = lane + 32 * q; y[q] = i < n ? A[n * ld + i] : 0.0; }
        for (int c = n - 1; c >= 0; --c) {
            const int q = c >> 5, owner = c & 31;
            double v = y[0];
#pragma unroll
            for (int qq = 1; qq < SLOTS; ++qq) v = q == qq ? y[qq] : v;
            v *= dinv[c];
            v = __shfl_sync(FULL_MASK, v, owner);
#pragma unroll
            for (int qq = 0; qq < SLOTS; ++qq) {
                const int i = lane + 32 * qq;
                if (i == c) y[qq] = v;
                else if (i < c) y[qq] -= A[c * ld + i] * v;
            }
        }
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) { const int i = lane + 32 * q; if (i < n) A[n * ld + i] = y[q]; }
    }
    __syncthreads();
}

// One-warp variant on PACKED lower-triangular storage: row i (i = 0..n, row n = right-hand side, n entries) starts at i (i + 1) / 2.
// Same panels, same arithmetic in the same order as block_chol_solve_blocked -- bit-identical factors -- but half the shared memory
// per system and no CTA barrier: the deferred solve kernels run one system per warp and keep 2-4x as many systems in flight per SM
// (the factorisation is a chain of short dependent steps; with 128 threads per system ncu showed 50 % barrier stalls and 43k clk
// per 40 x 40 system at 6 systems per SM).  Called by all 32 lanes of one warp; n + 1 <= 32 * SLOTS.
__device__ __forceinline__ int tri_off(int i) { return (i * (i + 1)) >> 1; }
template <int SLOTS>
__device__ __forceinline__ void warp_chol_solve_packed(double *A, double *dinv, int n) {
    const int lane = threadIdx.x & 31;
    const int tx = lane & 15, ty = lane >> 4;
    for (int c0 = 0; c0 < n; c0 += 8) {
        const int w = n - c0 < 8 ? n - c0 : 8;
        __syncwarp();   // trailing update of the previous panel finished
        {
            double p[SLOTS][8];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int row = c0 + lane + 32 * s;
#pragma unroll
                for (int q = 0; q < 8; ++q) p[s][q] = (row <= n && q < w && c0 + q <= row) ? A[tri_off(row) + c0 + q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q < w) {
                    const double d = __shfl_sync(FULL_MASK, p[0][q], q);      // pivot A[c][c], c = c0 + q
                    const double inv = rsqrt(d);
                    if (lane == 0) dinv[c0 + q] = inv;
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s)
                        if (lane + 32 * s > q) p[s][q] *= inv;                 // rows below the pivot: L[i][c]
#pragma unroll
                    for (int q2 = q + 1; q2 < 8; ++q2) {
                        const double lj = __shfl_sync(FULL_MASK, p[0][q], q2);  // L[c0+q2][c]
#pragma unroll
                        for (int s = 0; s < SLOTS; ++s)
                            if (lane + 32 * s >= q2) p[s][q2] -= p[s][q] * lj; // lower part of panel column q2
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int row = c0 + lane + 32 * s;
                if (row <= n) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (q < w && lane + 32 * s > q) A[tri_off(row) + c0 + q] = p[s][q];
                }
            }
        }
        __syncwarp();
        // trailing update: A[i][j] -= sum_q L[i][c0+q] L[j][c0+q]   (i in (c0+w .. n], j in [c0+w .. min(i, n-1)])
        const int t0 = c0 + w;
        for (int i = t0 + ty; i <= n; i += 2) {
            const double *ri = A + tri_off(i);
            double li[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) li[q] = q < w ? ri[c0 + q] : 0.0;
            const int jmax = i < n ? i : n - 1;
            for (int j = t0 + tx; j <= jmax; j += 16) {
                const double *rj = A + tri_off(j);
                double acc = ri[j];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc -= li[q] * (q < w ? rj[c0 + q] : 0.0);
                A[tri_off(i) + j] = acc;
            }
        }
    }
    __syncwarp();
    // backward substitution L^T x = y; lane owns x[lane + 32 q]
    {
        double y[SLOTS];
        const int rn = tri_off(n);
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) { const int i = lane + 32 * q; y[q] = i < n ? A[rn + i] : 0.0; }
        for (int c = n - 1; c >= 0; --c) {
            const int q = c >> 5, owner = c & 31;
            double v = y[0];
#pragma unroll
            for (int qq = 1; qq < SLOTS; ++qq) v = q == qq ? y[qq] : v;
            v *= dinv[c];
            v = __shfl_sync(FULL_MASK, v, owner);
            const double *rc = A + tri_off(c);
#pragma unroll
            for (int qq = 0; qq < SLOTS; ++qq) {
                const int i = lane + 32 * qq;
                if (i == c) y[qq] = v;
                else if (i < c) y[qq] -= rc[i] * v;
            }
        }
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) { const int i = lane + 32 * q; if (i < n) A[rn + i] = y[q]; }
    }
    __syncwarp();
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
