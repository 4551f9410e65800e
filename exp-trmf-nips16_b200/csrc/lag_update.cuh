// lag_update.cuh -- K7: the lag_val (Theta) least-squares update.
//
// Replaces l2r_autoregressive_solver::solve (reference trmf.cpp:455-484): for
// every latent dimension t, with x = W[:,t] and the window i in [mid, T):
//     G_ab  = sum_i x[i-lag_a] x[i-lag_b]  (a <= b, fp64 sums, trmf.cpp:447-453)
//     rhs_a = sum_i x[i] x[i-lag_a]
//     (G + lambdaLag I) theta_t = rhs            -> column t of the col-major Theta
// Stage 1 accumulates partial Grams over chunks of the window (one CTA per
// (t, chunk), the needed slice of column t staged in shared memory as fp64);
// stage 2 adds the partials in chunk order (deterministic) and solves.
#pragma once
#include "common.cuh"
#include "x_update.cuh"

// pair index p over (a,b), 0 <= a <= b <= L, "lag index 0" = the target x[i]
// (lag 0), index l >= 1 = lag_set[l-1].  Row-major upper triangle of (L+1)^2.
__device__ __forceinline__ void lag_pair_decode(int p, int L1, int &a, int &b) {
    a = 0;
    while (p >= L1 - a) { p -= L1 - a; ++a; }
    b = a + p;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
lag_gram_kernel(const V *__restrict__ W, LagSet ls, size_t T, int k, int chunk, int nchunks,
                double *__restrict__ partial /* [k][nchunks][npairs] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *x = reinterpret_cast<double *>(smem_raw);   // chunk + mid values of column t
    const int t = blockIdx.x, c = blockIdx.y;
    const int L1 = ls.L + 1, npairs = L1 * (L1 + 1) / 2;
    const size_t i0 = (size_t)ls.mid + (size_t)c * chunk;
    const size_t i1 = (i0 + chunk < T) ? i0 + chunk : T;
    double *dst = partial + ((size_t)t * nchunks + c) * npairs;
    if (i0 >= T) {
        for (int p = threadIdx.x; p < npairs; p += THREADS) dst[p] = 0.0;
        return;
    }
    const size_t s0 = i0 - ls.mid;   // first staged time stamp
    const int len = (int)(i1 - s0);
    for (int q = threadIdx.x; q < len; q += THREADS) x[q] = (double)W[(s0 + q) * k + t];
    __syncthreads();
    const int w = (int)(i1 - i0);
    for (int p = threadIdx.x; p < npairs; p += THREADS) {
        int a, b;
        lag_pair_decode(p, L1, a, b);
        const int la = a == 0 ? 0 : (int)ls.lags[a - 1];
        const int lb = b == 0 ? 0 : (int)ls.lags[b - 1];
        const double *xa = x + (ls.mid - la), *xb = x + (ls.mid - lb);
        double acc = 0.0;
        for (int i = 0; i < w; ++i) acc += xa[i] * xb[i];
        dst[p] = acc;
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
lag_solve_kernel(const double *__restrict__ partial, int L, int nchunks, double lambda,
                 V *__restrict__ theta /* L x k col-major */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int t = blockIdx.x;
    const int L1 = L + 1, npairs = L1 * (L1 + 1) / 2, ld = L + 1;
    double *A = reinterpret_cast<double *>(smem_raw);   // (L+1) x ld
    double *dinv = A + (size_t)(L + 1) * ld;
    for (int p = threadIdx.x; p < npairs; p += THREADS) {
        double v = 0.0;
        for (int c = 0; c < nchunks; ++c) v += partial[((size_t)t * nchunks + c) * npairs + p];
        int a, b;
        lag_pair_decode(p, L1, a, b);
        if (a == 0) {
            if (b > 0) A[L * ld + (b - 1)] = v;                 // rhs_b  (row L of the augmented system)
        } else {
            A[(b - 1) * ld + (a - 1)] = v + (a == b ? lambda : 0.0);   // lower triangle: row b-1 >= col a-1
        }
    }
    block_chol_solve(A, ld, dinv, L);
    for (int l = threadIdx.x; l < L; l += THREADS) theta[(size_t)L * t + l] = (V)A[L * ld + l];
}
