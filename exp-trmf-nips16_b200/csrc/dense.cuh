// dense.cuh -- K8: kernels for the fully-observed mode (`missing == 0`).
//
// Replaces arr_ls_fY_IX (reference trmf.cpp:155-215) and l2r_ls_fY_IX_chol
// (trmf.cpp:299-351): all contractions with Y collapse to two tall-skinny
// products per outer iteration (Y^T W : n x k and Y H : T x k) plus k x k
// Grams, after which grad / Hv / fun only touch T x k and k x k objects.
// The contractions are tiny next to the sparse path (<= 2*T*n*k flop, once per
// phase) and are bound by reading Y once; they are done with fp64 accumulation
// so that the fp32 build stays ~1e-7 from the float64 reference.
#pragma once
#include "common.cuh"

// C[m][c] = sum_kappa A(m,kappa) * B[kappa][c],  m < M, c < N (N = k <= 128), kappa < K
// A is addressed as A[m*sm + kappa*sk] (covers row-major, col-major and
// transposed views); B is row-major K x N.  Split-K over gridDim.y: partial
// sums go to Cpart[split][M][N] (fp64) and are combined in split order by
// gemm_finish_kernel, so the result is independent of scheduling.
#define GT_M 64
#define GT_K 32
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemm_partial_kernel(const TA *__restrict__ A, size_t sm, size_t sk, const TB *__restrict__ B,
                    size_t M, int N, size_t K, size_t kchunk, double *__restrict__ Cpart, const int *gate) {
    if (gate != nullptr && *gate == 0) return;   // device-side CG control, see CG_GATE in x_update.cuh
    __shared__ double As[GT_K][GT_M + 1];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Bs = reinterpret_cast<double *>(smem_raw);   // GT_K x N
    const int tid = threadIdx.x;
    const size_t m0 = (size_t)blockIdx.x * GT_M;
    const size_t k0 = (size_t)blockIdx.y * kchunk;
    const size_t k1 = (k0 + kchunk < K) ? k0 + kchunk : K;
    const int mloc = tid & 63, cg = tid >> 6;   // 4 column groups
    double acc[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0.0;
    for (size_t kk = k0; kk < k1; kk += GT_K) {
        __syncthreads();
        // A tile GT_M x GT_K; walk the unit-stride axis fastest
        if (sk <= sm) {
            for (int p = tid; p < GT_M * GT_K; p += 256) {
                const int kq = p % GT_K, mq = p / GT_K;
                const size_t m = m0 + mq, kap = kk + kq;
                As[kq][mq] = (m < M && kap < k1) ? (double)A[m * sm + kap * sk] : 0.0;
            }
        } else {
            for (int p = tid; p < GT_M * GT_K; p += 256) {
                const int mq = p % GT_M, kq = p / GT_M;
                const size_t m = m0 + mq, kap = kk + kq;
                As[kq][mq] = (m < M && kap < k1) ? (double)A[m * sm + kap * sk] : 0.0;
            }
        }
        for (int p = tid; p < GT_K * N; p += 256) {
            const int c = p % N, kq = p / N;
            const size_t kap = kk + kq;
            Bs[kq * N + c] = kap < k1 ? (double)B[kap * N + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kq = 0; kq < GT_K; ++kq) {
            const double a = As[kq][mloc];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int c = cg + 4 * q;
                if (c < N) acc[q] += a * Bs[kq * N + c];
            }
        }
    }
    const size_t m = m0 + mloc;
    if (m < M) {
        double *dst = Cpart + ((size_t)blockIdx.y * M + m) * N;
#pragma unroll
        for (int q = 0; q < 32; ++q) { const int c = cg + 4 * q; if (c < N) dst[c] = acc[q]; }
    }
}

// out = alpha * sum_splits(Cpart) + beta * addend (addend may be NULL) [+ diag on the diagonal when M == N]
template <typename TO>
__global__ void gemm_finish_kernel(const double *__restrict__ Cpart, int splits, size_t total, int N,
                                   double alpha, const V *__restrict__ addend, double beta, double diag,
                                   TO *__restrict__ out, const int *gate) {
    if (gate != nullptr && *gate == 0) return;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        for (int s = 0; s < splits; ++s) v += Cpart[(size_t)s * total + p];
        v *= alpha;
        if (addend) v += beta * (double)addend[p];
        if (diag != 0.0 && (p / N) == (p % N)) v += diag;
        out[p] = (TO)v;
    }
}

// out[p] = a*x[p] + b*y[p] + c*z[p]  (y, z may be NULL)
__global__ void axpbypcz_kernel(double a, const V *x, double b, const V *y, double c, const V *z, V *out, size_t n) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        double v = a * (double)x[p];
        if (y) v += b * (double)y[p];
        if (z) v += c * (double)z[p];
        out[p] = (V)v;
    }
}

__global__ void dotd_kernel(const double *__restrict__ a, const double *__restrict__ b, size_t n,
                            double *part, unsigned *ticket, double *out) {
    __shared__ double red[32];
    double v = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        v += a[p] * b[p];
    v = block_sum(v, red);
    grid_sum_commit(v, part, ticket, out, 1.0, red);
}

// Dense F-update solve (l2r_ls_fY_IX_chol::solve, trmf.cpp:328-337): one k x k
// system G = W^T W + lambda I (fp64, in `G`, full symmetric) shared by all n
// right-hand sides YtW[j,:].  Each CTA factors G in shared memory (k^3/6 flop,
// negligible) and then solves 256 right-hand sides, one per thread.
template <int KMAX>
__global__ void __launch_bounds__(256)
dense_f_solve_kernel(const double *__restrict__ G, const double *__restrict__ rhs /* n x k */, size_t n, int k,
                     V *__restrict__ F) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = k + 1;
    double *A = reinterpret_cast<double *>(smem_raw);   // (k+1) x ld, row k unused rhs (zeros)
    double *dinv = A + (size_t)(k + 1) * ld;
    for (int p = threadIdx.x; p < (k + 1) * ld; p += 256) {
        const int i = p / ld, j = p % ld;
        A[p] = (i < k && j < k) ? G[i * k + j] : 0.0;
    }
    block_chol_solve(A, ld, dinv, k);   // factor only matters; the dummy rhs row is zero
    const size_t j = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    double b[KMAX];   // dynamically indexed -> local memory; this kernel is not on the critical path
    for (int c = 0; c < k; ++c) b[c] = rhs[j * k + c];
    for (int c = 0; c < k; ++c) {          // L y = b   (A[i][c] = L[i][c] for i > c, dinv[c] = 1/L[c][c])
        const double yc = b[c] * dinv[c];
        b[c] = yc;
        for (int i = c + 1; i < k; ++i) b[i] -= A[i * ld + c] * yc;
    }
    for (int c = k - 1; c >= 0; --c) {     // L^T x = y
        const double xc = b[c] * dinv[c];
        b[c] = xc;
        for (int i = 0; i < c; ++i) b[i] -= A[c * ld + i] * xc;
    }
    for (int c = 0; c < k; ++c) F[j * k + c] = (V)b[c];
}
