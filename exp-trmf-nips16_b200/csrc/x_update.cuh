// x_update.cuh -- K2..K6: the AR-regularised update of the temporal factor ("X rows").
//
// The reference takes ONE inexact Newton step per outer iteration
// (arr_solver::solve trmf.h:175-186 -> TRON::tron_trustregion rf_tron.h:135-254
// with max_iter = 1, pure CG trcg rf_tron.h:412-505, <= 20 steps, stop when
// ||r|| <= 0.1 ||g||, accept iff actred > 1e-4 prered).  The objective pieces:
//   arr_base_IX   (trmf.cpp:70-149)   0.5 lI |W|^2 + 0.5 lAR sum_{i>=m} |rho_i|^2
//   arr_ls_pY_IX  (trmf.cpp:231-288)  0.5 sum_Omega (Y_ij - <W_i,H_j>)^2
// Kernels below evaluate fun / grad / Hv of these on the device.  Layouts:
// W,S,G,... are T x k row-major; H is n x k row-major; Theta (lag_val) is
// L x k column-major (Theta(l,t) = th[L*t + l]); Y is CSR by time stamp.
#pragma once
#include "common.cuh"

#define TRMF_MAX_LAGS 128   // block_chol_solve keeps the solution in 4 registers per lane: systems up to 128 x 128

// Device-side CG control (trmf_b200_x_update): the kernels of CG step j take `gate` = the address of that step's
// go flag and return at once when it is 0, so all <= 20 steps are enqueued without a host round trip per step.
// gate == nullptr: always run.  The test is grid-uniform and precedes every barrier / ticket operation.
#define CG_GATE(gate) do { if ((gate) != nullptr && *(gate) == 0) return; } while (0)

struct LagSet {           // passed by value (kernel parameter space)
    int L;
    int mid;              // max lag = last element of the sorted set (trmf.cpp:79)
    const uint32_t *lags; // device
};

// rho[i,t] = S[i,t] - sum_l Theta[l,t] S[i-lag_l,t]   (i >= mid), 0 otherwise.
// Residual in fp64 exactly like the reference (trmf.cpp:110-114).  One thread
// per (i,t), t fastest => coalesced over the row-major T x k layout.
__global__ void ar_rho_kernel(const V *__restrict__ S, const V *__restrict__ th, LagSet ls,
                              double *__restrict__ rho, size_t T, int k, const int *gate) {
    CG_GATE(gate);
    const size_t total = T * (size_t)k;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = p / k;
        const int t = (int)(p - i * k);
        double r = 0.0;
        if (i >= (size_t)ls.mid) {
            r = (double)S[p];
            const V *tht = th + (size_t)ls.L * t;
            for (int l = 0; l < ls.L; ++l) r -= (double)tht[l] * (double)S[p - (size_t)ls.lags[l] * k];
        }
        rho[p] = r;
    }
}

// out[j,t] = lI * S[j,t] + lAR * ( rho[j,t] - sum_l Theta[l,t] rho[j+lag_l,t] )
// = the closed form of the reference's forward-residual / adjoint-scatter loop
// (trmf.cpp:102-121 grad, 128-147 Hv).  Terms with j+lag_l >= T are dropped;
// rho is stored as 0 below mid.
__global__ void ar_apply_kernel(const V *__restrict__ S, const V *__restrict__ th, LagSet ls,
                                const double *__restrict__ rho, V *__restrict__ out,
                                size_t T, int k, double lambdaI, double lambdaAR, const int *gate) {
    CG_GATE(gate);
    const size_t total = T * (size_t)k;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const size_t j = p / k;
        const int t = (int)(p - j * k);
        double a = rho[p];
        const V *tht = th + (size_t)ls.L * t;
        for (int l = 0; l < ls.L; ++l) {
            const size_t jj = j + ls.lags[l];
            if (jj < T) a -= (double)tht[l] * rho[jj * k + t];
        }
        out[p] = (V)(lambdaI * (double)S[p] + lambdaAR * a);
    }
}

// scal[slot] = 0.5*lI*|S|^2 + 0.5*lAR*|rho|^2   (arr_base_IX::fun, trmf.cpp:70-97)
__global__ void base_fun_kernel(const V *__restrict__ S, const double *__restrict__ rho, size_t total,
                                double lambdaI, double lambdaAR, double *part, unsigned *ticket, double *out) {
    __shared__ double red[32];
    double a = 0.0, b = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const double s = (double)S[p], r = rho[p];
        a += s * s;
        b += r * r;
    }
    double v = 0.5 * lambdaI * a + 0.5 * lambdaAR * b;
    v = block_sum(v, red);
    grid_sum_commit(v, part, ticket, out, 1.0, red);
}

// ---------------------------------------------------------------------------
// One pass over Omega (by-time CSR): one warp per time stamp i.
//   MODE_FUN : out scalar += sum_j (Y_ij - <S_i,H_j>)^2      (trmf.cpp:231-245, fp64 sum)
//   MODE_GRAD: out_i += sum_j (<S_i,H_j> - Y_ij) H_j          (trmf.cpp:247-267)
//   MODE_HV  : out_i += sum_j  <S_i,H_j> H_j                  (trmf.cpp:269-288)
//   MODE_SPMM: out_i  = sum_j  Y_ij H_j                       (smat_x_dmat, dense-mode YH with sparse storage)
// 32 entries at a time: their H rows are gathered (coalesced along k) into a
// per-warp shared tile with an odd row stride; phase A: lane e forms the dot
// for entry e; phase B: lane t accumulates column t over the 32 entries.
// fp32 build: fp32 FMAs within a 32-entry tile, fp64 across tiles.
// ---------------------------------------------------------------------------
//   MODE_GRADFUN: MODE_GRAD and MODE_FUN from the same residuals in one pass (fun(w) and grad(w) are
//                 always evaluated at the same point at the start of the Newton step, rf_tron.h:154-158)
enum { MODE_FUN = 0, MODE_GRAD = 1, MODE_HV = 2, MODE_SPMM = 3, MODE_GRADFUN = 4 };

template <int MODE, int KR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
sparse_pass_kernel(const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ col,
                   const V *__restrict__ val, const V *__restrict__ H, const V *__restrict__ S,
                   V *__restrict__ out, int k, size_t T, bool accum,
                   double *part, unsigned *ticket, double *fout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ks = k | 1;   // odd stride: conflict-free column walks
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    V *tile = reinterpret_cast<V *>(smem_raw) + (size_t)wid * (32 * ks + k);
    V *svec = tile + 32 * ks;
    __shared__ double red[32];
    double fsum = 0.0;

    for (size_t i = (size_t)blockIdx.x * WARPS + wid; i < T; i += (size_t)gridDim.x * WARPS) {
        const uint64_t lo = ptr[i], hi = ptr[i + 1];
        if (MODE != MODE_SPMM) {
#pragma unroll
            for (int q = 0; q < KR; ++q) { const int t = lane + 32 * q; if (t < k) svec[t] = S[i * k + t]; }
        }
        double accd[KR];
#pragma unroll
        for (int q = 0; q < KR; ++q) accd[q] = 0.0;
        for (uint64_t base = lo; base < hi; base += 32) {
            const uint64_t e = base + lane;
            const bool valid = e < hi;
            const uint32_t j = valid ? col[e] : 0u;
            const V y = (valid && MODE != MODE_HV) ? val[e] : (V)0;
            const int cnt = (int)((hi - base) < 32 ? (hi - base) : 32);
            __syncwarp();
            for (int r = 0; r < cnt; ++r) {
                const uint32_t jr = __shfl_sync(FULL_MASK, j, r);
#pragma unroll
                for (int q = 0; q < KR; ++q) { const int t = lane + 32 * q; if (t < k) tile[r * ks + t] = H[(size_t)jr * k + t]; }
            }
            __syncwarp();
            V z;
            if (MODE == MODE_SPMM) {
                z = y;
            } else {
                z = (V)0;
                if (valid) {
                    const V *hr = tile + lane * ks;
                    for (int t = 0; t < k; ++t) z += svec[t] * hr[t];
                }
                if (MODE == MODE_FUN || MODE == MODE_GRADFUN) { const double rr = (double)y - (double)z; if (valid) fsum += rr * rr; }
                if (MODE == MODE_GRAD || MODE == MODE_GRADFUN) z -= y;
            }
            if (MODE != MODE_FUN) {
                V accf[KR];
#pragma unroll
                for (int q = 0; q < KR; ++q) accf[q] = (V)0;
                for (int r = 0; r < cnt; ++r) {
                    const V zr = __shfl_sync(FULL_MASK, z, r);
#pragma unroll
                    for (int q = 0; q < KR; ++q) { const int t = lane + 32 * q; if (t < k) accf[q] += zr * tile[r * ks + t]; }
                }
#pragma unroll
                for (int q = 0; q < KR; ++q) accd[q] += (double)accf[q];
            }
        }
        if (MODE != MODE_FUN) {
#pragma unroll
            for (int q = 0; q < KR; ++q) {
                const int t = lane + 32 * q;
                if (t < k) {
                    if (MODE == MODE_SPMM || !accum) out[i * k + t] = (V)accd[q];
                    else out[i * k + t] = (V)((double)out[i * k + t] + accd[q]);
                }
            }
        }
        __syncwarp();
    }
    if (MODE == MODE_FUN || MODE == MODE_GRADFUN) {
        double v = block_sum(fsum, red);
        grid_sum_commit(v, part, ticket, fout, 0.5, red);
    }
}

static inline size_t sparse_pass_smem(int k, int warps) {
    return sizeof(V) * (size_t)warps * (32 * (k | 1) + k);
}

// ---------------------------------------------------------------------------
// CG / TRON vector algebra (rf_tron.h:424-502), fused, scalars resident in
// device memory (double), reductions deterministic.
// ---------------------------------------------------------------------------
enum {   // slots of the device scalar block
    SC_F = 0, SC_FBASE, SC_FLOSS, SC_FNEW, SC_GG, SC_RTR, SC_DHD, SC_RNEW, SC_GS, SC_SR, SC_TMP, SC_TMP2, SC_SS,
    SC_CGTOL, SC_RNORM, SC_CGIT,   // device-side CG control: eps_cg*||g||, the last ||r||, CG steps taken
    SC_COUNT = 16
};

// out[slot] = <a,b>
__global__ void dot_kernel(const V *__restrict__ a, const V *__restrict__ b, size_t n,
                           double *part, unsigned *ticket, double *out, const int *gate) {
    CG_GATE(gate);
    __shared__ double red[32];
    double v = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        v += (double)a[p] * (double)b[p];
    v = block_sum(v, red);
    grid_sum_commit(v, part, ticket, out, 1.0, red);
}

// Multi-GPU Hessian-vector product, epilogue of the all-reduce: acc += reduced (the sum over ranks of the per-slab partials), and,
// in the same pass, out[slot] = <d, acc> (the d'Hd of the CG step).  Same arithmetic and summation order as axpbypcz_kernel (a = b = 1)
// followed by dot_kernel on the same grid, so the iterates are bit-identical to the two-kernel version.
__global__ void add_dot_kernel(V *__restrict__ acc, const V *__restrict__ reduced, const V *__restrict__ d, size_t n,
                               double *part, unsigned *ticket, double *out, const int *gate) {
    CG_GATE(gate);
    __shared__ double red[32];
    double v = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const V a = (V)(1.0 * (double)acc[p] + 1.0 * (double)reduced[p]);
        acc[p] = a;
        if (d != nullptr) v += (double)d[p] * (double)a;
    }
    if (out != nullptr) {
        v = block_sum(v, red);
        grid_sum_commit(v, part, ticket, out, 1.0, red);
    }
}

// s = 0, r = d = -g ; rTr = <g,g>         (rf_tron.h:424-436)
__global__ void cg_init_kernel(const V *__restrict__ g, V *__restrict__ s, V *__restrict__ r, V *__restrict__ d,
                               size_t n, double *part, unsigned *ticket, double *rtr) {
    __shared__ double red[32];
    double v = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const V gv = g[p];
        s[p] = (V)0;
        r[p] = -gv;
        d[p] = -gv;
        v += (double)gv * (double)gv;
    }
    v = block_sum(v, red);
    grid_sum_commit(v, part, ticket, rtr, 1.0, red);
}

// Device-side CG control, before the first step: cgtol = eps_cg * ||g||, and step 1 runs unless ||r0|| = ||g|| <= cgtol,
// i.e. unless g == 0 (rf_tron.h:170,441-449).  ctl[1..21] were zeroed by the host.
__global__ void cg_gate0_kernel(double *scal, int slot, int *ctl, double eps_cg) {
    const double gnorm = sqrt(scal[slot]);
    scal[SC_CGTOL] = eps_cg * gnorm;
    scal[SC_RNORM] = gnorm;
    scal[SC_CGIT] = 0.0;
    ctl[1] = (gnorm > 0.0 && !(gnorm <= eps_cg * gnorm)) ? 1 : 0;
}

// alpha = rTr / dHd ; s += alpha d ; r -= alpha Hd ; rnew = <r,r>   (rf_tron.h:460-493)
__global__ void cg_step1_kernel(V *__restrict__ s, V *__restrict__ r, const V *__restrict__ d, const V *__restrict__ Hd,
                                size_t n, double *scal, int cur, int nxt, double *part, unsigned *ticket, const int *gate) {
    CG_GATE(gate);
    __shared__ double red[32];
    const double alpha = scal[cur] / scal[SC_DHD];
    const V a = (V)alpha;
    double v = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        s[p] = s[p] + a * d[p];
        const V rv = r[p] - a * Hd[p];
        r[p] = rv;
        v += (double)rv * (double)rv;
    }
    v = block_sum(v, red);
    grid_sum_commit(v, part, ticket, scal + nxt, 1.0, red);
}

// beta = rnew / rTr ; d += (beta-1) d ; d += r                      (rf_tron.h:494-502)
// (rTr <- rnew is done by swapping the two scalar slots `cur`/`nxt` on the host)
// With device-side control (`ctl` != nullptr, this being step `j`): one thread also decides whether step j+1 runs --
// the reference's loop head, rf_tron.h:441-456: stop when ||r|| <= cgtol (the step cap is the number of steps enqueued).
__global__ void cg_step2_kernel(V *__restrict__ d, const V *__restrict__ r, size_t n, double *scal, int cur, int nxt,
                                int *ctl, int j) {
    if (ctl != nullptr) {
        if (ctl[j] == 0) return;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const double rnorm = sqrt(scal[nxt]);
            scal[SC_RNORM] = rnorm;
            scal[SC_CGIT] = (double)j;
            ctl[j + 1] = rnorm <= scal[SC_CGTOL] ? 0 : 1;
        }
    }
    const double beta = scal[nxt] / scal[cur];
    const V bm1 = (V)beta - (V)1;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        V dv = d[p];
        dv = dv + bm1 * dv;
        d[p] = dv + r[p];
    }
}

// w_new = w + s ; gs = <g,s> ; sr = <s,r>                           (rf_tron.h:183-190)
__global__ void tron_trial_kernel(const V *__restrict__ w, const V *__restrict__ s, const V *__restrict__ g,
                                  const V *__restrict__ r, V *__restrict__ wnew, size_t n,
                                  double *part, unsigned *ticket, double *gs_out, double *sr_out) {
    __shared__ double red[32];
    double a = 0.0, b = 0.0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const V sv = s[p];
        wnew[p] = w[p] + sv;
        a += (double)g[p] * (double)sv;
        b += (double)sv * (double)r[p];
    }
    a = block_sum(a, red);
    b = block_sum(b, red);
    // two commits share the ticket: do them through two partial arrays
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = a;
        part[gridDim.x + blockIdx.x] = b;
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double va = 0.0, vb = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) { va += __ldcg(part + i); vb += __ldcg(part + gridDim.x + i); }
        va = block_sum(va, red);
        vb = block_sum(vb, red);
        if (threadIdx.x == 0) { *gs_out = va; *sr_out = vb; *ticket = 0u; }
    }
}

// ---------------------------------------------------------------------------
// Hessian-vector product of the loss through precomputed per-time-stamp Grams:
//   (Hv)_i = ( sum_{j in Omega_i} h_j h_j^T ) s_i = G_i s_i        (same map as trmf.cpp:269-288)
// H is fixed during the X-update, so G_i (k x k, fp32, full symmetric square, built once per
// X-update by f_update_tiled_kernel<.., SOLVE=false>) turns every CG step from a walk over Omega
// (N (4+4k) gathered bytes) into a stream over T k^2 floats.  One warp per time stamp; lane t owns
// output component(s) t, t+32 and walks the rows u of the symmetric G_i, so every load is a
// coalesced row segment; fp64 accumulation.  out_i = (accum ? out_i : 0) + G_i s_i, and the CG
// scalar d^T(Hd) is reduced in the same pass (deterministic two-level sum) when dhd != nullptr.
// ---------------------------------------------------------------------------
#ifdef TRMF_F32
// Bandwidth-shaped variant for the fp32 build (k % 4 == 0): a warp pulls a whole k x k Gram with ceil(k*k/128)
// independent 16-byte loads per lane (the loop over rows below issues 2 dependent-free loads per iteration and sits at
// ~1.2 TB/s), forms the fp64 partial dot of every float4 with the matching piece of s_i, and lane t folds the k/4
// partials of row t (the Gram is stored as a bitwise-symmetric square, so row sums = column sums).
template <int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
gram_matvec4_kernel(const float *__restrict__ Gm, const float *__restrict__ S, float *__restrict__ out, size_t T, bool accum,
                    double *part, unsigned *ticket, double *dhd, const int *gate) {
    CG_GATE(gate);
    constexpr int NV = K * K / 4, NIT = (NV + 31) / 32, CPR = K / 4;
    __shared__ double red[32];
    __shared__ __align__(16) float sS[WARPS][K];
    __shared__ double sP[WARPS][NIT * 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double dsum = 0.0;
    for (size_t i = (size_t)blockIdx.x * WARPS + wid; i < T; i += (size_t)gridDim.x * WARPS) {
        const float4 *Gi = reinterpret_cast<const float4 *>(Gm + i * (size_t)K * K);
        float4 g[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int p = lane + 32 * it;
            g[it] = p < NV ? __ldg(Gi + p) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int t = lane; t < K; t += 32) sS[wid][t] = S[i * K + t];
        __syncwarp();
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int p = lane + 32 * it;
            if (p < NV) {
                const int u = p / CPR, c4 = p - u * CPR;
                const float4 sv = *reinterpret_cast<const float4 *>(&sS[wid][4 * c4]);
                sP[wid][p] = ((double)g[it].x * (double)sv.x + (double)g[it].y * (double)sv.y) +
                             ((double)g[it].z * (double)sv.z + (double)g[it].w * (double)sv.w);
            }
        }
        __syncwarp();
        for (int t = lane; t < K; t += 32) {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < CPR; ++c) a += sP[wid][t * CPR + c];
            const double o = (accum ? (double)out[i * K + t] : 0.0) + a;
            const float ov = (float)o;
            out[i * K + t] = ov;
            dsum += (double)sS[wid][t] * (double)ov;
        }
        __syncwarp();
    }
    if (dhd != nullptr) {
        double v = block_sum(dsum, red);
        grid_sum_commit(v, part, ticket, dhd, 1.0, red);
    }
}
#endif

template <int KR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
gram_matvec_kernel(const V *__restrict__ Gm, const V *__restrict__ S, V *__restrict__ out, int k, size_t T, bool accum,
                   double *part, unsigned *ticket, double *dhd, const int *gate) {
    CG_GATE(gate);
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double dsum = 0.0;
    for (size_t i = (size_t)blockIdx.x * WARPS + wid; i < T; i += (size_t)gridDim.x * WARPS) {
        const V *Gi = Gm + i * (size_t)k * k;
        V sv[KR];
        double acc[KR];
#pragma unroll
        for (int q = 0; q < KR; ++q) {
            const int t = lane + 32 * q;
            sv[q] = t < k ? S[i * k + t] : (V)0;
            acc[q] = 0.0;
        }
        for (int u = 0; u < k; ++u) {
            const V su = __shfl_sync(FULL_MASK, sv[u >> 5], u & 31);   // KR <= 2 here: see launcher
            const V *row = Gi + (size_t)u * k;
#pragma unroll
            for (int q = 0; q < KR; ++q) {
                const int t = lane + 32 * q;
                if (t < k) acc[q] += (double)row[t] * (double)su;
            }
        }
#pragma unroll
        for (int q = 0; q < KR; ++q) {
            const int t = lane + 32 * q;
            if (t < k) {
                const double o = (accum ? (double)out[i * k + t] : 0.0) + acc[q];
                const V ov = (V)o;
                out[i * k + t] = ov;
                dsum += (double)sv[q] * (double)ov;
            }
        }
    }
    if (dhd != nullptr) {
        double v = block_sum(dsum, red);
        grid_sum_commit(v, part, ticket, dhd, 1.0, red);
    }
}
