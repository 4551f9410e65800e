// rolling.cuh -- rolling-window sessions: rolling_validate with Y resident in HBM.
//
// The reference's experiment driver (python/trmf/trmf.py:303-329, run_electricity.py:15-17) fits nr_windows
// successive models, window w on the prefix Y[:T_w] of the time axis with T_w growing by window_size, each
// warm-started from the previous one (trmf.py:237-246).  Every fit there is a fresh c_trmf_train call: Y is
// converted (csr_matrix + both orientations, rf_util.py:88-98) and handed over again although consecutive windows
// differ by window_size rows out of tens of thousands, and all factors travel both ways.
//
// Here one session keeps the longest prefix any window trains on in HBM and exposes the window [0, T_w) to the three
// updates without touching the host:
//   * by-time CSR: row_ptr[0..T_w] / col_idx / val_t of the full matrix ARE the CSR of the prefix (no copy);
//   * by-series CSC: rows are ascending inside a series, so the window's entries are a prefix of every series'
//     run: count by binary search, exclusive scan, warp-per-series copy (roll_count / roll_compact) -- the result is
//     bit-identical to what a fresh ingest of Y[:T_w] produces (tests/test_rolling_gpu.py);
//   * dense row-major Y: the first T_w rows;
//   * NormalizedTransform (trmf.py:82-96; a fresh one per window, trmf.py:247-248): the per-series scale / offset
//     (n values each, computed by the host from the window's statistics like the reference does) are applied while
//     copying, y*a + b as two separately rounded operations -- the same bits NumPy's `Y * a + b` gives;
//   * factors: W[:T_prev], H and lag_val of the previous window are already in place; the host only sends the
//     window_size new rows of W (its AR roll-out, trmf.py:244) with trmf_b200_upload_W_rows.
// Buffers are sized once for the full length (S::T_cap); T-dependent plans (lag_plan) are redone per window, so a
// window trains exactly like a fresh session on Y[:T_w] started from the same factors.
#pragma once
#include "common.cuh"
#pragma push_macro("ValueType")
#undef ValueType
#include <cub/device/device_scan.cuh>
#pragma pop_macro("ValueType")

// y*a + b with each operation rounded on its own (no FMA contraction): NumPy evaluates `Y * a + b` as two ufuncs
__device__ __forceinline__ float roll_affine(float y, float a, float b) { return __fadd_rn(__fmul_rn(y, a), b); }
__device__ __forceinline__ double roll_affine(double y, double a, double b) { return __dadd_rn(__dmul_rn(y, a), b); }

// cnt[j] = number of entries of series j with time stamp < Tw (rows ascending inside a series); cnt[n] = 0
__global__ void roll_count_kernel(const uint64_t *__restrict__ col_ptr, const uint32_t *__restrict__ row_idx, uint64_t n,
                                  uint64_t Tw, uint64_t *__restrict__ cnt) {
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j <= n; j += (uint64_t)gridDim.x * blockDim.x) {
        if (j == n) { cnt[j] = 0; continue; }
        const uint64_t base = col_ptr[j];
        uint64_t lo = base, hi = col_ptr[j + 1];
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((uint64_t)row_idx[mid] < Tw) lo = mid + 1; else hi = mid;
        }
        cnt[j] = lo - base;
    }
}

// one warp per series: the first (wptr[j+1] - wptr[j]) entries of the resident series j -> the window's CSC
template <typename VT>
__global__ void roll_compact_kernel(const uint64_t *__restrict__ col_ptr, const uint32_t *__restrict__ row_idx,
                                    const VT *__restrict__ val, uint64_t n, const uint64_t *__restrict__ wptr,
                                    uint32_t *__restrict__ wrow, VT *__restrict__ wval, const VT *__restrict__ a,
                                    const VT *__restrict__ b) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = warp; j < n; j += nwarps) {
        const uint64_t src = col_ptr[j], dst = wptr[j], len = wptr[j + 1] - dst;
        const bool tr = a != nullptr;
        const VT aj = tr ? a[j] : (VT)1, bj = tr ? b[j] : (VT)0;
        for (uint64_t e = lane; e < len; e += 32) {
            wrow[dst + e] = row_idx[src + e];
            const VT y = val[src + e];
            wval[dst + e] = tr ? roll_affine(y, aj, bj) : y;
        }
    }
}

// by-time values of the window with the per-series transform: out[e] = val_t[e] * a[col_idx[e]] + b[col_idx[e]]
template <typename VT>
__global__ void roll_affine_csr_kernel(const VT *__restrict__ val_t, const uint32_t *__restrict__ col_idx, uint64_t nnz,
                                       const VT *__restrict__ a, const VT *__restrict__ b, VT *__restrict__ out) {
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < nnz; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t j = col_idx[e];
        out[e] = roll_affine(val_t[e], a[j], b[j]);
    }
}

// dense row-major window with the per-series transform
template <typename VT>
__global__ void roll_affine_dense_kernel(const VT *__restrict__ Y, uint64_t cells, uint64_t n, const VT *__restrict__ a,
                                         const VT *__restrict__ b, VT *__restrict__ out) {
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < cells; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = e % n;
        out[e] = roll_affine(Y[e], a[j], b[j]);
    }
}

// ---- dense -> sparse at ingest: the non-zero cells of a row-major T x n array as a canonical CSR, i.e. what the
// reference's rolling_validate gets from `csr_matrix(Y_trn)` when missing=True (trmf.py:320-321: exact zeros are
// unobserved; -0.0 counts as zero, NaN as a value -- both as scipy decides them).  One warp per time stamp.
template <typename VT>
__global__ void roll_dense_count_kernel(const VT *__restrict__ Y, uint64_t T, uint64_t n, uint64_t *__restrict__ cnt /* T + 1 */) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (warp == 0 && lane == 0) cnt[T] = 0;
    for (uint64_t i = warp; i < T; i += nwarps) {
        const VT *row = Y + i * n;
        unsigned c = 0;
        for (uint64_t j = lane; j < n; j += 32) c += row[j] != (VT)0 ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
        if (lane == 0) cnt[i] = c;
    }
}

template <typename VT>
__global__ void roll_dense_fill_kernel(const VT *__restrict__ Y, uint64_t T, uint64_t n, const uint64_t *__restrict__ row_ptr,
                                       uint32_t *__restrict__ col_idx, VT *__restrict__ val_t) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < T; i += nwarps) {
        const VT *row = Y + i * n;
        uint64_t pos = row_ptr[i];
        for (uint64_t j0 = 0; j0 < n; j0 += 32) {       // ballot keeps the columns of a time stamp in ascending order
            const uint64_t j = j0 + lane;
            const VT y = j < n ? row[j] : (VT)0;
            const bool keep = j < n && y != (VT)0;
            const unsigned m = __ballot_sync(FULL_MASK, keep);
            if (keep) {
                const uint64_t e = pos + __popc(m & ((1u << lane) - 1u));
                col_idx[e] = (uint32_t)j;
                val_t[e] = y;
            }
            pos += __popc(m);
        }
    }
}

static int roll_window_impl(S *s, uint64_t Tw, const void *a_host, const void *b_host);

static int roll_create_impl(S *s, const PyMatrix *Y, const uint32_t *lag_set, uint32_t lag_size, uint32_t k, int missing,
                            int device) {
    g_last_error.clear();
    s->device = device;
    s->rolling = true;
    s->T = s->T_cap = Y->rows;
    s->n = Y->cols;
    s->k = (int)k;
    s->missing = missing != 0;
    bool dense_to_sparse = false;
    if (k < 1 || k > 128) return fail("rank k = %u outside the supported range 1..128", k);
    if (s->T == 0 || s->n == 0) return fail("rolling session needs a non-empty Y");
    if (s->T >= (1ull << 32) || s->n >= (1ull << 32)) return fail("T and n must fit uint32 indices");
    if (check_lags(s, lag_set, lag_size)) return 1;
    if (Y->type == TRMF_SPARSE) {
        s->sparse_storage = true;
        s->nnz = s->R_nnz = Y->nnz;
    } else if (Y->type == TRMF_DENSE_ROWMAJOR) {
        // missing != 0 with a dense array: its non-zero cells are the observations (rolling_validate's
        // csr_matrix(Y_trn), trmf.py:320-321), sparsified on the device below
        dense_to_sparse = s->missing;
        s->sparse_storage = s->missing;
        s->dense_type = Y->type;
        s->nnz = s->T * s->n;
    } else if (Y->type == TRMF_DENSE_COLMAJOR) {
        return fail("rolling session: a dense Y must be row-major (a prefix of the time axis must be contiguous)");
    } else {
        return fail("unsupported PyMatrix type %d for Y", Y->type);
    }
    if (session_common_init(s)) return 1;
    s->own_factors = true;
    const size_t tk = s->T * (size_t)s->k, nk = s->n * (size_t)s->k, lk = (size_t)s->L * s->k;
    if (dev_alloc(&s->W, tk) || dev_alloc(&s->H, nk) || dev_alloc(&s->th, lk)) return 1;
    CUDA_TRY(cudaMemsetAsync(s->W, 0, tk * sizeof(V), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->H, 0, nk * sizeof(V), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->th, 0, lk * sizeof(V), s->stream));
    if (s->sparse_storage) {
        const size_t scan_items = std::max(s->T, s->n) + 1;
        if (dev_alloc(&s->row_ptr, s->T + 1) || dev_alloc(&s->win_cnt, scan_items)) return 1;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, s->scan_tmp_bytes, s->win_cnt, s->row_ptr, (int64_t)scan_items, s->stream));
        CUDA_TRY(cudaMallocAsync(&s->scan_tmp, s->scan_tmp_bytes ? s->scan_tmp_bytes : 1, s->stream));
        V *dense = nullptr;
        if (dense_to_sparse) {
            const unsigned grid = (unsigned)(s->num_sms * 8);
            if (dev_alloc(&dense, s->T * s->n)) return 1;
            CUDA_TRY(cudaMemcpyAsync(dense, Y->val, s->T * s->n * sizeof(V), cudaMemcpyHostToDevice, s->stream));
            LAUNCH(s, roll_dense_count_kernel<V>, grid, 256, 0, dense, (uint64_t)s->T, (uint64_t)s->n, s->win_cnt);
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->scan_tmp, s->scan_tmp_bytes, s->win_cnt, s->row_ptr, (int64_t)(s->T + 1), s->stream));
            s->launches++;
            uint64_t total = 0;
            CUDA_TRY(cudaMemcpyAsync(&total, s->row_ptr + s->T, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            s->nnz = s->R_nnz = total;
        }
        const size_t nnz = s->R_nnz;
        const bool have_csr = dense_to_sparse || (Y->row_ptr && (nnz == 0 || (Y->col_idx && Y->val_t)));
        const bool have_csc = !dense_to_sparse && Y->col_ptr && (nnz == 0 || (Y->row_idx && Y->val));
        if (!have_csr && !have_csc) return fail("sparse Y carries neither a complete CSR nor a complete CSC half");
        // may the complement formulation be used on the windows (cm_on)?  Not if some cell occurs twice in Y.
        s->idx_strict = dense_to_sparse ? true
                        : (sizeof(V) == 4 && s->missing) ? (have_csc ? host_lists_strict(Y->col_ptr, Y->row_idx, s->n) : host_lists_strict(Y->row_ptr, Y->col_idx, s->T))
                                                         : false;
        if (dev_alloc(&s->col_idx, nnz) || dev_alloc(&s->R_val_t, nnz) ||
            dev_alloc(&s->R_col_ptr, s->n + 1) || dev_alloc(&s->R_row_idx, nnz) || dev_alloc(&s->R_val, nnz) ||
            dev_alloc(&s->col_ptr, s->n + 1) || dev_alloc(&s->row_idx, nnz) || dev_alloc(&s->val, nnz))
            return 1;
        // one orientation crosses PCIe (or comes out of the dense array), the other is the stable device transpose
        // (ingest.cuh), like any session
        if (dense_to_sparse) {
            LAUNCH(s, roll_dense_fill_kernel<V>, (unsigned)(s->num_sms * 8), 256, 0, dense, (uint64_t)s->T, (uint64_t)s->n, s->row_ptr,
                   s->col_idx, s->R_val_t);
            dev_free(dense);
            CUDA_TRY(csr_from_csc_device<V>(s->stream, s->num_sms, s->n, s->T, nnz, s->row_ptr, s->col_idx, s->R_val_t, s->R_col_ptr,
                                            s->R_row_idx, s->R_val));
        } else if (have_csc) {
            CUDA_TRY(cudaMemcpyAsync(s->R_col_ptr, Y->col_ptr, (s->n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
            if (nnz) {
                CUDA_TRY(cudaMemcpyAsync(s->R_row_idx, Y->row_idx, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
                CUDA_TRY(cudaMemcpyAsync(s->R_val, Y->val, nnz * sizeof(V), cudaMemcpyHostToDevice, s->stream));
            }
            CUDA_TRY(csr_from_csc_device<V>(s->stream, s->num_sms, s->T, s->n, nnz, s->R_col_ptr, s->R_row_idx, s->R_val, s->row_ptr,
                                            s->col_idx, s->R_val_t));
        } else {
            CUDA_TRY(cudaMemcpyAsync(s->row_ptr, Y->row_ptr, (s->T + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
            if (nnz) {
                CUDA_TRY(cudaMemcpyAsync(s->col_idx, Y->col_idx, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
                CUDA_TRY(cudaMemcpyAsync(s->R_val_t, Y->val_t, nnz * sizeof(V), cudaMemcpyHostToDevice, s->stream));
            }
            CUDA_TRY(csr_from_csc_device<V>(s->stream, s->num_sms, s->n, s->T, nnz, s->row_ptr, s->col_idx, s->R_val_t, s->R_col_ptr,
                                            s->R_row_idx, s->R_val));
        }
        s->launches += 5;
    } else {
        if (dev_alloc(&s->R_Yd, s->T * s->n)) return 1;
        CUDA_TRY(cudaMemcpyAsync(s->R_Yd, Y->val, s->T * s->n * sizeof(V), cudaMemcpyHostToDevice, s->stream));
    }
    // the host buffers may go away once this returns
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return roll_window_impl(s, s->T_cap, nullptr, nullptr);
}

static int roll_window_impl(S *s, uint64_t Tw, const void *a_host, const void *b_host) {
    if (!s->rolling) return fail("trmf_b200_roll_window: not a rolling session (create it with trmf_b200_roll_create)");
    if (Tw == 0 || Tw > s->T_cap) return fail("window length %llu outside 1..%zu", (unsigned long long)Tw, s->T_cap);
    if ((a_host == nullptr) != (b_host == nullptr)) return fail("the transform needs both the scales and the offsets");
    CUDA_TRY(cudaSetDevice(s->device));
    s->T = Tw;
    s->prev_cg = -1;   // a window starts like a fresh session
    if (lag_plan(s)) return 1;
    const bool tr = a_host != nullptr;
    if (tr) {
        if (!s->aff_a && (dev_alloc(&s->aff_a, s->n) || dev_alloc(&s->aff_b, s->n))) return 1;
        // (pageable host memory: the copies are staged before the call returns; the sync below covers the rest)
        CUDA_TRY(cudaMemcpyAsync(s->aff_a, a_host, s->n * sizeof(V), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->aff_b, b_host, s->n * sizeof(V), cudaMemcpyHostToDevice, s->stream));
    }
    const unsigned grid = (unsigned)(s->num_sms * 8);
    s->valh_ok = false;      // the window's by-series values are rewritten below: their fp16 split is stale
    cm_reset(s);             // ... and so is everything the complement formulation derived from Y
    if (s->sparse_storage) {
        uint64_t nnz_w = 0;
        CUDA_TRY(cudaMemcpyAsync(&nnz_w, s->row_ptr + Tw, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
        LAUNCH(s, roll_count_kernel, (unsigned)std::min<uint64_t>((s->n + 256) / 256, grid), 256, 0, s->R_col_ptr, s->R_row_idx,
               (uint64_t)s->n, Tw, s->win_cnt);
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->scan_tmp, s->scan_tmp_bytes, s->win_cnt, s->col_ptr, (int64_t)(s->n + 1), s->stream));
        s->launches++;
        LAUNCH(s, roll_compact_kernel<V>, grid, 256, 0, s->R_col_ptr, s->R_row_idx, s->R_val, (uint64_t)s->n, s->col_ptr, s->row_idx,
               s->val, tr ? s->aff_a : (const V *)nullptr, tr ? s->aff_b : (const V *)nullptr);
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        s->nnz = nnz_w;
        if (tr) {
            if (!s->win_val_t && dev_alloc(&s->win_val_t, s->R_nnz)) return 1;
            LAUNCH(s, roll_affine_csr_kernel<V>, grid, 256, 0, s->R_val_t, s->col_idx, nnz_w, s->aff_a, s->aff_b, s->win_val_t);
            s->val_t = s->win_val_t;
        } else {
            s->val_t = s->R_val_t;
        }
    } else {
        if (tr) {
            if (!s->win_Yd && dev_alloc(&s->win_Yd, s->T_cap * s->n)) return 1;
            LAUNCH(s, roll_affine_dense_kernel<V>, grid, 256, 0, s->R_Yd, Tw * (uint64_t)s->n, (uint64_t)s->n, s->aff_a, s->aff_b, s->win_Yd);
            s->Yd = s->win_Yd;
        } else {
            s->Yd = s->R_Yd;
        }
        s->nnz = Tw * s->n;
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));   // a_host / b_host are the caller's again
    return 0;
}

extern "C" S *trmf_b200_roll_create(const PyMatrix *Y, const uint32_t *lag_set, uint32_t lag_size, uint32_t k, int32_t missing,
                                    int32_t device) {
    S *s = new S();
    if (roll_create_impl(s, Y, lag_set, lag_size, k, missing, device)) {
        std::string keep = g_last_error;
        trmf_b200_destroy(s);
        g_last_error = keep;
        return nullptr;
    }
    return s;
}

// Per-series mean and standard deviation over the first T_w time stamps of a DENSE resident Y -- the statistics of the reference's
// NormalizedTransform (trmf.py:84-88: Yd.mean(axis=0), Yd.std(axis=0)), bit for bit: NumPy reduces along axis 0 of a C-ordered
// array by adding row after row into one accumulator per column, in the array's own precision; std is sqrt(mean((y - mean)^2))
// with every operation rounded separately.  One thread per series walks the rows in that order (loads coalesced across series).
// (separately rounded multiply / add: nvcc must not contract d * d + sq into one FMA)
__device__ __forceinline__ float stat_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double stat_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float stat_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double stat_add(double a, double b) { return __dadd_rn(a, b); }
template <typename VT>
__global__ void roll_stats_kernel(const VT *__restrict__ Y, uint64_t Tw, uint64_t n, VT *__restrict__ mean_out, VT *__restrict__ std_out) {
    const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    VT acc = (VT)0;
    for (uint64_t i = 0; i < Tw; ++i) acc = stat_add(acc, Y[i * n + j]);
    const VT mean = acc / (VT)Tw;
    VT sq = (VT)0;
    for (uint64_t i = 0; i < Tw; ++i) {
        const VT d = stat_add(Y[i * n + j], -mean);
        sq = stat_add(sq, stat_mul(d, d));
    }
    mean_out[j] = mean;
    std_out[j] = (VT)sqrt((double)(sq / (VT)Tw));   // (sqrt of a float rounded from the correctly rounded double sqrt = sqrtf)
}

extern "C" int trmf_b200_roll_stats(S *s, uint64_t T_window, void *mean_host, void *std_host) {
    g_last_error.clear();
    if (!s->rolling || s->R_Yd == nullptr) return fail("roll_stats: needs a rolling session with a dense resident Y");
    if (T_window == 0 || T_window > s->T_cap) return fail("roll_stats: window of %llu time stamps outside the resident %zu", (unsigned long long)T_window, s->T_cap);
    CUDA_TRY(cudaSetDevice(s->device));
    V *d = nullptr;
    if (dev_alloc(&d, 2 * s->n)) return 1;
    LAUNCH(s, roll_stats_kernel<V>, (unsigned)((s->n + 63) / 64), 64, 0, s->R_Yd, T_window, (uint64_t)s->n, d, d + s->n);
    CUDA_TRY(cudaMemcpyAsync(mean_host, d, s->n * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(std_host, d + s->n, s->n * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    dev_free(d);
    return 0;
}

extern "C" int trmf_b200_roll_window(S *s, uint64_t T_window, const void *scale, const void *offset) {
    g_last_error.clear();
    return roll_window_impl(s, T_window, scale, offset);
}

// rows [row0, row0 + nrows) of W from / to a host buffer of nrows x k values (any session)
extern "C" int trmf_b200_upload_W_rows(S *s, uint64_t row0, uint64_t nrows, const void *src) {
    g_last_error.clear();
    if (row0 + nrows > s->T) return fail("upload_W_rows: rows [%llu, %llu) outside W (%zu rows)", (unsigned long long)row0,
                                         (unsigned long long)(row0 + nrows), s->T);
    CUDA_TRY(cudaSetDevice(s->device));
    if (nrows) CUDA_TRY(cudaMemcpyAsync(s->W + row0 * (size_t)s->k, src, nrows * (size_t)s->k * sizeof(V), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
extern "C" int trmf_b200_download_W_rows(S *s, uint64_t row0, uint64_t nrows, void *dst) {
    g_last_error.clear();
    if (row0 + nrows > s->T) return fail("download_W_rows: rows [%llu, %llu) outside W (%zu rows)", (unsigned long long)row0,
                                         (unsigned long long)(row0 + nrows), s->T);
    CUDA_TRY(cudaSetDevice(s->device));
    if (nrows) CUDA_TRY(cudaMemcpyAsync(dst, s->W + row0 * (size_t)s->k, nrows * (size_t)s->k * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

// The window's matrices back on the host (tests: bit-exactness of the windowed index work).  Sparse sessions only;
// any pointer may be NULL.  Sizes: row_ptr T_w+1, col_ptr n+1, the rest nnz_w = trmf_b200_roll_nnz().
extern "C" uint64_t trmf_b200_roll_nnz(S *s) { return (uint64_t)s->nnz; }
extern "C" int trmf_b200_roll_export(S *s, uint64_t *row_ptr, uint32_t *col_idx, void *val_t, uint64_t *col_ptr, uint32_t *row_idx,
                                     void *val) {
    g_last_error.clear();
    if (!s->rolling || !s->sparse_storage) return fail("roll_export: needs a sparse rolling session");
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t nz = s->nnz;
    if (row_ptr) CUDA_TRY(cudaMemcpyAsync(row_ptr, s->row_ptr, (s->T + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    if (col_ptr) CUDA_TRY(cudaMemcpyAsync(col_ptr, s->col_ptr, (s->n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    if (nz) {
        if (col_idx) CUDA_TRY(cudaMemcpyAsync(col_idx, s->col_idx, nz * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        if (val_t) CUDA_TRY(cudaMemcpyAsync(val_t, s->val_t, nz * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
        if (row_idx) CUDA_TRY(cudaMemcpyAsync(row_idx, s->row_idx, nz * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        if (val) CUDA_TRY(cudaMemcpyAsync(val, s->val, nz * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
