/*
 * trmf_b200.h -- C ABI of the B200-native TRMF ALS solver.
 *
 * Two shared libraries are built from the same sources, exactly like the
 * reference (corelib/Makefile:32-33, setup.py:42-58):
 *
 *     trmf_float32.so   (ValueType = float)
 *     trmf_float64.so   (ValueType = double)
 *
 * Both export every symbol below.  `c_trmf_train` is THE drop-in entry point:
 * same name, same 17 arguments, same in-place/void/stderr conventions as the
 * reference's `extern "C"` block (python/trmf/corelib/trmf.h:203-210,
 * definition trmf.cpp:696-725, ctypes prototype python/trmf/trmf.py:23-43), so
 * the reference's `trmf.py` can `CDLL` these libraries unchanged.  All other
 * symbols are additive (`trmf_b200_*`): a device-resident session API used by
 * the benchmark, the multi-GPU path and the per-phase parity tests.
 *
 * No torch / C++ types cross this boundary: plain pointers, sizes, scalars.
 * There is no CPU fallback anywhere behind these symbols; if no CUDA device is
 * usable every compute entry point fails loudly (message on stderr, non-zero
 * status / `trmf_b200_last_error()`).
 */
#ifndef TRMF_B200_H
#define TRMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- PyMatrix: bit-for-bit the reference POD (rf_matrix.h:3399-3415,
 *      rf_util.py:35-52).  80 bytes; offsets 0/8/16/24/32/40/48/56/64/72. ---- */
enum {
    TRMF_DENSE_ROWMAJOR = 1,
    TRMF_DENSE_COLMAJOR = 2,
    TRMF_SPARSE = 3,
    TRMF_EYE = 4,
    /* additive (Y only; never produced by the reference's rf_util.py): a sparse matrix whose CSC half carries the row
     * indices as one BITMAP per series instead of nnz uint32 indices -- `row_idx` points at cols x ceil(rows / 32) uint32
     * words, bit (i & 31) of word [j][i >> 5] set iff entry (i, j) is observed; `col_ptr` / `val` as in TRMF_SPARSE, the
     * CSR half absent (NULL).  A 10 %-missing panel uploads 4.1 instead of 8 bytes per observed entry; the library expands
     * the bitmap to the identical row_idx array in HBM (csrc/ingest.cuh). */
    TRMF_SPARSE_BITMAP = 5
};

typedef struct {
    uint64_t rows, cols, nnz;
    uint64_t *row_ptr;   /* CSR of the matrix:  row_ptr[rows+1], col_idx[nnz], val_t[nnz] */
    uint64_t *col_ptr;   /* CSC of the matrix:  col_ptr[cols+1], row_idx[nnz], val[nnz]   */
    uint32_t *row_idx;
    uint32_t *col_idx;
    void *val;           /* dense: the rows*cols buffer; sparse: CSC values */
    void *val_t;         /* sparse: CSR values */
    int32_t type;
} PyMatrix;

/* ---- The drop-in entry point (replaces trmf.h:205-209 / trmf.cpp:696-725) ----
 * Y: T x n (sparse required when missing != 0; dense row/col-major or sparse
 * when missing == 0).  W: T x k dense ROW-major, H: n x k dense ROW-major,
 * lag_val: L x k dense COL-major, lag_set: sorted uint32[L].  W, H, lag_val are
 * updated in place; Y is read-only.  Dimension / layout errors print the
 * reference's "[ERR MSG]: ..." lines to stderr and return without training
 * (trmf.cpp:561-596,632-634).  `threads` is accepted and ignored (the OpenMP
 * thread count of the reference has no meaning on a GPU).  `warm_start == 0`
 * draws W,H ~ U(0,1), lag_val ~ N(0,1) on the host first (trmf.cpp:547-559). */
void c_trmf_train(const PyMatrix *pyY, uint32_t *py_lag_set, uint32_t py_lag_size,
                  PyMatrix *pyW, PyMatrix *pyH, PyMatrix *pylag_val, int warm_start,
                  double lambdaI, double lambdaAR, double lambdaLag,
                  int32_t max_iter, int32_t period_W, int32_t period_H, int32_t period_Lag,
                  int32_t threads, int32_t missing, int32_t verbose);

/* ---- library info ---- */
int         trmf_b200_value_bytes(void);        /* 4 or 8: sizeof(ValueType) of this library */
const char *trmf_b200_version(void);
const char *trmf_b200_last_error(void);         /* "" when the last call on this thread succeeded */
int         trmf_b200_device_count(void);       /* <= 0: no usable CUDA device */
/* Host threads a host-buffer session may use to pack row indices into bitmaps (cores / LOCAL_WORLD_SIZE, TRMF_B200_PACK_THREADS);
 * below 4 the plain indices are uploaded instead. */
int         trmf_b200_pack_threads(void);

/* ---- device-resident session (additive) ----
 * A session owns Y (both orientations), the factors and all work vectors in
 * HBM.  Phases correspond one-to-one to the reference's three sub-solvers:
 *   f_update   = H_solver.solve   (trmf.cpp:369-397 sparse, 319-337 dense)
 *   x_update   = W_solver.solve   (rf_tron.h:135-254,412-505 over trmf.cpp:70-288)
 *   lag_update = LV_solver.solve  (trmf.cpp:455-484)
 * All functions return 0 on success, non-zero on failure. */
typedef struct trmf_b200_session trmf_b200_session;

/* Create from HOST buffers (copies H2D).  `device` = CUDA ordinal. */
trmf_b200_session *trmf_b200_create(const PyMatrix *Y, const uint32_t *lag_set, uint32_t lag_size,
                                    const PyMatrix *W, const PyMatrix *H, const PyMatrix *lag_val,
                                    int32_t missing, int32_t device);
/* By default trmf_b200_create has read everything it needs from Y's host buffers when it returns.  A caller that keeps those
 * buffers alive and unchanged until its first trmf_b200_sync / trmf_b200_download (or trmf_b200_destroy) may set
 * trmf_b200_feed_mode(1) on the creating thread: a large sparse Y is then packed and enqueued slab by slab on a feeder thread
 * while the caller already enqueues the first update, which runs under the upload instead of behind it (what c_trmf_train does
 * internally).  The mode is per thread and stays until changed; 0 restores the default. */
void trmf_b200_feed_mode(int32_t async_feed);

/* Create around arrays that ALREADY live in device memory (no copies; the
 * caller keeps ownership and must keep them alive).  Sparse Y only: by-time
 * CSR (row_ptr/col_idx/val_t) and by-series CSC (col_ptr/row_idx/val).  Both
 * halves must be canonical -- index lists strictly ascending, no cell twice (what
 * scipy's conversions and csr_from_csc produce): a mostly observed Y is solved over
 * its MISSING cells, which counts every cell once.
 * `n_total`/`col_offset`: this rank's slab [col_offset, col_offset+n) of a
 * global series axis of n_total (single GPU: n_total = n, col_offset = 0). */
trmf_b200_session *trmf_b200_create_device(uint64_t T, uint64_t n, uint64_t nnz, uint32_t k,
                                           const uint64_t *d_row_ptr, const uint32_t *d_col_idx, const void *d_val_t,
                                           const uint64_t *d_col_ptr, const uint32_t *d_row_idx, const void *d_val,
                                           const uint32_t *lag_set_host, uint32_t lag_size,
                                           void *d_W, void *d_H, void *d_lag_val, int32_t device);

void trmf_b200_destroy(trmf_b200_session *s);

int trmf_b200_set_params(trmf_b200_session *s, double lambdaI, double lambdaAR, double lambdaLag);
/* Run all device work of this session on `cuda_stream` (a cudaStream_t passed
 * as an integer/pointer, e.g. torch.cuda.current_stream().cuda_stream). */
int trmf_b200_set_stream(trmf_b200_session *s, void *cuda_stream);

int trmf_b200_f_update(trmf_b200_session *s);
int trmf_b200_x_update(trmf_b200_session *s);
int trmf_b200_lag_update(trmf_b200_session *s);
/* The reference's loop (trmf.cpp:647-693): for iter = 1..max_iter, each phase
 * iff iter % period == 0, order F -> X -> lag. */
int trmf_b200_train(trmf_b200_session *s, int32_t max_iter, int32_t period_W, int32_t period_H,
                    int32_t period_Lag, int32_t verbose);

/* Copy factors back to HOST buffers laid out as in c_trmf_train (any may be NULL). */
int trmf_b200_download(trmf_b200_session *s, void *W, void *H, void *lag_val);
/* Replace factors from HOST buffers (any may be NULL). */
int trmf_b200_upload(trmf_b200_session *s, const void *W, const void *H, const void *lag_val);
int trmf_b200_sync(trmf_b200_session *s);
/* Keep a device-side copy of the current factors / put it back (benchmark:
 * every timed step restarts from the same (W, H, lag_val)). */
int trmf_b200_save_factors(trmf_b200_session *s);
int trmf_b200_restore_factors(trmf_b200_session *s);

/* Introspection for parity tests and the benchmark. */
enum {
    TRMF_STAT_CG_ITERS = 0,      /* CG steps of the last x_update                         */
    TRMF_STAT_ACCEPTED = 1,      /* 1 if the last Newton step was accepted                */
    TRMF_STAT_F = 2,             /* objective before the step (double)                    */
    TRMF_STAT_FNEW = 3,          /* objective after  the step                             */
    TRMF_STAT_GNORM = 4,         /* ||g||_2                                               */
    TRMF_STAT_KERNEL_LAUNCHES = 5, /* kernels launched by this session so far             */
    TRMF_STAT_F_MS = 6,          /* device ms (CUDA events) of the last f_update          */
    TRMF_STAT_X_MS = 7,          /* ... last x_update                                     */
    TRMF_STAT_LAG_MS = 8,        /* ... last lag_update                                   */
    TRMF_STAT_F_KERNEL_MS = 9,   /* ... the Gram+Cholesky kernel alone inside f_update    */
    TRMF_STAT_PRERED = 10,
    TRMF_STAT_ACTRED = 11,
    TRMF_STAT_COLLECTIVES = 12,  /* NCCL collectives issued by this session so far            */
    TRMF_STAT_X_GRAM_MS = 13,    /* device ms of the Gram build (+ fused loss gradient) inside the last x_update */
    TRMF_STAT_FORMULATION = 14,  /* 1 = the session works over the MISSING cells of a mostly observed Y (csrc/complement.cuh),
                                    0 = it walks the observed entries like the reference; decided at the first sparse update */
    TRMF_STAT_CM_GRAM_MS = 15,   /* complement formulation, last whole-Y f_update: device ms of the Gram pass over the missing
                                    cells (factor pre-split + the gather kernel) ...                                       */
    TRMF_STAT_CM_PRODUCT_MS = 16,/* ... and of the fp64 tall-skinny product Y0^T W (split-K kernel + its finish)            */
    TRMF_STAT_CM_MISSING = 17    /* number of missing cells the formulation walks (0 when it is off)                        */
};
double trmf_b200_stat(trmf_b200_session *s, int32_t which);
/* Enable per-phase CUDA-event timing (off by default: it inserts stream syncs). */
int trmf_b200_enable_timing(trmf_b200_session *s, int32_t on);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) ----
 * Series slabs: rank r owns Y[:, slab_r] and the matching rows of H.  Every
 * Omega-proportional X-update pass produces a partial T x k result that is
 * summed with ncclAllReduce; W, lag_val and all CG vectors are replicated.
 * `unique_id` is the 128-byte ncclUniqueId produced by rank 0 with
 * trmf_b200_nccl_unique_id() and broadcast by the host program
 * (torch.distributed in bench.py). */
int trmf_b200_nccl_unique_id(void *out128);
int trmf_b200_dist_init(trmf_b200_session *s, int32_t rank, int32_t world, const void *unique_id128);
/* Share the communicator of `owner` with another session of this process (non-owning). */
int trmf_b200_dist_attach(trmf_b200_session *s, trmf_b200_session *owner);
/* all-gather of the per-rank H slabs into a host/device buffer on every rank
 * (rows in global series order); `counts[r]` = rows of rank r. */
int trmf_b200_allgather_H(trmf_b200_session *s, void *d_H_full, const uint64_t *counts);

/* ---- synthetic data, generated directly in HBM (bench / large configs) ----
 * Y_ij = <W*_i, H*_j> + noise * N(0,1), (i,j) observed iff hash(i,j,seed) < p.
 * Counter-based (stateless) so that any slab of columns can be generated
 * independently on any rank and reproduced on the host.  Produces the by-time
 * CSR and by-series CSC of Y[:, col_offset : col_offset+n].  Buffers are
 * allocated by the library and released by trmf_b200_free_synth(). */
typedef struct {
    uint64_t T, n, nnz;
    uint64_t *d_row_ptr; uint32_t *d_col_idx; void *d_val_t;
    uint64_t *d_col_ptr; uint32_t *d_row_idx; void *d_val;
} trmf_b200_synth;
int  trmf_b200_synth_generate(trmf_b200_synth *out, uint64_t T, uint64_t n, uint64_t n_total, uint64_t col_offset,
                              uint32_t rank_true, double p_observed, double noise, uint64_t seed, int32_t device);
void trmf_b200_free_synth(trmf_b200_synth *s);
/* plain cudaMemcpy device -> host (so that host programs need no CUDA binding of their own) */
int  trmf_b200_copy_to_host(void *dst_host, const void *src_device, uint64_t bytes);

/* Ingest primitive (replaces the second scipy conversion of reference rf_util.py:88-98): the by-time CSR of Y
 * (row_ptr u64[T+1], col_idx u32[nnz], val_t ValueType[nnz]) from its by-series CSC (col_ptr u64[n+1], row_idx
 * u32[nnz], val ValueType[nnz]) by a stable device sort; host arrays in, host arrays out, bit-identical to
 * scipy's csc.tocsr().  c_trmf_train / trmf_b200_create use the same transpose internally and upload only the
 * CSC half of a sparse PyMatrix (row_ptr / col_idx / val_t may then be NULL); TRMF_B200_HOST_CSR=1 makes them
 * upload the caller's CSR arrays instead. */
int  trmf_b200_csr_from_csc(uint64_t T, uint64_t n, uint64_t nnz, const uint64_t *col_ptr, const uint32_t *row_idx,
                            const void *val, uint64_t *row_ptr, uint32_t *col_idx, void *val_t, int32_t device);

/* Ingest primitive, host half: the per-series bitmaps (n x ceil(T / 32) uint32 words, bit (i & 31) of word [j][i >> 5] set iff
 * (i, j) is stored) that c_trmf_train packs out of plain row indices on the host cores before they cross PCIe.  No device
 * involved.  Returns 0 = packed, 1 = declined (row indices of some series not strictly ascending, or >= T: the reference's core
 * accepts those, a bitmap cannot carry them, and c_trmf_train then uploads the plain indices).  Replaces nothing in the
 * reference (rf_util.py:88-98 hands both index arrays over as they are). */
int  trmf_b200_pack_bitmap_host(uint64_t T, uint64_t n, const uint64_t *col_ptr, const uint32_t *row_idx, uint32_t *bitmap);

/* Ingest primitive: the row_idx array (uint32[nnz], ascending within every series) of a TRMF_SPARSE_BITMAP matrix,
 * expanded on the device; host arrays in, host array out (parity tests of the packed ingest). */
int  trmf_b200_bitmap_expand(uint64_t T, uint64_t n, uint64_t nnz, const uint64_t *col_ptr, const uint32_t *bitmap,
                             uint32_t *row_idx, int32_t device);

/* ---- rolling-window sessions (replaces the per-window re-ingest of the reference's
 *      rolling_validate, python/trmf/trmf.py:303-329) ----
 * trmf_b200_roll_create uploads Y ONCE (all the time stamps any window will train on: sparse with either or both
 * orientations, or dense ROW-major) and allocates factors and work space for the full length.  A dense Y with
 * missing != 0 is sparsified on the device: its non-zero cells become the observed entries, bit-identical to the
 * `csr_matrix(Y_trn)` of the reference's loop (trmf.py:320-321).  The factors start as
 * zeros: send them with trmf_b200_upload() after the first trmf_b200_roll_window().
 * trmf_b200_roll_window(s, T_w, scale, offset) makes the session train on the prefix Y[:T_w]: by-time CSR = prefix of
 * the resident one, by-series CSC re-compacted on the device (bit-identical to a fresh ingest of Y[:T_w]).  scale /
 * offset (host arrays of n values of the library's ValueType, or both NULL) apply the reference's
 * NormalizedTransform.preprocess (trmf.py:90-92) per series, y*scale[j] + offset[j], two roundings like NumPy.
 * W[:T_prev], H and lag_val of the previous window stay in place (the warm start of trmf.py:237-246); the host sends
 * the new rows of W with trmf_b200_upload_W_rows().  All other session calls (train, phases, download, stat) work
 * on the current window. */
trmf_b200_session *trmf_b200_roll_create(const PyMatrix *Y, const uint32_t *lag_set, uint32_t lag_size, uint32_t k,
                                         int32_t missing, int32_t device);
int trmf_b200_roll_window(trmf_b200_session *s, uint64_t T_window, const void *scale, const void *offset);
/* Per-series mean and standard deviation over Y[:T_window] of a rolling session with a DENSE resident Y (n values of the
 * library's ValueType each, host buffers): the statistics of the reference's NormalizedTransform (trmf.py:84-88), computed on the
 * device in NumPy's summation order, bit-identical to Yd.mean(axis=0) / Yd.std(axis=0). */
int trmf_b200_roll_stats(trmf_b200_session *s, uint64_t T_window, void *mean_host, void *std_host);
/* rows [row0, row0+nrows) of W from / to a host buffer of nrows x k values (any session) */
int trmf_b200_upload_W_rows(trmf_b200_session *s, uint64_t row0, uint64_t nrows, const void *src);
int trmf_b200_download_W_rows(trmf_b200_session *s, uint64_t row0, uint64_t nrows, void *dst);
/* the current window's matrices back on the host (parity tests of the windowed index work); sparse rolling sessions;
 * any pointer may be NULL; row_ptr T_w+1, col_ptr n+1, the others trmf_b200_roll_nnz() elements */
uint64_t trmf_b200_roll_nnz(trmf_b200_session *s);
int trmf_b200_roll_export(trmf_b200_session *s, uint64_t *row_ptr, uint32_t *col_idx, void *val_t, uint64_t *col_ptr,
                          uint32_t *row_idx, void *val);

#ifdef __cplusplus
}
#endif
#endif /* TRMF_B200_H */
