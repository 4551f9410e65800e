// trmf_b200.cu -- host driver + C ABI of the B200-native TRMF ALS solver.
//
// Built twice (-DValueType=float / double) into trmf_float32.so / trmf_float64.so,
// mirroring the reference's two libraries (corelib/Makefile:32-33).  The ALS loop
// and the TRON/CG control flow follow the reference's *semantics*
// (trmf.cpp:599-694, rf_tron.h:135-254,412-505); every array lives in HBM for the
// whole call and every inner update is a CUDA kernel from the .cuh files next to
// this one.  There is no CPU compute path.
#include "../../include/trmf_b200.h"
#include "common.cuh"
#include "f_update.cuh"
#include "f_update_tiled.cuh"
#include "f_update_mma.cuh"
#include "f_update_tc.cuh"
#include "f_update_mma2.cuh"
#include "complement.cuh"
#include "x_update.cuh"
#include "x_pass_fast.cuh"
#include "lag_update.cuh"
#include "dense.cuh"
#include "ingest.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>
#include <ctime>
#include <random>
#include <string>
#include <vector>

#define TRMF_B200_VERSION "trmf-b200 0.1 (sm_100a)"

// --------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------
static thread_local std::string g_last_error;
// set by c_trmf_train around its trmf_b200_create: the caller's buffers outlive the session, so the host-side packing and the
// enqueueing of the upload slabs may continue on a feeder thread after create has returned
static thread_local bool g_async_feed = false;
static thread_local bool g_on_feeder = false;     // this thread is a session's feeder: it waits politely, the caller needs a core

static int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    fprintf(stderr, "[trmf-b200 ERROR]: %s\n", buf);
    fflush(stderr);
    return 1;
}

#define CUDA_TRY(...)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (__VA_ARGS__);                                                   \
        if (e__ != cudaSuccess)                                                            \
            return fail("%s failed at %s:%d: %s", #__VA_ARGS__, __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

// --------------------------------------------------------------------------
// session
// --------------------------------------------------------------------------
struct trmf_b200_session {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // host-buffer sessions: the by-time CSR (first needed by the X-update) is uploaded on a second stream so
    // that the copy overlaps the F-update, which only reads the by-series CSC
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // host-packed ingest: the index bitmaps travel on their own stream
    cudaEvent_t csr_ready = nullptr;
    bool csr_pending = false;
    // host-buffer sessions with a device-built CSR: the CSC arrives in nnz-balanced series slabs on the copy stream;
    // the first F-update starts on slab b as soon as it has landed (slab_ev[b]), the CSR is built once all have
    std::vector<size_t> slab_j;          // slab b = series [slab_j[b], slab_j[b+1])
    std::vector<cudaEvent_t> slab_ev;
    bool slabs_pending = false;          // the F-update has not consumed the slab events yet
    // host-packed ingest inside c_trmf_train: a feeder thread packs + enqueues slab after slab while the calling thread already
    // enqueues the first F-update; slab b may be waited for once slabs_published > b
    std::thread feeder;
    std::atomic<size_t> slabs_published{0};
    std::atomic<int> feed_state{0};      // 0 = running / not used, 1 = finished, -1 = failed (feed_err)
    size_t feed_plain_from = (size_t)-1; // slabs from this one on carry plain row indices (the bitmaps could not: unsorted input)
    std::string feed_err;
    bool csr_deferred = false;           // the device transpose has not been issued yet
    uint32_t *pack_buf = nullptr;        // pinned staging of the host-packed index bitmap (back to the pool at destroy)
    uint32_t *bm_dev = nullptr;          // host-packed ingest: the bitmaps in HBM until the first consumer has expanded them
    uint32_t bm_words = 0;
    size_t pack_bytes = 0;

    size_t T = 0, n = 0, nnz = 0;
    int k = 0, L = 0, mid = 0;
    bool missing = true;          // ARR_LS_MISSING (31) vs ARR_LS_FULL (30), trmf.cpp:717
    bool sparse_storage = true;
    int dense_type = 0;           // TRMF_DENSE_ROWMAJOR / COLMAJOR when !sparse_storage

    // Y in HBM: both orientations when sparse
    uint64_t *row_ptr = nullptr, *col_ptr = nullptr;
    uint32_t *col_idx = nullptr, *row_idx = nullptr;
    V *val_t = nullptr, *val = nullptr;
    V *Yd = nullptr;
    bool own_Y = false;

    // rolling-window sessions (rolling.cuh): the whole time axis of Y stays in HBM, the solver sees the prefix
    // [0, T) of it.  R_* = the resident full-length arrays; the by-time CSR's row_ptr / col_idx serve every
    // window as they are (a prefix of a CSR is the CSR of the prefix), the by-series CSC is re-compacted per window.
    bool rolling = false;
    size_t T_cap = 0;                 // time stamps the buffers are sized for (>= every window's T)
    size_t R_nnz = 0;                 // observed entries of the resident matrix
    uint64_t *R_col_ptr = nullptr;
    uint32_t *R_row_idx = nullptr;
    V *R_val = nullptr, *R_val_t = nullptr, *R_Yd = nullptr;
    V *win_val_t = nullptr, *win_Yd = nullptr;   // per-series affine transform applied (NormalizedTransform)
    V *aff_a = nullptr, *aff_b = nullptr;        // its n scales / offsets
    uint64_t *win_cnt = nullptr;                 // n + 1 per-series counts of a window
    void *scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;

    // factors
    V *W = nullptr, *H = nullptr, *th = nullptr;
    V *W_sv = nullptr, *H_sv = nullptr, *th_sv = nullptr;   // save_factors() snapshot
    bool own_factors = false;
    std::vector<uint32_t> lags;
    uint32_t *lags_dev = nullptr;

    // X-update work space (TRON's s, r, w_new, g, d, Hd: rf_tron.h:52-54)
    V *g = nullptr, *s = nullptr, *r = nullptr, *d = nullptr, *Hd = nullptr, *wnew = nullptr;
    double *rho = nullptr;
    double *scal = nullptr, *part = nullptr;
    unsigned *ticket = nullptr;
    int *cgctl = nullptr;         // device-side CG control: go flags of the (at most 20) CG steps of an X-update
    unsigned *queue = nullptr;    // tiled kernel: [0] series counter, [1..] per-SM CTA arrival counters
    double *h_scal = nullptr;     // pinned mirror of `scal`

    // dense-mode work space
    V *YH = nullptr;              // T x k
    V *tmp_nk = nullptr;          // n x k (sparse-storage dense mode)
    double *HTH = nullptr, *WTW = nullptr, *YtW = nullptr, *Cpart = nullptr;
    size_t Cpart_elems = 0;
    double trYTY = 0.0;

    // lag update work space
    double *lag_partial = nullptr;
    int lag_chunk = 0, lag_nchunks = 0;

    // multi-GPU
    int rank = 0, world = 1;
    void *nccl_comm = nullptr;
    bool own_comm = false;
    V *part_tk = nullptr;         // this rank's partial of a T x k pass, all-reduced in place
    // X-update through per-time-stamp Grams (fp32 build, k % 4 == 0, k <= 64)
    V *Gt = nullptr, *bt = nullptr;   // T x k x k Grams of the series factor, T x k rhs
    V *Xs = nullptr;                  // column-scaled copy of the factor the mma Gram kernel reads (max(T, n) x k)
    float *invs = nullptr;            // its k inverse scales
    float *ysc = nullptr;             // power-of-two scale of the pre-split weights (or of the tcgen05 kernel's) and its inverse
    uint32_t *valh = nullptr;         // F-update of f_update_mma2.cuh: the CSC values as fp16 pairs (nnz words, indexed like val)
    size_t valh_cap = 0;
    bool valh_ok = false;             // valh / ysc describe the whole of the current val (cleared whenever Y changes)
    // complement formulation for a mostly observed Y (complement.cuh): 0 = undecided, 1 = in use, -1 = not
    int cm_state = 0;
    uint64_t *cm_ptr[2] = {nullptr, nullptr};   // missing cells per series [0] / per time stamp [1]
    uint32_t *cm_idx[2] = {nullptr, nullptr};
    float *cm_Y0 = nullptr;           // Y zero-filled at the missing cells, T x n row-major
    double *cm_yy = nullptr;          // per time stamp: sum of squares of its observed values
    double *cm_FtF = nullptr;         // k x k Gram of the gathered factor over ALL its rows
    V *cm_Xr = nullptr;               // the gathered factor as the fp16 split carries it (max(T, n) x k), see presplit_kernel
    double *cm_rhs = nullptr;         // max(T, n) x k: Y0^T W resp. Y0 H
    double *cm_cpart = nullptr;       // split-K scratch of the tall-skinny products
    size_t cm_cpart_elems = 0;
    uint32_t *cm_bm0 = nullptr;       // while the lists are built: per series, bitmap of its missing time stamps ...
    uint32_t *cm_bmT = nullptr;       // ... and per time stamp, bitmap of its missing series
    size_t cm_placed = 0;             // the series [0, cm_placed) are in cm_idx[0], cm_Y0 and cm_bmT
    // The formulation counts every cell of a row once, so it may only be used when no cell occurs twice in Y (duplicates are
    // legal input: the reference treats them as separate observations, rf_util.py:100-119).  Known where Y comes in: host index
    // lists are scanned for strict ascent (by the bitmap packer, or host_lists_strict), bitmaps and dense arrays cannot hold
    // duplicates, device-resident arrays are canonical by the API's contract.
    bool idx_strict = false;
    bool cm_ready = false;            // by-time list and cm_yy are built
    double *frow = nullptr;           // per-time-stamp loss values of the fused Gram + gradient kernel (T)
    double *sys = nullptr;            // F-update: assembled fp64 systems of one batch of series, solved by chol_solve_kernel
    size_t sys_batch = 0;             // series per batch
    int gram_state = 0;               // 0 = not decided, 1 = enabled, -1 = disabled
    int prev_cg = -1;                 // CG steps of the previous x_update of this session (-1: none yet)
    bool gram_now = false;            // this x_update goes through the Grams
    unsigned long long collectives = 0;

    double lambdaI = 0.1, lambdaAR = 0.1, lambdaLag = 0.1;

    // stats
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr, ev5 = nullptr, ev6 = nullptr, ev7 = nullptr, ev8 = nullptr;
    bool cm_timed = false;
    float ms_cmg = 0, ms_cmp = 0;        // complement F-update: Gram over the missing cells / tall-skinny product (last whole-Y pass)
    double st_cg = 0, st_acc = 0, st_f = 0, st_fnew = 0, st_gnorm = 0, st_prered = 0, st_actred = 0;
    double ms_f = 0, ms_x = 0, ms_lag = 0, ms_fk = 0, ms_xg = 0;
    unsigned long long launches = 0;
    double st_delta = 0, st_rnorm = 0;
    bool xg_timed = false;
};
typedef trmf_b200_session S;
// time stamps to size lazily allocated T-proportional buffers for (rolling sessions grow T window by window)
static inline size_t Tcap(const S *s) { return std::max(s->T, s->T_cap); }

// multi-GPU helpers, defined in extras.cuh
static int dist_allreduce_v(S *s, V *buf, size_t count);
static int dist_allreduce_f64(S *s, double *buf, size_t count);
static void dist_teardown(S *s);

// `kern` may be a parenthesised template-id; it is bound to a pointer first.
#define LAUNCH(s, kern, grid, block, smem, ...)                                           \
    do {                                                                                   \
        auto kfn__ = kern;                                                                 \
        kfn__<<<(grid), (block), (smem), (s)->stream>>>(__VA_ARGS__);                      \
        (s)->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess)                                                            \
            return fail("launch of %s failed at %s:%d: %s", #kern, __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

static inline unsigned ew_grid(const S *s, size_t n, int threads = 256) {
    size_t b = (n + threads - 1) / threads;
    size_t cap = (size_t)s->num_sms * 8;
    return (unsigned)std::max<size_t>(1, std::min(b, cap));
}

// Device memory comes from the device's default stream-ordered pool with the release threshold
// lifted, so the buffers of one c_trmf_train call are recycled by the next one instead of going
// through cudaMalloc/cudaFree (measured: 50-640 ms per call for the 1.5 GB of a C2 session).
static int pool_setup(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return 0;
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done[device] = true;
    return 0;
}

template <typename T>
static int dev_alloc_on(cudaStream_t st, T **p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    CUDA_TRY(cudaMallocAsync((void **)p, count * sizeof(T), st));
    return 0;
}
#define dev_alloc(ptr, count) dev_alloc_on(s->stream, ptr, count)
#define dev_free(ptr) do { if (ptr) cudaFreeAsync((void *)(ptr), s->stream); } while (0)

// pinned host scratch (the scalar mirror) is recycled across sessions as well
static std::vector<double *> g_pinned_free;
static std::mutex g_pinned_mu;
static int pinned_get(double **p) {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (!g_pinned_free.empty()) { *p = g_pinned_free.back(); g_pinned_free.pop_back(); return 0; }
    }
    CUDA_TRY(cudaMallocHost((void **)p, SC_COUNT * sizeof(double)));
    return 0;
}
static void pinned_put(double *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned_free.push_back(p);
}

// Host side of the packed ingest: row indices of a by-series CSC as one bitmap per series (TRMF_SPARSE_BITMAP layout), written
// into a pinned staging buffer by all host cores.  A 10 %-missing panel then sends 1/8 byte instead of 4 per observed cell for its
// indices: the scan of the caller's row_idx array runs at host-memory speed while the values are already crossing PCIe.
static std::vector<std::pair<uint32_t *, size_t>> g_bitmap_free;
static uint32_t *bitmap_buf_get(size_t bytes, size_t *got) {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        for (size_t i = 0; i < g_bitmap_free.size(); ++i)
            if (g_bitmap_free[i].second >= bytes) {
                uint32_t *p = g_bitmap_free[i].first;
                *got = g_bitmap_free[i].second;
                g_bitmap_free.erase(g_bitmap_free.begin() + i);
                return p;
            }
    }
    uint32_t *p = nullptr;
    if (cudaMallocHost((void **)&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *got = bytes;
    return p;
}
static void bitmap_buf_put(uint32_t *p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_bitmap_free.push_back(std::make_pair(p, bytes));
}
static void pack_series_scalar(const uint32_t *r, size_t cnt, uint32_t *w) {
    for (size_t i = 0; i < cnt; ++i) w[r[i] >> 5] |= 1u << (r[i] & 31);
}
#if defined(__x86_64__)
#include <immintrin.h>
// eight sorted indices at a time: they fall into one bitmap word (75 % of the time at 10 % missing) or two adjacent ones; the
// lane bits are OR-reduced in registers and cost one read-modify-write per word instead of one per entry (0.65 vs 1.35 ns/entry)
__attribute__((target("avx2"))) static inline uint32_t pack_hor_avx2(__m256i a) {
    __m128i x = _mm_or_si128(_mm256_castsi256_si128(a), _mm256_extracti128_si256(a, 1));
    x = _mm_or_si128(x, _mm_shuffle_epi32(x, 0x4e));
    x = _mm_or_si128(x, _mm_shuffle_epi32(x, 0xb1));
    return (uint32_t)_mm_cvtsi128_si32(x);
}
__attribute__((target("avx2"))) static void pack_series_avx2(const uint32_t *r, size_t cnt, uint32_t *w) {
    size_t i = 0;
    const __m256i one = _mm256_set1_epi32(1), m31 = _mm256_set1_epi32(31);
    for (; i + 8 <= cnt; i += 8) {
        const uint32_t w0 = r[i] >> 5, w7 = r[i + 7] >> 5;
        const __m256i v = _mm256_loadu_si256((const __m256i *)(r + i));
        const __m256i bits = _mm256_sllv_epi32(one, _mm256_and_si256(v, m31));
        if (w0 == w7) {
            w[w0] |= pack_hor_avx2(bits);
        } else if (w7 == w0 + 1) {
            const __m256i first = _mm256_cmpeq_epi32(_mm256_srli_epi32(v, 5), _mm256_set1_epi32((int)w0));
            w[w0] |= pack_hor_avx2(_mm256_and_si256(bits, first));
            w[w7] |= pack_hor_avx2(_mm256_andnot_si256(first, bits));
        } else {
            pack_series_scalar(r + i, 8, w);
        }
    }
    pack_series_scalar(r + i, cnt - i, w);
}
#endif
// TRMF_B200_TRACE: host wall-clock marks inside a host-buffer call (stderr), relative to the first mark
static void trace_pt(const char *what) {
    static const bool on = getenv("TRMF_B200_TRACE") != nullptr;
    if (!on) return;
    static double t0 = 0;
    struct timespec tw;
    clock_gettime(CLOCK_MONOTONIC, &tw);
    const double t = tw.tv_sec * 1e3 + tw.tv_nsec * 1e-6;
    if (!strcmp(what, "begin")) t0 = t;
    fprintf(stderr, "[trmf-b200 trace] %8.3f ms  %s\n", t - t0, what);
}
// ... and the device side of the same call: timing events recorded on the streams at the points named below, printed (relative to the
// first one) by trace_dev_dump() once the call has synchronised.  Nothing is recorded unless TRMF_B200_TRACE is set.
static std::vector<std::pair<std::string, cudaEvent_t>> g_trace_ev;
static std::mutex g_trace_mu;
static void trace_dev(cudaStream_t st, const char *fmt, ...) {
    static const bool on = getenv("TRMF_B200_TRACE") != nullptr;
    if (!on) return;
    char buf[160];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    std::lock_guard<std::mutex> lk(g_trace_mu);
    g_trace_ev.emplace_back(buf, ev);
}
static void trace_dev_dump() {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    if (g_trace_ev.empty()) return;
    for (auto &pe : g_trace_ev) {
        float ms = 0;
        cudaEventSynchronize(pe.second);
        cudaEventElapsedTime(&ms, g_trace_ev.front().second, pe.second);
        fprintf(stderr, "[trmf-b200 trace] device %8.3f ms  %s\n", ms, pe.first.c_str());
    }
    for (auto &pe : g_trace_ev) cudaEventDestroy(pe.second);
    g_trace_ev.clear();
}
// Branch-free packing at any density: one output word (32 rows) at a time.  The entries of word w are a prefix of what is left of
// the (ascending) list, at most 32 of them: four 8-lane compares find and count them, a variable shift turns them into bits.
// No data-dependent branch, so 10 % randomly missing rows cost nothing in mispredictions (a scan for the gaps between runs of
// consecutive rows paid ~1 per 8 entries and ran at half the speed: profiles/r02_packbench.txt).  0.60 ns per entry and thread on
// the B200 box's host, 0.71 for round 2's first packer (OR of 8 entries per step, pack_series_avx2).
// The caller has checked that the indices ascend strictly; `r` must be readable up to r[cnt + 31] (see pack_one_series).
#if defined(__x86_64__)
__attribute__((target("avx2,popcnt"))) static void pack_series_words(const uint32_t *r, size_t cnt, uint32_t *w, uint32_t words) {
    const __m256i one = _mm256_set1_epi32(1), m31 = _mm256_set1_epi32(31);
    const __m256i lane = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
    size_t p = 0;
    for (uint32_t wi = 0; wi < words && p < cnt; ++wi) {
        const __m256i wv = _mm256_set1_epi32((int)wi);
        const long long left = (long long)(cnt - p);
        const __m256i leftv = _mm256_set1_epi32((int)(left > 64 ? 64 : left));
        __m256i acc = _mm256_setzero_si256();
        unsigned taken = 0;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const __m256i idx = _mm256_loadu_si256((const __m256i *)(r + p + 8 * v));
            const __m256i in = _mm256_and_si256(_mm256_cmpeq_epi32(_mm256_srli_epi32(idx, 5), wv),
                                                _mm256_cmpgt_epi32(leftv, _mm256_add_epi32(lane, _mm256_set1_epi32(8 * v))));
            acc = _mm256_or_si256(acc, _mm256_and_si256(_mm256_sllv_epi32(one, _mm256_and_si256(idx, m31)), in));
            taken += (unsigned)__builtin_popcount((unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(in)));
        }
        w[wi] = pack_hor_avx2(acc);
        p += taken;
    }
}
// The same for NS series in lock step, with the strict-ascent check folded in.  One series alone is bound by the latency of its
// dependency chain (position -> load -> compare -> movemask -> popcount -> position: ~50 clk per word) and, with the ascent check
// as a pass of its own, reads the index array twice -- a packer thread is then bound by what one core can stream (~10 GB/s on
// the build container's host: 0.87 ns per entry for the two passes).  NS independent chains fill the pipeline and the one pass
// halves the bytes.  Every pair (i, i + 1) of a series lies in some window [p, p + 32] the loop visits, so comparing each window
// position with its successor (signed: indices below 2^31, see the caller) covers them all.  Every series' words are all written
// (no memset needed); r[s] must be readable up to r[s][cnt[s] + 32].  Returns false when some series is not strictly ascending.
template <int NS>
__attribute__((target("avx2,popcnt"))) static bool pack_series_words_multi(const uint32_t *const *r, const size_t *cnt, uint32_t *const *w,
                                                                           uint32_t words) {
    const __m256i one = _mm256_set1_epi32(1), m31 = _mm256_set1_epi32(31);
    const __m256i lane = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
    size_t p[NS];
    __m256i badv = _mm256_setzero_si256();
    for (int q = 0; q < NS; ++q) p[q] = 0;
    for (uint32_t wi = 0; wi < words; ++wi) {
        const __m256i wv = _mm256_set1_epi32((int)wi);
#pragma GCC unroll 8
        for (int q = 0; q < NS; ++q) {
            const long long left = (long long)(cnt[q] - p[q]);
            const __m256i leftv = _mm256_set1_epi32((int)(left > 64 ? 64 : left));
            __m256i acc = _mm256_setzero_si256();
            unsigned taken = 0;
#pragma GCC unroll 4
            for (int v = 0; v < 4; ++v) {
                const __m256i pos = _mm256_add_epi32(lane, _mm256_set1_epi32(8 * v));
                const __m256i idx = _mm256_loadu_si256((const __m256i *)(r[q] + p[q] + 8 * v));
                const __m256i nxt = _mm256_loadu_si256((const __m256i *)(r[q] + p[q] + 8 * v + 1));
                const __m256i in = _mm256_and_si256(_mm256_cmpeq_epi32(_mm256_srli_epi32(idx, 5), wv), _mm256_cmpgt_epi32(leftv, pos));
                // pair (pos, pos + 1) exists iff pos + 1 < left; it is in order iff nxt > idx
                const __m256i pair = _mm256_cmpgt_epi32(leftv, _mm256_add_epi32(pos, one));
                badv = _mm256_or_si256(badv, _mm256_andnot_si256(_mm256_cmpgt_epi32(nxt, idx), pair));
                acc = _mm256_or_si256(acc, _mm256_and_si256(_mm256_sllv_epi32(one, _mm256_and_si256(idx, m31)), in));
                taken += (unsigned)__builtin_popcount((unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(in)));
            }
            w[q][wi] = pack_hor_avx2(acc);
            p[q] += taken;
        }
    }
    // Windows [p, p + 32] tile the whole list only if every entry was consumed (an entry beyond the last word, or out of order,
    // stalls its series): both conditions together say "strictly ascending and inside the bitmap".
    bool all = true;
    for (int q = 0; q < NS; ++q) all &= p[q] == cnt[q];
    return all && _mm256_testz_si256(badv, badv) != 0;
}
// strictly ascending?  (signed compares: the caller guarantees indices below 2^31)
__attribute__((target("avx2"))) static bool ascending_avx2(const uint32_t *r, size_t cnt) {
    __m256i ok = _mm256_set1_epi32(-1);
    size_t i = 0;
    for (; i + 9 <= cnt; i += 8)
        ok = _mm256_and_si256(ok, _mm256_cmpgt_epi32(_mm256_loadu_si256((const __m256i *)(r + i + 1)), _mm256_loadu_si256((const __m256i *)(r + i))));
    unsigned good = _mm256_movemask_ps(_mm256_castsi256_ps(ok)) == 0xff;
    for (; i + 1 < cnt; ++i) good &= (unsigned)(r[i] < r[i + 1]);
    return good != 0;
}
#endif
// Bitmap of one series (words = ceil(rows / 32), zeroed here).  Returns false when the row indices are not strictly ascending
// or reach past `rows` (unsorted arrays or duplicate entries are legal input for the reference's core, which never looks at
// their order): a bitmap cannot carry those, the caller uploads plain indices instead.
// `tail_ok`: r[cnt .. cnt + 31] may be read (the series is not among the last of the index array).  algo: 0 = word-wise,
// 2 = OR per 8 entries (TRMF_B200_PACK_ALGO=or8).
static bool pack_one_series(const uint32_t *r, size_t cnt, uint32_t *w, uint32_t words, uint64_t rows, bool avx2, bool tail_ok, int algo) {
    memset(w, 0, (size_t)words * sizeof(uint32_t));
    if (cnt == 0) return true;
    if ((uint64_t)r[cnt - 1] >= rows) return false;
#if defined(__x86_64__)
    if (avx2 && algo == 0 && tail_ok && rows < (1ull << 31)) {
        if (!ascending_avx2(r, cnt)) return false;
        pack_series_words(r, cnt, w, words);
        return true;
    }
#endif
    unsigned ok = 1;
    for (size_t i = 0; i + 1 < cnt; ++i) ok &= (unsigned)(r[i] < r[i + 1]);
    if (!ok) return false;
#if defined(__x86_64__)
    if (avx2) { pack_series_avx2(r, cnt, w); return true; }
#endif
    pack_series_scalar(r, cnt, w);
    return true;
}
// every index list of a compressed sparse half strictly ascending?  (host arrays; up to 8 threads for large inputs)
static bool host_lists_strict(const uint64_t *ptr, const uint32_t *idx, uint64_t nlists) {
    if (!ptr || !idx) return false;
    auto range_ok = [&](uint64_t l0, uint64_t l1) {
        for (uint64_t l = l0; l < l1; ++l) {
            const uint32_t *r = idx + ptr[l];
            const size_t cnt = (size_t)(ptr[l + 1] - ptr[l]);
#if defined(__x86_64__)
            if (__builtin_cpu_supports("avx2")) { if (!ascending_avx2(r, cnt)) return false; continue; }
#endif
            for (size_t i = 0; i + 1 < cnt; ++i) if (!(r[i] < r[i + 1])) return false;
        }
        return true;
    };
    const uint64_t nnz = ptr[nlists];
    unsigned nt = nnz >= (1u << 22) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
    if (nt <= 1) return range_ok(0, nlists);
    std::vector<char> ok(nt, 1);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { ok[t] = range_ok(nlists * t / nt, nlists * (t + 1) / nt) ? 1 : 0; });
    for (auto &x : th) x.join();
    for (char c : ok) if (!c) return false;
    return true;
}
// The index bitmaps of all series, packed by the host cores IN SLAB ORDER while the calling thread waits for slab after slab and
// hands each to `slab_done(b)` (which enqueues its copies): the first F-update launch only waits for the first slab's bitmap and
// values, and the copy engine never idles behind the packing.  slab_j = series bounds of the slabs.  Returns false when some
// series cannot be carried by a bitmap (slab_done may already have been called for earlier slabs) or slab_done failed (*rc = 1).
// threads a host-buffer call may use for packing (packers + the coordinating thread)
static unsigned pack_thread_budget() {
    unsigned nt = std::thread::hardware_concurrency();
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) nt /= (unsigned)std::max(1, atoi(e));   // one process per GPU: share the host cores
    if (const char *e = getenv("TRMF_B200_PACK_THREADS")) nt = (unsigned)std::max(1, atoi(e));
    return std::max(2u, std::min(nt, 32u));
}
template <class F>
static bool pack_bitmap_slabs(const uint64_t *col_ptr, const uint32_t *row_idx, const std::vector<size_t> &slab_j, uint32_t words,
                              uint64_t rows, uint32_t *out, F slab_done, int *rc) {
    unsigned nt = pack_thread_budget();        // (nt - 1 packers + the calling thread)
    // on a feeder thread the session's calling thread is busy enqueueing kernels at the same time: leave it a core
    if (g_on_feeder && !getenv("TRMF_B200_PACK_THREADS") && nt > 4) nt -= 1;
    const bool polite = g_on_feeder;
    bool avx2 = false;
#if defined(__x86_64__)
    avx2 = __builtin_cpu_supports("avx2") && !getenv("TRMF_B200_PACK_SCALAR");
#endif
    int algo = 0;
    if (const char *e = getenv("TRMF_B200_PACK_ALGO")) algo = !strcmp(e, "or8") ? 2 : 0;
    const uint64_t nnz_all = col_ptr[slab_j.back()];
    const size_t nsl = slab_j.size() - 1, chunk = 16;
    struct Chunk { size_t j0, j1, b; };
    std::vector<Chunk> chunks;
    for (size_t b = 0; b < nsl; ++b)
        for (size_t j = slab_j[b]; j < slab_j[b + 1]; j += chunk) chunks.push_back({j, std::min(slab_j[b + 1], j + chunk), b});
    std::unique_ptr<std::atomic<size_t>[]> done(new std::atomic<size_t>[nsl]);
    for (size_t b = 0; b < nsl; ++b) done[b].store(0);
    std::atomic<size_t> next(0);
    std::atomic<bool> bad(false);
    auto work = [&]() {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= chunks.size() || bad.load(std::memory_order_relaxed)) return;
            size_t j = chunks[c].j0;
#if defined(__x86_64__)
            // four series at a time (pack_series_words_multi) while all four may be read 32 entries past their end
            while (avx2 && algo == 0 && rows < (1ull << 31) && j + 4 <= chunks[c].j1 && col_ptr[j + 4] + 33 <= nnz_all) {
                const uint32_t *r4[4];
                size_t c4[4];
                uint32_t *w4[4];
                for (int q = 0; q < 4; ++q) {
                    r4[q] = row_idx + col_ptr[j + q];
                    c4[q] = (size_t)(col_ptr[j + q + 1] - col_ptr[j + q]);
                    w4[q] = out + (j + q) * (size_t)words;
                    if (c4[q] && (uint64_t)r4[q][c4[q] - 1] >= rows) {
                        bad.store(true);
                        return;
                    }
                }
                if (!pack_series_words_multi<4>(r4, c4, w4, words)) {      // some series not strictly ascending
                    bad.store(true);
                    return;
                }
                j += 4;
            }
#endif
            for (; j < chunks[c].j1; ++j)
                if (!pack_one_series(row_idx + col_ptr[j], (size_t)(col_ptr[j + 1] - col_ptr[j]), out + j * (size_t)words, words, rows, avx2,
                                     col_ptr[j + 1] + 32 <= nnz_all, algo)) {
                    bad.store(true);
                    return;
                }
            done[chunks[c].b].fetch_add(chunks[c].j1 - chunks[c].j0, std::memory_order_release);
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
    *rc = 0;
    for (size_t b = 0; b < nsl && !*rc; ++b) {
        const size_t want = slab_j[b + 1] - slab_j[b];
        while (done[b].load(std::memory_order_acquire) < want && !bad.load(std::memory_order_relaxed)) {
            if (polite) { std::this_thread::sleep_for(std::chrono::microseconds(10)); continue; }
#if defined(__x86_64__)
            _mm_pause();
#endif
        }
        if (bad.load()) break;
        if (slab_done(b)) { *rc = 1; bad.store(true); }
    }
    for (auto &t : th) t.join();
    return !bad.load();
}

// --------------------------------------------------------------------------
// creation / destruction
// --------------------------------------------------------------------------
// lag update: chunk the window [mid, T) so that (chunk + mid) fp64 values fit in shared memory; depends on T, so a
// rolling session re-plans per window (same formula: a window trains exactly like a fresh session of that length)
static int lag_plan(S *s) {
    const size_t budget = 160 * 1024 / sizeof(double);
    if ((size_t)s->mid + 256 > budget)
        return fail("max lag %d too large for the lag_val kernel's shared-memory staging (limit %zu)", s->mid, budget - 256);
    size_t chunk = std::min<size_t>(4096, budget - s->mid);
    size_t win = s->T > (size_t)s->mid ? s->T - s->mid : 0;
    // enough CTAs to fill the machine: k * nchunks >= 2 * SMs when the window allows
    size_t want = std::max<size_t>(1, (2 * (size_t)s->num_sms + s->k - 1) / s->k);
    size_t c2 = std::max<size_t>(256, (win + want - 1) / want);
    chunk = std::min(chunk, c2);
    s->lag_chunk = (int)chunk;
    s->lag_nchunks = (int)std::max<size_t>(1, (win + chunk - 1) / chunk);
    const size_t L1 = s->L + 1, npairs = L1 * (L1 + 1) / 2;
    dev_free(s->lag_partial);
    s->lag_partial = nullptr;
    return dev_alloc(&s->lag_partial, (size_t)s->k * s->lag_nchunks * npairs);
}

static int session_common_init(S *s) {
    CUDA_TRY(cudaSetDevice(s->device));
    {   // (cudaGetDeviceProperties costs milliseconds: two attribute queries, cached per device)
        static int sms_cache[64] = {0}, major_cache[64] = {0};
        const int d = s->device;
        if (d < 0 || d >= 64) return fail("device index %d out of range", d);
        if (sms_cache[d] == 0) {
            int sms = 0, major = 0;
            CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
            CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d));
            major_cache[d] = major;
            sms_cache[d] = sms;
        }
        s->num_sms = sms_cache[d];
        if (major_cache[d] < 10)
            return fail("device %d (compute capability %d.x) is not a Blackwell (sm_100a) GPU; this library has no other code path",
                        d, major_cache[d]);
    }
    if (pool_setup(s->device)) return 1;
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->own_stream = true;
    const size_t tk = s->T * (size_t)s->k;
    if (dev_alloc(&s->g, tk) || dev_alloc(&s->s, tk) || dev_alloc(&s->r, tk) || dev_alloc(&s->d, tk) ||
        dev_alloc(&s->Hd, tk) || dev_alloc(&s->wnew, tk) || dev_alloc(&s->rho, tk))
        return 1;
    if (dev_alloc(&s->scal, SC_COUNT) || dev_alloc(&s->part, 8192) || dev_alloc(&s->ticket, 4) ||
        dev_alloc(&s->queue, 1024))
        return 1;
    CUDA_TRY(cudaMemsetAsync(s->scal, 0, SC_COUNT * sizeof(double), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->ticket, 0, 4 * sizeof(unsigned), s->stream));
    if (pinned_get(&s->h_scal)) return 1;
    if (dev_alloc(&s->lags_dev, s->lags.size())) return 1;
    CUDA_TRY(cudaMemcpyAsync(s->lags_dev, s->lags.data(), s->lags.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
    if (lag_plan(s)) return 1;
    if (!s->missing) {
        const size_t kk = (size_t)s->k * s->k;
        if (dev_alloc(&s->YH, tk) || dev_alloc(&s->HTH, kk) || dev_alloc(&s->WTW, kk) ||
            dev_alloc(&s->YtW, s->n * (size_t)s->k))
            return 1;
        if (s->sparse_storage && dev_alloc(&s->tmp_nk, s->n * (size_t)s->k)) return 1;
        // split-K scratch: sized for the largest product (see gemm())
        s->Cpart_elems = std::max(tk, s->n * (size_t)s->k) * 64 + kk * 1024;
        size_t cap = ((size_t)1 << 31) / sizeof(double);   // never more than 2 GB
        s->Cpart_elems = std::min(s->Cpart_elems, std::max(cap, std::max(tk, s->n * (size_t)s->k)));
        if (dev_alloc(&s->Cpart, s->Cpart_elems)) return 1;
    }
    CUDA_TRY(cudaEventCreate(&s->ev0));
    CUDA_TRY(cudaEventCreate(&s->ev1));
    CUDA_TRY(cudaEventCreate(&s->ev2));
    CUDA_TRY(cudaEventCreate(&s->ev3));
    CUDA_TRY(cudaEventCreate(&s->ev4));
    CUDA_TRY(cudaEventCreate(&s->ev5));
    CUDA_TRY(cudaEventCreate(&s->ev6));
    CUDA_TRY(cudaEventCreate(&s->ev7));
    CUDA_TRY(cudaEventCreate(&s->ev8));
    return 0;
}

static int check_lags(S *s, const uint32_t *lag_set, uint32_t lag_size) {
    if (lag_set == nullptr || lag_size == 0) return fail("lag_set must be a non-empty sorted uint32 array");
    if (lag_size > TRMF_MAX_LAGS) return fail("lag_set larger than %d is not supported", TRMF_MAX_LAGS);
    s->lags.assign(lag_set, lag_set + lag_size);
    s->L = (int)lag_size;
    s->mid = (int)s->lags.back();   // trmf.cpp:79 -- "supposed to be the max index in lag_set"
    for (uint32_t l : s->lags)
        if ((int)l > s->mid) return fail("lag_set must be sorted ascending (last element is taken as the max lag)");
    return 0;
}

extern "C" void trmf_b200_destroy(S *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    if (s->aux_stream) cudaStreamSynchronize(s->aux_stream);
    if (s->feeder.joinable()) s->feeder.join();
    if (s->stream) cudaStreamSynchronize(s->stream);
    dist_teardown(s);
    dev_free(s->part_tk);
    dev_free(s->Gt); dev_free(s->bt); dev_free(s->Xs); dev_free(s->invs); dev_free(s->ysc); dev_free(s->frow); dev_free(s->sys); dev_free(s->valh); dev_free(s->bm_dev);
    dev_free(s->cm_ptr[0]); dev_free(s->cm_ptr[1]); dev_free(s->cm_idx[0]); dev_free(s->cm_idx[1]); dev_free(s->cm_Y0); dev_free(s->cm_yy);
    dev_free(s->cm_FtF); dev_free(s->cm_rhs); dev_free(s->cm_cpart); dev_free(s->cm_Xr); dev_free(s->cm_bm0); dev_free(s->cm_bmT);
    if (s->own_Y) {
        dev_free(s->row_ptr); dev_free(s->col_ptr); dev_free(s->col_idx); dev_free(s->row_idx);
        dev_free(s->val_t); dev_free(s->val); dev_free(s->Yd);
    }
    if (s->rolling) {   // (val_t / Yd alias R_val_t / R_Yd or the transformed window copies: freed through those)
        dev_free(s->row_ptr); dev_free(s->col_idx); dev_free(s->R_val_t); dev_free(s->win_val_t);
        dev_free(s->R_col_ptr); dev_free(s->R_row_idx); dev_free(s->R_val);
        dev_free(s->col_ptr); dev_free(s->row_idx); dev_free(s->val);
        dev_free(s->R_Yd); dev_free(s->win_Yd); dev_free(s->aff_a); dev_free(s->aff_b); dev_free(s->win_cnt);
        dev_free(s->scan_tmp);
    }
    if (s->own_factors) { dev_free(s->W); dev_free(s->H); dev_free(s->th); }
    dev_free(s->lags_dev);
    dev_free(s->W_sv); dev_free(s->H_sv); dev_free(s->th_sv);
    dev_free(s->g); dev_free(s->s); dev_free(s->r); dev_free(s->d); dev_free(s->Hd); dev_free(s->wnew);
    dev_free(s->rho); dev_free(s->scal); dev_free(s->part); dev_free(s->ticket); dev_free(s->queue); dev_free(s->cgctl);
    dev_free(s->YH); dev_free(s->tmp_nk); dev_free(s->HTH); dev_free(s->WTW); dev_free(s->YtW); dev_free(s->Cpart);
    dev_free(s->lag_partial);
    pinned_put(s->h_scal);
    bitmap_buf_put(s->pack_buf, s->pack_bytes);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev2) cudaEventDestroy(s->ev2);
    if (s->ev3) cudaEventDestroy(s->ev3);
    if (s->ev4) cudaEventDestroy(s->ev4);
    if (s->ev5) cudaEventDestroy(s->ev5);
    if (s->ev6) cudaEventDestroy(s->ev6);
    if (s->ev7) cudaEventDestroy(s->ev7);
    if (s->ev8) cudaEventDestroy(s->ev8);
    if (s->stream) cudaStreamSynchronize(s->stream);   // the frees above are ordered on the stream
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
    if (s->csr_ready) cudaEventDestroy(s->csr_ready);
    for (cudaEvent_t e : s->slab_ev) cudaEventDestroy(e);
    delete s;
}

template <typename T>
static int h2d_new(S *s, T **dst, const void *src, size_t count) {
    if (dev_alloc(dst, count)) return 1;
    if (count) CUDA_TRY(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    return 0;
}

static int create_host_impl(S *s, const PyMatrix *Y, const uint32_t *lag_set, uint32_t lag_size,
                            const PyMatrix *W, const PyMatrix *H, const PyMatrix *lag_val, int missing, int device) {
    g_last_error.clear();
    s->device = device;
    s->T = Y->rows;
    s->n = Y->cols;
    s->k = (int)W->cols;
    s->missing = missing != 0;
    if (s->k < 1 || s->k > 128) return fail("rank k = %d outside the supported range 1..128", s->k);
    if (check_lags(s, lag_set, lag_size)) return 1;
    const bool bitmap = Y->type == TRMF_SPARSE_BITMAP;
    s->idx_strict = bitmap;      // (plain index lists: settled below, where the upload path is chosen)
    if (Y->type == TRMF_SPARSE || bitmap) {
        s->sparse_storage = true;
        s->nnz = Y->nnz;
        if (bitmap && (!Y->col_ptr || !Y->row_idx || (s->nnz && !Y->val)))
            return fail("TRMF_SPARSE_BITMAP needs col_ptr, the bitmap words in row_idx and val");
        if (bitmap && s->T >= (1ull << 32)) return fail("TRMF_SPARSE_BITMAP: rows must fit uint32");
    } else if (Y->type == TRMF_DENSE_ROWMAJOR || Y->type == TRMF_DENSE_COLMAJOR) {
        if (s->missing) return fail("missing != 0 requires a sparse Y (the reference asserts in get_sparse(), trmf.cpp:229)");
        s->sparse_storage = false;
        s->dense_type = Y->type;
        s->nnz = s->T * s->n;
    } else {
        return fail("unsupported PyMatrix type %d for Y", Y->type);
    }
    trace_pt("create: begin common init");
    if (session_common_init(s)) return 1;
    trace_pt("create: common init done");
    // factors first: the H2D engine serves copies in submission order, and the F-update must not queue
    // behind the side-stream upload of the by-time CSR
    s->own_factors = true;
    if (h2d_new(s, &s->W, W->val, s->T * (size_t)s->k) || h2d_new(s, &s->H, H->val, s->n * (size_t)s->k) ||
        h2d_new(s, &s->th, lag_val->val, (size_t)s->L * s->k))
        return 1;
    s->own_Y = true;
    if (s->sparse_storage) {
        const bool have_host_csr = !bitmap && Y->row_ptr && (s->nnz == 0 || (Y->col_idx && Y->val_t));
        const bool have_host_csc = Y->col_ptr && (s->nnz == 0 || (Y->row_idx && Y->val));
        if (!have_host_csc && !have_host_csr) return fail("sparse Y carries neither a complete CSR nor a complete CSC half");
        // would the complement formulation be considered at all (cm_on)?  Only then is it worth knowing whether the lists are strict
        const bool cm_candidate = sizeof(V) == 4 && s->missing && (double)s->nnz >= 0.7 * (double)s->T * (double)s->n;
        if (!have_host_csc) {
            if (cm_candidate) s->idx_strict = host_lists_strict(Y->row_ptr, Y->col_idx, s->T);
            // CSR-only PyMatrix (trmf.rf_util.PyMatrix(..., twin=False) of a csr_matrix): upload it and derive the
            // by-series CSC on the device -- the same stable transpose with the roles of rows and columns swapped
            if (h2d_new(s, &s->row_ptr, Y->row_ptr, s->T + 1) || h2d_new(s, &s->col_idx, Y->col_idx, s->nnz) ||
                h2d_new(s, &s->val_t, Y->val_t, s->nnz))
                return 1;
            if (dev_alloc(&s->col_ptr, s->n + 1) || dev_alloc(&s->row_idx, s->nnz) || dev_alloc(&s->val, s->nnz)) return 1;
            CUDA_TRY(csr_from_csc_device<V>(s->stream, s->num_sms, s->n, s->T, s->nnz, s->row_ptr, s->col_idx, s->val_t, s->col_ptr,
                                            s->row_idx, s->val));
            s->launches += 5;
            return 0;
        }
        const bool device_csr = !have_host_csr || !getenv("TRMF_B200_HOST_CSR");
        const bool slabs = device_csr && s->nnz >= (1u << 22) && s->n >= 16 && !getenv("TRMF_B200_NO_SLAB_UPLOAD");
        if (h2d_new(s, &s->col_ptr, Y->col_ptr, s->n + 1)) return 1;
        if (dev_alloc(&s->row_ptr, s->T + 1) || dev_alloc(&s->col_idx, s->nnz) || dev_alloc(&s->val_t, s->nnz)) return 1;
        // Caller sent plain row indices of a mostly-observed matrix: pack them into bitmaps on the host cores (see pack_bitmap_host)
        // instead of pushing 4 bytes per entry over PCIe.  Worth it when the bitmaps are at most 1/4 of the index bytes.
        const uint32_t bm_words = (uint32_t)((s->T + 31) / 32);
        // ... and when there are host cores to pack with: a packer does 0.65 ns per entry, the 4 bytes it saves cross PCIe in
        // 0.08-0.2 ns -- with fewer than three packers (8 ranks sharing 16 cores) the plain indices arrive sooner.  Measured per
        // rank at C2 (ms per host-buffer session): N = 1 (15 packers) 10.4, N = 2 (7) 11.9, N = 4 (3) 20.7 against 33.6 plain.
        const bool host_pack = !bitmap && slabs && s->T < (1ull << 32) && (size_t)s->n * bm_words * 4 <= s->nnz &&
                               !getenv("TRMF_B200_NO_HOST_PACK") && (pack_thread_budget() >= 4 || getenv("TRMF_B200_PACK_THREADS"));
        uint32_t *bm_dev = nullptr;
        if (bitmap) {
            // the row indices travel as one bitmap per series (cols x ceil(T/32) words) and are expanded here, on the session
            // stream, while the values are still crossing PCIe on the copy stream
            const uint32_t words = (uint32_t)((s->T + 31) / 32);
            if (h2d_new(s, &bm_dev, Y->row_idx, s->n * (size_t)words)) return 1;
            if (dev_alloc(&s->row_idx, s->nnz)) return 1;
            if (s->nnz) {
                bitmap_expand_kernel<<<(unsigned)(s->num_sms * 8), 256, 0, s->stream>>>(s->col_ptr, bm_dev, s->n, s->T, words, s->row_idx);
                s->launches++;
                CUDA_TRY(cudaGetLastError());
            }
            dev_free(bm_dev);
        }
        // plain index lists that the packer will not look at: scan them here (the packer's own check covers the host-packed path:
        // a declined series flips the session to the walk in cm_f_update)
        if (!bitmap && cm_candidate) s->idx_strict = host_pack ? true : host_lists_strict(Y->col_ptr, Y->row_idx, s->n);
        if (!slabs) {
            if (!bitmap && h2d_new(s, &s->row_idx, Y->row_idx, s->nnz)) return 1;
            if (h2d_new(s, &s->val, Y->val, s->nnz)) return 1;
        } else {
            // CSC in ~8 nnz-balanced series slabs on the copy stream: the F-update of slab b overlaps the upload of b+1...
            const bool idx_by_bitmap = bitmap || host_pack;
            if ((!bitmap && dev_alloc(&s->row_idx, s->nnz)) || dev_alloc(&s->val, s->nnz)) return 1;
            CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&s->csr_ready, cudaEventDisableTiming));
            CUDA_TRY(cudaEventRecord(s->csr_ready, s->stream));                 // allocations + factor uploads are ordered on s->stream
            CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->csr_ready, 0));
            const int nslab = 8;
            s->slab_j.assign(1, 0);
            for (int b = 1; b < nslab; ++b) {
                const uint64_t target = s->nnz / nslab * b;
                const uint64_t *lo = std::lower_bound(Y->col_ptr, Y->col_ptr + s->n + 1, target);
                size_t j = (size_t)(lo - Y->col_ptr);
                if (j > s->n) j = s->n;
                if (j > s->slab_j.back()) s->slab_j.push_back(j);
            }
            if (s->slab_j.back() < s->n) s->slab_j.push_back(s->n);
            const size_t nsl = s->slab_j.size() - 1;
            // (everything below may run on the feeder thread after this function has returned: captures by value, events created here)
            const uint64_t *h_col_ptr = Y->col_ptr;
            const uint32_t *h_row_idx = Y->row_idx;
            const V *h_val = (const V *)Y->val;
            s->slab_ev.resize(nsl);
            for (size_t b = 0; b < nsl; ++b) CUDA_TRY(cudaEventCreateWithFlags(&s->slab_ev[b], cudaEventDisableTiming));
            auto issue_vals = [=](size_t b) -> int {
                const uint64_t e0 = h_col_ptr[s->slab_j[b]], e1 = h_col_ptr[s->slab_j[b + 1]];
                if (e1 > e0) {
                    if (!idx_by_bitmap) CUDA_TRY(cudaMemcpyAsync(s->row_idx + e0, h_row_idx + e0, (e1 - e0) * sizeof(uint32_t), cudaMemcpyHostToDevice, s->copy_stream));
                    CUDA_TRY(cudaMemcpyAsync(s->val + e0, h_val + e0, (e1 - e0) * sizeof(V), cudaMemcpyHostToDevice, s->copy_stream));
                }
                return 0;
            };
            auto publish_slab = [=](size_t b) -> int {
                CUDA_TRY(cudaEventRecord(s->slab_ev[b], s->copy_stream));
                trace_dev(s->copy_stream, "copy stream: slab %zu landed", b);
                s->slabs_published.store(b + 1, std::memory_order_release);
                return 0;
            };
            trace_pt("create: allocations done, issuing slabs");
            trace_dev(s->copy_stream, "copy stream: first copy issued (host mark 'issuing slabs')");
            s->slabs_published.store(0);
            s->feed_state.store(0);
            if (!host_pack) {
                for (size_t b = 0; b < nsl; ++b) if (issue_vals(b) || publish_slab(b)) return 1;
                s->feed_state.store(1);
            } else {
                // host-packed indices, slab by slab: the host cores pack the bitmaps of slab b while the copy engine is busy with
                // the VALUES of slab b (which need no packing and go out one slab ahead: the copy engine starts at once and never
                // waits for a packer); each slab's bitmap follows its values, and the F-update of slab b expands the bitmap on
                // the device as soon as both have landed (trmf_b200_f_update / slab_wait).  Inside c_trmf_train all of this runs
                // on a FEEDER THREAD, so that the calling thread can enqueue the first F-update's per-slab kernels while the
                // slabs are still being packed -- otherwise the compute stream stays empty until the last slab is packed and
                // the slab-wise F-update runs behind the upload instead of under it (measured: 5 ms of an 13 ms call).
                // (Round-2 history: all bitmaps packed first = 17.5 ms end to end at C2; bitmap in front of its values = the
                // copy engine idled through the first slab's packing; F-update enqueued after the packing = 13.2 ms.)
                const size_t bytes = (size_t)s->n * bm_words * sizeof(uint32_t);
                s->pack_buf = bitmap_buf_get(bytes, &s->pack_bytes);
                if (!s->pack_buf) return fail("cannot allocate %zu bytes of pinned memory for the index bitmaps", bytes);
                if (dev_alloc(&s->bm_dev, (size_t)s->n * bm_words)) return 1;
                s->bm_words = bm_words;
                CUDA_TRY(cudaEventRecord(s->csr_ready, s->stream));             // (bm_dev's allocation is ordered on s->stream)
                CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->csr_ready, 0));
                const uint64_t rowsT = s->T, nnz_all = s->nnz;
                auto feed = [=]() -> int {
                    if (cudaSetDevice(s->device) != cudaSuccess) return fail("feeder: cudaSetDevice failed");
                    int rc = 0;
                    auto slab_packed = [=](size_t b) -> int {
                        const size_t j0 = s->slab_j[b], j1 = s->slab_j[b + 1];
                        CUDA_TRY(cudaMemcpyAsync(s->bm_dev + j0 * (size_t)bm_words, s->pack_buf + j0 * (size_t)bm_words,
                                                 (j1 - j0) * (size_t)bm_words * sizeof(uint32_t), cudaMemcpyHostToDevice, s->copy_stream));
                        if (publish_slab(b)) return 1;
                        return b + 1 < nsl ? issue_vals(b + 1) : 0;
                    };
                    if (issue_vals(0)) return 1;
                    const bool packed = pack_bitmap_slabs(h_col_ptr, h_row_idx, s->slab_j, bm_words, rowsT, s->pack_buf, slab_packed, &rc);
                    trace_pt("feeder: host pack done, every slab enqueued");
                    if (rc) return 1;
                    if (!packed) {
                        // unsorted or duplicate indices from some slab on: the whole row_idx array goes over as it is, then the
                        // values of the slabs that are still missing; their events are recorded behind it, so nothing reads those
                        // indices before they are complete (the slabs published so far were carried by their bitmaps -- same
                        // indices -- and may already have been consumed)
                        const size_t done = s->slabs_published.load();
                        s->feed_plain_from = done;
                        CUDA_TRY(cudaMemcpyAsync(s->row_idx, h_row_idx, nnz_all * sizeof(uint32_t), cudaMemcpyHostToDevice, s->copy_stream));
                        for (size_t b = done; b < nsl; ++b) if (issue_vals(b) || publish_slab(b)) return 1;
                    }
                    return 0;
                };
                const bool async = g_async_feed;
                auto feed_main = [=]() {
                    g_on_feeder = async;
                    const int rc = feed();
                    g_on_feeder = false;
                    if (rc) s->feed_err = g_last_error.empty() ? std::string("feeder thread failed") : g_last_error;
                    s->feed_state.store(rc ? -1 : 1, std::memory_order_release);
                };
                if (g_async_feed) {
                    s->feeder = std::thread(feed_main);
                } else {
                    feed_main();      // (a caller of trmf_b200_create may release its buffers when the call returns)
                    if (s->feed_state.load() < 0) return fail("%s", s->feed_err.c_str());
                }
            }
            s->slabs_pending = true;
        }
        if (device_csr) {
            // by-time CSR derived in HBM from the by-series CSC (ingest.cuh): halves the PCIe traffic.  Issued by the
            // first consumer (need_csr) so that the F-update is not queued behind it.
            s->csr_deferred = true;
        } else {
            // caller's CSR arrays as given: copy on the side stream, publish with an event
            CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&s->csr_ready, cudaEventDisableTiming));
            CUDA_TRY(cudaEventRecord(s->csr_ready, s->stream));                 // the allocations are ordered on s->stream
            CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->csr_ready, 0));
            CUDA_TRY(cudaMemcpyAsync(s->row_ptr, Y->row_ptr, (s->T + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->copy_stream));
            if (s->nnz) {
                CUDA_TRY(cudaMemcpyAsync(s->col_idx, Y->col_idx, s->nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, s->copy_stream));
                CUDA_TRY(cudaMemcpyAsync(s->val_t, Y->val_t, s->nnz * sizeof(V), cudaMemcpyHostToDevice, s->copy_stream));
            }
            CUDA_TRY(cudaEventRecord(s->csr_ready, s->copy_stream));
            s->csr_pending = true;
        }
    } else {
        if (h2d_new(s, &s->Yd, Y->val, s->T * s->n)) return 1;
    }
    return 0;
}

extern "C" void trmf_b200_feed_mode(int32_t async_feed) { g_async_feed = async_feed != 0; }
extern "C" S *trmf_b200_create(const PyMatrix *Y, const uint32_t *lag_set, uint32_t lag_size, const PyMatrix *W,
                               const PyMatrix *H, const PyMatrix *lag_val, int32_t missing, int32_t device) {
    S *s = new S();
    if (create_host_impl(s, Y, lag_set, lag_size, W, H, lag_val, missing, device)) {
        std::string keep = g_last_error;
        trmf_b200_destroy(s);
        g_last_error = keep;
        return nullptr;
    }
    return s;
}

extern "C" S *trmf_b200_create_device(uint64_t T, uint64_t n, uint64_t nnz, uint32_t k, const uint64_t *d_row_ptr,
                                      const uint32_t *d_col_idx, const void *d_val_t, const uint64_t *d_col_ptr,
                                      const uint32_t *d_row_idx, const void *d_val, const uint32_t *lag_set_host,
                                      uint32_t lag_size, void *d_W, void *d_H, void *d_lag_val, int32_t device) {
    g_last_error.clear();
    S *s = new S();
    s->device = device;
    s->T = T; s->n = n; s->nnz = nnz; s->k = (int)k;
    s->missing = true;
    s->sparse_storage = true;
    int rc = 0;
    if (k < 1 || k > 128) rc = fail("rank k = %u outside the supported range 1..128", k);
    if (!rc) rc = check_lags(s, lag_set_host, lag_size);
    if (!rc) rc = session_common_init(s);
    if (rc) {
        std::string keep = g_last_error;
        trmf_b200_destroy(s);
        g_last_error = keep;
        return nullptr;
    }
    s->row_ptr = const_cast<uint64_t *>(d_row_ptr); s->col_idx = const_cast<uint32_t *>(d_col_idx);
    s->val_t = (V *)const_cast<void *>(d_val_t);
    s->col_ptr = const_cast<uint64_t *>(d_col_ptr); s->row_idx = const_cast<uint32_t *>(d_row_idx);
    s->val = (V *)const_cast<void *>(d_val);
    s->W = (V *)d_W; s->H = (V *)d_H; s->th = (V *)d_lag_val;
    s->idx_strict = true;      // the API's contract: canonical CSR / CSC (strictly ascending index lists, no cell twice)
    return s;
}

// Host CSC in, host CSR out, through the device transpose (ingest.cuh): the ingest primitive on its own.
extern "C" int trmf_b200_csr_from_csc(uint64_t T, uint64_t n, uint64_t nnz, const uint64_t *col_ptr, const uint32_t *row_idx,
                                      const void *val, uint64_t *row_ptr, uint32_t *col_idx, void *val_t, int32_t device) {
    g_last_error.clear();
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (T >= (1ull << 32) || n >= (1ull << 32)) return fail("T and n must fit uint32 indices");
    cudaStream_t st;
    CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint64_t *d_cp = nullptr, *d_rp = nullptr;
    uint32_t *d_ri = nullptr, *d_ci = nullptr;
    V *d_v = nullptr, *d_vt = nullptr;
    const size_t nz = nnz ? nnz : 1;
    int rc = 0;
    if (cudaMalloc((void **)&d_cp, (n + 1) * sizeof(uint64_t)) || cudaMalloc((void **)&d_rp, (T + 1) * sizeof(uint64_t)) ||
        cudaMalloc((void **)&d_ri, nz * sizeof(uint32_t)) || cudaMalloc((void **)&d_ci, nz * sizeof(uint32_t)) ||
        cudaMalloc((void **)&d_v, nz * sizeof(V)) || cudaMalloc((void **)&d_vt, nz * sizeof(V)))
        rc = fail("csr_from_csc: out of device memory");
    if (!rc) {
        cudaMemcpyAsync(d_cp, col_ptr, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
        if (nnz) {
            cudaMemcpyAsync(d_ri, row_idx, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_v, val, nnz * sizeof(V), cudaMemcpyHostToDevice, st);
        }
        cudaError_t e = csr_from_csc_device<V>(st, prop.multiProcessorCount, T, n, nnz, d_cp, d_ri, d_v, d_rp, d_ci, d_vt);
        if (e == cudaSuccess) {
            cudaMemcpyAsync(row_ptr, d_rp, (T + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
            if (nnz) {
                cudaMemcpyAsync(col_idx, d_ci, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
                cudaMemcpyAsync(val_t, d_vt, nnz * sizeof(V), cudaMemcpyDeviceToHost, st);
            }
            e = cudaStreamSynchronize(st);
        }
        if (e != cudaSuccess) rc = fail("csr_from_csc failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d_cp); cudaFree(d_rp); cudaFree(d_ri); cudaFree(d_ci); cudaFree(d_v); cudaFree(d_vt);
    cudaStreamDestroy(st);
    return rc;
}

// the host half of the packed ingest on its own (no device involved): what c_trmf_train does to plain row indices of a mostly-observed
// matrix before they cross PCIe.  Returns 0 = packed, 1 = declined (indices not strictly ascending within a series, or >= T).
extern "C" int trmf_b200_pack_bitmap_host(uint64_t T, uint64_t n, const uint64_t *col_ptr, const uint32_t *row_idx, uint32_t *bitmap) {
    g_last_error.clear();
    if (T >= (1ull << 32)) return fail("T must fit uint32 indices");
    // a few slabs, so that the slab-ordered hand-over is exercised too
    std::vector<size_t> slab_j(1, 0);
    for (int b = 1; b <= 3; ++b) { const size_t j = (size_t)(n * b / 3); if (j > slab_j.back()) slab_j.push_back(j); }
    if (slab_j.back() < n) slab_j.push_back((size_t)n);
    if (slab_j.size() < 2) return 0;
    int rc = 0;
    size_t seen = 0;
    const bool ok = pack_bitmap_slabs(col_ptr, row_idx, slab_j, (uint32_t)((T + 31) / 32), T, bitmap, [&](size_t b) { seen += b == seen; return 0; }, &rc);
    if (ok && seen != slab_j.size() - 1) return fail("slabs were not handed over in order");
    return ok ? 0 : 1;
}

extern "C" int trmf_b200_bitmap_expand(uint64_t T, uint64_t n, uint64_t nnz, const uint64_t *col_ptr, const uint32_t *bitmap,
                                       uint32_t *row_idx, int32_t device) {
    g_last_error.clear();
    CUDA_TRY(cudaSetDevice(device));
    if (T >= (1ull << 32)) return fail("T must fit uint32 indices");
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const uint32_t words = (uint32_t)((T + 31) / 32);
    uint64_t *d_cp = nullptr;
    uint32_t *d_bm = nullptr, *d_ri = nullptr;
    int rc = 0;
    if (cudaMalloc((void **)&d_cp, (n + 1) * sizeof(uint64_t)) || cudaMalloc((void **)&d_bm, std::max<size_t>(1, n * (size_t)words) * sizeof(uint32_t)) ||
        cudaMalloc((void **)&d_ri, std::max<uint64_t>(1, nnz) * sizeof(uint32_t)))
        rc = fail("bitmap_expand: out of device memory");
    if (!rc) {
        cudaMemcpy(d_cp, col_ptr, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice);
        cudaMemcpy(d_bm, bitmap, n * (size_t)words * sizeof(uint32_t), cudaMemcpyHostToDevice);
        cudaMemset(d_ri, 0xff, std::max<uint64_t>(1, nnz) * sizeof(uint32_t));
        if (nnz) bitmap_expand_kernel<<<(unsigned)(sms * 8), 256>>>(d_cp, d_bm, n, T, words, d_ri);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(row_idx, d_ri, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail("bitmap_expand failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d_cp); cudaFree(d_bm); cudaFree(d_ri);
    return rc;
}

extern "C" int trmf_b200_set_params(S *s, double lambdaI, double lambdaAR, double lambdaLag) {
    s->lambdaI = lambdaI; s->lambdaAR = lambdaAR; s->lambdaLag = lambdaLag;
    return 0;
}

static int wait_slabs(S *s);
extern "C" int trmf_b200_set_stream(S *s, void *cuda_stream) {
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->copy_stream) CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    if (s->aux_stream) CUDA_TRY(cudaStreamSynchronize(s->aux_stream));
    s->csr_pending = false;
    if (wait_slabs(s)) return 1;   // (everything has landed: expands host-packed bitmaps; a deferred CSR build simply runs on the new stream)
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->own_stream) { cudaStreamDestroy(s->stream); s->own_stream = false; }
    s->stream = (cudaStream_t)cuda_stream;
    return 0;
}

extern "C" int trmf_b200_enable_timing(S *s, int32_t on) { s->timing = on != 0; return 0; }
extern "C" int trmf_b200_sync(S *s) {
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

// the whole by-series CSC has landed (a consumer other than the slab-wise F-update calls this)
// host-packed ingest: row indices of the series [j0, j1) out of their bitmaps (the slab's copies are already waited for)
static int expand_slab_bitmaps(S *s, size_t j0, size_t j1) {
    if (!s->bm_dev || j1 <= j0) return 0;
    if (s->feed_plain_from != (size_t)-1) {      // the feeder fell back to plain indices from that slab on
        const size_t jp = s->slab_j[std::min(s->feed_plain_from, s->slab_j.size() - 1)];
        if (j0 >= jp) return 0;
        j1 = std::min(j1, jp);
    }
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((size_t)s->num_sms * 8, (j1 - j0 + 7) / 8));
    bitmap_expand_kernel<<<grid, 256, 0, s->stream>>>(s->col_ptr + j0, s->bm_dev + j0 * (size_t)s->bm_words, j1 - j0, s->T, s->bm_words, s->row_idx);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
// slab b of a slab-wise upload is on its way: make the session stream wait for it (a feeder thread may still be packing it)
static int slab_wait(S *s, size_t b) {
    while (s->slabs_published.load(std::memory_order_acquire) <= b) {
        if (s->feed_state.load(std::memory_order_acquire) < 0) return fail("%s", s->feed_err.c_str());
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->slab_ev[b], 0));
    return 0;
}
static int wait_slabs(S *s) {
    if (!s->slabs_pending) return 0;
    if (s->feeder.joinable()) s->feeder.join();
    if (s->feed_state.load() < 0) return fail("%s", s->feed_err.c_str());
    for (cudaEvent_t e : s->slab_ev) CUDA_TRY(cudaStreamWaitEvent(s->stream, e, 0));
    if (s->bm_dev) {
        if (expand_slab_bitmaps(s, 0, s->n)) return 1;
        dev_free(s->bm_dev);
        s->bm_dev = nullptr;
    }
    s->slabs_pending = false;
    return 0;
}

// by-time CSR ready on the session stream (first consumer calls this): either wait for the side-stream upload of
// the caller's arrays, or build it now from the by-series CSC (ingest.cuh)
static int need_csr(S *s) {
    if (s->csr_deferred) {
        if (wait_slabs(s)) return 1;
        CUDA_TRY(csr_from_csc_device<V>(s->stream, s->num_sms, s->T, s->n, s->nnz, s->col_ptr, s->row_idx, s->val, s->row_ptr,
                                        s->col_idx, s->val_t));
        s->launches += 5;
        s->csr_deferred = false;
    }
    if (!s->csr_pending) return 0;
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->csr_ready, 0));
    s->csr_pending = false;
    return 0;
}

static int read_scalars(S *s) {
    CUDA_TRY(cudaMemcpyAsync(s->h_scal, s->scal, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

// Which Gram kernel serves a sparse F-update / Gram build of rank k over the factor at `X`:
//   mma   3xTF32 mma.sync tiles (f_update_mma.cuh)   -- default where supported
//   ffma  register-tiled FFMA kernel (f_update_tiled.cuh)
//   generic  any k <= 128, any alignment, fp64 build (f_update.cuh)
// TRMF_B200_F_KERNEL=mma|ffma|generic pins the choice (tests, before/after profiles).
enum { F_KERNEL_GENERIC = 0, F_KERNEL_FFMA = 1, F_KERNEL_MMA = 2 };
// TRMF_B200_F_KERNEL=tc routes the Gram passes of a supported rank through the tcgen05 pipeline (f_update_tc.cuh) instead of the
// mma.sync kernel; everything around them (scaling, deferred Cholesky, Gram mat-vecs) is shared, so it counts as F_KERNEL_MMA.
static bool use_tc(int k) {
    const char *e = getenv("TRMF_B200_F_KERNEL");
    return e && !strcmp(e, "tc") && f_update_tc_supported(k);
}
// TRMF_B200_F_KERNEL=mma1 keeps the first-generation mma.sync kernel (fp32 gather, split in the hot loop: f_update_mma.cuh); the default
// is the pre-split / ldmatrix kernel of f_update_mma2.cuh.  Both count as F_KERNEL_MMA.
static bool use_mma2() {
    const char *e = getenv("TRMF_B200_F_KERNEL");
    return !(e && !strcmp(e, "mma1"));
}
static int f_kernel_choice(int k, const V *X) {
    const char *e = getenv("TRMF_B200_F_KERNEL");
    if (getenv("TRMF_B200_GENERIC_F") || (e && !strcmp(e, "generic"))) return F_KERNEL_GENERIC;
    if ((((uintptr_t)X) & 15) != 0) return F_KERNEL_GENERIC;
    const bool want_ffma = e && !strcmp(e, "ffma");
    if (!want_ffma && f_update_mma_supported(k)) return F_KERNEL_MMA;
    if (f_update_tiled_supported(k)) return F_KERNEL_FFMA;
    return F_KERNEL_GENERIC;
}

// scratch of the mma Gram kernel: the column-scaled factor copy and its inverse scales
static int mma_scratch(S *s) {
    if (s->Xs) return 0;
    // (rows of the pre-split fp16 copy are padded to 8 columns per split part: 8 * ceil(k / 8) floats' worth of bytes)
    if (dev_alloc(&s->Xs, std::max(Tcap(s), s->n) * (size_t)(8 * ((s->k + 7) / 8))) || dev_alloc(&s->invs, 128) || dev_alloc(&s->ysc, 2) || dev_alloc(&s->frow, Tcap(s))) return 1;
    return 0;
}

// --------------------------------------------------------------------------
// building blocks
// --------------------------------------------------------------------------
static int dot(S *s, const V *a, const V *b, size_t n, int slot, const int *gate = nullptr) {
    LAUNCH(s, dot_kernel, ew_grid(s, n), 256, 0, a, b, n, s->part, s->ticket, s->scal + slot, gate);
    return 0;
}

// C (M x N, fp64 or V) = alpha * A(M x K, strided) * B (K x N row-major) + beta*addend [+ diag*I]
template <typename TA, typename TB, typename TO>
static int gemm(S *s, const TA *A, size_t sm, size_t sk, const TB *B, size_t M, int N, size_t K, double alpha,
                const V *addend, double beta, double diag, TO *out, const int *gate = nullptr) {
    const size_t mt = (M + GT_M - 1) / GT_M;
    size_t splits = std::max<size_t>(1, (2 * (size_t)s->num_sms + mt - 1) / mt);
    splits = std::min(splits, std::max<size_t>(1, K / (GT_K * 2)));
    splits = std::min(splits, std::max<size_t>(1, s->Cpart_elems / (M * (size_t)N)));
    size_t kchunk = (K + splits - 1) / splits;
    kchunk = (kchunk + GT_K - 1) / GT_K * GT_K;
    splits = (K + kchunk - 1) / kchunk;
    if (M * (size_t)N * splits > s->Cpart_elems) return fail("internal: split-K scratch too small");
    dim3 grid((unsigned)mt, (unsigned)splits);
    LAUNCH(s, (gemm_partial_kernel<TA, TB>), grid, 256, GT_K * N * sizeof(double), A, sm, sk, B, M, N, K, kchunk, s->Cpart, gate);
    LAUNCH(s, (gemm_finish_kernel<TO>), ew_grid(s, M * (size_t)N), 256, 0, s->Cpart, (int)splits, M * (size_t)N, N, alpha,
           addend, beta, diag, out, gate);
    return 0;
}

template <int MODE>
static int sparse_pass(S *s, const uint64_t *ptr, const uint32_t *col, const V *val, const V *Hm, const V *Sv, V *out,
                       size_t rows, int fslot, bool accum = true) {
    const int k = s->k;
    if (MODE != MODE_SPMM && x_pass_fast_supported(k) && ((((uintptr_t)Hm) | ((uintptr_t)Sv) | ((uintptr_t)out)) & 15) == 0 &&
        rows < 0xffffffffull && !getenv("TRMF_B200_GENERIC_PASS")) {
        double *fo = s->scal + (fslot >= 0 ? fslot : SC_TMP);
        if (x_pass_fast_launch<MODE>(s->stream, s->num_sms, ptr, col, val, Hm, Sv, out, k, (uint32_t)rows, accum, s->queue,
                                     s->part, s->ticket, fo))
            return fail("x_pass_fast launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        s->launches++;
        return 0;
    }
    const int WARPS = 4;
    const size_t smem = sparse_pass_smem(k, WARPS);
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((rows + WARPS - 1) / WARPS, (size_t)s->num_sms * 8));
    double *fout = s->scal + (fslot >= 0 ? fslot : SC_TMP);
#define SP_LAUNCH(KR)                                                                                         \
    do {                                                                                                      \
        CUDA_TRY(cudaFuncSetAttribute((sparse_pass_kernel<MODE, KR, WARPS>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        LAUNCH(s, (sparse_pass_kernel<MODE, KR, WARPS>), grid, WARPS * 32, smem, ptr, col, val, Hm, Sv, out, k, rows, accum, \
               s->part, s->ticket, fout);                                                                     \
    } while (0)
    if (k <= 32) SP_LAUNCH(1);
    else if (k <= 64) SP_LAUNCH(2);
    else if (k <= 96) SP_LAUNCH(3);
    else SP_LAUNCH(4);
#undef SP_LAUNCH
    return 0;
}

static LagSet lagset(const S *s) {
    LagSet ls;
    ls.L = s->L; ls.mid = s->mid; ls.lags = s->lags_dev;
    return ls;
}

// arr_base_IX::grad / ::Hv : out = lI*v + lAR*A^T A v   (trmf.cpp:99-149)
static int base_apply(S *s, const V *v, V *out, const int *gate = nullptr) {
    const size_t tk = s->T * (size_t)s->k;
    LAUNCH(s, ar_rho_kernel, ew_grid(s, tk), 256, 0, v, s->th, lagset(s), s->rho, s->T, s->k, gate);
    LAUNCH(s, ar_apply_kernel, ew_grid(s, tk), 256, 0, v, s->th, lagset(s), s->rho, out, s->T, s->k, s->lambdaI, s->lambdaAR, gate);
    return 0;
}

// dense-mode init(): YH = Y H, HTH = H^T H   (trmf.cpp:183-187)
static int dense_loss_init(S *s) {
    const int k = s->k;
    if (s->sparse_storage) {
        if (sparse_pass<MODE_SPMM>(s, s->row_ptr, s->col_idx, s->val_t, s->H, nullptr, s->YH, s->T, -1)) return 1;
    } else {
        const bool rm = s->dense_type == TRMF_DENSE_ROWMAJOR;
        if (gemm<V, V, V>(s, s->Yd, rm ? s->n : 1, rm ? 1 : s->T, s->H, s->T, k, s->n, 1.0, nullptr, 0.0, 0.0, s->YH)) return 1;
    }
    if (gemm<V, V, double>(s, s->H, 1, (size_t)k, s->H, (size_t)k, k, s->n, 1.0, nullptr, 0.0, 0.0, s->HTH)) return 1;
    // tr(Y^T Y)  (trmf.cpp:184)
    if (s->sparse_storage) { if (dot(s, s->val_t, s->val_t, s->nnz, SC_TMP2)) return 1; }
    else { if (dot(s, s->Yd, s->Yd, s->T * s->n, SC_TMP2)) return 1; }
    return 0;
}

// objective value at v -> scal[slot]; dense-mode pieces are combined on the host after read_scalars()
static int fun_launch(S *s, const V *v) {
    const size_t tk = s->T * (size_t)s->k;
    LAUNCH(s, ar_rho_kernel, ew_grid(s, tk), 256, 0, v, s->th, lagset(s), s->rho, s->T, s->k, (const int *)nullptr);
    LAUNCH(s, base_fun_kernel, ew_grid(s, tk), 256, 0, v, s->rho, tk, s->lambdaI, s->lambdaAR, s->part, s->ticket, s->scal + SC_FBASE);
    if (s->missing) {
        if (sparse_pass<MODE_FUN>(s, s->row_ptr, s->col_idx, s->val_t, s->H, v, nullptr, s->T, SC_FLOSS)) return 1;
        if (dist_allreduce_f64(s, s->scal + SC_FLOSS, 1)) return 1;   // sum of the per-slab losses
    } else {
        // 0.5 trYTY + 0.5 <W^T W, HTH> - <YH, W>    (trmf.cpp:189-197)
        const int k = s->k;
        if (gemm<V, V, double>(s, v, 1, (size_t)k, v, (size_t)k, k, s->T, 1.0, nullptr, 0.0, 0.0, s->WTW)) return 1;
        LAUNCH(s, dotd_kernel, 1, 256, 0, s->WTW, s->HTH, (size_t)k * k, s->part, s->ticket, s->scal + SC_TMP);
        if (dot(s, s->YH, v, tk, SC_FLOSS)) return 1;
    }
    return 0;
}
static double fun_combine(const S *s) {
    const double *h = s->h_scal;
    if (s->missing) return h[SC_FLOSS] + h[SC_FBASE];
    return h[SC_FBASE] + 0.5 * h[SC_TMP2] + 0.5 * h[SC_TMP] - h[SC_FLOSS];
}

// loss part of grad / Hv over this rank's series slab; with several ranks the T x k
// partials are summed by one ncclAllReduce and then added to the (replicated) base term
template <int MODE>
static int loss_pass(S *s, const V *v, V *out) {
    if (s->world == 1) return sparse_pass<MODE>(s, s->row_ptr, s->col_idx, s->val_t, s->H, v, out, s->T, -1, true);
    const size_t tk = s->T * (size_t)s->k;
    if (sparse_pass<MODE>(s, s->row_ptr, s->col_idx, s->val_t, s->H, v, s->part_tk, s->T, -1, false)) return 1;
    if (dist_allreduce_v(s, s->part_tk, tk)) return 1;
    LAUNCH(s, axpbypcz_kernel, ew_grid(s, tk), 256, 0, 1.0, out, 1.0, s->part_tk, 0.0, (const V *)nullptr, out, tk);
    return 0;
}

static int grad_launch(S *s, const V *w, V *g) {
    if (base_apply(s, w, g)) return 1;
    if (s->missing) return loss_pass<MODE_GRAD>(s, w, g);
    // G += -YH + W HTH   (trmf.cpp:204-206)
    const size_t tk = s->T * (size_t)s->k;
    LAUNCH(s, axpbypcz_kernel, ew_grid(s, tk), 256, 0, 1.0, g, -1.0, s->YH, 0.0, (const V *)nullptr, g, tk);
    return gemm<V, double, V>(s, w, (size_t)s->k, 1, s->HTH, s->T, s->k, (size_t)s->k, 1.0, g, 1.0, 0.0, g);
}

static int hv_launch(S *s, const V *d, V *Hd, const int *gate = nullptr) {
    if (base_apply(s, d, Hd, gate)) return 1;
    if (s->missing) return loss_pass<MODE_HV>(s, d, Hd);   // (walks over Omega are not gated: host-controlled CG only)
    return gemm<V, double, V>(s, d, (size_t)s->k, 1, s->HTH, s->T, s->k, (size_t)s->k, 1.0, Hd, 1.0, 0.0, Hd, gate);
}

// fun(w) and grad(w) at the same point from one walk over Omega (rf_tron.h:154-158 evaluates both
// at w before the CG solve): base value + base gradient share rho, loss value + loss gradient
// share the residuals.
static int fun_grad_launch(S *s, const V *w, V *g) {
    const size_t tk = s->T * (size_t)s->k;
    LAUNCH(s, ar_rho_kernel, ew_grid(s, tk), 256, 0, w, s->th, lagset(s), s->rho, s->T, s->k, (const int *)nullptr);
    LAUNCH(s, base_fun_kernel, ew_grid(s, tk), 256, 0, w, s->rho, tk, s->lambdaI, s->lambdaAR, s->part, s->ticket, s->scal + SC_FBASE);
    LAUNCH(s, ar_apply_kernel, ew_grid(s, tk), 256, 0, w, s->th, lagset(s), s->rho, g, s->T, s->k, s->lambdaI, s->lambdaAR, (const int *)nullptr);
    if (s->world == 1) {
        if (sparse_pass<MODE_GRADFUN>(s, s->row_ptr, s->col_idx, s->val_t, s->H, w, g, s->T, SC_FLOSS, true)) return 1;
    } else {
        if (sparse_pass<MODE_GRADFUN>(s, s->row_ptr, s->col_idx, s->val_t, s->H, w, s->part_tk, s->T, SC_FLOSS, false)) return 1;
        if (dist_allreduce_v(s, s->part_tk, tk)) return 1;
        if (dist_allreduce_f64(s, s->scal + SC_FLOSS, 1)) return 1;
        LAUNCH(s, axpbypcz_kernel, ew_grid(s, tk), 256, 0, 1.0, g, 1.0, s->part_tk, 0.0, (const V *)nullptr, g, tk);
    }
    return 0;
}

// Decide once per session whether the CG Hessian-vector products go through per-time-stamp
// Grams (needs T*k*k values of HBM; fp32 build with a tiled-kernel-compatible k).
static int gram_prepare(S *s) {
    if (s->gram_state != 0) return 0;
    int want = 1;
    if (!s->missing || f_kernel_choice(s->k, s->H) == F_KERNEL_GENERIC || getenv("TRMF_B200_NO_GRAM_HV")) want = 0;
    const size_t need = (Tcap(s) * (size_t)s->k * s->k + Tcap(s) * (size_t)s->k) * sizeof(V);
    if (want) {
        // memory the allocation can draw on: what the driver reports as free plus what this device's stream-ordered
        // pool holds in reserve from earlier sessions of the process (pool_setup() never releases it)
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        cudaMemPool_t pool;
        unsigned long long reserved = 0, used = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, s->device) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
            free_b += (size_t)(reserved - used);
        if (need > free_b / 2) {
            want = 0;
            if (getenv("TRMF_B200_VERBOSE"))
                fprintf(stderr, "[trmf-b200] note: per-time-stamp Grams (%.1f GB) do not fit (%.1f GB available): the X-update walks Omega\n",
                        need / 1e9, free_b / 1e9);
        }
    }
    if (s->world > 1) {
        // The choice changes the sequence (count, dtype) of NCCL collectives in x_update, so it must be the same on
        // every rank: Grams only if every rank wants them, and the kernel family must agree.
        const int fk = f_kernel_choice(s->k, s->H);
        double h[4] = {(double)want, (double)(fk == F_KERNEL_MMA), (double)(fk == F_KERNEL_FFMA), (double)(fk == F_KERNEL_GENERIC)};
        CUDA_TRY(cudaMemcpyAsync(s->scal + SC_TMP, h, sizeof h, cudaMemcpyHostToDevice, s->stream));
        if (dist_allreduce_f64(s, s->scal + SC_TMP, 4)) return 1;
        CUDA_TRY(cudaMemcpyAsync(h, s->scal + SC_TMP, sizeof h, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        for (int q = 1; q < 4; ++q)
            if (h[q] != 0.0 && h[q] != (double)s->world)
                return fail("ranks disagree on the Gram kernel family (factor alignment or TRMF_B200_F_KERNEL differs between ranks)");
        if (h[0] != (double)s->world) want = 0;
    }
    s->gram_state = -1;
    if (!want) return 0;
    if (dev_alloc(&s->Gt, Tcap(s) * (size_t)s->k * s->k) || dev_alloc(&s->bt, Tcap(s) * (size_t)s->k)) return 1;
    s->gram_state = 1;
    return 0;
}

// loss part of the Hessian-vector product from the per-time-stamp Grams: out (+)= G_i v_i for every time stamp
static int gram_matvec(S *s, const V *v, V *out, bool accum, double *dhd, const int *gate = nullptr) {
    const int k = s->k;
    const int WARPS = 8;
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((s->T + WARPS - 1) / WARPS, (size_t)s->num_sms * 8));
#ifdef TRMF_F32
    if (!getenv("TRMF_B200_GENERIC_GRAM_MATVEC")) {
#define GM_CASE(KK) case KK: LAUNCH(s, (gram_matvec4_kernel<KK, WARPS>), grid, WARPS * 32, 0, s->Gt, v, out, s->T, accum, s->part, s->ticket, dhd, gate); return 0;
        // k >= 56: four warps per CTA (the per-warp fp64 partials of a k x k Gram must fit 48 KB of static shared memory)
#define GM_CASE4(KK) case KK: { const unsigned g4 = (unsigned)std::max<size_t>(1, std::min<size_t>((s->T + 3) / 4, (size_t)s->num_sms * 12)); \
        LAUNCH(s, (gram_matvec4_kernel<KK, 4>), g4, 128, 0, s->Gt, v, out, s->T, accum, s->part, s->ticket, dhd, gate); return 0; }
        switch (k) { GM_CASE(8) GM_CASE(12) GM_CASE(16) GM_CASE(20) GM_CASE(24) GM_CASE(28) GM_CASE(32) GM_CASE(36) GM_CASE(40) GM_CASE(44) GM_CASE(48)
                     GM_CASE4(52) GM_CASE4(56) GM_CASE4(60) GM_CASE4(64) default: break; }
#undef GM_CASE4
#undef GM_CASE
    }
#endif
    if (k <= 32) LAUNCH(s, (gram_matvec_kernel<1, WARPS>), grid, WARPS * 32, 0, s->Gt, v, out, k, s->T, accum, s->part, s->ticket, dhd, gate);
    else LAUNCH(s, (gram_matvec_kernel<2, WARPS>), grid, WARPS * 32, 0, s->Gt, v, out, k, s->T, accum, s->part, s->ticket, dhd, gate);
    return 0;
}

static int gram_hv_launch(S *s, const V *d, V *Hd, bool want_dhd, const int *gate = nullptr) {
    const size_t tk = s->T * (size_t)s->k;
    if (base_apply(s, d, Hd, gate)) return 1;
    if (s->world == 1) {
        if (gram_matvec(s, d, Hd, true, want_dhd ? s->scal + SC_DHD : nullptr, gate)) return 1;
    } else {
        // per-slab partial -> one ncclAllReduce -> fused epilogue (Hd += sum, d'Hd in the same pass).  The collective itself cannot
        // be gated: a gated-off step all-reduces whatever part_tk holds (every rank alike) and nothing reads the result.
        if (gram_matvec(s, d, s->part_tk, false, nullptr, gate)) return 1;
        if (dist_allreduce_v(s, s->part_tk, tk)) return 1;
        LAUNCH(s, add_dot_kernel, ew_grid(s, tk), 256, 0, Hd, s->part_tk, want_dhd ? d : (const V *)nullptr, tk, s->part, s->ticket,
               want_dhd ? s->scal + SC_DHD : (double *)nullptr, gate);
    }
    return 0;
}


// --------------------------------------------------------------------------
// complement formulation (complement.cuh): mostly observed Y, fp32 build
// --------------------------------------------------------------------------
#ifdef TRMF_F32
static int need_csr(S *s);
static int mma_f_range(S *s, size_t j0, size_t j1, bool rescale);
static void cm_reset(S *s);
static bool cm_supported_k(int k) { return f_update_mma_supported(k); }
// decided once per session: a sparse session of the mma2 kernel family whose Y observes >= 70 % of its cells and whose
// dense copy fits comfortably.  TRMF_B200_COMPLEMENT=0 / 1 pins the choice (1 still needs the kernel family).
static bool cm_on(S *s) {
    if (s->cm_state != 0) return s->cm_state > 0;
    s->cm_state = -1;
    const char *e = getenv("TRMF_B200_COMPLEMENT");
    if (e && atoi(e) == 0) return false;
    if (!s->missing || !s->sparse_storage || !use_mma2() || use_tc(s->k) || !cm_supported_k(s->k)) return false;
    if (!s->idx_strict) return false;      // some cell may occur twice: only the walk counts it twice, like the reference
    if (f_kernel_choice(s->k, s->W) != F_KERNEL_MMA || f_kernel_choice(s->k, s->H) != F_KERNEL_MMA) return false;
    if (s->T >= (1ull << 32) || s->n >= (1ull << 32) || s->T == 0 || s->n == 0) return false;
    const double cells = (double)s->T * (double)s->n;
    if (!(e && atoi(e) == 1) && (double)s->nnz < 0.7 * cells) return false;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
    if (cells * 4.0 * 1.5 > (double)free_b * 0.5) return false;
    s->cm_state = 1;
    return true;
}
// Everything the formulation derives from Y comes out of the by-series CSC alone, series range by series range, so that a
// host-buffer session can place each slab while the next one is still crossing PCIe and never needs the by-time CSR:
//   cm_begin   allocations, cm_ptr[0] (a scan over T - the series' lengths: needs col_ptr only), zeroed Y0 and bitmaps
//   cm_place   one range of series: its missing time stamps (cm_idx[0]), its column block of Y0, its bits of the per-time-stamp bitmaps
//   cm_finish  per time stamp: the missing series (cm_ptr[1] / cm_idx[1], ascending) out of the bitmaps, sum of squares of Y0's row
static int cm_scan(S *s, const uint64_t *cnt, uint64_t *out, uint64_t items) {
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, out, (int64_t)items, s->stream));
    CUDA_TRY(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s->stream));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, out, (int64_t)items, s->stream));
    s->launches++;
    CUDA_TRY(cudaFreeAsync(tmp, s->stream));
    return 0;
}
static int cm_begin(S *s) {
    if (s->cm_Y0) return 0;
    const size_t rmax = std::max(Tcap(s), s->n), k = (size_t)s->k;
    const uint64_t nmiss = s->T * s->n - s->nnz;
    const uint32_t w0 = (uint32_t)((s->T + 31) / 32), wT = (uint32_t)((s->n + 31) / 32);
    uint64_t *cnt = nullptr;
    if (dev_alloc(&s->cm_ptr[0], s->n + 1) || dev_alloc(&s->cm_idx[0], std::max<uint64_t>(nmiss, 1)) || dev_alloc(&s->cm_ptr[1], s->T + 1) ||
        dev_alloc(&s->cm_idx[1], std::max<uint64_t>(nmiss, 1)) || dev_alloc(&cnt, s->n + 1) || dev_alloc(&s->cm_bm0, s->n * (size_t)w0) ||
        dev_alloc(&s->cm_bmT, s->T * (size_t)wT))
        return 1;
    if (dev_alloc(&s->cm_Y0, s->T * s->n) || dev_alloc(&s->cm_yy, s->T) || dev_alloc(&s->cm_FtF, k * k) || dev_alloc(&s->cm_rhs, rmax * k) ||
        dev_alloc(&s->cm_Xr, rmax * k))
        return 1;
    s->cm_cpart_elems = rmax * k * 16;
    if (dev_alloc(&s->cm_cpart, s->cm_cpart_elems)) return 1;
    if (!s->Cpart) {       // (the fp64 split-K product of dense.cuh serves the k x k Grams)
        s->Cpart_elems = k * k * 512;
        if (dev_alloc(&s->Cpart, s->Cpart_elems)) return 1;
    }
    const unsigned grid = (unsigned)(s->num_sms * 8);
    LAUNCH(s, cm::count_missing_kernel, (unsigned)std::min<uint64_t>((s->n + 256) / 256, grid), 256, 0, s->col_ptr, s->n, s->T, cnt);
    if (cm_scan(s, cnt, s->cm_ptr[0], s->n + 1)) return 1;
    dev_free(cnt);
    CUDA_TRY(cudaMemsetAsync(s->cm_Y0, 0, s->T * s->n * sizeof(float), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->cm_bmT, 0, s->T * (size_t)wT * sizeof(uint32_t), s->stream));
    s->cm_placed = 0;
    s->cm_ready = false;
    return 0;
}
static int cm_place(S *s, size_t j0, size_t j1) {
    if (j1 <= j0) return 0;
    if (j0 != s->cm_placed) return fail("internal: complement ranges out of order (%zu after %zu)", j0, s->cm_placed);
    const uint64_t rows = j1 - j0;
    const uint32_t w0 = (uint32_t)((s->T + 31) / 32), wT = (uint32_t)((s->n + 31) / 32);
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)s->num_sms * 8, (rows + 7) / 8));
    const bool have_bm = s->bm_dev && s->bm_words == w0 &&
                         (s->feed_plain_from == (size_t)-1 || j1 <= s->slab_j[std::min(s->feed_plain_from, s->slab_j.size() - 1)]);
    if (have_bm) {
        // host-packed ingest: the bitmap of the series' OBSERVED time stamps is still in HBM -- its clear bits are the list
        LAUNCH(s, bitmap_expand_inverted_kernel, grid, 256, 0, s->cm_ptr[0] + j0, s->bm_dev + j0 * (size_t)w0, rows, s->T, w0, s->cm_idx[0]);
    } else {
        uint32_t *bm = s->cm_bm0 + j0 * (size_t)w0;
        LAUNCH(s, cm::missing_bitmap_kernel, grid, 256, 0, s->col_ptr + j0, s->row_idx, rows, w0, bm);
        LAUNCH(s, bitmap_expand_kernel, grid, 256, 0, s->cm_ptr[0] + j0, bm, rows, s->T, w0, s->cm_idx[0]);
    }
    LAUNCH(s, cm::scatter_dense_csc_kernel, dim3((unsigned)((rows + 31) / 32), cm::SC_CHUNKS / 8), dim3(32, 8), 0, s->col_ptr + j0, s->row_idx,
           s->val, rows, (uint64_t)s->n, s->cm_Y0 + j0);
    LAUNCH(s, cm::transpose_missing_kernel, grid, 256, 0, s->cm_ptr[0] + j0, s->cm_idx[0], rows, (uint64_t)j0, wT, s->cm_bmT);
    s->cm_placed = j1;
    return 0;
}
static int cm_finish(S *s) {
    if (s->cm_ready) return 0;
    if (s->cm_placed != s->n) return fail("internal: complement finished with %zu of %zu series placed", s->cm_placed, (size_t)s->n);
    const uint32_t wT = (uint32_t)((s->n + 31) / 32);
    const unsigned grid = (unsigned)(s->num_sms * 8);
    uint64_t *cnt = nullptr;
    if (dev_alloc(&cnt, s->T + 1)) return 1;
    LAUNCH(s, cm::count_bits_kernel, grid, 256, 0, s->cm_bmT, (uint64_t)s->T, (uint64_t)s->n, wT, cnt);
    if (cm_scan(s, cnt, s->cm_ptr[1], s->T + 1)) return 1;
    LAUNCH(s, bitmap_expand_kernel, grid, 256, 0, s->cm_ptr[1], s->cm_bmT, (uint64_t)s->T, (uint64_t)s->n, wT, s->cm_idx[1]);
    LAUNCH(s, cm::row_sumsq_dense_kernel, grid, 256, 0, s->cm_Y0, (uint64_t)s->T, (uint64_t)s->n, s->cm_yy);
    dev_free(cnt);
    dev_free(s->cm_bm0); s->cm_bm0 = nullptr;
    dev_free(s->cm_bmT); s->cm_bmT = nullptr;
    s->cm_ready = true;
    return 0;
}
// the whole of it at once (Y resident, or whatever a slab-wise F-update has not placed yet)
static int cm_build(S *s) {
    if (s->cm_ready) return 0;
    if (cm_begin(s)) return 1;
    if (s->cm_placed < s->n && (wait_slabs(s) || cm_place(s, s->cm_placed, s->n))) return 1;
    return cm_finish(s);
}
template <int NB>
static int cm_gemm_t(S *s, const float *A, size_t sm, size_t sk, const V *B, size_t M, size_t K, double *out) {
    const int N = s->k;
    auto kfn = cm::gemm64_partial_kernel<NB>;
    const size_t smem = cm::gemm64_smem<NB>();
    CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // split-K by K alone (up to 16 parts of at least 256): a row's sums then do not depend on how many rows the call covers, and
    // a series slab of a host-buffer session gets the same right-hand sides as the whole Y at once
    const size_t mt = (M + cm::GM - 1) / cm::GM;
    size_t kchunk = std::max<size_t>(256, (K + 15) / 16);
    kchunk = (kchunk + cm::GK - 1) / cm::GK * cm::GK;
    const size_t splits = (K + kchunk - 1) / kchunk;
    if (M * (size_t)N * splits > s->cm_cpart_elems) return fail("internal: split-K scratch too small (%zu rows)", M);
    dim3 grid((unsigned)mt, (unsigned)splits);
    LAUNCH(s, kfn, grid, 128, smem, A, sm, sk, B, M, N, K, kchunk, s->cm_cpart);
    LAUNCH(s, (gemm_finish_kernel<double>), ew_grid(s, M * (size_t)N), 256, 0, s->cm_cpart, (int)splits, M * (size_t)N, N, 1.0,
           (const V *)nullptr, 0.0, 0.0, out, (const int *)nullptr);
    return 0;
}
// out (M x k, fp64) = A B with A(m, kappa) = A[m * sm + kappa * sk] (a block of Y0)
static int cm_gemm(S *s, const float *A, size_t sm, size_t sk, const V *B, size_t M, size_t K, double *out) {
    switch ((s->k + 7) / 8) {      // 8-column accumulator blocks
        case 1: return cm_gemm_t<1>(s, A, sm, sk, B, M, K, out);
        case 2: return cm_gemm_t<2>(s, A, sm, sk, B, M, K, out);
        case 3: return cm_gemm_t<3>(s, A, sm, sk, B, M, K, out);
        case 4: return cm_gemm_t<4>(s, A, sm, sk, B, M, K, out);
        case 5: return cm_gemm_t<5>(s, A, sm, sk, B, M, K, out);
        case 6: return cm_gemm_t<6>(s, A, sm, sk, B, M, K, out);
        case 7: return cm_gemm_t<7>(s, A, sm, sk, B, M, K, out);
        default: return cm_gemm_t<8>(s, A, sm, sk, B, M, K, out);
    }
}
static int ensure_sys(S *s) {
    if (s->sys) return 0;
    const size_t sysd = f_update_mma_sys_doubles(s->k);
    s->sys_batch = std::max<size_t>(1, std::min<size_t>(std::max(s->n, Tcap(s)), ((size_t)1 << 30) / (sysd * sizeof(double))));
    return dev_alloc(&s->sys, s->sys_batch * sysd);
}
// F-update of the series [j0, j1) through the complement (their range is placed: cm_place); `first` = first range of this F-update:
// W is split into fp16 pairs and its k x k Gram summed once
static int cm_f_range(S *s, size_t j0, size_t j1, bool first) {
    const int k = s->k;
    if (ensure_sys(s)) return 1;
    const size_t smem = sizeof(double) * ((size_t)(k + 1) * (k + 1) + k);
    // threads per system of the deferred solve: 64 / 128 = one CTA per system, 32 = one warp per system on packed triangular storage.
    // Measured at C2 (10 000 systems of 40): 0.949 / 0.955 / 0.992 ms per F-update for 64 / 32 / 128 -- identical factors.
    int solve_nt = 64;
    if (const char *e = getenv("TRMF_B200_SOLVE_THREADS")) solve_nt = atoi(e);
    for (size_t b0 = j0; b0 < j1; b0 += s->sys_batch) {
        const size_t b1 = std::min(j1, b0 + s->sys_batch);
        const bool head = first && b0 == j0;
        const bool timed = s->timing && b0 == 0 && b1 == s->n;
        if (timed) CUDA_TRY(cudaEventRecord(s->ev6, s->stream));
        // (the first launch splits W into fp16 pairs and leaves the values the split carries in cm_Xr: the products below read those)
        if (f_update_mma2_launch<fm::MODE_GONLY>(s->stream, s->num_sms, s->cm_ptr[0] + b0, s->cm_idx[0], (const V *)nullptr, s->W, s->T, s->Xs,
                                                 s->invs, (V *)nullptr, (V *)nullptr, k, 0.0, (uint32_t)(b1 - b0), s->queue, &s->launches,
                                                 nullptr, 0, s->sys, head, nullptr, s->cm_Xr, (uint32_t)std::min<size_t>(s->n, s->sys_batch)))
            return fail("complement Gram launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (timed) CUDA_TRY(cudaEventRecord(s->ev7, s->stream));
        if (b0 == j0 && cm_gemm(s, s->cm_Y0 + j0, 1, s->n, s->cm_Xr, j1 - j0, s->T, s->cm_rhs + j0 * (size_t)k)) return 1;
        if (timed) CUDA_TRY(cudaEventRecord(s->ev8, s->stream));
        s->cm_timed = timed;
        if (head && gemm<V, V, double>(s, s->cm_Xr, 1, (size_t)k, s->cm_Xr, (size_t)k, k, s->T, 1.0, nullptr, 0.0, 0.0, s->cm_FtF)) return 1;
        unsigned grid = 1;
#define CM_SOLVE_NT(KK, NT)                                                                                                  \
    do {                                                                                                                     \
        CUDA_TRY(cudaFuncSetAttribute(cm::solve_kernel<KK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        int per_sm = 0;                                                                                                      \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cm::solve_kernel<KK, NT>, NT, smem));                 \
        grid = (unsigned)std::max<size_t>(1, std::min<size_t>(b1 - b0, (size_t)s->num_sms * (size_t)std::max(per_sm, 1)));    \
        LAUNCH(s, (cm::solve_kernel<KK, NT>), grid, NT, smem, s->col_ptr + b0, s->cm_ptr[0] + b0, s->sys, s->cm_FtF,          \
               s->cm_rhs + b0 * (size_t)k, s->H + b0 * (size_t)k, s->lambdaI, (uint32_t)(b1 - b0));                          \
    } while (0)
#define CM_SOLVE_WARP(KK)                                                                                                    \
    do {                                                                                                                     \
        const size_t smw = cm::solve_warp_smem<KK>();                                                                        \
        CUDA_TRY(cudaFuncSetAttribute(cm::solve_warp_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smw));     \
        int per_sm = 0;                                                                                                      \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cm::solve_warp_kernel<KK>, 32, smw));                 \
        grid = (unsigned)std::max<size_t>(1, std::min<size_t>(b1 - b0, (size_t)s->num_sms * (size_t)std::max(per_sm, 1)));    \
        LAUNCH(s, (cm::solve_warp_kernel<KK>), grid, 32, smw, s->col_ptr + b0, s->cm_ptr[0] + b0, s->sys, s->cm_FtF,          \
               s->cm_rhs + b0 * (size_t)k, s->H + b0 * (size_t)k, s->lambdaI, (uint32_t)(b1 - b0));                          \
    } while (0)
#define CM_SOLVE(KK)                                                                                                         \
    case KK:                                                                                                                 \
        if (solve_nt == 32) CM_SOLVE_WARP(KK); else if (solve_nt == 64) CM_SOLVE_NT(KK, 64); else CM_SOLVE_NT(KK, 128);      \
        break;
        switch (k) {
            CM_SOLVE(8) CM_SOLVE(12) CM_SOLVE(16) CM_SOLVE(20) CM_SOLVE(24) CM_SOLVE(28) CM_SOLVE(32) CM_SOLVE(36) CM_SOLVE(40) CM_SOLVE(44)
            CM_SOLVE(48) CM_SOLVE(52) CM_SOLVE(56) CM_SOLVE(60) CM_SOLVE(64)
            default: return fail("internal: complement solve for k = %d", k);
        }
#undef CM_SOLVE_NT
#undef CM_SOLVE_WARP
#undef CM_SOLVE
    }
    return 0;
}
// F-update of every series.  A host-buffer session whose series slabs are still on their way places and solves slab after slab,
// each as soon as it has landed (the by-time CSR is not needed, hence never built for a call that stays on this path).
static int cm_f_update(S *s) {
    if (s->cm_ready || !s->slabs_pending) return cm_build(s) || cm_f_range(s, 0, s->n, true);
    if (cm_begin(s)) return 1;
    trace_pt("F-update: complement scratch enqueued");
    for (size_t b = 0; b + 1 < s->slab_j.size(); ++b) {
        const size_t j0 = s->slab_j[b], j1 = s->slab_j[b + 1];
        if (slab_wait(s, b)) return 1;
        if (s->feed_plain_from != (size_t)-1 && b >= s->feed_plain_from) {
            // the packer declined a series of this slab: its index list is not strictly ascending -- unsorted, or the same cell
            // twice.  The formulation is dropped for good and every series is (re)done by the walk over the observed entries.
            cm_reset(s);
            s->cm_state = -1;
            s->idx_strict = false;
            if (getenv("TRMF_B200_VERBOSE"))
                fprintf(stderr, "[trmf-b200] note: row indices of some series are not strictly ascending (unsorted or duplicate "
                                "cells): walking the observed entries instead of the complement\n");
            return wait_slabs(s) || mma_f_range(s, 0, s->n, true);
        }
        if (b < 2) trace_pt("F-update: slab published, enqueueing its kernels");
        if (expand_slab_bitmaps(s, j0, j1)) return 1;
        if (b == 0) trace_dev(s->stream, "F-update: slab 0 expanded (scratch zeroed before it)");
        if (j1 > s->cm_placed && cm_place(s, j0, j1)) return 1;
        if (b == 0) trace_dev(s->stream, "F-update: slab 0 placed");
        if (cm_f_range(s, j0, j1, b == 0)) return 1;
        trace_dev(s->stream, "F-update: slab %zu placed and solved", b);
    }
    if (s->bm_dev) { dev_free(s->bm_dev); s->bm_dev = nullptr; }
    s->slabs_pending = false;
    if (cm_finish(s)) return 1;
    trace_dev(s->stream, "F-update: complement lists by time stamp finished");
    return 0;
}
// X-update: per-time-stamp Grams (fp32, s->Gt) + loss gradient into gout (added when gaccum) + sum of squared residuals per row
static int cm_x_gram(S *s, V *gout, int gaccum) {
    const int k = s->k;
    if (cm_build(s) || ensure_sys(s)) return 1;
    for (size_t b0 = 0; b0 < s->T; b0 += s->sys_batch) {
        const size_t b1 = std::min(s->T, b0 + s->sys_batch);
        if (f_update_mma2_launch<fm::MODE_GONLY>(s->stream, s->num_sms, s->cm_ptr[1] + b0, s->cm_idx[1], (const V *)nullptr, s->H, s->n, s->Xs,
                                                 s->invs, (V *)nullptr, (V *)nullptr, k, 0.0, (uint32_t)(b1 - b0), s->queue, &s->launches,
                                                 nullptr, 0, s->sys, b0 == 0, nullptr, s->cm_Xr))
            return fail("complement Gram launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (b0 == 0) {
            if (gemm<V, V, double>(s, s->cm_Xr, 1, (size_t)k, s->cm_Xr, (size_t)k, k, s->n, 1.0, nullptr, 0.0, 0.0, s->cm_FtF)) return 1;
            if (cm_gemm(s, s->cm_Y0, s->n, 1, s->cm_Xr, s->T, s->n, s->cm_rhs)) return 1;
        }
        const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(b1 - b0, (size_t)s->num_sms * 8));
#define CM_XGRAM(KK)                                                                                                         \
    case KK:                                                                                                                 \
        LAUNCH(s, cm::xgram_kernel<KK>, grid, 128, 0, (uint64_t)s->n, s->cm_ptr[1] + b0, s->sys, s->cm_FtF, s->cm_rhs + b0 * (size_t)k, \
               s->cm_yy + b0, s->W + b0 * (size_t)k, gout + b0 * (size_t)k, s->Gt + b0 * (size_t)k * k, gaccum, s->frow + b0,  \
               (uint32_t)(b1 - b0));                                                                                         \
        break;
        switch (k) {
            CM_XGRAM(8) CM_XGRAM(12) CM_XGRAM(16) CM_XGRAM(20) CM_XGRAM(24) CM_XGRAM(28) CM_XGRAM(32) CM_XGRAM(36) CM_XGRAM(40) CM_XGRAM(44)
            CM_XGRAM(48) CM_XGRAM(52) CM_XGRAM(56) CM_XGRAM(60) CM_XGRAM(64)
            default: return fail("internal: complement Gram assembly for k = %d", k);
        }
#undef CM_XGRAM
    }
    return 0;
}
// Y changed (a rolling session moved its window): everything derived from it goes, the decision is taken again
static void cm_reset(S *s) {
    dev_free(s->cm_ptr[0]); dev_free(s->cm_ptr[1]); dev_free(s->cm_idx[0]); dev_free(s->cm_idx[1]); dev_free(s->cm_Y0); dev_free(s->cm_yy);
    dev_free(s->cm_FtF); dev_free(s->cm_rhs); dev_free(s->cm_cpart); dev_free(s->cm_Xr); dev_free(s->cm_bm0); dev_free(s->cm_bmT);
    s->cm_bm0 = s->cm_bmT = nullptr; s->cm_placed = 0; s->cm_ready = false;
    s->cm_ptr[0] = s->cm_ptr[1] = nullptr; s->cm_idx[0] = s->cm_idx[1] = nullptr;
    s->cm_Y0 = nullptr; s->cm_yy = nullptr; s->cm_FtF = nullptr; s->cm_rhs = nullptr; s->cm_cpart = nullptr; s->cm_Xr = nullptr;
    s->cm_cpart_elems = 0;
    s->cm_state = 0;
}
#else
static void cm_reset(S *) {}
static bool cm_on(S *) { return false; }
static int cm_f_update(S *) { return 1; }
static int cm_x_gram(S *, V *, int) { return 1; }
#endif

// sparse F-update of the series [j0, j1) with the mma kernel.  Default: the kernel only assembles the fp64 systems
// (MODE_DEFER) in batches of at most ~1 GB of scratch and chol_solve_kernel factors each batch; the in-kernel solve
// (TRMF_B200_INLINE_SOLVE) gives bit-identical factors.
static int mma_f_range(S *s, size_t j0, size_t j1, bool rescale) {
    const int k = s->k;
    const bool m2 = use_mma2() && !use_tc(k);
    if (m2) {
        // weights as fp16 pairs: once per Y (valh_ok), or per slab while the first F-update follows the upload
        if (s->valh_cap < s->nnz) {
            dev_free(s->valh);
            s->valh = nullptr;
            s->valh_cap = 0;
            s->valh_ok = false;
            if (dev_alloc(&s->valh, std::max<size_t>(s->nnz, 1))) return 1;
            s->valh_cap = std::max<size_t>(s->nnz, 1);
        }
        if (!s->valh_ok) {
            if (f_update_mma2_split_y(s->stream, s->num_sms, s->val, s->col_ptr + j0, (uint32_t)(j1 - j0), s->valh, s->ysc, s->queue + 300, &s->launches))
                return fail("weight split launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            s->valh_ok = j0 == 0 && j1 == s->n;
        }
    }
    if (getenv("TRMF_B200_INLINE_SOLVE")) {
        if (m2 ? f_update_mma2_launch<fm::MODE_SOLVE>(s->stream, s->num_sms, s->col_ptr + j0, s->row_idx, reinterpret_cast<const V *>(s->valh), s->W, s->T,
                                                      s->Xs, s->invs, s->H + j0 * (size_t)k, (V *)nullptr, k, s->lambdaI, (uint32_t)(j1 - j0), s->queue,
                                                      &s->launches, nullptr, 0, nullptr, rescale, s->ysc)
               : f_update_mma_launch<fm::MODE_SOLVE>(s->stream, s->num_sms, s->col_ptr + j0, s->row_idx, s->val, s->W, s->T, s->Xs, s->invs,
                                                s->H + j0 * (size_t)k, (V *)nullptr, k, s->lambdaI, (uint32_t)(j1 - j0), s->queue,
                                                &s->launches, nullptr, 0, nullptr, rescale))
            return fail("f_update_mma launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    const size_t sysd = f_update_mma_sys_doubles(k);
    if (!s->sys) {
        s->sys_batch = std::max<size_t>(1, std::min<size_t>(s->n, ((size_t)1 << 30) / (sysd * sizeof(double))));
        if (dev_alloc(&s->sys, s->sys_batch * sysd)) return 1;
    }
    for (size_t b0 = j0; b0 < j1; b0 += s->sys_batch) {
        const size_t b1 = std::min(j1, b0 + s->sys_batch);
        if (use_tc(k)) {
            const int rc = f_update_tc_launch<fm::MODE_DEFER>(s->stream, s->num_sms, s->col_ptr + b0, s->row_idx, s->val, s->W, s->T, s->Xs, s->invs,
                                                              s->H + b0 * (size_t)k, (V *)nullptr, k, (uint32_t)(b1 - b0), s->queue, s->ysc,
                                                              &s->launches, nullptr, 0, s->sys, rescale && b0 == j0);
            if (rc == 2) return fail("tcgen05 Gram kernel: ptxas' register allocation differs from the kernel's setmaxnreg plan; unset TRMF_B200_F_KERNEL");
            if (rc || f_update_mma_solve(s->stream, s->num_sms, s->col_ptr + b0, s->sys, s->H + b0 * (size_t)k, k, s->lambdaI, (uint32_t)(b1 - b0),
                                         &s->launches))
                return fail("f_update_tc launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            continue;
        }
        if ((m2 ? f_update_mma2_launch<fm::MODE_DEFER>(s->stream, s->num_sms, s->col_ptr + b0, s->row_idx, reinterpret_cast<const V *>(s->valh), s->W,
                                                       s->T, s->Xs, s->invs, s->H + b0 * (size_t)k, (V *)nullptr, k, s->lambdaI, (uint32_t)(b1 - b0),
                                                       s->queue, &s->launches, nullptr, 0, s->sys, rescale && b0 == j0, s->ysc)
                : f_update_mma_launch<fm::MODE_DEFER>(s->stream, s->num_sms, s->col_ptr + b0, s->row_idx, s->val, s->W, s->T, s->Xs, s->invs,
                                                s->H + b0 * (size_t)k, (V *)nullptr, k, s->lambdaI, (uint32_t)(b1 - b0), s->queue,
                                                &s->launches, nullptr, 0, s->sys, rescale && b0 == j0)) ||
            f_update_mma_solve(s->stream, s->num_sms, s->col_ptr + b0, s->sys, s->H + b0 * (size_t)k, k, s->lambdaI, (uint32_t)(b1 - b0),
                               &s->launches))
            return fail("f_update_mma launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

// --------------------------------------------------------------------------
// the three phases
// --------------------------------------------------------------------------
extern "C" int trmf_b200_f_update(S *s) {
    g_last_error.clear();
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->timing) CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    const int k = s->k;
    if (s->missing) {
        if (s->timing) CUDA_TRY(cudaEventRecord(s->ev2, s->stream));
        const int fk = f_kernel_choice(k, s->W);
        if (fk != F_KERNEL_MMA && wait_slabs(s)) return 1;
        if (fk == F_KERNEL_MMA) {
            if (mma_scratch(s)) return 1;
            if (cm_on(s)) {
                // mostly observed Y: the complement formulation (slab by slab while a host-buffer session's upload is under way)
                if (cm_f_update(s)) return 1;
            } else if (s->slabs_pending) {
                // first F-update of a host-buffer session: one launch per series slab, each as soon as its slab has landed
                for (size_t b = 0; b + 1 < s->slab_j.size(); ++b) {
                    if (slab_wait(s, b)) return 1;
                    if (expand_slab_bitmaps(s, s->slab_j[b], s->slab_j[b + 1])) return 1;
                    if (mma_f_range(s, s->slab_j[b], s->slab_j[b + 1], b == 0)) return 1;
                }
                if (s->bm_dev) { dev_free(s->bm_dev); s->bm_dev = nullptr; }
                s->slabs_pending = false;
            } else if (mma_f_range(s, 0, s->n, true)) return 1;
        } else if (fk == F_KERNEL_FFMA) {
            if (f_update_tiled_launch<true>(s->stream, s->num_sms, s->col_ptr, s->row_idx, s->val, s->W, s->H, (V *)nullptr, k,
                                            s->lambdaI, (uint32_t)s->n, s->queue, &s->launches))
                return fail("f_update_tiled launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
            const int ENT = 32;
            const size_t smem = f_update_generic_smem(k, ENT);
            const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(s->n, (size_t)s->num_sms * 4));
#define FG_LAUNCH(MAXP)                                                                                        \
    do {                                                                                                       \
        CUDA_TRY(cudaFuncSetAttribute((f_update_generic_kernel<MAXP, 256, ENT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        LAUNCH(s, (f_update_generic_kernel<MAXP, 256, ENT>), grid, 256, smem, s->col_ptr, s->row_idx, s->val, s->W, s->H, k, \
               s->lambdaI, (uint32_t)s->n);                                                                    \
    } while (0)
            if (k <= 40) FG_LAUNCH(4);
            else if (k <= 64) FG_LAUNCH(9);
            else FG_LAUNCH(33);
#undef FG_LAUNCH
        }
        if (s->timing) CUDA_TRY(cudaEventRecord(s->ev3, s->stream));
    } else {
        // dense mode: YtW = Y^T W, G = W^T W + lI I, one factorisation, n solves (trmf.cpp:319-337)
        if (wait_slabs(s)) return 1;
        if (s->sparse_storage) {
            if (sparse_pass<MODE_SPMM>(s, s->col_ptr, s->row_idx, s->val, s->W, nullptr, s->tmp_nk, s->n, -1)) return 1;
            // widen to fp64 for the solve
            LAUNCH(s, (gemm_finish_kernel<double>), ew_grid(s, s->n * (size_t)k), 256, 0, (const double *)nullptr, 0,
                   s->n * (size_t)k, k, 0.0, s->tmp_nk, 1.0, 0.0, s->YtW, (const int *)nullptr);
        } else {
            const bool rm = s->dense_type == TRMF_DENSE_ROWMAJOR;
            if (gemm<V, V, double>(s, s->Yd, rm ? 1 : s->T, rm ? s->n : 1, s->W, s->n, k, s->T, 1.0, nullptr, 0.0, 0.0, s->YtW)) return 1;
        }
        if (gemm<V, V, double>(s, s->W, 1, (size_t)k, s->W, (size_t)k, k, s->T, 1.0, nullptr, 0.0, s->lambdaI, s->WTW)) return 1;
        const size_t smem = sizeof(double) * ((size_t)(k + 1) * (k + 1) + k);
        CUDA_TRY(cudaFuncSetAttribute(dense_f_solve_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(s, dense_f_solve_kernel<128>, (unsigned)((s->n + 255) / 256), 256, smem, s->WTW, s->YtW, s->n, k, s->H);
    }
    if (s->timing) {
        CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
        CUDA_TRY(cudaEventSynchronize(s->ev1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->ms_f = ms;
        if (s->missing) { CUDA_TRY(cudaEventElapsedTime(&ms, s->ev2, s->ev3)); s->ms_fk = ms; }
        if (s->cm_timed) {
            CUDA_TRY(cudaEventElapsedTime(&ms, s->ev6, s->ev7)); s->ms_cmg = ms;
            CUDA_TRY(cudaEventElapsedTime(&ms, s->ev7, s->ev8)); s->ms_cmp = ms;
            s->cm_timed = false;
        }
    }
    return 0;
}

// One TRON step with pure CG: rf_tron.h:135-254 (max_iter = 1) + trcg 412-505.
// Solver constants: eps_cg = 0.1, CG cap = max_cg_iter*max_tron_iter = 20
// (trmf.h:90-93, trmf.cpp:603-606), clamped to the number of variables
// (trmf.cpp:523-526); eta0 = 1e-4 (rf_tron.h:138).
extern "C" int trmf_b200_x_update(S *s) {
    g_last_error.clear();
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->timing) CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    // The by-time CSR serves the walks over Omega.  The complement formulation with per-time-stamp Grams (Gram build, gradient,
    // objective, Hessian products and f(w + s) all come from the Grams) never touches it: a host-buffer session on that path does
    // not even build it.
    if (s->slabs_pending && s->feeder.joinable()) {
        // no F-update has consumed the upload yet: let the feeder finish, it may have found index lists the formulation cannot take
        s->feeder.join();
        if (s->feed_plain_from != (size_t)-1) {
            s->idx_strict = false;
            if (s->cm_state > 0) cm_reset(s);
            s->cm_state = 0;
        }
    }
    bool csr_free = s->missing && cm_on(s) && !getenv("TRMF_B200_NO_FUSED_GRAD") && !getenv("TRMF_B200_WALK_FNEW") &&
                    !getenv("TRMF_B200_NO_GRAM_HV");
    if (csr_free) {
        if (gram_prepare(s)) return 1;
        csr_free = s->gram_state == 1;
    }
    if (!csr_free && need_csr(s)) return 1;
    const size_t tk = s->T * (size_t)s->k;
    const unsigned eg = ew_grid(s, tk);
    const double eps_cg = 0.1, eta0 = 1e-4, eta1 = 0.25, eta2 = 0.75, sigma1 = 0.25, sigma2 = 0.5, sigma3 = 4.0;
    const size_t max_cg = std::min<size_t>(20, tk);

    if (!s->missing && dense_loss_init(s)) return 1;   // fun_obj->init(), trmf.h:166-173
    int cur = SC_RTR, nxt = SC_RNEW;
    if (s->missing) {
        if (gram_prepare(s)) return 1;
        // With the mma kernel the Gram build also delivers fun(w), grad(w) and (through the quadratic identity) fun(w+s):
        // 3.4 ms at C2 against 2.3 + 1.5 ms for the two walks it replaces, before a single CG step is counted -- always
        // taken.  The FFMA Gram kernel only replaces the Hv walks (6.3 ms vs 2.1 ms each): worth it from ~4 CG steps
        // on, the previous X-update's step count being the predictor.  TRMF_B200_FORCE_GRAM_HV pins the choice for tests.
        const bool fusable = f_kernel_choice(s->k, s->H) == F_KERNEL_MMA && !getenv("TRMF_B200_NO_FUSED_GRAD");
        s->gram_now = s->gram_state == 1 && (fusable || s->prev_cg < 0 || s->prev_cg >= 4 || getenv("TRMF_B200_FORCE_GRAM_HV"));
        bool fused = false;   // fun(w), grad(w) came out of the Gram build's own gather
        s->xg_timed = false;
        s->ms_xg = 0;
        if (s->gram_now) {   // Grams of the (fixed) series factor over every time stamp's observed set
            int rc;
            if (f_kernel_choice(s->k, s->H) == F_KERNEL_MMA) {
                fused = !getenv("TRMF_B200_NO_FUSED_GRAD");
                rc = mma_scratch(s);
                if (!rc && fused) {
                    // base value + base gradient first (the fused kernel adds the loss gradient on top of g)
                    LAUNCH(s, ar_rho_kernel, ew_grid(s, tk), 256, 0, s->W, s->th, lagset(s), s->rho, s->T, s->k, (const int *)nullptr);
                    LAUNCH(s, base_fun_kernel, ew_grid(s, tk), 256, 0, s->W, s->rho, tk, s->lambdaI, s->lambdaAR, s->part, s->ticket, s->scal + SC_FBASE);
                    LAUNCH(s, ar_apply_kernel, ew_grid(s, tk), 256, 0, s->W, s->th, lagset(s), s->rho, s->g, s->T, s->k, s->lambdaI, s->lambdaAR, (const int *)nullptr);
                    const bool one = s->world == 1;
                    if (s->timing) CUDA_TRY(cudaEventRecord(s->ev4, s->stream));
                    if (cm_on(s))
                        rc = cm_x_gram(s, one ? s->g : s->part_tk, one ? 1 : 0);
                    else if (use_tc(s->k))
                        rc = f_update_tc_launch<fm::MODE_GRAD>(s->stream, s->num_sms, s->row_ptr, s->col_idx, s->val_t, s->H, s->n, s->Xs, s->invs,
                                                               one ? s->g : s->part_tk, s->Gt, s->k, (uint32_t)s->T, s->queue, s->ysc, &s->launches,
                                                               s->W, one ? 1 : 0, s->frow);
                    else if (use_mma2())
                        rc = f_update_mma2_launch<fm::MODE_GRAD>(s->stream, s->num_sms, s->row_ptr, s->col_idx, s->val_t, s->H, s->n, s->Xs, s->invs,
                                                                 one ? s->g : s->part_tk, s->Gt, s->k, 0.0, (uint32_t)s->T, s->queue, &s->launches,
                                                                 s->W, one ? 1 : 0, s->frow);
                    else
                        rc = f_update_mma_launch<fm::MODE_GRAD>(s->stream, s->num_sms, s->row_ptr, s->col_idx, s->val_t, s->H, s->n, s->Xs, s->invs,
                                                                one ? s->g : s->part_tk, s->Gt, s->k, 0.0, (uint32_t)s->T, s->queue, &s->launches,
                                                                s->W, one ? 1 : 0, s->frow);
                    if (s->timing) CUDA_TRY(cudaEventRecord(s->ev5, s->stream));
                    s->xg_timed = s->timing;
                    if (!rc) {
                        LAUNCH(s, fm::sum_rows_kernel, ew_grid(s, s->T), 256, 0, s->frow, s->T, 0.5, s->part, s->ticket, s->scal + SC_FLOSS);
                        if (!one) {
                            if (dist_allreduce_v(s, s->part_tk, tk)) return 1;
                            if (dist_allreduce_f64(s, s->scal + SC_FLOSS, 1)) return 1;
                            LAUNCH(s, axpbypcz_kernel, ew_grid(s, tk), 256, 0, 1.0, s->g, 1.0, s->part_tk, 0.0, (const V *)nullptr, s->g, tk);
                        }
                    }
                } else if (!rc) {
                    rc = use_mma2() && !use_tc(s->k)
                             ? f_update_mma2_launch<fm::MODE_STORE>(s->stream, s->num_sms, s->row_ptr, s->col_idx, s->val_t, s->H, s->n, s->Xs,
                                                                    s->invs, s->bt, s->Gt, s->k, 0.0, (uint32_t)s->T, s->queue, &s->launches)
                             : f_update_mma_launch<fm::MODE_STORE>(s->stream, s->num_sms, s->row_ptr, s->col_idx, s->val_t, s->H, s->n, s->Xs,
                                                                   s->invs, s->bt, s->Gt, s->k, 0.0, (uint32_t)s->T, s->queue, &s->launches);
                }
            } else {
                rc = f_update_tiled_launch<false>(s->stream, s->num_sms, s->row_ptr, s->col_idx, s->val_t, s->H, s->bt, s->Gt, s->k, 0.0,
                                                  (uint32_t)s->T, s->queue, &s->launches);
            }
            if (rc) return fail("gram build launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        if (!fused && fun_grad_launch(s, s->W, s->g)) return 1;
    } else {
        if (fun_launch(s, s->W)) return 1;
        if (grad_launch(s, s->W, s->g)) return 1;
    }
    LAUNCH(s, cg_init_kernel, eg, 256, 0, s->g, s->s, s->r, s->d, tk, s->part, s->ticket, s->scal + cur);
    if (read_scalars(s)) return 1;
    const double f = fun_combine(s);
    const double gg = s->h_scal[cur];
    const double gnorm = std::sqrt(gg);
    s->st_f = f; s->st_fnew = f; s->st_gnorm = gnorm; s->st_cg = 0; s->st_acc = 0; s->st_prered = 0; s->st_actred = 0;
    double delta = gnorm;
    if (gnorm > 0.0) {   // rf_tron.h:170: `gnorm <= eps*gnorm1` can only hold for gnorm == 0
        const double cgtol = eps_cg * gnorm;
        double rTr = gg;
        size_t cg_iter = 0;
        double rnorm = gnorm;
        // Device-side CG control: when every kernel of a CG step is one of ours on this stream (Gram-based or dense-mode
        // Hessian products, one GPU) the steps are enqueued in chunks, each step gated by the flag the previous one left
        // in ctl[] -- the loop head of rf_tron.h:441-456 evaluated on the device with the same fp64 scalars -- and the host
        // looks at the scalars once per chunk instead of once per step.  Same kernels, same arguments, same order as the
        // host-driven loop, hence bit-identical iterates.  A step that is gated off still costs its launches (~3 us per
        // kernel, measured), a host round trip ~18 us: chunks of 4 steps bound the waste to 3 idle steps, and when the
        // previous X-update ran into the step cap the whole budget goes out at once.  Walks over Omega and multi-GPU
        // steps (NCCL collectives cannot be gated) keep the host loop; TRMF_B200_HOST_CG=1 forces it,
        // TRMF_B200_CG_CHUNK=<n> sets the chunk length.
        const bool device_cg = (s->world == 1 ? (!s->missing || s->gram_now) : (s->missing && s->gram_now)) && !getenv("TRMF_B200_HOST_CG");
        if (device_cg) {
            if (!s->cgctl && dev_alloc(&s->cgctl, 32)) return 1;
            CUDA_TRY(cudaMemsetAsync(s->cgctl, 0, 32 * sizeof(int), s->stream));
            LAUNCH(s, cg_gate0_kernel, 1, 1, 0, s->scal, cur, s->cgctl, eps_cg);
            size_t chunk = s->prev_cg >= (int)max_cg ? max_cg : 4;
            if (const char *e = getenv("TRMF_B200_CG_CHUNK")) chunk = (size_t)std::max(1, atoi(e));
            size_t j = 1;
            bool go = true;
            while (go && j <= max_cg) {
                const size_t jend = std::min(max_cg, j + chunk - 1);
                for (; j <= jend; ++j) {
                    const int *gate = s->cgctl + j;
                    if (s->missing) {
                        if (gram_hv_launch(s, s->d, s->Hd, true, gate)) return 1;
                    } else {
                        if (hv_launch(s, s->d, s->Hd, gate)) return 1;
                        if (dot(s, s->d, s->Hd, tk, SC_DHD, gate)) return 1;
                    }
                    LAUNCH(s, cg_step1_kernel, eg, 256, 0, s->s, s->r, s->d, s->Hd, tk, s->scal, cur, nxt, s->part, s->ticket, gate);
                    LAUNCH(s, cg_step2_kernel, eg, 256, 0, s->d, s->r, tk, s->scal, cur, nxt, s->cgctl, (int)j);
                    std::swap(cur, nxt);
                }
                if (j <= max_cg) {   // budget left: go on iff the chunk ran to its end and its last loop head said so (= ctl[j])
                    if (read_scalars(s)) return 1;
                    go = (size_t)s->h_scal[SC_CGIT] == jend && !(s->h_scal[SC_RNORM] <= s->h_scal[SC_CGTOL]);
                }
            }
        }
        while (!device_cg) {
            rnorm = std::sqrt(rTr);
            if (rnorm <= cgtol) break;
            if (cg_iter >= max_cg) break;
            ++cg_iter;
            if (s->missing && s->gram_now) {
                if (gram_hv_launch(s, s->d, s->Hd, true)) return 1;
            } else {
                if (hv_launch(s, s->d, s->Hd)) return 1;
                if (dot(s, s->d, s->Hd, tk, SC_DHD)) return 1;
            }
            LAUNCH(s, cg_step1_kernel, eg, 256, 0, s->s, s->r, s->d, s->Hd, tk, s->scal, cur, nxt, s->part, s->ticket, (const int *)nullptr);
            LAUNCH(s, cg_step2_kernel, eg, 256, 0, s->d, s->r, tk, s->scal, cur, nxt, (int *)nullptr, 0);
            if (read_scalars(s)) return 1;
            rTr = s->h_scal[nxt];
            std::swap(cur, nxt);
        }
        // trial point and the acceptance test (rf_tron.h:183-236)
        LAUNCH(s, tron_trial_kernel, eg, 256, 0, s->W, s->s, s->g, s->r, s->wnew, tk, s->part, s->ticket,
               s->scal + SC_GS, s->scal + SC_SR);
        if (dot(s, s->s, s->s, tk, SC_SS)) return 1;   // |s|^2, only feeds the (inert) trust radius / verbose line
        // fun(w + s), rf_tron.h:191.  The objective is exactly quadratic in W for fixed H, so with the per-time-stamp
        // Grams in hand f(w+s) = f(w) + g's + 0.5 s'Hs costs one more Gram-based Hessian-vector product (0.08 ms at C2)
        // instead of a third walk over Omega (1.5 ms), and its rounding error scales with the reduction itself, not
        // with f (two independent evaluations cancel ~7 digits).  TRMF_B200_WALK_FNEW evaluates it by the walk.
        const bool quad_fnew = s->missing && s->gram_now && !getenv("TRMF_B200_WALK_FNEW");
        if (quad_fnew) {
            if (gram_hv_launch(s, s->s, s->Hd, true)) return 1;   // scal[SC_DHD] = s'Hs
        }
        if (read_scalars(s)) return 1;
        if (device_cg) { cg_iter = (size_t)s->h_scal[SC_CGIT]; rnorm = s->h_scal[SC_RNORM]; }
        const double gs = s->h_scal[SC_GS], sr = s->h_scal[SC_SR], snorm = std::sqrt(s->h_scal[SC_SS]);
        const double prered = -0.5 * (gs - sr);
        double fnew;
        if (quad_fnew) {
            fnew = f + gs + 0.5 * s->h_scal[SC_DHD];
        } else {
            if (fun_launch(s, s->wnew)) return 1;
            if (read_scalars(s)) return 1;
            fnew = fun_combine(s);
        }
        const double actred = f - fnew;
        // trust-radius bookkeeping (rf_tron.h:196-217); inert under pure_cg but printed at verbose >= 2
        delta = std::min(delta, snorm);
        double alpha;
        if (fnew - f - gs <= 0) alpha = sigma3;
        else alpha = std::max(sigma1, -0.5 * (gs / (fnew - f - gs)));
        if (actred < eta0 * prered) delta = std::min(std::max(alpha, sigma1) * snorm, sigma2 * delta);
        else if (actred < eta1 * prered) delta = std::max(sigma1 * delta, std::min(alpha * snorm, sigma2 * delta));
        else if (actred < eta2 * prered) delta = std::max(sigma1 * delta, std::min(alpha * snorm, sigma3 * delta));
        else delta = std::max(delta, std::min(alpha * snorm, sigma3 * delta));
        s->prev_cg = (int)cg_iter;
        s->st_cg = (double)cg_iter; s->st_fnew = fnew; s->st_prered = prered; s->st_actred = actred;
        s->st_delta = delta; s->st_rnorm = rnorm;
        if (actred > eta0 * prered) {
            s->st_acc = 1;
            CUDA_TRY(cudaMemcpyAsync(s->W, s->wnew, tk * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
            // (the reference recomputes the gradient here, rf_tron.h:229; its value is never used
            //  because max_iter == 1 ends the loop -- not reproduced)
        }
    }
    if (s->timing) {
        CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
        CUDA_TRY(cudaEventSynchronize(s->ev1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->ms_x = ms;
        if (s->xg_timed) { CUDA_TRY(cudaEventElapsedTime(&ms, s->ev4, s->ev5)); s->ms_xg = ms; }
    }
    return 0;
}

extern "C" int trmf_b200_lag_update(S *s) {
    g_last_error.clear();
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->timing) CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    const size_t smem1 = sizeof(double) * ((size_t)s->lag_chunk + s->mid);
    CUDA_TRY(cudaFuncSetAttribute(lag_gram_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    dim3 grid((unsigned)s->k, (unsigned)s->lag_nchunks);
    LAUNCH(s, lag_gram_kernel<256>, grid, 256, smem1, s->W, lagset(s), s->T, s->k, s->lag_chunk, s->lag_nchunks, s->lag_partial);
    const size_t smem2 = sizeof(double) * ((size_t)(s->L + 1) * (s->L + 1) + s->L);
    CUDA_TRY(cudaFuncSetAttribute(lag_solve_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    LAUNCH(s, lag_solve_kernel<128>, (unsigned)s->k, 128, smem2, s->lag_partial, s->L, s->lag_nchunks, s->lambdaLag, s->th);
    if (s->timing) {
        CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
        CUDA_TRY(cudaEventSynchronize(s->ev1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->ms_lag = ms;
    }
    return 0;
}

static int sq_norm(S *s, const V *a, size_t n, double *out) {
    if (dot(s, a, a, n, SC_TMP)) return 1;
    if (read_scalars(s)) return 1;
    *out = s->h_scal[SC_TMP];
    return 0;
}

// trmf_train's loop, trmf.cpp:647-693 (verbose trace formats 660-688, rf_tron.h:219)
extern "C" int trmf_b200_train(S *s, int32_t max_iter, int32_t period_W, int32_t period_H, int32_t period_Lag, int32_t verbose) {
    if (period_W <= 0 || period_H <= 0 || period_Lag <= 0) return fail("periods must be positive");
    const size_t tk = s->T * (size_t)s->k, nk = s->n * (size_t)s->k, lk = (size_t)s->L * s->k;
    for (int iter = 1; iter <= max_iter; ++iter) {
        double nv = 0;
        if (iter % period_H == 0) {
            if (trmf_b200_f_update(s)) return 1;
            trace_pt("train: F-update enqueued");
            if (getenv("TRMF_B200_TRACE_SYNC")) { cudaStreamSynchronize(s->stream); trace_pt("train: F-update finished on the device"); }
            if (verbose) {
                if (dot(s, s->H, s->H, nk, SC_TMP) || dist_allreduce_f64(s, s->scal + SC_TMP, 1) || read_scalars(s)) return 1;
                fprintf(stderr, ">> iter %d F %g\n", iter, s->h_scal[SC_TMP]);
            }
        }
        if (iter % period_W == 0) {
            if (trmf_b200_x_update(s)) return 1;
            trace_pt("train: X-update returned (its scalars were read back)");
            trace_dev(s->stream, "X-update done");
            if (verbose >= 2) {
                fprintf(stdout, "iter %2d act %5.3e pre %5.3e delta %5.3e f %5.3e |g| %5.3e CG %3d |g| %5.3e\n", 1, s->st_actred,
                        s->st_prered, s->st_delta, s->st_f, s->st_gnorm, (int)s->st_cg, s->st_rnorm);
                fflush(stdout);
            }
            if (verbose) { if (sq_norm(s, s->W, tk, &nv)) return 1; fprintf(stderr, ">> iter %d X %g\n", iter, nv); }
        }
        if (iter % period_Lag == 0) {
            if (verbose) { if (sq_norm(s, s->th, lk, &nv)) return 1; fprintf(stderr, ">> iter %d LV(%d %d) %g\n", iter, s->L, s->k, nv); }
            if (trmf_b200_lag_update(s)) return 1;
            if (verbose) { if (sq_norm(s, s->th, lk, &nv)) return 1; fprintf(stderr, ">> iter %d LV %g\n", iter, nv); }
        }
    }
    return 0;
}

extern "C" int trmf_b200_download(S *s, void *W, void *H, void *lag_val) {
    CUDA_TRY(cudaSetDevice(s->device));
    if (W) CUDA_TRY(cudaMemcpyAsync(W, s->W, s->T * (size_t)s->k * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
    if (H) CUDA_TRY(cudaMemcpyAsync(H, s->H, s->n * (size_t)s->k * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
    if (lag_val) CUDA_TRY(cudaMemcpyAsync(lag_val, s->th, (size_t)s->L * s->k * sizeof(V), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int trmf_b200_upload(S *s, const void *W, const void *H, const void *lag_val) {
    CUDA_TRY(cudaSetDevice(s->device));
    if (W) CUDA_TRY(cudaMemcpyAsync(s->W, W, s->T * (size_t)s->k * sizeof(V), cudaMemcpyHostToDevice, s->stream));
    if (H) CUDA_TRY(cudaMemcpyAsync(s->H, H, s->n * (size_t)s->k * sizeof(V), cudaMemcpyHostToDevice, s->stream));
    if (lag_val) CUDA_TRY(cudaMemcpyAsync(s->th, lag_val, (size_t)s->L * s->k * sizeof(V), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int trmf_b200_save_factors(S *s) {
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t tk = s->T * (size_t)s->k, nk = s->n * (size_t)s->k, lk = (size_t)s->L * s->k;
    if (!s->W_sv && (dev_alloc(&s->W_sv, Tcap(s) * (size_t)s->k) || dev_alloc(&s->H_sv, nk) || dev_alloc(&s->th_sv, lk))) return 1;
    CUDA_TRY(cudaMemcpyAsync(s->W_sv, s->W, tk * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->H_sv, s->H, nk * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->th_sv, s->th, lk * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}
extern "C" int trmf_b200_restore_factors(S *s) {
    CUDA_TRY(cudaSetDevice(s->device));
    if (!s->W_sv) return fail("restore_factors called before save_factors");
    const size_t tk = s->T * (size_t)s->k, nk = s->n * (size_t)s->k, lk = (size_t)s->L * s->k;
    CUDA_TRY(cudaMemcpyAsync(s->W, s->W_sv, tk * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->H, s->H_sv, nk * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->th, s->th_sv, lk * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}

extern "C" double trmf_b200_stat(S *s, int32_t which) {
    switch (which) {
        case TRMF_STAT_CG_ITERS: return s->st_cg;
        case TRMF_STAT_ACCEPTED: return s->st_acc;
        case TRMF_STAT_F: return s->st_f;
        case TRMF_STAT_FNEW: return s->st_fnew;
        case TRMF_STAT_GNORM: return s->st_gnorm;
        case TRMF_STAT_KERNEL_LAUNCHES: return (double)s->launches;
        case TRMF_STAT_F_MS: return s->ms_f;
        case TRMF_STAT_X_MS: return s->ms_x;
        case TRMF_STAT_LAG_MS: return s->ms_lag;
        case TRMF_STAT_F_KERNEL_MS: return s->ms_fk;
        case TRMF_STAT_PRERED: return s->st_prered;
        case TRMF_STAT_ACTRED: return s->st_actred;
        case TRMF_STAT_COLLECTIVES: return (double)s->collectives;
        case TRMF_STAT_X_GRAM_MS: return s->ms_xg;
        case TRMF_STAT_FORMULATION: return s->cm_state > 0 ? 1.0 : 0.0;
        case TRMF_STAT_CM_GRAM_MS: return s->ms_cmg;
        case TRMF_STAT_CM_PRODUCT_MS: return s->ms_cmp;
        case TRMF_STAT_CM_MISSING: return s->cm_state > 0 ? (double)(s->T * s->n - s->nnz) : 0.0;
    }
    return NAN;
}

// --------------------------------------------------------------------------
// library info
// --------------------------------------------------------------------------
extern "C" int trmf_b200_value_bytes(void) { return (int)sizeof(V); }
extern "C" int trmf_b200_pack_threads(void) { return (int)pack_thread_budget(); }
extern "C" const char *trmf_b200_version(void) { return TRMF_B200_VERSION; }
extern "C" const char *trmf_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int trmf_b200_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

// --------------------------------------------------------------------------
// the drop-in entry point
// --------------------------------------------------------------------------
static bool is_rowmajor(const PyMatrix *m) { return m->type == TRMF_DENSE_ROWMAJOR; }
static bool is_colmajor(const PyMatrix *m) { return m->type == TRMF_DENSE_COLMAJOR; }

// check_dimension, trmf.cpp:561-596 -- same messages, same "print and return" behaviour
static bool check_dimension(const PyMatrix *Y, uint32_t lag_size, const PyMatrix *W, const PyMatrix *H, const PyMatrix *lv) {
    bool pass = true;
    if (Y->rows != W->rows) { fprintf(stderr, "[ERR MSG]: Y.rows (%ld) != W.rows (%ld)\n", (long)Y->rows, (long)W->rows); pass = false; }
    if (Y->cols != H->rows) { fprintf(stderr, "[ERR MSG]: Y.cols (%ld) != H.rows (%ld)\n", (long)Y->cols, (long)H->rows); pass = false; }
    if (W->cols != H->cols) { fprintf(stderr, "[ERR MSG]: W.cols (%ld) != H.cols (%ld)\n", (long)W->cols, (long)H->cols); pass = false; }
    if (lag_size != lv->rows) { fprintf(stderr, "[ERR MSG]: lag_set.size(%ld) != lag_val.rows(%ld)\n", (long)lag_size, (long)lv->rows); pass = false; }
    if (W->cols != lv->cols) { fprintf(stderr, "[ERR MSG]: W.cols(%ld) != lag_val.cols(%ld)\n", (long)W->cols, (long)lv->cols); pass = false; }
    if (!is_rowmajor(W)) { fprintf(stderr, "[ERR MSG]: W should be rowmajored\n"); pass = false; }
    if (!is_rowmajor(H)) { fprintf(stderr, "[ERR MSG]: H should be rowmajored\n"); pass = false; }
    if (!is_colmajor(lv)) { fprintf(stderr, "[ERR MSG]: lag_val should be colmajored\n"); pass = false; }
    fflush(stderr);
    return pass;
}

extern "C" void c_trmf_train(const PyMatrix *pyY, uint32_t *py_lag_set, uint32_t py_lag_size, PyMatrix *pyW, PyMatrix *pyH,
                             PyMatrix *pylag_val, int warm_start, double lambdaI, double lambdaAR, double lambdaLag,
                             int32_t max_iter, int32_t period_W, int32_t period_H, int32_t period_Lag, int32_t threads,
                             int32_t missing, int32_t verbose) {
    g_last_error.clear();
    const int solver_type = missing != 0 ? 31 : 30;
    if (verbose > 0) {   // parameter dump, trmf.cpp:607-629 (max_tron_iter already merged: 603-606)
        fprintf(stdout, ">> param.solver_type %d\n", solver_type);
        fprintf(stdout, ">> param.max_iter %d\n", max_iter);
        fprintf(stdout, ">> param.lambdaI %g\n", lambdaI);
        fprintf(stdout, ">> param.lambdaAR %g\n", lambdaAR);
        fprintf(stdout, ">> param.lambdaLag %g\n", lambdaLag);
        fprintf(stdout, ">> param.period_W %d\n", period_W);
        fprintf(stdout, ">> param.period_H %d\n", period_H);
        fprintf(stdout, ">> param.period_Lag %d\n", period_Lag);
        fprintf(stdout, ">> param.threads %d\n", threads);
        fprintf(stdout, ">> param.verbose %d\n", verbose);
        fprintf(stdout, ">> param.eps %g\n", 0.1);
        fprintf(stdout, ">> param.eps_cg %g\n", 0.1);
        fprintf(stdout, ">> param.max_tron_iter %d\n", 1);
        fprintf(stdout, ">> param.max_cg_iter %d\n", 20);
        fprintf(stdout, ">> prob.lag_size %ld:  ", (long)py_lag_size);
        for (uint32_t i = 0; i < py_lag_size; ++i) fprintf(stdout, " %d", (int)py_lag_set[i]);
        fprintf(stdout, "\n");
        fflush(stdout);
    }
    if (!warm_start) {
        // trmf_initialization, trmf.cpp:547-559: W,H ~ U(0,1), lag_val ~ N(0,1) from a default-seeded
        // Mersenne twister.  (Never reached from Python: trmf.py:259 always passes warm_start=True.)
        std::mt19937 rng;
        std::uniform_real_distribution<double> U(0.0, 1.0);
        std::normal_distribution<double> N(0.0, 1.0);
        V *w = (V *)pyW->val, *h = (V *)pyH->val, *lv = (V *)pylag_val->val;
        for (size_t i = 0; i < pyW->rows * pyW->cols; ++i) w[i] = (V)U(rng);
        for (size_t i = 0; i < pyH->rows * pyH->cols; ++i) h[i] = (V)U(rng);
        for (size_t i = 0; i < pylag_val->rows * pylag_val->cols; ++i) lv[i] = (V)N(rng);
        pyW->type = TRMF_DENSE_ROWMAJOR; pyH->type = TRMF_DENSE_ROWMAJOR; pylag_val->type = TRMF_DENSE_COLMAJOR;
    }
    if (!check_dimension(pyY, py_lag_size, pyW, pyH, pylag_val)) {
        g_last_error = "dimension / layout check failed (see the [ERR MSG] lines on stderr); nothing was trained";
        return;
    }
    (void)threads;   // OpenMP thread count of the reference (trmf.cpp:636): no meaning here
    const bool trace = getenv("TRMF_B200_TRACE") != nullptr;   // host-side wall-clock split of the call, to stderr
    auto now_ms = []() {
        struct timespec tw;
        clock_gettime(CLOCK_MONOTONIC, &tw);
        return tw.tv_sec * 1e3 + tw.tv_nsec * 1e-6;
    };
    const double t0 = now_ms();
    trace_pt("begin");
    // device: TRMF_B200_DEVICE if set, else the calling thread's current CUDA device (e.g. torch.cuda.set_device);
    // the caller's current device is put back before returning
    int prev_dev = 0, dev = 0;
    const bool have_prev = cudaGetDevice(&prev_dev) == cudaSuccess;
    if (!have_prev) cudaGetLastError();
    dev = have_prev ? prev_dev : 0;
    if (const char *e = getenv("TRMF_B200_DEVICE")) dev = atoi(e);
    const bool feed_before = g_async_feed;
    g_async_feed = !getenv("TRMF_B200_SYNC_FEED");      // (the caller's buffers stay valid until this call returns)
    S *s = trmf_b200_create(pyY, py_lag_set, py_lag_size, pyW, pyH, pylag_val, missing, dev);
    g_async_feed = feed_before;
    if (!s) { if (have_prev) cudaSetDevice(prev_dev); return; }   // message already on stderr
    if (verbose > 0 && s->missing && f_kernel_choice(s->k, s->W) != F_KERNEL_MMA)
        fprintf(stderr, "[trmf-b200] note: k = %d is outside the tensor-core Gram kernels' ranks (or a factor is not 16-byte aligned): "
                        "the slower %s kernel runs\n", s->k, f_kernel_choice(s->k, s->W) == F_KERNEL_FFMA ? "FFMA" : "generic");
    double t1 = now_ms();
    if (trace) { cudaStreamSynchronize(s->stream); t1 = now_ms(); }
    trmf_b200_set_params(s, lambdaI, lambdaAR, lambdaLag);
    trace_pt("train: begin");
    const int rc = trmf_b200_train(s, max_iter, period_W, period_H, period_Lag, verbose);
    trace_pt("train: returned");
    double t2 = now_ms();
    if (trace) { cudaStreamSynchronize(s->stream); t2 = now_ms(); }
    if (trace) trace_dev(s->stream, "lag_val update done (train finished)");
    if (rc == 0) trmf_b200_download(s, pyW->val, pyH->val, pylag_val->val);
    if (trace) { trace_dev(s->stream, "factors downloaded"); trace_dev_dump(); }
    const double t3 = now_ms();
    std::string keep = g_last_error;
    trmf_b200_destroy(s);
    g_last_error = keep;
    if (have_prev) cudaSetDevice(prev_dev);
    if (trace)
        fprintf(stderr, "[trmf-b200 trace] create+H2D %.2f ms, train %.2f ms, D2H %.2f ms, destroy %.2f ms\n", t1 - t0, t2 - t1,
                t3 - t2, now_ms() - t3);
}

#include "extras.cuh"   // multi-GPU (NCCL) and on-device synthetic data
#include "rolling.cuh"  // rolling-window sessions (rolling_validate with Y resident in HBM)
