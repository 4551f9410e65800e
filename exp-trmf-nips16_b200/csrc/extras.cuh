// extras.cuh -- multi-GPU plumbing (NCCL over NVLink) and on-device synthetic data.
// Included at the end of trmf_b200.cu (uses its session struct and error helpers).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "synth.cuh"

// --------------------------------------------------------------------------
// NCCL, resolved at run time (dlopen) so that single-GPU users need no NCCL at
// all.  Under torchrun the process has already loaded torch's bundled
// libnccl.so.2; otherwise the system library is used.
// --------------------------------------------------------------------------
struct NcclApi {
    bool ok = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.ok) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail("cannot load libnccl.so.2 (%s); import torch first or add NCCL to LD_LIBRARY_PATH", dlerror());
#define NCCL_SYM(field, name)                                              \
    *(void **)(&g_nccl.field) = dlsym(h, name);                            \
    if (!g_nccl.field) return fail("libnccl.so.2 lacks symbol %s", name);
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(Broadcast, "ncclBroadcast")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    g_nccl.ok = true;
    return 0;
}

#define NCCL_TRY(...)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (__VA_ARGS__);                                                      \
        if (r__ != ncclSuccess)                                                                \
            return fail("%s failed at %s:%d: %s", #__VA_ARGS__, __FILE__, __LINE__, g_nccl.GetErrorString(r__)); \
    } while (0)

static const ncclDataType_t kNcclV = sizeof(V) == 8 ? ncclFloat64 : ncclFloat32;

extern "C" int trmf_b200_nccl_unique_id(void *out128) {
    g_last_error.clear();
    if (nccl_load()) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return 0;
}

extern "C" int trmf_b200_dist_init(S *s, int32_t rank, int32_t world, const void *unique_id128) {
    g_last_error.clear();
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank/world %d/%d", rank, world);
    if (!s->missing) return fail("multi-GPU sharding is implemented for the sparse (missing != 0) path only");
    s->rank = rank;
    s->world = world;
    if (world == 1) return 0;
    if (nccl_load()) return 1;
    CUDA_TRY(cudaSetDevice(s->device));
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof id);
    ncclComm_t comm;
    NCCL_TRY(g_nccl.CommInitRank(&comm, world, id, rank));
    s->nccl_comm = (void *)comm;
    s->own_comm = true;
    if (dev_alloc(&s->part_tk, Tcap(s) * (size_t)s->k)) return 1;
    return 0;
}

extern "C" int trmf_b200_dist_attach(S *s, S *owner) {
    g_last_error.clear();
    if (!s->missing) return fail("multi-GPU sharding is implemented for the sparse (missing != 0) path only");
    if (s->T != owner->T || s->k != owner->k) return fail("dist_attach: sessions disagree on T or k");
    s->rank = owner->rank;
    s->world = owner->world;
    s->nccl_comm = owner->nccl_comm;
    s->own_comm = false;
    if (s->world > 1 && !s->part_tk && dev_alloc(&s->part_tk, Tcap(s) * (size_t)s->k)) return 1;
    return 0;
}

static void dist_teardown(S *s) {
    if (s->nccl_comm && s->own_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)s->nccl_comm);
    s->nccl_comm = nullptr;
}

extern "C" int trmf_b200_copy_to_host(void *dst_host, const void *src_device, uint64_t bytes) {
    g_last_error.clear();
    if (bytes) CUDA_TRY(cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

// in-place sum over ranks of a T x k partial (ValueType) / of `count` fp64 scalars
static int dist_allreduce_v(S *s, V *buf, size_t count) {
    if (s->world == 1) return 0;
    NCCL_TRY(g_nccl.AllReduce(buf, buf, count, kNcclV, ncclSum, (ncclComm_t)s->nccl_comm, s->stream));
    s->collectives++;
    return 0;
}
static int dist_allreduce_f64(S *s, double *buf, size_t count) {
    if (s->world == 1) return 0;
    NCCL_TRY(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)s->nccl_comm, s->stream));
    s->collectives++;
    return 0;
}

extern "C" int trmf_b200_allgather_H(S *s, void *d_H_full, const uint64_t *counts) {
    g_last_error.clear();
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t k = (size_t)s->k;
    if (s->world == 1) {
        CUDA_TRY(cudaMemcpyAsync(d_H_full, s->H, s->n * k * sizeof(V), cudaMemcpyDeviceToDevice, s->stream));
        return 0;
    }
    if (counts[s->rank] != s->n) return fail("allgather_H: counts[rank] (%llu) != local series (%zu)",
                                             (unsigned long long)counts[s->rank], s->n);
    // slabs are nnz-balanced, hence of unequal length: one broadcast per slab, grouped
    size_t off = 0;
    NCCL_TRY(g_nccl.GroupStart());
    for (int r = 0; r < s->world; ++r) {
        V *dst = (V *)d_H_full + off * k;
        const void *src = r == s->rank ? (const void *)s->H : (const void *)dst;
        ncclResult_t rc = g_nccl.Broadcast(src, dst, counts[r] * k, kNcclV, r, (ncclComm_t)s->nccl_comm, s->stream);
        if (rc != ncclSuccess) { g_nccl.GroupEnd(); return fail("ncclBroadcast failed: %s", g_nccl.GetErrorString(rc)); }
        off += counts[r];
    }
    NCCL_TRY(g_nccl.GroupEnd());
    s->collectives++;
    return 0;
}

// --------------------------------------------------------------------------
// synthetic data
// --------------------------------------------------------------------------
static int synth_orientation(cudaStream_t st, int sms, uint64_t nlines, uint64_t len, bool by_time, uint64_t n_total,
                             uint64_t col_offset, uint64_t seed, uint32_t thresh24, int r, double noise,
                             uint64_t **ptr_out, uint32_t **idx_out, V **val_out, uint64_t *nnz_out) {
    uint64_t *counts = nullptr, *ptr = nullptr;
    CUDA_TRY(cudaMalloc((void **)&counts, (nlines + 1) * sizeof(uint64_t)));
    CUDA_TRY(cudaMalloc((void **)&ptr, (nlines + 1) * sizeof(uint64_t)));
    CUDA_TRY(cudaMemsetAsync(counts, 0, (nlines + 1) * sizeof(uint64_t), st));
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nlines + 7) / 8, (uint64_t)sms * 16));
    synth_line_kernel<false><<<grid, 256, 0, st>>>(nlines, len, by_time, n_total, col_offset, seed, thresh24, r, noise,
                                                   counts, nullptr, nullptr, nullptr);
    CUDA_TRY(cudaGetLastError());
    synth_scan_kernel<<<1, 1024, 0, st>>>(counts, ptr, nlines);
    CUDA_TRY(cudaGetLastError());
    uint64_t nnz = 0;
    CUDA_TRY(cudaMemcpyAsync(&nnz, ptr + nlines, sizeof nnz, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    uint32_t *idx = nullptr;
    V *val = nullptr;
    CUDA_TRY(cudaMalloc((void **)&idx, std::max<uint64_t>(nnz, 1) * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc((void **)&val, std::max<uint64_t>(nnz, 1) * sizeof(V)));
    synth_line_kernel<true><<<grid, 256, 0, st>>>(nlines, len, by_time, n_total, col_offset, seed, thresh24, r, noise,
                                                  nullptr, ptr, idx, val);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(counts);
    *ptr_out = ptr; *idx_out = idx; *val_out = val; *nnz_out = nnz;
    return 0;
}

extern "C" int trmf_b200_synth_generate(trmf_b200_synth *out, uint64_t T, uint64_t n, uint64_t n_total, uint64_t col_offset,
                                        uint32_t rank_true, double p_observed, double noise, uint64_t seed, int32_t device) {
    g_last_error.clear();
    memset(out, 0, sizeof *out);
    if (rank_true < 1 || rank_true > 64) return fail("rank_true must be in 1..64");
    if (T >= (1ull << 32) || n >= (1ull << 32)) return fail("T and n must fit uint32 indices");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    double pt = p_observed * 16777216.0;
    const uint32_t thresh24 = (uint32_t)(pt < 0 ? 0 : (pt > 16777216.0 ? 16777216.0 : pt));
    uint64_t nnz1 = 0, nnz2 = 0;
    V *vt = nullptr, *v = nullptr;
    if (synth_orientation(nullptr, prop.multiProcessorCount, T, n, true, n_total, col_offset, seed, thresh24, (int)rank_true,
                          noise, &out->d_row_ptr, &out->d_col_idx, &vt, &nnz1))
        return 1;
    if (synth_orientation(nullptr, prop.multiProcessorCount, n, T, false, n_total, col_offset, seed, thresh24, (int)rank_true,
                          noise, &out->d_col_ptr, &out->d_row_idx, &v, &nnz2))
        return 1;
    if (nnz1 != nnz2) return fail("internal: orientations disagree on nnz (%llu vs %llu)", (unsigned long long)nnz1, (unsigned long long)nnz2);
    out->d_val_t = vt; out->d_val = v;
    out->T = T; out->n = n; out->nnz = nnz1;
    return 0;
}

extern "C" void trmf_b200_free_synth(trmf_b200_synth *s) {
    if (!s) return;
    cudaFree(s->d_row_ptr); cudaFree(s->d_col_idx); cudaFree(s->d_val_t);
    cudaFree(s->d_col_ptr); cudaFree(s->d_row_idx); cudaFree(s->d_val);
    memset(s, 0, sizeof *s);
}
