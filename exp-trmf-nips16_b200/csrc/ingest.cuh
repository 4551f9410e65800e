// ingest.cuh -- by-time CSR of Y built on the device from the by-series CSC.
//
// The reference's Python glue builds BOTH orientations on the host with two scipy conversions
// (rf_util.py:88-98) and its C core adopts them by pointer (rf_matrix.h:3430-3446, transpose = pointer swap
// 1633-1640).  A host-buffer call here would have to push both over PCIe -- 16 bytes per observed entry,
// 1.44 GB at BASELINE config 2, which is most of the end-to-end time.  Instead only the CSC half (col_ptr /
// row_idx / val: the F-update needs it first) is uploaded and the CSR half is derived in HBM:
//
//   a stable sort of the CSC entries by row index.  CSC order is (column, row) ascending, so a STABLE sort by
//   row yields (row, column) ascending -- exactly the canonical CSR scipy produces (csc.tocsr() is the same
//   counting sort); index arrays come out bit-identical (tests/test_ingest_gpu.py).
//
// The sort is cub::DeviceRadixSort (least-significant-digit radix sort is stable) over only the bits T needs,
// with the (column, value) pair as payload; row_ptr is a binary search over the sorted keys.  This is ingest,
// not one of the three ALS updates, hence a library primitive; TRMF_B200_HOST_CSR=1 uploads the caller's CSR
// arrays instead.
#pragma once
#include "common.cuh"
// the reference-compatible build flag -DValueType=... is a macro; CUB uses that word as a template parameter name
#pragma push_macro("ValueType")
#undef ValueType
#include <cub/device/device_radix_sort.cuh>
#pragma pop_macro("ValueType")

template <typename VT> struct CsrPayload { uint32_t col; VT val; };

// one warp per series: payload[e] = (j, val[e]) for e in [col_ptr[j], col_ptr[j+1])
template <typename VT>
__global__ void ingest_expand_kernel(const uint64_t *__restrict__ col_ptr, const VT *__restrict__ val, uint64_t n,
                                     CsrPayload<VT> *__restrict__ payload) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = warp; j < n; j += nwarps) {
        const uint64_t lo = col_ptr[j], hi = col_ptr[j + 1];
        for (uint64_t e = lo + lane; e < hi; e += 32) {
            CsrPayload<VT> p;
            p.col = (uint32_t)j;
            p.val = val[e];
            payload[e] = p;
        }
    }
}

template <typename VT>
__global__ void ingest_split_kernel(const CsrPayload<VT> *__restrict__ payload, uint64_t nnz, uint32_t *__restrict__ col_idx,
                                    VT *__restrict__ val_t) {
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < nnz; e += (uint64_t)gridDim.x * blockDim.x) {
        const CsrPayload<VT> p = payload[e];
        col_idx[e] = p.col;
        val_t[e] = p.val;
    }
}

// row_ptr[i] = number of sorted keys < i  (i = 0 .. T)
__global__ void ingest_rowptr_kernel(const uint32_t *__restrict__ keys, uint64_t nnz, uint64_t T, uint64_t *__restrict__ row_ptr) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= T; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((uint64_t)keys[mid] < i) lo = mid + 1; else hi = mid;
        }
        row_ptr[i] = lo;
    }
}

// TRMF_SPARSE_BITMAP ingest: row_idx[col_ptr[j] ..] = the set bits of series j's bitmap (words = ceil(T / 32) per series), ascending.
// One warp per series, 32 words per step: popc + warp prefix sum place every lane's bits.  Bit-exact by construction; a bitmap whose
// population disagrees with col_ptr is clipped to the series' range (never writes outside it).
template <bool INVERT>
__device__ __forceinline__ void bitmap_expand_body(const uint64_t *__restrict__ col_ptr, const uint32_t *__restrict__ bitmap, uint64_t n, uint64_t T,
                                     uint32_t words, uint32_t *__restrict__ row_idx) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = warp; j < n; j += nwarps) {
        uint64_t pos = col_ptr[j];
        const uint64_t end = col_ptr[j + 1];
        const uint32_t *bm = bitmap + j * (uint64_t)words;
        for (uint32_t w0 = 0; w0 < words; w0 += 32) {
            const uint32_t wi = w0 + lane;
            uint32_t bits = wi < words ? (INVERT ? ~__ldg(bm + wi) : __ldg(bm + wi)) : 0u;
            if (wi == words - 1 && (T & 31)) bits &= (1u << (T & 31)) - 1u;     // padding bits of the last word
            const uint32_t cnt = __popc(bits);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += up;
            }
            uint64_t dst = pos + (incl - cnt);
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                if (dst < end) row_idx[dst] = wi * 32u + (uint32_t)b;
                ++dst;
            }
            pos += __shfl_sync(FULL_MASK, incl, 31);
        }
    }
}

__global__ void bitmap_expand_kernel(const uint64_t *__restrict__ col_ptr, const uint32_t *__restrict__ bitmap, uint64_t n, uint64_t T,
                                     uint32_t words, uint32_t *__restrict__ row_idx) {
    bitmap_expand_body<false>(col_ptr, bitmap, n, T, words, row_idx);
}
// the CLEAR bits instead (the complement formulation's list of a series' missing time stamps, straight from the bitmap of its
// observed ones: `col_ptr` then holds the offsets of the missing-cell lists)
__global__ void bitmap_expand_inverted_kernel(const uint64_t *__restrict__ col_ptr, const uint32_t *__restrict__ bitmap, uint64_t n, uint64_t T,
                                              uint32_t words, uint32_t *__restrict__ row_idx) {
    bitmap_expand_body<true>(col_ptr, bitmap, n, T, words, row_idx);
}

// All arrays on the device; temporaries come from (and return to) the stream-ordered pool.  Returns a cudaError_t.
template <typename VT>
static cudaError_t csr_from_csc_device(cudaStream_t st, int num_sms, uint64_t T, uint64_t n, uint64_t nnz, const uint64_t *col_ptr,
                                       const uint32_t *row_idx, const VT *val, uint64_t *row_ptr, uint32_t *col_idx, VT *val_t) {
    cudaError_t e;
    if (nnz == 0) return cudaMemsetAsync(row_ptr, 0, (T + 1) * sizeof(uint64_t), st);
    typedef CsrPayload<VT> P;
    P *pin = nullptr, *pout = nullptr;
    uint32_t *kout = nullptr;
    void *temp = nullptr;
    size_t temp_bytes = 0;
    int bits = 1;
    while (bits < 32 && (T - 1) >> bits) ++bits;
    if ((e = cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, row_idx, kout, pin, pout, (int64_t)nnz, 0, bits, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync((void **)&pin, nnz * sizeof(P), st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync((void **)&pout, nnz * sizeof(P), st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync((void **)&kout, nnz * sizeof(uint32_t), st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&temp, temp_bytes ? temp_bytes : 1, st)) != cudaSuccess) return e;
    const unsigned grid = (unsigned)(num_sms * 8);
    ingest_expand_kernel<VT><<<grid, 256, 0, st>>>(col_ptr, val, n, pin);
    if ((e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, row_idx, kout, pin, pout, (int64_t)nnz, 0, bits, st)) != cudaSuccess) return e;
    ingest_split_kernel<VT><<<grid, 256, 0, st>>>(pout, nnz, col_idx, val_t);
    ingest_rowptr_kernel<<<(unsigned)std::min<uint64_t>((T + 256) / 256, (uint64_t)grid), 256, 0, st>>>(kout, nnz, T, row_ptr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    cudaFreeAsync(pin, st); cudaFreeAsync(pout, st); cudaFreeAsync(kout, st); cudaFreeAsync(temp, st);
    return cudaSuccess;
}
