// synth.cuh -- synthetic Y generated directly in HBM (benchmark input, large configs).
//
// Modelled on the reference's own generator Model.syn_gen (python/trmf/trmf.py:195-220:
// Y = W* H*^T, Gaussian factors) plus the Bernoulli observation mask of SURVEY 8(d).
// Everything is a pure function of (seed, i, j): any column slab can be produced
// on any rank, in either orientation, and reproduced bit-for-bit on the host
// (bench.py:host_synth) -- integer hashing, then fp64 multiplies/adds issued with
// explicit round-to-nearest intrinsics so that no FMA contraction changes a bit.
//
//   w*(i,q), h*(j,q), z(i,j) ~ "N(0,1)": sqrt(3) * (sum of four 16-bit uniforms - 2)
//   Y_ij = sum_q w*(i,q) h*(j,q) + noise * z(i,j)          (q < rank_true, fp64, cast to ValueType)
//   (i,j) observed  iff  top 24 bits of hash(i,j) < p * 2^24
#pragma once
#include "common.cuh"

__host__ __device__ __forceinline__ uint64_t synth_mix(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ uint64_t synth_key(uint64_t seed, uint64_t stream, uint64_t idx) {
    return synth_mix(seed + stream * 0x9E3779B97F4A7C15ULL + synth_mix(idx + 0x632BE59BD9B4E019ULL * (stream + 1)));
}
__device__ __forceinline__ double synth_normal(uint64_t h) {
    const long long ssum = (long long)(h & 0xffff) + (long long)((h >> 16) & 0xffff) + (long long)((h >> 32) & 0xffff) +
                           (long long)((h >> 48) & 0xffff) - 131072LL;
    return __dmul_rn((double)ssum, 1.7320508075688772 / 65536.0);
}
__device__ __forceinline__ bool synth_observed(uint64_t seed, uint64_t i, uint64_t j, uint64_t n_total, uint32_t thresh24) {
    return (uint32_t)(synth_key(seed, 4, i * n_total + j) >> 40) < thresh24;
}
__device__ __forceinline__ double synth_value(uint64_t seed, uint64_t i, uint64_t j, uint64_t n_total, int r, double noise) {
    double acc = 0.0;
    for (int q = 0; q < r; ++q) {
        const double w = synth_normal(synth_key(seed, 1, i * 64 + q));
        const double h = synth_normal(synth_key(seed, 2, j * 64 + q));
        acc = __dadd_rn(acc, __dmul_rn(w, h));
    }
    const double z = synth_normal(synth_key(seed, 3, i * n_total + j));
    return __dadd_rn(acc, __dmul_rn(noise, z));
}

// One warp per line (a time stamp when by_time, a series otherwise) walks the
// other axis 32 cells at a time.  pass 0: count; pass 1: fill (ptr known).
template <bool FILL>
__global__ void synth_line_kernel(uint64_t nlines, uint64_t len, bool by_time, uint64_t n_total, uint64_t col_offset,
                                  uint64_t seed, uint32_t thresh24, int r, double noise, uint64_t *__restrict__ counts,
                                  const uint64_t *__restrict__ ptr, uint32_t *__restrict__ idx, V *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t line = warp; line < nlines; line += nwarps) {
        uint64_t pos = FILL ? ptr[line] : 0, cnt = 0;
        for (uint64_t base = 0; base < len; base += 32) {
            const uint64_t o = base + lane;
            const uint64_t i = by_time ? line : o;
            const uint64_t j = (by_time ? o : line) + col_offset;   // global series index
            const bool obs = o < len && synth_observed(seed, i, j, n_total, thresh24);
            const unsigned m = __ballot_sync(FULL_MASK, obs);
            if (FILL) {
                if (obs) {
                    const uint64_t dst = pos + __popc(m & ((1u << lane) - 1));
                    idx[dst] = (uint32_t)o;   // local index along the walked axis
                    val[dst] = (V)synth_value(seed, i, j, n_total, r, noise);
                }
                pos += __popc(m);
            } else {
                cnt += __popc(m);
            }
        }
        if (!FILL && lane == 0) counts[line] = cnt;
    }
}

// ptr[0..n] = exclusive prefix sum of counts[0..n-1] (ptr[n] = total).  One CTA of 1024
// threads, each owning a contiguous segment; n is at most a few million and this runs
// once per generated matrix.  (CUB is not used: the -DValueType=... build macro collides
// with its template parameter names.)
__global__ void __launch_bounds__(1024) synth_scan_kernel(const uint64_t *__restrict__ counts, uint64_t *__restrict__ ptr, uint64_t n) {
    __shared__ uint64_t sums[1024];
    const uint64_t seg = (n + 1023) / 1024;
    const uint64_t lo = threadIdx.x * seg, hi = lo + seg < n ? lo + seg : n;
    uint64_t local = 0;
    for (uint64_t i = lo; i < hi; ++i) local += counts[i];
    sums[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int t = 0; t < 1024; ++t) { const uint64_t v = sums[t]; sums[t] = run; run += v; }
        ptr[n] = run;
    }
    __syncthreads();
    uint64_t run = sums[threadIdx.x];
    for (uint64_t i = lo; i < hi; ++i) { ptr[i] = run; run += counts[i]; }
}
