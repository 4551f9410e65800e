#include <immintrin.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
__attribute__((target("avx2"))) static inline uint32_t pack_hor_avx2(__m256i a) {
    __m128i x = _mm_or_si128(_mm256_castsi256_si128(a), _mm256_extracti128_si256(a, 1));
    x = _mm_or_si128(x, _mm_shuffle_epi32(x, 0x4e));
    x = _mm_or_si128(x, _mm_shuffle_epi32(x, 0xb1));
    return (uint32_t)_mm_cvtsi128_si32(x);
}
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
#pragma GCC unroll 4
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
// v2: no popcount chain: the number of entries in word wi = first position (among the next 33) whose word index is > wi
__attribute__((target("avx2,bmi,popcnt"))) static void pack_series_words2(const uint32_t *r, size_t cnt, uint32_t *w, uint32_t words) {
    const __m256i one = _mm256_set1_epi32(1), m31 = _mm256_set1_epi32(31);
    size_t p = 0;
    for (uint32_t wi = 0; wi < words && p < cnt; ++wi) {
        const __m256i wv = _mm256_set1_epi32((int)wi);
        __m256i acc = _mm256_setzero_si256();
        uint32_t mask = 0;
#pragma GCC unroll 4
        for (int v = 0; v < 4; ++v) {
            const __m256i idx = _mm256_loadu_si256((const __m256i *)(r + p + 8 * v));
            const __m256i in = _mm256_cmpeq_epi32(_mm256_srli_epi32(idx, 5), wv);
            acc = _mm256_or_si256(acc, _mm256_and_si256(_mm256_sllv_epi32(one, _mm256_and_si256(idx, m31)), in));
            mask |= (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(in)) << (8 * v);
        }
        // entries of this word are a prefix: count = trailing ones; clip to what is left of the series
        unsigned taken = (unsigned)__builtin_ctz(~mask | 0u) ;
        if (mask == 0xffffffffu) taken = 32;
        const size_t left = cnt - p;
        uint32_t bits = pack_hor_avx2(acc);
        if (taken > left) {  // ran into the next series: rebuild scalar
            taken = (unsigned)left; bits = 0; for (size_t i = 0; i < left; ++i) bits |= 1u << (r[p + i] & 31);
        }
        w[wi] = bits;
        p += taken;
    }
}
int main() {
    const size_t T = 10000, n = 3000;
    std::vector<uint32_t> idx; std::vector<uint64_t> ptr(1, 0);
    srand(1);
    for (size_t j = 0; j < n; ++j) { for (uint32_t t = 0; t < T; ++t) if (rand() % 10) idx.push_back(t); ptr.push_back(idx.size()); }
    idx.resize(idx.size() + 64, 0);
    const uint32_t words = (T + 31) / 32;
    std::vector<uint32_t> out(n * words), out2(n * words);
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        for (size_t j = 0; j < n; ++j) { memset(&out[j * words], 0, words * 4); pack_series_words(&idx[ptr[j]], ptr[j + 1] - ptr[j], &out[j * words], words); }
        auto t1 = std::chrono::steady_clock::now();
        for (size_t j = 0; j < n; ++j) { memset(&out2[j * words], 0, words * 4); pack_series_words2(&idx[ptr[j]], ptr[j + 1] - ptr[j], &out2[j * words], words); }
        auto t2 = std::chrono::steady_clock::now();
        printf("words: %.3f ns/entry, words2: %.3f ns/entry, same %d\n", std::chrono::duration<double>(t1 - t0).count() * 1e9 / ptr[n],
               std::chrono::duration<double>(t2 - t1).count() * 1e9 / ptr[n], (int)(out == out2));
    }
}
