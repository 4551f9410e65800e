// f_update_tiled.cuh -- register-tiled Gram + Cholesky kernel for the F-update (placeholder:
// the generic kernel in f_update.cuh serves every k until this one is enabled).
#pragma once
#include "common.cuh"
static inline bool f_update_tiled_supported(int k) { (void)k; return false; }
static inline int f_update_tiled_launch(cudaStream_t, int, const uint64_t *, const uint32_t *, const V *, const V *, V *,
                                        int, double, uint32_t, unsigned long long *) { return 1; }
