// extras.cuh -- multi-GPU (NCCL) and on-device synthetic data entry points.
#pragma once
extern "C" int trmf_b200_nccl_unique_id(void *) { return fail("multi-GPU support is not built into this library yet"); }
extern "C" int trmf_b200_dist_init(S *, int32_t, int32_t, const void *) { return fail("multi-GPU support is not built into this library yet"); }
extern "C" int trmf_b200_allgather_H(S *, void *, const uint64_t *) { return fail("multi-GPU support is not built into this library yet"); }
extern "C" int trmf_b200_synth_generate(trmf_b200_synth *, uint64_t, uint64_t, uint64_t, uint64_t, uint32_t, double, double, uint64_t, int32_t) {
    return fail("on-device synthetic data is not built into this library yet");
}
extern "C" void trmf_b200_free_synth(trmf_b200_synth *) {}
