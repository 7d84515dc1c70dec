// TEMPORARY: entry points not written yet (replaced by eval.cu / cma.cu / awgn.cu).
#include "common.cuh"
#define NI(name) vaeq::set_error(name " not implemented yet"); return VAEQ_ENODEV
extern "C" {
int vaeq_find_shift(const float *, int64_t, const float *, int64_t, const uint16_t *, int64_t, const float *, int32_t, int32_t, int32_t, float *, int16_t *, int32_t *, void *, void *) { NI("vaeq_find_shift"); }
int vaeq_ser_iqflip(const float *, int64_t, const uint16_t *, int64_t, int32_t, int32_t, int32_t *, float *, void *) { NI("vaeq_ser_iqflip"); }
int vaeq_ser_constell(float *, int64_t, const uint16_t *, int64_t, const float *, const float *, float, int32_t, int32_t, int32_t *, float *, void *, void *) { NI("vaeq_ser_constell"); }
int vaeq_gmi(const float *, int64_t, const uint16_t *, int64_t, const float *, int32_t, int32_t, float *, void *, void *) { NI("vaeq_gmi"); }
size_t vaeq_cma_scratch_bytes(int32_t, int32_t, int32_t) { return 0; }
int vaeq_cma(int32_t, const float *, int32_t, float, float *, int32_t, float, int32_t, int32_t, int32_t, int32_t, float *, float *, int32_t, void *, void *) { NI("vaeq_cma"); }
size_t vaeq_cpe_scratch_bytes(int32_t) { return 0; }
int vaeq_cpe(const float *, int32_t, float *, void *, void *) { NI("vaeq_cpe"); }
size_t vaeq_awgn_workspace_bytes(int32_t, int32_t, int32_t) { return 0; }
size_t vaeq_adam_state_floats_awgn(int32_t) { return 0; }
int vaeq_awgn_forward(const vaeq_awgn_desc *, void *) { NI("vaeq_awgn_forward"); }
int vaeq_awgn_forward_backward(const vaeq_awgn_desc *, void *) { NI("vaeq_awgn_forward_backward"); }
int vaeq_awgn_train_step(const vaeq_awgn_desc *, float, float, void *) { NI("vaeq_awgn_train_step"); }
}
