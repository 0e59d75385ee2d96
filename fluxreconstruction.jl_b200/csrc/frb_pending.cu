// Entry points declared in include/frb200.h whose kernels are not written yet.
// They fail loudly (no CPU fallback, no silent success).
#include "frb_internal.cuh"

int frb_launch_ns2d(frb_prob_t, const double *, const double *, double *, FrbStage) {
  frb_set_error("ns2d (gas-kinetic) kernels are not implemented yet");
  return FRB_ERR_STATE;
}

extern "C" int32_t frb_ns2d_create(frb_ctx_t, int32_t, int32_t, const frb_operators *, double, double,
                                   double, double, double, double, double, double, double,
                                   frb_prob_t *) {
  frb_set_error("frb_ns2d_create: not implemented yet");
  return FRB_ERR_STATE;
}

extern "C" int32_t frb_rhs_pipelined(frb_prob_t, const double *, double *, int32_t) {
  frb_set_error("frb_rhs_pipelined: not implemented yet");
  return FRB_ERR_STATE;
}

