// Entry points declared in include/frb200.h whose kernels are not written yet.
// They fail loudly (no CPU fallback, no silent success).
#include "frb_internal.cuh"

extern "C" int32_t frb_rhs_pipelined(frb_prob_t, const double *, double *, int32_t) {
  frb_set_error("frb_rhs_pipelined: not implemented yet");
  return FRB_ERR_STATE;
}

