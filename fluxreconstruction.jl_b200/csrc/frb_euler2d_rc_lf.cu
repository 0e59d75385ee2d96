// Row-chunk stage kernel of the 2-D Euler path with the LF common flux (frb_set_flux): the same kernel as
// frb_euler2d_rc.cu, the flux functor is a template parameter (frb_euler2d_rc_impl.cuh, frb_physics.cuh).
#include "frb_euler2d_rc_impl.cuh"

int frb_rc_dispatch_lf(frb_prob_t p, const frbrc::RcParams &rp, const MarchOps &mo, bool usea, bool samej) {
  return frbrc::dispatch_rc_flux<FRB_FLUX_LF>(p, rp, mo, usea, samej);
}
