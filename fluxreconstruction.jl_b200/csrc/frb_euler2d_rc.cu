// Fused 2-D Euler residual + RK stage on the row-chunk layout (frb_rc.cuh): the roofline path
// of the resident time loop (frb_step).
//
// Same arithmetic as frb_euler2d_march.cu (reference: dudt! of example/euler2d_wave.jl:35-107
// + the OrdinaryDiffEq stage axpys): a CTA of NSP warps marches over the rows of one strip,
// x pass (warp = point row l) -> shared memory -> y pass (warp = point column k), HLL once per
// face, bottom-face flux carried in registers.  What changes is how bytes move:
//   * u row j of the strip = ONE contiguous chunk -> one cp.async.bulk (UBLKCP) into the ring;
//   * u_n row j = one more bulk copy into its own smem buffer, issued a whole row step before
//     its use: no LDG latency on the critical path, no registers held for it;
//   * u' goes out with full-line st.global.cs (a warp writes one aligned 256-B row of a plane),
//     lanes 1 and 30 also refresh the duplicate of their column in the neighbouring chunk.
// Shared memory (p3): 24-B stage 2 tiles + u_n + xd + xrp = 72 KB, 16-B stage 3 tiles + xd + xrp
// = 72 KB -> 3 CTAs/SM either way.
#include "frb_euler2d_rc_impl.cuh"

using frbrc::RcParams;

int frb_rc_dispatch_hll(frb_prob_t p, const RcParams &rp, const MarchOps &mo, bool usea, bool samej) {
  return frbrc::dispatch_rc_flux<FRB_FLUX_HLL>(p, rp, mo, usea, samej);
}


bool frb_euler2d_rc_supported(frb_prob_t p) {
  return p->kind == K_EULER2D && !p->curv_iJ && (p->nsp == 4 || p->nsp == 3);
}

// u, ua, out are RC buffers (frb_rc.cuh); stage semantics as in frb_launch_euler2d_march
int frb_launch_euler2d_rc(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st,
                          const RcHalo *halo) {
  if (!frb_euler2d_rc_supported(p)) {
    frb_set_error("row-chunk kernel needs euler2d with deg 2 or 3");
    return FRB_ERR_ARG;
  }
  RcParams rp;
  rp.u = u;
  rp.ua = ua;
  rp.out = out;
  rp.g = rc_geom(p->nx, p->ny, p->nsp);
  rp.rows_per_seg = 0;
  rp.gamma = p->gamma;
  if (halo) rp.h = *halo;
  else memset(&rp.h, 0, sizeof rp.h);
  double cxs, cys;  // -cdt/Jx, -cdt/Jy  (L = -(dF/dr / Jx + dG/ds / Jy))
  bool usea;
  if (st.rhs_only) {
    rp.ca = 0.0; rp.cb = 0.0; usea = false;
    cxs = -1.0 / p->Jx; cys = -1.0 / p->Jy;
  } else {
    const double cdt = st.nested ? st.cb * st.cdt : st.cdt;
    rp.ca = st.ca; rp.cb = st.cb; usea = st.use_a != 0;
    cxs = -cdt / p->Jx; cys = -cdt / p->Jy;
  }
  MarchOps mo;
  const int n = p->nsp;
  for (int k = 0; k < 4; ++k) {
    const bool in = k < n;
    mo.ll[k] = in ? p->ops.ll[k] : 0.0;
    mo.lr[k] = in ? p->ops.lr[k] : 0.0;
    mo.glx[k] = in ? cxs * p->ops.dgl[k] : 0.0;
    mo.grx[k] = in ? cxs * p->ops.dgr[k] : 0.0;
    mo.gly[k] = in ? cys * p->ops.dgl[k] : 0.0;
    mo.gry[k] = in ? cys * p->ops.dgr[k] : 0.0;
    for (int q = 0; q < 4; ++q) {
      const double d = (in && q < n) ? p->ops.dmod[k * FRB_NSPMAX + q] : 0.0;
      mo.dmx[k * 4 + q] = cxs * d;
      mo.dmy[k * 4 + q] = cys * d;
    }
  }
  const bool samej = cxs == cys;
  // the common flux is a compile-time choice of the kernel (frb_set_flux)
  if (p->flux == FRB_FLUX_LF) return frb_rc_dispatch_lf(p, rp, mo, usea, samej);
  if (p->flux == FRB_FLUX_ROE) return frb_rc_dispatch_roe(p, rp, mo, usea, samej);
  return frb_rc_dispatch_hll(p, rp, mo, usea, samej);
}
