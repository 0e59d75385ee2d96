// TEST HARNESS (not part of libfrb200.so): runs the __host__ __device__ element routines of
// csrc/frb_euler2d_curv_elem.cuh in plain CPU loops, so that the arithmetic and the index maps of the
// curvilinear kernels can be checked against the NumPy oracle on a box without a GPU
// (tests/test_curv_host.py).  Built by tests/harness/build.py with nvcc as host code only.
#include <string.h>

#include "../../fluxreconstruction.jl_b200/csrc/frb_euler2d_curv_elem.cuh"

// IX / FOLD: the instantiation the kernels launch (unsigned indices, folded operators) or the literal one
template <int NSP, typename IX, bool FOLD>
static void run(const double *u, const double *ua, double *out, double *fx, double *fy, const CurvGeom &g,
                double gamma, const FrbOps &ops, const FrbStage &st) {
  for (int j = 1; j <= g.ny + 1; ++j)
    for (int i = 1; i <= g.nx + 1; ++i)
      for (int p = 0; p < NSP; ++p)
        frbcurv::face_xy<NSP, IX>(i, j, p, j <= g.ny, i <= g.nx, u, fx, fy, g, gamma, ops);
  double tile[NSP * NSP * 4], fyt[2 * NSP * 4];
  FrbStage sk = st;  // the launcher's mapping of rhs_only (frb_launch_euler2d_curv)
  if (sk.rhs_only) { sk.ca = 0.0; sk.cb = 0.0; sk.cdt = 1.0; sk.use_a = 0; }
  if (!sk.use_a) sk.ca = 0.0;
  frbcurv::RowCarry<NSP> c[NSP];
  for (int j = 1; j <= g.ny; ++j)
    for (int i = 1; i <= g.nx; ++i) {  // the block barrier of the kernel = the boundary between the two loops
      for (int l = 0; l < NSP; ++l)
        frbcurv::row_xpass<NSP, IX, FOLD>(i, j, l, u, fx, fy, g, gamma, ops, tile, fyt, 1, c[l]);
      for (int l = 0; l < NSP; ++l)
        frbcurv::row_ypass<NSP, IX, FOLD>(i, j, l, ua, out, g, ops, sk, tile, fyt, 1, c[l]);
    }
}

// operators as the ABI takes them (lpdm column-major nsp x nsp); stage = (ca, cb, cdt, use_a, rhs_only)
extern "C" int curv_host_stage(int nx, int ny, int nsp, const double *u, const double *ua, double *out, double *fx,
                               double *fy, const double *iJ, const double *n1, const double *n2, const double *fpc,
                               const double *vert, const double *r, int flags, double gamma, const double *ll, const double *lr, const double *lpdm,
                               const double *dgl, const double *dgr, double ca, double cb, double cdt, int use_a,
                               int rhs_only) {
  FrbOps ops;
  memset(&ops, 0, sizeof ops);
  for (int q = 0; q < nsp; ++q) {
    ops.ll[q] = ll[q]; ops.lr[q] = lr[q]; ops.dgl[q] = dgl[q]; ops.dgr[q] = dgr[q];
    for (int k = 0; k < nsp; ++k) ops.lpdm[q * FRB_NSPMAX + k] = lpdm[q + nsp * k];
  }
  for (int k = 0; k < nsp; ++k)  // as make_operators of frb_api.cu
    for (int q = 0; q < nsp; ++q)
      ops.dmod[k * FRB_NSPMAX + q] = ops.lpdm[k * FRB_NSPMAX + q] - dgl[k] * ll[q] - dgr[k] * lr[q];
  CurvGeom g;
  g.nx = nx; g.ny = ny; g.iJ = iJ; g.n1 = n1; g.n2 = n2; g.fpc = fpc; g.vert = vert;
  for (int q = 0; q < FRB_NSPMAX; ++q) g.r[q] = (r && q < nsp) ? r[q] : 0.0;
  g.fy_row = (flags & FRB_CURV_FY_ROW_INDEX) ? 1 : 0;
  g.wall_xlo = (flags & FRB_CURV_WALL_XLO) ? 1 : 0;
  g.flux = (flags >> 8) & 3;  // test-only: the flux kind rides in bits 8..9
  FrbStage st = {ca, cb, cdt, use_a, rhs_only, 0};
  const bool fast = (flags >> 10) & 1;  // test-only: the kernels' instantiation
  switch (nsp * 2 + fast) {
    case 4: run<2, size_t, false>(u, ua, out, fx, fy, g, gamma, ops, st); break;
    case 6: run<3, size_t, false>(u, ua, out, fx, fy, g, gamma, ops, st); break;
    case 8: run<4, size_t, false>(u, ua, out, fx, fy, g, gamma, ops, st); break;
    case 5: run<2, unsigned, true>(u, ua, out, fx, fy, g, gamma, ops, st); break;
    case 7: run<3, unsigned, true>(u, ua, out, fx, fy, g, gamma, ops, st); break;
    case 9: run<4, unsigned, true>(u, ua, out, fx, fy, g, gamma, ops, st); break;
    default: return -1;
  }
  return 0;
}
