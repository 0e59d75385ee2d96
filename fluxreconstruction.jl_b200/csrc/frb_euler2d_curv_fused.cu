// 2-D Euler on curvilinear structured quadrilaterals, one launch per stage WITHOUT a march: every
// block owns 30 elements of one element row and computes the common fluxes of its own faces itself --
//   y faces  thread (lane, p): the traces of rows j-1 (top), j (both) and j+1 (bottom) at flux point p, 48 loads
//            issued together; the rows above and below are in L2 (their own blocks run a few waves earlier /
//            later), so the state crosses DRAM about once; both y fluxes of the element -> fyt
//   x faces  the row's traces are in the row owner's registers, the neighbours' come by __shfl (lanes 0 / 31 are
//            the duplicate edge lanes of the rectangular row-chunk kernel)
// then the x pass / y pass of the element kernel.  The column owners leave the block's own row in a shared-memory
// tile, so the row owners do not load it a second time, and every global load of the block (metric, normals, the
// three rows) is issued before the first use: one exposed DRAM latency per block.  Against the two-kernel
// form (frb_euler2d_curv.cu: 2.71 GB of DRAM traffic per 16-B stage at 1024^2 p3) the common fluxes never leave
// the SM and the state is read from DRAM once (1.6 GB algorithmic); every y face is evaluated by both rows that
// share it, which costs about what the face kernel's second pass over the state cost in instructions.  Against
// the marching kernel (frb_euler2d_curv_march.cu) there is no loop-carried state and no ring: 12 warps / SM with
// thousands of independent blocks hide the latencies the march exposes (profiles/r02_summary.md, section F).
//
// Reference semantics: dudt! of dev/parallelogram.jl:80-165 and dev/cylinder2.jl:52-164 as restated in
// frb_euler2d_curv_elem.cuh (same layouts, same factors, flux_normal_t, the FOLD form of the correction).
#include "frb_euler2d_curv_elem.cuh"

namespace {

constexpr int kOwn = 30;
#ifndef FRB_CURV_FUSED_MINB
#define FRB_CURV_FUSED_MINB 3  // blocks per SM the register budget is cut for (168 registers at p3)
#endif

template <int NSP, typename IX, int FLUX>
__global__ void __launch_bounds__(32 * NSP, FRB_CURV_FUSED_MINB)
euler2d_curv_fused_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
                          CurvGeom g, double gamma, FrbOps ops, FrbStage st) {
  using frbcurv::plane;
  using frbcurv::W4;
  // the block's row of elements, [plane(k, l, m)][lane]: first the state (written by the column owners, read by the
  // row owners), then f2 (every row owner overwrites the values it has just read)
  __shared__ double tile[NSP * NSP * 4 * 32];
  __shared__ double fyt[2 * NSP * 4 * 32];  // y common fluxes below / above the elements, [side][p][m][lane]
  const int lane = threadIdx.x, l = threadIdx.y;
  const int nx = g.nx, ny = g.ny;
  const IX NXG = nx + 2, NE = NXG * (IX)(ny + 2);
  const int iraw = blockIdx.x * kOwn + lane;
  const int i = iraw <= nx + 1 ? iraw : nx + 1;  // lanes past the mesh re-read the last ghost column
  const bool owned = lane >= 1 && lane <= kOwn && iraw <= nx;
  const int j = blockIdx.y + 1;
  const IX e = i + NXG * j;
  const int ifx = min(max(i, 1), nx + 1) - 1;  // x face left of element i (clamped for the lanes that own none)
  const int ify = min(max(i, 1), nx) - 1;      // column of the y faces
  const IX s1 = (IX)(nx + 1) * ny, s2 = (IX)nx * (ny + 1), sfp = (IX)nx * ny;
  const IX i1 = ifx + (IX)(nx + 1) * (j - 1), i2 = ify + (IX)nx * (j - 1);
  const double gm1 = gamma - 1.0;
  double *T = tile + lane;

  // ---- every global load of the x pass and of the y faces first: one exposed DRAM latency per block
  double a[NSP][4];  // the metric of point row l: a11, a21, a12, a22
#pragma unroll
  for (int k = 0; k < NSP; ++k)
#pragma unroll
    for (int m = 0; m < 4; ++m) a[k][m] = owned ? g.iJ[e + NE * plane<NSP>(k, l, m)] : 0.0;
  const double n1c = g.n1[i1], n1s = g.n1[i1 + s1];
  const double nbc = g.n2[i2], nbs = g.n2[i2 + s2], ntc = g.n2[i2 + nx], nts = g.n2[i2 + nx + s2];

  // ---- y faces j (below) and j+1 (above) at flux point p = l
  {
    double vm[NSP][4], v0[NSP][4], vp[NSP][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int q = 0; q < NSP; ++q) {
        const IX o = NE * plane<NSP>(l, q, m);
        vm[q][m] = u[e - NXG + o];
        v0[q][m] = u[e + o];
        vp[q][m] = u[e + NXG + o];
      }
    double tm[4], b0[4], t0[4], bp[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double x = 0, b = 0, c = 0, d = 0;  // the summation order of load_trace_y
#pragma unroll
      for (int q = 0; q < NSP; ++q) {
        T[plane<NSP>(l, q, m) * 32] = v0[q][m];  // point (k = l, row q) of the element
        x = fma(vm[q][m], ops.lr[q], x);
        b = fma(v0[q][m], ops.ll[q], b);
        c = fma(v0[q][m], ops.lr[q], c);
        d = fma(vp[q][m], ops.ll[q], d);
      }
      tm[m] = x; b0[m] = b; t0[m] = c; bp[m] = d;
    }
    const W4 hb = frbcurv::flux_normal_t<FLUX>(g.flux, {tm[0], tm[1], tm[2], tm[3]}, {b0[0], b0[1], b0[2], b0[3]},
                                               nbc, nbs, gamma, 0);
    const W4 ht = frbcurv::flux_normal_t<FLUX>(g.flux, {t0[0], t0[1], t0[2], t0[3]}, {bp[0], bp[1], bp[2], bp[3]},
                                               ntc, nts, gamma, 0);
    double *F = fyt + lane;
    F[((0 * NSP + l) * 4 + 0) * 32] = hb.a; F[((0 * NSP + l) * 4 + 1) * 32] = hb.b;
    F[((0 * NSP + l) * 4 + 2) * 32] = hb.c; F[((0 * NSP + l) * 4 + 3) * 32] = hb.d;
    F[((1 * NSP + l) * 4 + 0) * 32] = ht.a; F[((1 * NSP + l) * 4 + 1) * 32] = ht.b;
    F[((1 * NSP + l) * 4 + 2) * 32] = ht.c; F[((1 * NSP + l) * 4 + 3) * 32] = ht.d;
  }
  __syncthreads();  // the row of elements is in the tile

  // ---- x pass of point row l (the formulas of frbcurv::row_xpass, FOLD form), every lane: the edge lanes supply
  // traces and the flux of their left face
  double d[NSP][4], w[NSP][4], un[NSP][4], gyL[NSP], gyR[NSP];
  {
    double cxL[NSP], cxR[NSP];
#pragma unroll
    for (int k = 0; k < NSP; ++k)
#pragma unroll
      for (int m = 0; m < 4; ++m) w[k][m] = T[plane<NSP>(k, l, m) * 32];
    if (g.fpc) {  // cylinder2.jl:155-158
      const IX ifp = (IX)(owned ? i - 1 : 0) + (IX)nx * (j - 1);
      const double xl = g.fpc[ifp + sfp * (l + NSP * 0)], xr = g.fpc[ifp + sfp * (l + NSP * 1)];
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        cxL[k] = xl * ops.dgl[k];
        cxR[k] = xr * ops.dgr[k];
        gyL[k] = g.fpc[ifp + sfp * (k + NSP * 2)] * ops.dgl[l];
        gyR[k] = g.fpc[ifp + sfp * (k + NSP * 3)] * ops.dgr[l];
      }
    } else {  // parallelogram.jl:145-148
      const double nrc = __shfl_down_sync(0xffffffffu, n1c, 1), nrs = __shfl_down_sync(0xffffffffu, n1s, 1);
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        cxL[k] = fma(a[k][2], n1s, a[k][0] * n1c) * ops.dgl[k];
        cxR[k] = fma(a[k][2], nrs, a[k][0] * nrc) * ops.dgr[k];
        gyL[k] = fma(a[k][3], nbs, a[k][1] * nbc) * ops.dgl[l];
        gyR[k] = fma(a[k][3], nts, a[k][1] * ntc) * ops.dgr[l];
      }
    }
    double tl[4], tr[4], FxL[4], FxR[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double x = 0, y = 0;  // the summation order of load_trace_x
#pragma unroll
      for (int q = 0; q < NSP; ++q) {
        x = fma(w[q][m], ops.ll[q], x);
        y = fma(w[q][m], ops.lr[q], y);
      }
      tl[m] = x;
      tr[m] = __shfl_up_sync(0xffffffffu, y, 1);  // u_face[i-1, j, 2, l, m]
    }
    {
      const W4 hx = frbcurv::flux_normal_t<FLUX>(g.flux, {tr[0], tr[1], tr[2], tr[3]}, {tl[0], tl[1], tl[2], tl[3]},
                                                 n1c, n1s, gamma, g.wall_xlo && i == 1);
      FxL[0] = hx.a; FxL[1] = hx.b; FxL[2] = hx.c; FxL[3] = hx.d;
#pragma unroll
      for (int m = 0; m < 4; ++m) FxR[m] = __shfl_down_sync(0xffffffffu, FxL[m], 1);
    }
    double f1[NSP][4];
#pragma unroll
    for (int k = 0; k < NSP; ++k) {
      const double w0 = w[k][0], w1 = w[k][1], w2 = w[k][2], w3 = w[k][3];
      const double r = frbcurv::rcp(w0), vx = w1 * r, vy = w2 * r;
      const double p = gm1 * (w3 - 0.5 * fma(w1, vx, w2 * vy));
      const double h = w3 + p;
      const double F[4] = {w1, fma(w1, vx, p), w1 * vy, h * vx};
      const double G[4] = {w2, w2 * vx, fma(w2, vy, p), h * vy};
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        f1[k][m] = fma(a[k][2], G[m], a[k][0] * F[m]);
        T[plane<NSP>(k, l, m) * 32] = fma(a[k][3], G[m], a[k][1] * F[m]);  // f2 over the state value just read
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        double x = f1[0][m] * ops.dmod[k * FRB_NSPMAX];
#pragma unroll
        for (int q = 1; q < NSP; ++q) x = fma(f1[q][m], ops.dmod[k * FRB_NSPMAX + q], x);
        x = fma(cxL[k], FxL[m], x);
        d[k][m] = fma(cxR[k], FxR[m], x);
      }
#pragma unroll
    for (int k = 0; k < NSP; ++k)
#pragma unroll
      for (int m = 0; m < 4; ++m) un[k][m] = (st.use_a && owned) ? ua[e + NE * plane<NSP>(k, l, m)] : 0.0;
  }
  __syncthreads();  // the tile holds f2

  // ---- y pass (the formulas of frbcurv::row_ypass, FOLD form) and the stage update
  if (owned) {
    const double *F = fyt + lane;
#pragma unroll
    for (int k = 0; k < NSP; ++k) {
      const int yi = g.fy_row ? l : k;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double x = d[k][m];
#pragma unroll
        for (int q = 0; q < NSP; ++q) x = fma(T[plane<NSP>(k, q, m) * 32], ops.dmod[l * FRB_NSPMAX + q], x);
        x = fma(gyL[k], F[((0 * NSP + yi) * 4 + m) * 32], x);
        x = fma(gyR[k], F[((1 * NSP + yi) * 4 + m) * 32], x);
        x = -x;
        out[e + NE * plane<NSP>(k, l, m)] = fma(st.ca, un[k][m], fma(st.cdt, x, st.cb * w[k][m]));
      }
    }
  }
}

}  // namespace

// st: the branch-free form of frb_launch_euler2d_curv (rhs_only and nested already mapped onto ca, cb, cdt)
int frb_launch_euler2d_curv_fused(frb_prob_t p, const double *u, const double *ua, double *out, const CurvGeom &g,
                                  const FrbStage &st) {
  const bool ix32 = (double)(p->nx + 2) * (p->ny + 2) * p->nsp * p->nsp * 4 < 4294967296.0;
  cudaStream_t s = p->ctx->stream;
  dim3 blk(32, p->nsp), grd((p->nx + kOwn - 1) / kOwn, p->ny);
  const int key = p->nsp * 100 + (ix32 ? 10 : 0) + g.flux;
  switch (key) {
#define FRB_CURV_CASE(N, IX, I, F)                                                                      \
  case N * 100 + I * 10 + F:                                                                            \
    euler2d_curv_fused_kernel<N, IX, F><<<grd, blk, 0, s>>>(u, ua, out, g, p->gamma, p->ops, st);       \
    break;
#define FRB_CURV_CASES(N)                                                                               \
  FRB_CURV_CASE(N, unsigned, 1, 0) FRB_CURV_CASE(N, unsigned, 1, 1) FRB_CURV_CASE(N, unsigned, 1, 2)   \
  FRB_CURV_CASE(N, size_t, 0, 0) FRB_CURV_CASE(N, size_t, 0, 1) FRB_CURV_CASE(N, size_t, 0, 2)
    FRB_CURV_CASES(2)
    FRB_CURV_CASES(3)
    FRB_CURV_CASES(4)
#undef FRB_CURV_CASES
#undef FRB_CURV_CASE
    default: frb_set_error("euler2d_curv: deg must be in 1..3, flux HLL / LF / ROE"); return FRB_ERR_ARG;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "euler2d_curv_fused_kernel", __FILE__, __LINE__);
  return 1;
}
