// 2-D Euler residual + RK stage of ONE element of a curvilinear structured quadrilateral mesh
// (SURVEY 8f-2): metric terms per solution point, unit normals per face, correction factors either from
// the solution-point inverse Jacobian or from a flux-point table.
//
// Reference semantics: dudt! of dev/parallelogram.jl:80-165 (factors (iJ[i,j][k,l] * n)[c], :145-148) and
// of dev/cylinder2.jl:52-164 (factors (inv(Ji[i,j][face,pt]) * n)[c], :155-158; mirror wall on the inner
// radial face, :100-120).  Both scripts index the y common flux with the row index l in the correction
// (parallelogram.jl:147-148) where the rectangular scripts use the flux-point index k
// (example/euler2d_wave.jl:100-103); CurvGeom::fy_row selects the scripts' form.
//
// Two passes per stage (the split the north star words: a face kernel over the face connectivity, an
// element kernel that fuses derivative, correction and the RK update):
//   face_xy          thread = (element, flux point p): the x face left of the element and the y face below it --
//                    traces from the neighbouring element blocks, common flux in the face frame, all 4
//                    components -> fx[nx+1, ny, nsp, 4], fy[nx, ny+1, nsp, 4]
//   row_xpass /      thread = (element, point row l), a block = 32 elements x NSP rows: point fluxes iJ [F; G] of
//   row_ypass        the row, r-derivative + x corrections in registers, f2 through a shared-memory tile, then the
//                    s-derivative + y corrections + stage update of the same row
//
// The routines are __host__ __device__ so that tests/harness/curv_host.cu can run the very same code on the
// CPU against the NumPy oracle (the build box has no GPU); the library only ever calls it from a kernel.
//
// Layouts (Julia column-major, one ghost ring, NE = (nx+2)(ny+2), e = i + (nx+2) j):
//   u  [nx+2, ny+2, nsp, nsp, 4]     u[e + NE (k + nsp (l + nsp m))]
//   iJ [nx+2, ny+2, nsp, nsp, 2, 2]  iJ[e + NE (k + nsp (l + nsp (a + 2 b)))] = ps.iJ[i,j][k,l][a,b]
//   n1 [nx+1, ny, 2]  unit normal of x face i (between elements i-1 and i) in row j: n1[i-1 + (nx+1)(j-1 + ny c)]
//   n2 [nx, ny+1, 2]  unit normal of y face j (between rows j-1 and j) in column i:  n2[i-1 + nx (j-1 + (ny+1) c)]
//   fpc[nx, ny, nsp, 4] (optional) flux-point factors (xL by l, xR by l, yL by k, yR by k):
//                     fpc[i-1 + nx (j-1 + ny (q + nsp c))]
#pragma once
#include <math.h>

#include "frb_internal.cuh"
#include "frb_physics.cuh"

#if defined(__CUDACC__)
#define FRB_HD __host__ __device__ __forceinline__
#else
#define FRB_HD inline
#endif

struct CurvGeom {
  int nx, ny;
  const double *iJ, *n1, *n2, *fpc;  // fpc == nullptr: solution-point factors
  const double *vert;                // != nullptr: iJ is evaluated from the cell vertices [nx+2, ny+2, 4, 2]
  double r[FRB_NSPMAX];              // solution points ps.xpl (used with vert)
  int fy_row;    // 1: y common flux indexed by the row l (the scripts' literal form)
  int wall_xlo;  // 1: x face 1 is the mirror wall of dev/cylinder2.jl:100-120
  int flux;      // FRB_FLUX_HLL (the scripts) | LF | ROE
};

int frb_launch_euler2d_curv_march(frb_prob_t p, const double *u, const double *ua, double *out, const CurvGeom &g,
                                  const FrbStage &st);
int frb_launch_euler2d_curv_fused(frb_prob_t p, const double *u, const double *ua, double *out, const CurvGeom &g,
                                  const FrbStage &st);

namespace frbcurv {

struct W4 {
  double a, b, c, d;
};

// local_frame -> common flux (flux_hll!(fw, wL, wR, gamma, 1.0) in the scripts; LF / Roe as defined in
// DESIGN.md section 2) -> global_frame about the unit normal (c, s) (parallelogram.jl:119-123).  wall != 0
// replaces the left state by the mirror state of the right one in the face frame (cylinder2.jl:103-114).
FRB_HD W4 flux_normal(int kind, W4 L, W4 R, double c, double s, double gamma, int wall) {
  const double gm1 = gamma - 1.0;
  double l0 = L.a, l1 = fma(L.c, s, L.b * c), l2 = fma(-L.b, s, L.c * c), l3 = L.d;
  const double r0 = R.a, r1 = fma(R.c, s, R.b * c), r2 = fma(-R.b, s, R.c * c), r3 = R.d;
  if (wall) {
    // prim = conserve_prim(ul) = (rho, U, V, lambda);  pn = ((1-t)/(1+t) rho, -U, V, 2 - lambda), t = lambda - 1
    const double U = r1 / r0, V = r2 / r0;
    const double lam = 0.5 * r0 / gm1 / (r3 - 0.5 * (r1 * r1 + r2 * r2) / r0);
    const double t = lam - 1.0;
    const double rn = (1.0 - t) / (1.0 + t) * r0, ln = 2.0 - lam;
    l0 = rn;
    l1 = rn * (-U);
    l2 = rn * V;
    l3 = 0.5 * rn / ln / gm1 + 0.5 * rn * (U * U + V * V);
  }
#if defined(__CUDA_ARCH__)
  // HLL on the device: the branch-free form of the roofline kernels (MUFU-seeded reciprocal / square root,
  // selects instead of early returns; within ~2 ulp of the IEEE form the host harness runs)
  const frb::Flux4 f = kind == FRB_FLUX_HLL ? frb::hll4_fast(l0, l1, l2, l3, r0, r1, r2, r3, gamma, gm1)
                                            : frb::riemann4(kind, l0, l1, l2, l3, r0, r1, r2, r3, gamma);
#else
  const frb::Flux4 f = frb::riemann4(kind, l0, l1, l2, l3, r0, r1, r2, r3, gamma);
#endif
  return {f.f0, fma(-f.f2, s, f.f1 * c), fma(f.f1, s, f.f2 * c), f.f3};
}

// the same with the flux a compile-time choice (FLUX < 0: g.flux at run time): on the device every common flux in
// its branch-free form (riemann4_fast) and the mirror state with MUFU-seeded reciprocals
template <int FLUX>
FRB_HD W4 flux_normal_t(int kind, W4 L, W4 R, double c, double s, double gamma, int wall) {
#if defined(__CUDA_ARCH__)
  if (FLUX < 0) return flux_normal(kind, L, R, c, s, gamma, wall);
  const double gm1 = gamma - 1.0;
  double l0 = L.a, l1 = fma(L.c, s, L.b * c), l2 = fma(-L.b, s, L.c * c), l3 = L.d;
  const double r0 = R.a, r1 = fma(R.c, s, R.b * c), r2 = fma(-R.b, s, R.c * c), r3 = R.d;
  if (wall) {
    const double ir = frb::rcp_fast(r0), U = r1 * ir, V = r2 * ir;
    const double lam = 0.5 * r0 * frb::rcp_fast(gm1 * (r3 - 0.5 * (r1 * r1 + r2 * r2) * ir));
    const double t = lam - 1.0;
    const double rn = (1.0 - t) * frb::rcp_fast(1.0 + t) * r0, ln = 2.0 - lam;
    l0 = rn;
    l1 = rn * (-U);
    l2 = rn * V;
    l3 = 0.5 * rn * frb::rcp_fast(ln * gm1) + 0.5 * rn * (U * U + V * V);
  }
  const frb::Flux4 f = frb::riemann4_fast<(FLUX < 0 ? 0 : FLUX)>(l0, l1, l2, l3, r0, r1, r2, r3, gamma, gm1);
  return {f.f0, fma(-f.f2, s, f.f1 * c), fma(f.f1, s, f.f2 * c), f.f3};
#else
  return flux_normal(FLUX < 0 ? kind : FLUX, L, R, c, s, gamma, wall);
#endif
}

template <int NSP>
FRB_HD int plane(int k, int l, int m) {
  return k + NSP * (l + NSP * m);
}

template <int NSP, typename IX>
FRB_HD void load_trace_x(const double *__restrict__ u, IX e, IX NE, int p, const double *lq, double w[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {  // dot(u[i,j,:,p,m], lq)
    double a = 0;
#pragma unroll
    for (int q = 0; q < NSP; ++q) a = fma(u[e + NE * plane<NSP>(q, p, m)], lq[q], a);
    w[m] = a;
  }
}
template <int NSP, typename IX>
FRB_HD void load_trace_y(const double *__restrict__ u, IX e, IX NE, int p, const double *lq, double w[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {  // dot(u[i,j,p,:,m], lq)
    double a = 0;
#pragma unroll
    for (int q = 0; q < NSP; ++q) a = fma(u[e + NE * plane<NSP>(p, q, m)], lq[q], a);
    w[m] = a;
  }
}

// Common fluxes: fx[i-1 + (nx+1)(j-1 + ny (p + NSP m))] for x face i (1..nx+1) of row j (parallelogram.jl:115-125),
//                fy[i-1 + nx (j-1 + (ny+1)(p + NSP m))] for y face j (1..ny+1) of column i (:126-136).
// The x face left of element (i, j) and the y face below it, flux point p, in one go: the four traces first
// (64 loads without control flow in between, clamped to valid cells for the threads that own only one of the two
// faces), then the two common fluxes.  do_x / do_y say which results exist (i <= nx + 1, j <= ny / i <= nx, j <= ny + 1).
// IX: the index type (size_t; unsigned when every array stays below 2^32 elements: a third fewer instructions
// in kernels that are bound by instruction issue).  FLUX: see flux_normal_t.
template <int NSP, typename IX = size_t, int FLUX = -1>
FRB_HD void face_xy(int i, int j, int p, bool do_x, bool do_y, const double *__restrict__ u,
                    double *__restrict__ fx, double *__restrict__ fy, const CurvGeom &g, double gamma,
                    const FrbOps &ops) {
  const IX NXG = g.nx + 2, NE = NXG * (IX)(g.ny + 2);
  const int jx = j <= g.ny ? j : g.ny, iy = i <= g.nx ? i : g.nx;
  const IX ex = i + NXG * jx, ey = iy + NXG * j;
  double xl[4], xr[4], yl[4], yr[4];
  load_trace_x<NSP, IX>(u, ex - 1, NE, p, ops.lr, xl);    // u_face[i-1, j, 2, p, :]
  load_trace_x<NSP, IX>(u, ex, NE, p, ops.ll, xr);        // u_face[i, j, 4, p, :]
  load_trace_y<NSP, IX>(u, ey - NXG, NE, p, ops.lr, yl);  // u_face[i, j-1, 3, p, :]
  load_trace_y<NSP, IX>(u, ey, NE, p, ops.ll, yr);        // u_face[i, j, 1, p, :]
  const IX f1 = (IX)(i - 1) + (IX)(g.nx + 1) * (jx - 1), s1 = (IX)(g.nx + 1) * g.ny;
  const IX f2 = (IX)(iy - 1) + (IX)g.nx * (j - 1), s2 = (IX)g.nx * (g.ny + 1);
  const double c1 = g.n1[f1], d1 = g.n1[f1 + s1], c2 = g.n2[f2], d2 = g.n2[f2 + s2];
  const W4 hx = flux_normal_t<FLUX>(g.flux, {xl[0], xl[1], xl[2], xl[3]}, {xr[0], xr[1], xr[2], xr[3]}, c1, d1,
                                    gamma, g.wall_xlo && i == 1);
  const W4 hy = flux_normal_t<FLUX>(g.flux, {yl[0], yl[1], yl[2], yl[3]}, {yr[0], yr[1], yr[2], yr[3]}, c2, d2,
                                    gamma, 0);
  if (do_x) {
    fx[f1 + s1 * (p + NSP * 0)] = hx.a;
    fx[f1 + s1 * (p + NSP * 1)] = hx.b;
    fx[f1 + s1 * (p + NSP * 2)] = hx.c;
    fx[f1 + s1 * (p + NSP * 3)] = hx.d;
  }
  if (do_y) {
    fy[f2 + s2 * (p + NSP * 0)] = hy.a;
    fy[f2 + s2 * (p + NSP * 1)] = hy.b;
    fy[f2 + s2 * (p + NSP * 2)] = hy.c;
    fy[f2 + s2 * (p + NSP * 3)] = hy.d;
  }
}

// 1 / x: on the device the branch-free MUFU seed + two Newton steps of frb_physics.cuh (within ~1 ulp; no
// slow-path call that would split the basic block and serialise the loads around it), IEEE on the host.
FRB_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
  return frb::rcp_fast(x);
#else
  return 1.0 / x;
#endif
}

// What a row owner carries from its x pass to its y pass (registers on the device).
template <int NSP>
struct RowCarry {
  double d[NSP][4];          // rhs1 + x-face corrections at (k, l = row), per variable
  double w[NSP][4];          // the row's state values (the stage update needs them again)
  double cyL[NSP], cyR[NSP]; // y correction factors at (k, l)
};

// Element (i, j), point row l -- x pass.  Every global load of the pass is issued before the first use
// (state, metric, common fluxes, normals / factors: one exposed DRAM latency per thread), then: point fluxes
// iJ [F; G] of the row's NSP points (parallelogram.jl:88-96), the r-derivative and the x-face corrections of f1
// in registers (:138-142,150-157), f2 of the row into the element's tile for the y pass,
// tile[((l NSP + k) 4 + m) ts] (ts = 32 lanes on the device, 1 on the host), and -- one flux point per row
// owner -- the y common fluxes into fyt[((side NSP + p) 4 + m) ts].
// FOLD: the flux traces of the correction folded into the derivative matrix (FrbOps::dmod, as in the rectangular
// marching kernels): sum_q lpdm[k][q] f[q] + (c F - f.ll) dgl[k] + (c F - f.lr) dgr[k]
//                  = sum_q dmod[k][q] f[q] + c dgl[k] F + c dgr[k] F  -- a fifth fewer FP64 instructions per row.
template <int NSP, typename IX = size_t, bool FOLD = false>
FRB_HD void row_xpass(int i, int j, int l, const double *__restrict__ u, const double *__restrict__ fx,
                      const double *__restrict__ fy, const CurvGeom &g, double gamma, const FrbOps &ops,
                      double *__restrict__ tile, double *__restrict__ fyt, int ts, RowCarry<NSP> &c) {
  const int nx = g.nx, ny = g.ny;
  const IX NXG = nx + 2, NE = NXG * (IX)(ny + 2);
  const IX e = i + NXG * j;
  const double gm1 = gamma - 1.0;
  const IX i1 = (IX)(i - 1) + (IX)(nx + 1) * (j - 1), s1 = (IX)(nx + 1) * ny;
  const IX i2 = (IX)(i - 1) + (IX)nx * (j - 1), s2 = (IX)nx * (ny + 1);
  const IX ifp = (IX)(i - 1) + (IX)nx * (j - 1), sfp = (IX)nx * ny;

  // ---- loads
  double a[NSP][4], FxL[4], FxR[4], FyB[4], FyT[4], cxL[NSP], cxR[NSP];
#pragma unroll
  for (int k = 0; k < NSP; ++k)
#pragma unroll
    for (int m = 0; m < 4; ++m) c.w[k][m] = u[e + NE * plane<NSP>(k, l, m)];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    FxL[m] = fx[i1 + s1 * (l + NSP * m)];
    FxR[m] = fx[i1 + 1 + s1 * (l + NSP * m)];
    FyB[m] = fy[i2 + s2 * (l + NSP * m)];
    FyT[m] = fy[i2 + nx + s2 * (l + NSP * m)];
  }
  if (g.vert) {
    // iJ[i,j][k,l] = inv(rs_jacobi(r_k, r_l, vertices[i,j])) (struct.jl:135-142, geo_jacobi.jl:77-88) from the 8
    // vertex coordinates of the element instead of 4 stored doubles per solution point
    double vx[4], vy[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      vx[q] = g.vert[e + NE * q];
      vy[q] = g.vert[e + NE * (q + 4)];
    }
    double rl = 0.0;  // g.r[l] by selects: a dynamic index would move the parameter array to local memory
#pragma unroll
    for (int q = 0; q < NSP; ++q) rl = q == l ? g.r[q] : rl;
    const double sm = rl - 1.0, sp = rl + 1.0;
    const double xr = 0.25 * (sm * vx[0] - sm * vx[1] + sp * vx[2] - sp * vx[3]);
    const double yr = 0.25 * (sm * vy[0] - sm * vy[1] + sp * vy[2] - sp * vy[3]);
#pragma unroll
    for (int k = 0; k < NSP; ++k) {
      const double rm = g.r[k] - 1.0, rp = g.r[k] + 1.0;
      const double xs = 0.25 * (rm * vx[0] - rp * vx[1] + rp * vx[2] - rm * vx[3]);
      const double ys = 0.25 * (rm * vy[0] - rp * vy[1] + rp * vy[2] - rm * vy[3]);
      const double id = rcp(xr * ys - xs * yr);
      a[k][0] = ys * id;   // a11
      a[k][1] = -yr * id;  // a21
      a[k][2] = -xs * id;  // a12
      a[k][3] = xr * id;   // a22
    }
  } else {
#pragma unroll
    for (int k = 0; k < NSP; ++k)
#pragma unroll
      for (int m = 0; m < 4; ++m) a[k][m] = g.iJ[e + NE * plane<NSP>(k, l, m)];  // a11, a21, a12, a22
  }
  if (g.fpc) {  // cylinder2.jl:155-158
    const double xl = g.fpc[ifp + sfp * (l + NSP * 0)], xr = g.fpc[ifp + sfp * (l + NSP * 1)];
#pragma unroll
    for (int k = 0; k < NSP; ++k) {
      cxL[k] = xl;
      cxR[k] = xr;
      c.cyL[k] = g.fpc[ifp + sfp * (k + NSP * 2)];
      c.cyR[k] = g.fpc[ifp + sfp * (k + NSP * 3)];
    }
  } else {  // parallelogram.jl:145-148
    const double nxl_c = g.n1[i1], nxl_s = g.n1[i1 + s1], nxr_c = g.n1[i1 + 1], nxr_s = g.n1[i1 + 1 + s1];
    const double nyb_c = g.n2[i2], nyb_s = g.n2[i2 + s2], nyt_c = g.n2[i2 + nx], nyt_s = g.n2[i2 + nx + s2];
#pragma unroll
    for (int k = 0; k < NSP; ++k) {
      cxL[k] = fma(a[k][2], nxl_s, a[k][0] * nxl_c);
      cxR[k] = fma(a[k][2], nxr_s, a[k][0] * nxr_c);
      c.cyL[k] = fma(a[k][3], nyb_s, a[k][1] * nyb_c);
      c.cyR[k] = fma(a[k][3], nyt_s, a[k][1] * nyt_c);
    }
  }

  // ---- point fluxes
  double f1[NSP][4];
#pragma unroll
  for (int k = 0; k < NSP; ++k) {
    const double w0 = c.w[k][0], w1 = c.w[k][1], w2 = c.w[k][2], w3 = c.w[k][3];
    const double r = rcp(w0), vx = w1 * r, vy = w2 * r;
    const double p = gm1 * (w3 - 0.5 * fma(w1, vx, w2 * vy));
    const double h = w3 + p;
    const double F[4] = {w1, fma(w1, vx, p), w1 * vy, h * vx};
    const double G[4] = {w2, w2 * vx, fma(w2, vy, p), h * vy};
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      f1[k][m] = fma(a[k][2], G[m], a[k][0] * F[m]);
      tile[(size_t)(((l * NSP + k) * 4) + m) * ts] = fma(a[k][3], G[m], a[k][1] * F[m]);
    }
  }
  if (FOLD) {
#pragma unroll
    for (int k = 0; k < NSP; ++k) {
      cxL[k] *= ops.dgl[k];
      cxR[k] *= ops.dgr[k];
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    if (FOLD) {
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        double d = f1[0][m] * ops.dmod[k * FRB_NSPMAX];
#pragma unroll
        for (int q = 1; q < NSP; ++q) d = fma(f1[q][m], ops.dmod[k * FRB_NSPMAX + q], d);
        d = fma(cxL[k], FxL[m], d);
        c.d[k][m] = fma(cxR[k], FxR[m], d);
      }
    } else {
      double t4 = 0, t2 = 0;  // f_face[i,j,4,l,m,1], f_face[i,j,2,l,m,1]
#pragma unroll
      for (int q = 0; q < NSP; ++q) {
        t4 = fma(f1[q][m], ops.ll[q], t4);
        t2 = fma(f1[q][m], ops.lr[q], t2);
      }
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        double d = f1[0][m] * ops.lpdm[k * FRB_NSPMAX];
#pragma unroll
        for (int q = 1; q < NSP; ++q) d = fma(f1[q][m], ops.lpdm[k * FRB_NSPMAX + q], d);
        d += (cxL[k] * FxL[m] - t4) * ops.dgl[k];
        d += (cxR[k] * FxR[m] - t2) * ops.dgr[k];
        c.d[k][m] = d;
      }
    }
    fyt[(size_t)(((0 * NSP + l) * 4) + m) * ts] = FyB[m];
    fyt[(size_t)(((1 * NSP + l) * 4) + m) * ts] = FyT[m];
  }
}

// y pass of the same row owner, after every row of the element has written its part of the tiles:
// s-derivative (:143-149), y-face corrections (:158-163; the common flux indexed by k, or by l in the scripts'
// literal form) and the stage update out = ca u_n + cb u + cdt L(u) (the launcher maps rhs_only onto
// ca = cb = 0, cdt = 1, so there is no branch here; u_n is fetched first, as one batch).
template <int NSP, typename IX = size_t, bool FOLD = false>
FRB_HD void row_ypass(int i, int j, int l, const double *__restrict__ ua, double *__restrict__ out,
                      const CurvGeom &g, const FrbOps &ops, const FrbStage &st, const double *__restrict__ tile,
                      const double *__restrict__ fyt, int ts, const RowCarry<NSP> &c) {
  const IX NXG = g.nx + 2, NE = NXG * (IX)(g.ny + 2);
  const IX e = i + NXG * j;
  double un[NSP][4];
#pragma unroll
  for (int k = 0; k < NSP; ++k)
#pragma unroll
    for (int m = 0; m < 4; ++m) un[k][m] = st.use_a ? ua[e + NE * plane<NSP>(k, l, m)] : 0.0;
#pragma unroll
  for (int k = 0; k < NSP; ++k) {
    const int yi = g.fy_row ? l : k;
    const double gyL = c.cyL[k] * ops.dgl[l], gyR = c.cyR[k] * ops.dgr[l];  // FOLD
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double FyB = fyt[(size_t)(((0 * NSP + yi) * 4) + m) * ts];
      const double FyT = fyt[(size_t)(((1 * NSP + yi) * 4) + m) * ts];
      double d = c.d[k][m];
      if (FOLD) {
#pragma unroll
        for (int q = 0; q < NSP; ++q)
          d = fma(tile[(size_t)(((q * NSP + k) * 4) + m) * ts], ops.dmod[l * FRB_NSPMAX + q], d);
        d = fma(gyL, FyB, d);
        d = fma(gyR, FyT, d);
      } else {
        double b = 0, t1 = 0, t3 = 0;  // rhs2, f_face[i,j,1,k,m,2], f_face[i,j,3,k,m,2]
#pragma unroll
        for (int q = 0; q < NSP; ++q) {
          const double f2 = tile[(size_t)(((q * NSP + k) * 4) + m) * ts];
          b = fma(f2, ops.lpdm[l * FRB_NSPMAX + q], b);
          t1 = fma(f2, ops.ll[q], t1);
          t3 = fma(f2, ops.lr[q], t3);
        }
        d += b;
        d += (c.cyL[k] * FyB - t1) * ops.dgl[l];
        d += (c.cyR[k] * FyT - t3) * ops.dgr[l];
      }
      d = -d;
      out[e + NE * plane<NSP>(k, l, m)] = fma(st.ca, un[k][m], fma(st.cdt, d, st.cb * c.w[k][m]));
    }
  }
}

}  // namespace frbcurv
