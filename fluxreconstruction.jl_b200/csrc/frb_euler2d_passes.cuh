// The arithmetic shared by the two marching stage kernels of the 2-D Euler path
// (frb_euler2d_rc.cu on the row-chunk layout, frb_euler2d_march.cu on the reference image).
// Both keep a row of a 30-element strip in a shared-memory tile laid out [plane][32 lanes],
// plane(k, l, m) = k + NSP*(l + NSP*m), and run
//   x pass  warp t = point row l   : F, x traces, left-face HLL (neighbour trace by shuffle),
//                                    cb*u + (-dt/Jx)(d/dr + correction)  -> xd;  v_y, p -> xrp
//   y pass  warp t = point column k: G, y traces, top-face HLL, the y chain on top of xd
// Reference: dudt! of example/euler2d_wave.jl:35-107.  Views: Ux = tile + 32*NSP*t + lane (row view),
// Uy = tile + 32*t + lane (column view); xd / xrp likewise.
#pragma once
#include "frb_physics.cuh"
#include "frb_ptx.cuh"

namespace frbpass {

// y traces of one column (k = t) of a tile: top (lr) or bottom (ll)
template <int NSP>
__device__ __forceinline__ void col_trace(const double *__restrict__ Uy, const double *l, double (&tr)[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double a = Uy[32 * NSP * (0 + NSP * m)] * l[0];
#pragma unroll
    for (int q = 1; q < NSP; ++q) a = fma(Uy[32 * NSP * (q + NSP * m)], l[q], a);
    tr[m] = a;
  }
}

// common flux on the face between row j (tile Ulo, top trace) and row j+1 (tile Uhi, bottom trace)
// FLUX: the common flux (FRB_FLUX_HLL = the reference's flux_hll!, LF, ROE), a compile-time choice
template <int NSP, int FLUX = 0>
__device__ __forceinline__ void face_flux_y(const double (&uT)[4], const double *__restrict__ Uhi,
                                            const MarchOps &ops, double gamma, double gm1, double (&h)[4]) {
  double uB[4];
  col_trace<NSP>(Uhi, ops.ll, uB);
  frb::Flux4 f = frb::riemann4_y_fast<FLUX>(uT[0], uT[1], uT[2], uT[3], uB[0], uB[1], uB[2], uB[3], gamma, gm1);
  h[0] = f.f0; h[1] = f.f1; h[2] = f.f2; h[3] = f.f3;
}

// The same face for a march in either direction (row-chunk kernel: odd segments march downwards so that the two
// segments sharing a boundary read its rows at the same time and the second read is an L2 hit).  uOwn = trace of
// the current row on the face towards the next row of the march, Unb = tile of that next row, lnb = the operator
// of ITS trace on the face (ll when it lies above, lr when below).  The lower row is always the left state, so the
// flux is bit for bit the one the upward march computes.
template <int NSP, int FLUX = 0>
__device__ __forceinline__ void face_flux_y_dir(const double (&uOwn)[4], const double *__restrict__ Unb,
                                                const double *lnb, bool down, double gamma, double gm1,
                                                double (&h)[4]) {
  double uN[4];
  col_trace<NSP>(Unb, lnb, uN);
  double L[4], R[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) { L[m] = down ? uN[m] : uOwn[m]; R[m] = down ? uOwn[m] : uN[m]; }
  frb::Flux4 f = frb::riemann4_y_fast<FLUX>(L[0], L[1], L[2], L[3], R[0], R[1], R[2], R[3], gamma, gm1);
  h[0] = f.f0; h[1] = f.f1; h[2] = f.f2; h[3] = f.f3;
}

// CB1: the stage has cb == 1 (no multiply)
template <int NSP, bool CB1, int FLUX = 0>
__device__ __forceinline__ void x_pass(const double *__restrict__ Ux, double *__restrict__ xdx,
                                       double *__restrict__ xrpx, const MarchOps &ops, double cb, double gamma,
                                       double gm1) {
  double w[NSP][4], f[NSP][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int k = 0; k < NSP; ++k) w[k][m] = Ux[32 * (k + NSP * NSP * m)];
#pragma unroll
  for (int k = 0; k < NSP; ++k) {
    double rr = frb::rcp_fast(w[k][0]);
    double vx = w[k][1] * rr, vy = w[k][2] * rr;
    double p = gm1 * fma(-0.5, fma(w[k][1], vx, w[k][2] * vy), w[k][3]);
    f[k][0] = w[k][1];
    f[k][1] = fma(w[k][1], vx, p);
    f[k][2] = w[k][1] * vy;
    f[k][3] = (w[k][3] + p) * vx;
    xrpx[32 * k] = vy;  // the y pass needs only v_y and p of the point
    xrpx[32 * (NSP * NSP + k)] = p;
  }
  double uL[4], uR[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double a = w[0][m] * ops.ll[0], b = w[0][m] * ops.lr[0];
#pragma unroll
    for (int q2 = 1; q2 < NSP; ++q2) {
      a = fma(w[q2][m], ops.ll[q2], a);
      b = fma(w[q2][m], ops.lr[q2], b);
    }
    uL[m] = a; uR[m] = b;
  }
  // left face: HLL(u_face[i-1,j,2,l,:], u_face[i,j,4,l,:])  (euler2d_wave.jl:69-74)
  double n0 = __shfl_up_sync(0xffffffffu, uR[0], 1), n1 = __shfl_up_sync(0xffffffffu, uR[1], 1);
  double n2 = __shfl_up_sync(0xffffffffu, uR[2], 1), n3 = __shfl_up_sync(0xffffffffu, uR[3], 1);
  frb::Flux4 hl = frb::riemann4_fast<FLUX>(n0, n1, n2, n3, uL[0], uL[1], uL[2], uL[3], gamma, gm1);
  const double hL[4] = {hl.f0, hl.f1, hl.f2, hl.f3};
  const double hR[4] = {__shfl_down_sync(0xffffffffu, hl.f0, 1), __shfl_down_sync(0xffffffffu, hl.f1, 1),
                        __shfl_down_sync(0xffffffffu, hl.f2, 1), __shfl_down_sync(0xffffffffu, hl.f3, 1)};
  // cb*u + (-cdt/Jx) * (d/dr + correction), flux traces folded into dmx (see MarchOps)
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int k = 0; k < NSP; ++k) {
      double d = CB1 ? w[k][m] : cb * w[k][m];
#pragma unroll
      for (int q2 = 0; q2 < NSP; ++q2) d = fma(f[q2][m], ops.dmx[k * 4 + q2], d);
      d = fma(hL[m], ops.glx[k], d);
      d = fma(hR[m], ops.grx[k], d);
      xdx[32 * (k + NSP * NSP * m)] = d;
    }
}

// first half of the y pass: G at the column's points and the top trace of the row
template <int NSP>
__device__ __forceinline__ void y_fluxes_dir(const double *__restrict__ Uy, const double *__restrict__ xrpy,
                                             const double *lown, double (&g)[NSP][4], double (&uT)[4]);

template <int NSP>
__device__ __forceinline__ void y_fluxes(const double *__restrict__ Uy, const double *__restrict__ xrpy,
                                         const MarchOps &ops, double (&g)[NSP][4], double (&uT)[4]) {
  y_fluxes_dir<NSP>(Uy, xrpy, ops.lr, g, uT);
}

// lown: the operator of the row's trace on the face towards the next row of the march (lr upwards, ll downwards)
template <int NSP>
__device__ __forceinline__ void y_fluxes_dir(const double *__restrict__ Uy, const double *__restrict__ xrpy,
                                             const double *lown, double (&g)[NSP][4], double (&uT)[4]) {
  double w[NSP][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int l = 0; l < NSP; ++l) w[l][m] = Uy[32 * NSP * (l + NSP * m)];
#pragma unroll
  for (int l = 0; l < NSP; ++l) {
    double vy = xrpy[32 * NSP * l];
    double p = xrpy[32 * NSP * (NSP + l)];
    g[l][0] = w[l][2];
    g[l][1] = w[l][1] * vy;
    g[l][2] = fma(w[l][2], vy, p);
    g[l][3] = (w[l][3] + p) * vy;
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double a = w[0][m] * lown[0];
#pragma unroll
    for (int q2 = 1; q2 < NSP; ++q2) a = fma(w[q2][m], lown[q2], a);
    uT[m] = a;
  }
}

// second half: value (l, m) of the stage = xd + (-dt/Jy)(d/ds + correction)
template <int NSP, bool SAMEJ>
__device__ __forceinline__ double y_value(double xd, const double (&g)[NSP][4], double hb, double ht,
                                          const MarchOps &ops, int l, int m) {
  double d = xd;
#pragma unroll
  for (int q2 = 0; q2 < NSP; ++q2) d = fma(g[q2][m], (SAMEJ ? ops.dmx : ops.dmy)[l * 4 + q2], d);
  d = fma(hb, (SAMEJ ? ops.glx : ops.gly)[l], d);
  d = fma(ht, (SAMEJ ? ops.grx : ops.gry)[l], d);
  return d;
}

}  // namespace frbpass
