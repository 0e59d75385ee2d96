// 2-D Navier-Stokes with the gas-kinetic (GKS) flux: dudt! + boundary! of
// example/ns_cavity.jl:147-344 with the in-script flux_gks! overloads (:49-145), fused with the
// explicit stage update.  State layout u[4, ns, nr, ny+2, nx+2] (variable fastest, element
// contiguous: 4*nsp^2 doubles; ns_cavity.jl:33).
//
// The path is FP64-compute-bound (erfc, exp, pow and ~2 k flops per interface point), not
// HBM-bound, so it is split the way the north star words it and the round trip of the common
// fluxes through HBM (268 MB at 1024^2, p3) is the cheap part:
//   ns_boundary_kernel   boundary!: mirrored isothermal-wall ghost states (per stage, as in :148)
//   ns_face_kernel       thread = (face, flux point): the two neighbours' traces and edge slopes
//                        (dll / dlr) and the two-state, time-averaged GKS flux -- each interface is
//                        evaluated ONCE, by a thread that owns nothing else
//   ns_elem_kernel       thread = (element, solution point): point fluxes (shared through smem
//                        inside the element), lpdm derivative, dgl/dgr correction with the stored
//                        common fluxes, 1/J, and the RK stage -- one 32-B row of the state per
//                        thread, fully coalesced.
// (The first version did everything with one thread per element: 255 registers, 2.6 KB of local
// memory, every interface flux computed by both neighbours -- 8.6 ms per stage at 1024^2 p3.)
//
// [KB] closures restated from KitBase.jl 0.9: gauss_moments, moments_conserve,
// moments_conserve_slope, pdf_slope, vhs_collision_time.  Reference quirks kept: the left
// cell's interface slope uses dll and the right cell's dlr (:213-214,:244-245); slopes are not
// rotated on y faces (:260); Mv of the LEFT state in the right state's time slope (:102).
#include "frb_internal.cuh"
#include "frb_physics.cuh"

namespace {

// Reciprocals: an IEEE FP64 division is a ~35-instruction subroutine on this part and the literal formulas
// divide ~45 times per interface point and ~27 times per solution point -- half of all executed instructions
// in the round-1 kernels (ncu: profiles/r02_cfg5.md).  Every state therefore carries 1/rho and 1/lambda, computed
// once with the MUFU-seeded Newton reciprocal of the 2-D Euler path (frb::rcp_fast, ~1 ulp), and the closures
// multiply.  The results move by an ulp or two: far inside the 1e-12 parity bound.
using frb::rcp_fast;

struct Prim4 {
  double rho, U, V, lam, ir, il;  // ir = 1 / rho, il = 1 / lambda
};
__device__ __forceinline__ Prim4 conserve_prim(const double *w, double gm1) {
  Prim4 p;
  p.rho = w[0];
  p.ir = rcp_fast(w[0]);
  p.U = w[1] * p.ir;
  p.V = w[2] * p.ir;
  p.il = 2.0 * gm1 * (w[3] - 0.5 * (w[1] * p.U + w[2] * p.V)) * p.ir;  // 1 / lambda = 2 (gamma - 1) rho e / rho
  p.lam = rcp_fast(p.il);
  return p;
}

// moments <u^n> of a Maxwellian: full (Mu), v (Mv) and the half-space ones (MuL: u>0, MuR: u<0)
struct Moments {
  double Mu[7], Mv[7], Mxi[3];
};
// (the recurrences multiply by h = 1/(2 lambda) instead of dividing by lambda: one reciprocal per
// state instead of fifteen IEEE divisions; well inside the 1e-12 parity bound)
__device__ __forceinline__ void moments_v(double V, double il, double *Mv) {
  const double h = 0.5 * il;
  Mv[0] = 1.0;
  Mv[1] = V;
#pragma unroll
  for (int i = 2; i <= 4; ++i) Mv[i] = V * Mv[i - 1] + (i - 1) * h * Mv[i - 2];  // <v^5>, <v^6> are never used
}
__device__ __forceinline__ void moments_half(double U, double lam, double il, double *MuL, double *MuR) {
  // one erfc per state: erfc of the non-negative argument (accurate where it is small), its mirror 2 - erfc
  const double sl = sqrt(lam), x = sl * U;
  const double ec = erfc(fabs(x)), eb = 2.0 - ec;
  const double e = 0.5 * exp(-lam * U * U) * (0.56418958354775628695 * sl * il);  // / sqrt(pi lambda) = sqrt(lambda) / (sqrt(pi) lambda)
  MuL[0] = 0.5 * (x >= 0.0 ? eb : ec);  // 0.5 erfc(-sqrt(lambda) U)
  MuL[1] = U * MuL[0] + e;
  MuR[0] = 0.5 * (x >= 0.0 ? ec : eb);  // 0.5 erfc(+sqrt(lambda) U)
  MuR[1] = U * MuR[0] - e;
  const double h = 0.5 * il;
#pragma unroll
  for (int i = 2; i <= 6; ++i) {
    MuL[i] = U * MuL[i - 1] + (i - 1) * h * MuL[i - 2];
    MuR[i] = U * MuR[i - 1] + (i - 1) * h * MuR[i - 2];
  }
}
__device__ __forceinline__ void mxi(double K, double il, double *M) {
  M[0] = 1.0;
  M[1] = 0.5 * K * il;
  M[2] = 0.25 * (K * K + 2.0 * K) * (il * il);
}
// [KB] moments_conserve(Mu, Mv, Mw, a, b, d)
__device__ __forceinline__ void mom_cons(const double *Mu, const double *Mv, const double *Mw, int a,
                                         int b, int d, double *uv) {
  uv[0] = Mu[a] * Mv[b] * Mw[d / 2];
  uv[1] = Mu[a + 1] * Mv[b] * Mw[d / 2];
  uv[2] = Mu[a] * Mv[b + 1] * Mw[d / 2];
  uv[3] = 0.5 * (Mu[a + 2] * Mv[b] * Mw[d / 2] + Mu[a] * Mv[b + 2] * Mw[d / 2] +
                 Mu[a] * Mv[b] * Mw[(d + 2) / 2]);
}
// [KB] moments_conserve_slope(sl, Mu, Mv, Mw, a, 0), factored.  Every call of the path has b = 0, so the v and xi
// moments enter only through five constants of the state and the sum of six moments_conserve terms
//   au = sl0 <psi u^a> + sl1 <psi u^(a+1)> + sl2 <psi v u^a> + sl3 <psi T u^a>,  T = (u^2 + v^2 + xi^2)/2,
// psi = (1, u, v, T), collapses to (m_i = Mu[a + i], V_k = Mv[k], X_k = Mxi[k]):
//   E_i = <T u^(a+i)>  = (m_(i+2) + A0 m_i)/2                 A0 = V2 + X1
//   F   = <T v u^a>    = (V1 m_2 + B0 m_0)/2                  B0 = V3 + V1 X1
//   Q   = <T^2 u^a>    = (m_4 + 2 A0 m_2 + C0 m_0)/4          C0 = V4 + X2 + 2 V2 X1
// -- ~35 flops instead of ~100 per call (six calls per interface point), the same value to rounding
// (checked against the literal form in tests/test_oracle_extras.py).
struct SlopeConst {
  double V1, V2, A0, B0, C0;
};
__device__ __forceinline__ SlopeConst slope_const(const double *Mv, const double *Mw) {
  SlopeConst c;
  c.V1 = Mv[1];
  c.V2 = Mv[2];
  c.A0 = Mv[2] + Mw[1];
  c.B0 = Mv[3] + Mv[1] * Mw[1];
  c.C0 = Mv[4] + Mw[2] + 2.0 * Mv[2] * Mw[1];
  return c;
}
__device__ __forceinline__ void mom_slope(const double *sl, const double *m, const SlopeConst &c, double *au) {
  const double E0 = 0.5 * (m[2] + c.A0 * m[0]), E1 = 0.5 * (m[3] + c.A0 * m[1]);
  const double F0 = 0.5 * (c.V1 * m[2] + c.B0 * m[0]);
  const double Q0 = 0.25 * (m[4] + 2.0 * c.A0 * m[2] + c.C0 * m[0]);
  const double k = sl[0] + sl[2] * c.V1;
  au[0] = k * m[0] + sl[1] * m[1] + sl[3] * E0;
  au[1] = k * m[1] + sl[1] * m[2] + sl[3] * E1;
  au[2] = (sl[0] * c.V1 + sl[2] * c.V2) * m[0] + sl[1] * c.V1 * m[1] + sl[3] * F0;
  au[3] = sl[0] * E0 + sl[1] * E1 + sl[2] * F0 + sl[3] * Q0;
}
// [KB] moments_conserve(Mu, Mv, Mw, a, 0, 0) in the same terms
__device__ __forceinline__ void mom_cons0(const double *m, const SlopeConst &c, double *uv) {
  uv[0] = m[0];
  uv[1] = m[1];
  uv[2] = m[0] * c.V1;
  uv[3] = 0.5 * (m[2] + c.A0 * m[0]);
}
// [KB] pdf_slope(prim, sw, K)
__device__ __forceinline__ void pdf_slope(const Prim4 &p, const double *sw, double K, double iK2, double *sl) {
  const double q2 = p.U * p.U + p.V * p.V, hk = 0.5 * (K + 2.0) * p.il, l2r = 2.0 * p.lam * p.ir;
  sl[3] = 4.0 * p.lam * p.lam * iK2 * p.ir *
          (2.0 * sw[3] - 2.0 * p.U * sw[1] - 2.0 * p.V * sw[2] + sw[0] * (q2 - hk));
  sl[2] = l2r * (sw[2] - p.V * sw[0]) - p.V * sl[3];
  sl[1] = l2r * (sw[1] - p.U * sw[0]) - p.U * sl[3];
  sl[0] = sw[0] * p.ir - p.U * sl[1] - p.V * sl[2] - 0.5 * (q2 + hk) * sl[3];
}

struct GasPar {
  double K, gamma, mu, omega, dt;
  double gm1, iK2, iJx, iJy;  // gamma - 1, 1 / (K + 2), 1 / Jx, 1 / Jy: set by the launcher
};

// flux_gks!(fw, w, K, gamma, mu, omega, zeros(4))  (ns_cavity.jl:49-73): with zero slopes
// a = A = 0, so fw = rho * <u psi>  (the Maxwellian's own moments: Mu[0] = 1, Mu[1] = U)
__device__ __forceinline__ void gks_point_fluxes(const double *w, const GasPar &g, double *F, double *G) {
  const Prim4 p = conserve_prim(w, g.gm1);
  const double h = 0.5 * p.il;
  const double U2 = p.U * p.U + h, V2 = p.V * p.V + h;      // <u^2>, <v^2>
  const double U3 = p.U * U2 + 2.0 * h * p.U;               // <u^3>
  const double V3 = p.V * V2 + 2.0 * h * p.V;
  const double X1 = g.K * h;                                 // <xi^2>
  F[0] = p.rho * p.U;
  F[1] = p.rho * U2;
  F[2] = p.rho * p.U * p.V;
  F[3] = p.rho * 0.5 * (U3 + p.U * V2 + p.U * X1);
  G[0] = p.rho * p.V;
  G[1] = p.rho * p.U * p.V;
  G[2] = p.rho * V2;
  G[3] = p.rho * 0.5 * (V3 + p.V * U2 + p.V * X1);
}

// flux_gks!(fw, wL, wR, K, gamma, mu, omega, dt, swL, swR)  (ns_cavity.jl:75-145), states in the
// face-normal frame
__device__ __forceinline__ void gks_face_flux(double *fw, const double *wL, const double *wR,
                                           const double *swL, const double *swR, GasPar g) {
  const Prim4 pL = conserve_prim(wL, g.gm1), pR = conserve_prim(wR, g.gm1);
  double MuL1[7], MuR1[7], Mu1[7], Mv1[5], Mxi1[3];
  double MuL2[7], MuR2[7], Mu2[7], Mv2[5], Mxi2[3];
  moments_half(pL.U, pL.lam, pL.il, MuL1, MuR1);
  moments_half(pR.U, pR.lam, pR.il, MuL2, MuR2);
#pragma unroll
  for (int i = 0; i <= 6; ++i) { Mu1[i] = MuL1[i] + MuR1[i]; Mu2[i] = MuL2[i] + MuR2[i]; }
  moments_v(pL.V, pL.il, Mv1);
  moments_v(pR.V, pR.il, Mv2);
  mxi(g.K, pL.il, Mxi1);
  mxi(g.K, pR.il, Mxi2);
  const SlopeConst cL = slope_const(Mv1, Mxi1), cR = slope_const(Mv2, Mxi2);
  const SlopeConst cLR = slope_const(Mv1, Mxi2);  // Mv1 with the right state's xi moments: as written in ns_cavity.jl:102
  double a0[4], b0[4], w[4];
  mom_cons0(MuL1, cL, a0);
  mom_cons0(MuR2, cR, b0);
#pragma unroll
  for (int m = 0; m < 4; ++m) w[m] = pL.rho * a0[m] + pR.rho * b0[m];
  const Prim4 pc = conserve_prim(w, g.gm1);
  // vhs_collision_time: mu 2 lambda^(1 - omega) / rho, the power as exp((1 - omega) log lambda)
  const double eL = pL.rho * pL.il, eR = pR.rho * pR.il;
  const double tau = g.mu * 2.0 * exp((1.0 - g.omega) * log(pc.lam)) * pc.ir + 2.0 * g.dt * fabs(eL - eR) * rcp_fast(eL + eR);
  double faL[4], faTL[4], faR[4], faTR[4], sw[4];
  pdf_slope(pL, swL, g.K, g.iK2, faL);
  mom_slope(faL, Mu1 + 1, cL, sw);
#pragma unroll
  for (int m = 0; m < 4; ++m) sw[m] = -pL.rho * sw[m];
  pdf_slope(pL, sw, g.K, g.iK2, faTL);
  pdf_slope(pR, swR, g.K, g.iK2, faR);
  mom_slope(faR, Mu2 + 1, cLR, sw);
#pragma unroll
  for (int m = 0; m < 4; ++m) sw[m] = -pR.rho * sw[m];
  pdf_slope(pR, sw, g.K, g.iK2, faTR);
  // Mt[1] = dt - Mt[4] = 0: the central-state term drops out; Mt[4] = dt cancels against the final / dt
  double MuvL[4], MauL[4], MauLT[4], MuvR[4], MauR[4], MauRT[4];
  mom_cons0(MuL1 + 1, cL, MuvL);
  mom_slope(faL, MuL1 + 2, cL, MauL);
  mom_slope(faTL, MuL1 + 1, cL, MauLT);
  mom_cons0(MuR2 + 1, cR, MuvR);
  mom_slope(faR, MuR2 + 2, cR, MauR);
  mom_slope(faTR, MuR2 + 1, cR, MauRT);
#pragma unroll
  for (int m = 0; m < 4; ++m)
    fw[m] = pL.rho * (MuvL[m] - tau * (MauL[m] + MauLT[m])) + pR.rho * (MuvR[m] - tau * (MauR[m] + MauRT[m]));
}

template <int NSP>
__device__ __forceinline__ size_t eoff(int j, int i, int nyg) {
  return (size_t)4 * NSP * NSP * ((size_t)j + (size_t)nyg * i);
}

// boundary!(u, p, lambda0): isothermal walls by mirrored ghost states; lid on the top wall
template <int NSP>
__global__ void ns_boundary_kernel(double *__restrict__ u, int nx, int ny, double gamma, double lam0,
                                   double lid, int halo_lo, int halo_hi) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int side = blockIdx.y;
  const int n1 = side < 2 ? ny : nx;
  if (a > n1) return;
  // slab-parallel runs: column 0 / nx+1 of an interior slab boundary is the neighbour's data, not a wall
  if ((side == 0 && halo_lo) || (side == 1 && halo_hi)) return;
  const int nyg = ny + 2;
  const double gm1 = gamma - 1.0;
  int is, js, id, jd;
  if (side == 0) { is = 1; js = a; id = 0; jd = a; }
  else if (side == 1) { is = nx; js = a; id = nx + 1; jd = a; }
  else if (side == 2) { is = a; js = 1; id = a; jd = 0; }
  else { is = a; js = ny; id = a; jd = ny + 1; }
  const double *src = u + eoff<NSP>(js, is, nyg);
  double *dst = u + eoff<NSP>(jd, id, nyg);
  for (int k = 0; k < NSP; ++k)
    for (int l = 0; l < NSP; ++l) {
      const double *w = src + 4 * (l + NSP * k);
      const Prim4 p = conserve_prim(w, gm1);
      const double lamb = 2.0 * lam0 - p.lam;
      const double tmp = (p.lam - lam0) / lam0;
      const double rb = (1.0 - tmp) / (1.0 + tmp) * p.rho;
      const double ub = side == 3 ? lid : -p.U, vb = -p.V;
      const int kd = side < 2 ? NSP - 1 - k : k, ld = side < 2 ? l : NSP - 1 - l;
      double *o = dst + 4 * (ld + NSP * kd);
      o[0] = rb;
      o[1] = rb * ub;
      o[2] = rb * vb;
      o[3] = 0.5 * rb / lamb / gm1 + 0.5 * rb * (ub * ub + vb * vb);
    }
}

// ---- common fluxes: one thread per (face, flux point) ------------------------------------
// x faces: i = 1..nx+1 (between cells i-1 | i), j = 1..ny, row l     -> fhx[m + 4*(l + NSP*(j-1 + ny*(i-1)))]
// y faces: i = 1..nx, j = 1..ny+1 (between cells j-1 | j), column k  -> fhy[m + 4*(k + NSP*(j-1 + (ny+1)*(i-1)))]
// ns_cavity.jl:208-237 and :239-262 (states rotated by local_frame(., 0, 1) on y faces; slopes are not)
#ifndef FRB_NS_FACE_MINB
#define FRB_NS_FACE_MINB 4  // kernel experiment switch (scripts/build_variants.py)
#endif
template <int NSP>
__global__ void __launch_bounds__(128, FRB_NS_FACE_MINB)
ns_face_kernel(const double *__restrict__ u, double *__restrict__ fhx, double *__restrict__ fhy, int nx, int ny,
               double Jx, double Jy, GasPar gas, FrbOps ops) {
  const long long nfx = (long long)NSP * ny * (nx + 1), nfy = (long long)NSP * (ny + 1) * nx;
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nfx + nfy) return;
  const int nyg = ny + 2;
  const bool yface = t >= nfx;
  if (yface) t -= nfx;
  const int r = (int)(t % NSP);  // flux point along the face
  t /= NSP;
  const int rows = yface ? ny + 1 : ny;
  const int j = (int)(t % rows) + 1, i = (int)(t / rows) + 1;
  // "left" = lower cell in the face-normal direction
  const double *eA = u + eoff<NSP>(yface ? j - 1 : j, yface ? i : i - 1, nyg);
  const double *eB = u + eoff<NSP>(j, i, nyg);
  // stride of the contracted index and offset of the fixed one inside an element block
  const int sq = yface ? 4 : 4 * NSP, base = yface ? 4 * NSP * r : 4 * r;
  const double iJ = yface ? gas.iJy : gas.iJx;
  // the four variables of a point are 32 contiguous, 32-byte aligned bytes: two 16-byte loads per point and cell, all
  // 4 NSP of them issued before the first use
  double vA[NSP][4], vB[NSP][4];
#pragma unroll
  for (int q = 0; q < NSP; ++q) {
    const double2 *pa = reinterpret_cast<const double2 *>(eA + base + sq * q);
    const double2 *pb = reinterpret_cast<const double2 *>(eB + base + sq * q);
    const double2 a01 = pa[0], a23 = pa[1], b01 = pb[0], b23 = pb[1];
    vA[q][0] = a01.x; vA[q][1] = a01.y; vA[q][2] = a23.x; vA[q][3] = a23.y;
    vB[q][0] = b01.x; vB[q][1] = b01.y; vB[q][2] = b23.x; vB[q][3] = b23.y;
  }
  double wA[4], wB[4], sA[4], sB[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double a = 0, b = 0, c = 0, d = 0;
#pragma unroll
    for (int q = 0; q < NSP; ++q) {
      const double va = vA[q][m], vb = vB[q][m];
      a += va * ops.lr[q];   // trace of the lower cell on its upper face
      b += vb * ops.ll[q];   // trace of the upper cell on its lower face
      c += va * ops.dll[q];  // swL (the left cell's slope uses dll, :213)
      d += vb * ops.dlr[q];  // swR (:214)
    }
    wA[m] = a; wB[m] = b; sA[m] = c * iJ; sB[m] = d * iJ;
  }
  // local_frame(w, 0, 1) = (w0, w2, -w1, w3) on y faces, global_frame(f, 0, 1) = (f0, -f2, f1, f3) on the way back:
  // by selects, so that the flux routine is instantiated once (half the code: the kernel lives in the I-cache)
  const double lA[4] = {wA[0], yface ? wA[2] : wA[1], yface ? -wA[1] : wA[2], wA[3]};
  const double lB[4] = {wB[0], yface ? wB[2] : wB[1], yface ? -wB[1] : wB[2], wB[3]};
  double h[4];
  gks_face_flux(h, lA, lB, sA, sB, gas);
  const double fw[4] = {h[0], yface ? -h[2] : h[1], yface ? h[1] : h[2], h[3]};
  double *o = yface ? fhy + 4 * (r + NSP * ((long long)(j - 1) + (long long)(ny + 1) * (i - 1)))
                    : fhx + 4 * (r + NSP * ((long long)(j - 1) + (long long)ny * (i - 1)));
  reinterpret_cast<double2 *>(o)[0] = make_double2(fw[0], fw[1]);
  reinterpret_cast<double2 *>(o)[1] = make_double2(fw[2], fw[3]);
}

// (Issuing every global load before the first use and / or 16-byte loads -- 80-116 registers -- were measured and
// dropped: every variant is 5-10 % slower than this form at 64 registers and 8 CTAs / SM, although the same change
// gained 16 % in the face kernel; profiles/r02_summary.md section D.)
// ---- element kernel: thread = (solution point, element); a block holds EPB consecutive cells of a
// column of the mesh (j fastest in memory)
template <int NSP, int EPB>
__global__ void __launch_bounds__(NSP * NSP * EPB)
ns_elem_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
               const double *__restrict__ fhx, const double *__restrict__ fhy, int nx, int ny, double Jx,
               double Jy, GasPar gas, FrbOps ops, FrbStage st) {
  constexpr int NP = NSP * NSP;
  __shared__ double sF[EPB][NP][4], sG[EPB][NP][4];  // point fluxes / J of the block's elements
  const int pt = threadIdx.x % NP, el = threadIdx.x / NP;
  const int l = pt % NSP, k = pt / NSP;  // state index m + 4*(l + NSP*k)
  const int j = blockIdx.x * EPB + el + 1, i = blockIdx.y + 1;
  const int nyg = ny + 2;
  const bool live = j <= ny;
  double w[4] = {1.0, 0.0, 0.0, 1.0};
  const size_t eo = eoff<NSP>(live ? j : ny, i, nyg) + 4 * pt;
  if (live) {
#pragma unroll
    for (int m = 0; m < 4; ++m) w[m] = u[eo + m];
  }
  {
    double F[4], G[4];  // ns_cavity.jl:169-187
    gks_point_fluxes(w, gas, F, G);
#pragma unroll
    for (int m = 0; m < 4; ++m) { sF[el][pt][m] = F[m] * gas.iJx; sG[el][pt][m] = G[m] * gas.iJy; }
  }
  __syncthreads();
  if (!live) return;
  const double *hl = fhx + 4 * (l + NSP * ((long long)(j - 1) + (long long)ny * (i - 1)));
  const double *hr = fhx + 4 * (l + NSP * ((long long)(j - 1) + (long long)ny * i));
  const double *hb = fhy + 4 * (k + NSP * ((long long)(j - 1) + (long long)(ny + 1) * (i - 1)));
  const double *ht = fhy + 4 * (k + NSP * ((long long)j + (long long)(ny + 1) * (i - 1)));
  const double *a0 = ua ? ua + eo : nullptr;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double r1 = 0.0, r2 = 0.0, fL = 0.0, fR = 0.0, gB = 0.0, gT = 0.0;
#pragma unroll
    for (int q = 0; q < NSP; ++q) {
      const double fx = sF[el][l + NSP * q][m], fy = sG[el][q + NSP * k][m];
      r1 += fx * ops.lpdm[k * FRB_NSPMAX + q];  // :264-267
      r2 += fy * ops.lpdm[l * FRB_NSPMAX + q];
      fL += fx * ops.ll[q]; fR += fx * ops.lr[q];
      gB += fy * ops.ll[q]; gT += fy * ops.lr[q];
    }
    double du = r1 + r2;
    du += (hl[m] * gas.iJx - fL) * ops.dgl[k] + (hr[m] * gas.iJx - fR) * ops.dgr[k];  // :271-274
    du += (hb[m] * gas.iJy - gB) * ops.dgl[l] + (ht[m] * gas.iJy - gT) * ops.dgr[l];  // :275-278
    const double d = -du;
    double r;
    if (st.rhs_only) r = d;
    else {
      r = st.nested ? st.cb * (w[m] + st.cdt * d) : st.cb * w[m] + st.cdt * d;
      if (st.use_a) r = st.ca * a0[m] + r;
    }
    out[eo + m] = r;
  }
}

// the stage combination on the ghost ring (du = 0 there): dst = ca*ua + cb*src on whole ghost
// elements, so that the new state carries the ghosts OrdinaryDiffEq's axpys would give it
template <int NSP>
__global__ void ns_ring_stage_kernel(const double *src, const double *ua, double *dst, int nx, int ny,
                                     double ca, double cb, int use_a, int halo_lo, int halo_hi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nxg = nx + 2, nyg = ny + 2, nring = 2 * nxg + 2 * ny;
  if (t >= nring) return;
  int i, j;
  if (t < nxg) { i = t; j = 0; }
  else if (t < 2 * nxg) { i = t - nxg; j = nyg - 1; }
  else if (t < 2 * nxg + ny) { i = 0; j = t - 2 * nxg + 1; }
  else { i = nxg - 1; j = t - 2 * nxg - ny + 1; }
  if ((i == 0 && halo_lo) || (i == nxg - 1 && halo_hi)) return;  // the neighbouring rank writes that column
  const size_t o = eoff<NSP>(j, i, nyg);
  for (int q = 0; q < 4 * NSP * NSP; ++q) {
    double v = cb * src[o + q];
    if (use_a) v = ca * ua[o + q] + v;
    dst[o + q] = v;
  }
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, what, __FILE__, __LINE__);
  return 0;
}

}  // namespace

#define FRB_NS_SWITCH(nsp, CALL)                        \
  switch (nsp) {                                        \
    case 2: { constexpr int N = 2; CALL; } break;       \
    case 3: { constexpr int N = 3; CALL; } break;       \
    case 4: { constexpr int N = 4; CALL; } break;       \
    default: frb_set_error("ns2d kernels support deg 1..3"); return FRB_ERR_ARG; \
  }

// boundary!(u) on the stage input, then the fused RHS + stage
int frb_launch_ns2d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  cudaStream_t s = p->ctx->stream;
  const int nmax = p->nx > p->ny ? p->nx : p->ny;
  dim3 bb(64), bg((nmax + 63) / 64, 4);
  double *uw = const_cast<double *>(u);  // boundary! rewrites the ghosts of the state it is given
  int nranks = 1;
  const int rank = frb_halo_rank(p, &nranks);
  const int halo_lo = frb_halo_active(p) && rank > 0, halo_hi = frb_halo_active(p) && rank < nranks - 1;
  FRB_NS_SWITCH(p->nsp, (ns_boundary_kernel<N><<<bg, bb, 0, s>>>(uw, p->nx, p->ny, p->gamma, p->lambda_wall, p->lid_u,
                                                                halo_lo, halo_hi)));
  if (int rc = check_launch("ns_boundary_kernel")) return rc;
  int n = 1;
  if (!st.rhs_only) {  // ghosts of the new state: the stage combination with du = 0
    dim3 rg((2 * (p->nx + 2) + 2 * p->ny + 63) / 64);
    FRB_NS_SWITCH(p->nsp, (ns_ring_stage_kernel<N><<<rg, bb, 0, s>>>(u, ua, out, p->nx, p->ny, st.ca, st.cb, st.use_a,
                                                                  halo_lo, halo_hi)));
    if (int rc = check_launch("ns_ring_stage_kernel")) return rc;
    n += 1;
  }
  GasPar gas = {p->gks_K, p->gamma, p->gks_mu, p->gks_omega, p->gks_dt,
                p->gamma - 1.0, 1.0 / (p->gks_K + 2.0), 1.0 / p->Jx, 1.0 / p->Jy};
  const long long nfx = (long long)p->nsp * p->ny * (p->nx + 1), nfy = (long long)p->nsp * (p->ny + 1) * p->nx;
  if (!p->ns_flux) FRB_CUDA(cudaMalloc(&p->ns_flux, sizeof(double) * 4 * (nfx + nfy)));
  double *fhx = p->ns_flux, *fhy = p->ns_flux + 4 * nfx;
  dim3 fb(128), fg((unsigned)((nfx + nfy + 127) / 128));
  FRB_NS_SWITCH(p->nsp, (ns_face_kernel<N><<<fg, fb, 0, s>>>(u, fhx, fhy, p->nx, p->ny, p->Jx, p->Jy, gas, p->ops)));
  if (int rc = check_launch("ns_face_kernel")) return rc;
  constexpr int EPB = 8;
  dim3 grd((p->ny + EPB - 1) / EPB, p->nx);
  FRB_NS_SWITCH(p->nsp, (ns_elem_kernel<N, EPB><<<grd, N * N * EPB, 0, s>>>(u, ua, out, fhx, fhy, p->nx, p->ny, p->Jx,
                                                                          p->Jy, gas, p->ops, st)));
  if (int rc = check_launch("ns_elem_kernel")) return rc;
  return n + 2;
}
