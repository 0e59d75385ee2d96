// Pointwise physics closures (device).  These restate what the reference calls in
// KitBase.jl on the hot path: euler_flux (eq_euler.jl:37,
// euler2d_wave.jl:46), flux_hll! (eq_euler.jl:53; euler2d_wave.jl:73,80),
// local_frame/global_frame (euler2d_wave.jl:78-81), conserve_prim, maxwellian.
// Algebra is arranged for the FP64 pipe: one reciprocal and one square root per state.
#pragma once
#include <cuda_runtime.h>

// the portable closures are also callable from host code: tests/harness runs the curvilinear element
// routines (frb_euler2d_curv_elem.cuh) on the CPU; device code generation is unaffected
#define FRB_PHYS_HD __host__ __device__ __forceinline__

namespace frb {

// 1-D Euler: w = (rho, rho*u, E).  Returns F(w); also u and a (sound speed) if wanted.
struct Flux3 {
  double f0, f1, f2;
};
__device__ __forceinline__ Flux3 euler_flux3(double w0, double w1, double w2, double gm1) {
  double r = 1.0 / w0;
  double v = w1 * r;
  double p = gm1 * (w2 - 0.5 * w1 * v);
  return {w1, fma(w1, v, p), (w2 + p) * v};
}

// flux_hll!(fw, wL, wR, gamma, 1.0), 3 components
__device__ __forceinline__ Flux3 hll3(double l0, double l1, double l2, double r0, double r1,
                                      double r2, double gamma) {
  const double gm1 = gamma - 1.0;
  double il = 1.0 / l0, ir = 1.0 / r0;
  double ul = l1 * il, ur = r1 * ir;
  double pl = gm1 * (l2 - 0.5 * l1 * ul), pr = gm1 * (r2 - 0.5 * r1 * ur);
  double al = sqrt(gamma * pl * il), ar = sqrt(gamma * pr * ir);
  double lmin = ul - al, lmax = ur + ar;
  double fl0 = l1, fl1 = fma(l1, ul, pl), fl2 = (l2 + pl) * ul;
  double fr0 = r1, fr1 = fma(r1, ur, pr), fr2 = (r2 + pr) * ur;
  if (lmin >= 0.0) return {fl0, fl1, fl2};
  if (lmax <= 0.0) return {fr0, fr1, fr2};
  double fac = 1.0 / (lmax - lmin), mm = lmax * lmin;
  return {fac * (lmax * fl0 - lmin * fr0 + mm * (r0 - l0)),
          fac * (lmax * fl1 - lmin * fr1 + mm * (r1 - l1)),
          fac * (lmax * fl2 - lmin * fr2 + mm * (r2 - l2))};
}

// 2-D Euler: w = (rho, rho*u, rho*v, E)
struct Flux4 {
  double f0, f1, f2, f3;
};

// F and G at a point: euler_flux(w, gamma) -> (F, G)
FRB_PHYS_HD void euler_flux4(double w0, double w1, double w2, double w3, double gm1,
                                            Flux4 &F, Flux4 &G) {
  double r = 1.0 / w0;
  double u = w1 * r, v = w2 * r;
  double p = gm1 * (w3 - 0.5 * fma(w1, u, w2 * v));
  double h = w3 + p;
  F = {w1, fma(w1, u, p), w1 * v, h * u};
  G = {w2, w2 * u, fma(w2, v, p), h * v};
}

// HLL flux normal to an x-face: states in the global frame, normal velocity = component 1.
FRB_PHYS_HD Flux4 hll4(double l0, double l1, double l2, double l3, double r0,
                                      double r1, double r2, double r3, double gamma) {
  const double gm1 = gamma - 1.0;
  double il = 1.0 / l0, ir = 1.0 / r0;
  double ul = l1 * il, vl = l2 * il, ur = r1 * ir, vr = r2 * ir;
  double pl = gm1 * (l3 - 0.5 * fma(l1, ul, l2 * vl));
  double pr = gm1 * (r3 - 0.5 * fma(r1, ur, r2 * vr));
  double al = sqrt(gamma * pl * il), ar = sqrt(gamma * pr * ir);
  double lmin = ul - al, lmax = ur + ar;
  double fl0 = l1, fl1 = fma(l1, ul, pl), fl2 = l1 * vl, fl3 = (l3 + pl) * ul;
  double fr0 = r1, fr1 = fma(r1, ur, pr), fr2 = r1 * vr, fr3 = (r3 + pr) * ur;
  if (lmin >= 0.0) return {fl0, fl1, fl2, fl3};
  if (lmax <= 0.0) return {fr0, fr1, fr2, fr3};
  double fac = 1.0 / (lmax - lmin), mm = lmax * lmin;
  return {fac * (lmax * fl0 - lmin * fr0 + mm * (r0 - l0)),
          fac * (lmax * fl1 - lmin * fr1 + mm * (r1 - l1)),
          fac * (lmax * fl2 - lmin * fr2 + mm * (r2 - l2)),
          fac * (lmax * fl3 - lmin * fr3 + mm * (r3 - l3))};
}

// HLL flux normal to a y-face, i.e. global_frame(flux_hll!(local_frame(wL,0,1),
// local_frame(wR,0,1)), 0, 1) of euler2d_wave.jl:76-82: local = (w0, w2, -w1, w3),
// global(f) = (f0, -f2, f1, f3).
__device__ __forceinline__ Flux4 hll4_y(double l0, double l1, double l2, double l3, double r0,
                                        double r1, double r2, double r3, double gamma) {
  Flux4 f = hll4(l0, l2, -l1, l3, r0, r2, -r1, r3, gamma);
  return {f.f0, -f.f2, f.f1, f.f3};
}

// ---- extra common fluxes (the reference's live flux is HLL; LF and Roe are named by
// the north star, have no reference implementation, and are specified in DESIGN.md section 2) ----
// Local Lax-Friedrichs (Rusanov): 0.5 (F_L + F_R) - 0.5 alpha (w_R - w_L), alpha = max(|u|+a).
FRB_PHYS_HD Flux4 lf4(double l0, double l1, double l2, double l3, double r0, double r1,
                                     double r2, double r3, double gamma) {
  const double gm1 = gamma - 1.0;
  double il = 1.0 / l0, ir = 1.0 / r0;
  double ul = l1 * il, vl = l2 * il, ur = r1 * ir, vr = r2 * ir;
  double pl = gm1 * (l3 - 0.5 * fma(l1, ul, l2 * vl));
  double pr = gm1 * (r3 - 0.5 * fma(r1, ur, r2 * vr));
  double al = sqrt(gamma * pl * il), ar = sqrt(gamma * pr * ir);
  double alpha = fmax(fabs(ul) + al, fabs(ur) + ar);
  double fl0 = l1, fl1 = fma(l1, ul, pl), fl2 = l1 * vl, fl3 = (l3 + pl) * ul;
  double fr0 = r1, fr1 = fma(r1, ur, pr), fr2 = r1 * vr, fr3 = (r3 + pr) * ur;
  return {0.5 * (fl0 + fr0) - 0.5 * alpha * (r0 - l0), 0.5 * (fl1 + fr1) - 0.5 * alpha * (r1 - l1),
          0.5 * (fl2 + fr2) - 0.5 * alpha * (r2 - l2), 0.5 * (fl3 + fr3) - 0.5 * alpha * (r3 - l3)};
}

// Roe flux with Harten's entropy fix on the acoustic waves (below 0.1 a~).
FRB_PHYS_HD double roe_fix(double lam, double d) {
  double a = fabs(lam);
  return a < d ? (lam * lam + d * d) / (2.0 * d) : a;
}
FRB_PHYS_HD Flux4 roe4(double l0, double l1, double l2, double l3, double r0, double r1,
                                      double r2, double r3, double gamma) {
  const double gm1 = gamma - 1.0;
  double ul = l1 / l0, vl = l2 / l0, ur = r1 / r0, vr = r2 / r0;
  double pl = gm1 * (l3 - 0.5 * l0 * (ul * ul + vl * vl));
  double pr = gm1 * (r3 - 0.5 * r0 * (ur * ur + vr * vr));
  double Hl = (l3 + pl) / l0, Hr = (r3 + pr) / r0;
  double R = sqrt(r0 / l0), iR = 1.0 / (1.0 + R);
  double ut = (ul + R * ur) * iR, vt = (vl + R * vr) * iR, Ht = (Hl + R * Hr) * iR;
  double q2 = ut * ut + vt * vt;
  double a2 = gm1 * (Ht - 0.5 * q2), at = sqrt(a2), rt = R * l0;
  double dr = r0 - l0, du = ur - ul, dv = vr - vl, dp = pr - pl;
  double c1 = (dp - rt * at * du) / (2.0 * a2), c2 = dr - dp / a2, c3 = rt * dv;
  double c4 = (dp + rt * at * du) / (2.0 * a2);
  double d = 0.1 * at;
  double e1 = roe_fix(ut - at, d) * c1, e2 = fabs(ut) * c2, e3 = fabs(ut) * c3, e4 = roe_fix(ut + at, d) * c4;
  double fl0 = l1, fl1 = l1 * ul + pl, fl2 = l1 * vl, fl3 = (l3 + pl) * ul;
  double fr0 = r1, fr1 = r1 * ur + pr, fr2 = r1 * vr, fr3 = (r3 + pr) * ur;
  return {0.5 * (fl0 + fr0) - 0.5 * (e1 + e2 + e4),
          0.5 * (fl1 + fr1) - 0.5 * (e1 * (ut - at) + e2 * ut + e4 * (ut + at)),
          0.5 * (fl2 + fr2) - 0.5 * (e1 * vt + e2 * vt + e3 + e4 * vt),
          0.5 * (fl3 + fr3) - 0.5 * (e1 * (Ht - ut * at) + e2 * (0.5 * q2) + e3 * vt + e4 * (Ht + ut * at))};
}

// common flux selector of the generic kernels: kind = FRB_FLUX_HLL / LF / ROE
FRB_PHYS_HD Flux4 riemann4(int kind, double l0, double l1, double l2, double l3, double r0,
                                          double r1, double r2, double r3, double gamma) {
  if (kind == 1) return lf4(l0, l1, l2, l3, r0, r1, r2, r3, gamma);
  if (kind == 2) return roe4(l0, l1, l2, l3, r0, r1, r2, r3, gamma);
  return hll4(l0, l1, l2, l3, r0, r1, r2, r3, gamma);
}
// normal to a y face: local_frame(., 0, 1) -> flux -> global_frame (euler2d_wave.jl:76-82)
__device__ __forceinline__ Flux4 riemann4_y(int kind, double l0, double l1, double l2, double l3, double r0,
                                            double r1, double r2, double r3, double gamma) {
  Flux4 f = riemann4(kind, l0, l2, -l1, l3, r0, r2, -r1, r3, gamma);
  return {f.f0, -f.f2, f.f1, f.f3};
}
// 1-D: the 4-component flux with zero tangential momentum
__device__ __forceinline__ Flux3 riemann3(int kind, double l0, double l1, double l2, double r0, double r1,
                                          double r2, double gamma) {
  Flux4 f = riemann4(kind, l0, l1, 0.0, l2, r0, r1, 0.0, r2, gamma);
  return {f.f0, f.f1, f.f3};
}

// ---- branch-free variants for the roofline kernel -----------------------------------------
// 1/x and sqrt(x) from the MUFU seeds (about 20 good bits) plus Newton steps on the FP64
// pipe: no IEEE slow-path subroutine, no divergent branch.  Results are within ~1 ulp for
// the normal, well-scaled arguments of this path (densities, pressures, wave speeds).
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double sqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx, y * y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx, y * y, 0.5);
  y = fma(y, e, y);
  double s = x * y;
  double r = fma(-s, s, x);
  return fma(0.5 * r, y, s);
}

// sqrt by two coupled Goldschmidt steps on (g ~ sqrt x, h ~ 1/(2 sqrt x)) from the MUFU seed:
// 7 FP64 instructions instead of 11, result within ~2 ulp (the seed's 2^-20 squares twice).
__device__ __forceinline__ double sqrt_gs(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  return fma(g, r, g);
}

__device__ __forceinline__ Flux4 hll4_fast(double l0, double l1, double l2, double l3, double r0,
                                           double r1, double r2, double r3, double gamma,
                                           double gm1) {
  double il = rcp_fast(l0), ir = rcp_fast(r0);
  double ul = l1 * il, vl = l2 * il, ur = r1 * ir, vr = r2 * ir;
  double pl = gm1 * fma(-0.5, fma(l1, ul, l2 * vl), l3);
  double pr = gm1 * fma(-0.5, fma(r1, ur, r2 * vr), r3);
  double al = sqrt_gs(gamma * pl * il), ar = sqrt_gs(gamma * pr * ir);
  double lmin = ul - al, lmax = ur + ar;
  double fl0 = l1, fl1 = fma(l1, ul, pl), fl2 = l1 * vl, fl3 = (l3 + pl) * ul;
  double fr0 = r1, fr1 = fma(r1, ur, pr), fr2 = r1 * vr, fr3 = (r3 + pr) * ur;
  double fac = rcp_fast(lmax - lmin), mm = lmax * lmin;
  double a = lmax * fac, b = lmin * fac, c = mm * fac;
  double m0 = fma(a, fl0, fma(-b, fr0, c * (r0 - l0)));
  double m1 = fma(a, fl1, fma(-b, fr1, c * (r1 - l1)));
  double m2 = fma(a, fl2, fma(-b, fr2, c * (r2 - l2)));
  double m3 = fma(a, fl3, fma(-b, fr3, c * (r3 - l3)));
  const bool sl = lmin >= 0.0, sr = lmax <= 0.0;
  return {sl ? fl0 : (sr ? fr0 : m0), sl ? fl1 : (sr ? fr1 : m1), sl ? fl2 : (sr ? fr2 : m2),
          sl ? fl3 : (sr ? fr3 : m3)};
}
__device__ __forceinline__ Flux4 hll4_y_fast(double l0, double l1, double l2, double l3, double r0,
                                             double r1, double r2, double r3, double gamma,
                                             double gm1) {
  Flux4 f = hll4_fast(l0, l2, -l1, l3, r0, r2, -r1, r3, gamma, gm1);
  return {f.f0, -f.f2, f.f1, f.f3};
}

// ---- the other common fluxes in the same branch-free form (marching kernels) ------------------------
// Same definitions as lf4 / roe4 above, reciprocals and roots from the MUFU seeds.
__device__ __forceinline__ Flux4 lf4_fast(double l0, double l1, double l2, double l3, double r0, double r1,
                                          double r2, double r3, double gamma, double gm1) {
  double il = rcp_fast(l0), ir = rcp_fast(r0);
  double ul = l1 * il, vl = l2 * il, ur = r1 * ir, vr = r2 * ir;
  double pl = gm1 * fma(-0.5, fma(l1, ul, l2 * vl), l3);
  double pr = gm1 * fma(-0.5, fma(r1, ur, r2 * vr), r3);
  double al = sqrt_gs(gamma * pl * il), ar = sqrt_gs(gamma * pr * ir);
  double ha = 0.5 * fmax(fabs(ul) + al, fabs(ur) + ar);
  double fl0 = l1, fl1 = fma(l1, ul, pl), fl2 = l1 * vl, fl3 = (l3 + pl) * ul;
  double fr0 = r1, fr1 = fma(r1, ur, pr), fr2 = r1 * vr, fr3 = (r3 + pr) * ur;
  return {fma(-ha, r0 - l0, 0.5 * (fl0 + fr0)), fma(-ha, r1 - l1, 0.5 * (fl1 + fr1)),
          fma(-ha, r2 - l2, 0.5 * (fl2 + fr2)), fma(-ha, r3 - l3, 0.5 * (fl3 + fr3))};
}
__device__ __forceinline__ double roe_fix_fast(double lam, double d, double i2d) {
  double a = fabs(lam);
  return a < d ? fma(lam, lam, d * d) * i2d : a;
}
__device__ __forceinline__ Flux4 roe4_fast(double l0, double l1, double l2, double l3, double r0, double r1,
                                           double r2, double r3, double gamma, double gm1) {
  double il = rcp_fast(l0), ir = rcp_fast(r0);
  double ul = l1 * il, vl = l2 * il, ur = r1 * ir, vr = r2 * ir;
  double pl = gm1 * fma(-0.5, fma(l1, ul, l2 * vl), l3);
  double pr = gm1 * fma(-0.5, fma(r1, ur, r2 * vr), r3);
  double Hl = (l3 + pl) * il, Hr = (r3 + pr) * ir;
  double R = sqrt_gs(r0 * il), iR = rcp_fast(1.0 + R);
  double ut = fma(R, ur, ul) * iR, vt = fma(R, vr, vl) * iR, Ht = fma(R, Hr, Hl) * iR;
  double q2 = fma(ut, ut, vt * vt);
  double a2 = gm1 * fma(-0.5, q2, Ht), at = sqrt_gs(a2), rt = R * l0;
  double ia2 = rcp_fast(a2);
  double dr = r0 - l0, du = ur - ul, dv = vr - vl, dp = pr - pl;
  double ra = rt * at * du;
  double c1 = (dp - ra) * (0.5 * ia2), c2 = fma(-dp, ia2, dr), c3 = rt * dv, c4 = (dp + ra) * (0.5 * ia2);
  double d = 0.1 * at, i2d = rcp_fast(2.0 * d);
  double au = fabs(ut);
  double e1 = roe_fix_fast(ut - at, d, i2d) * c1, e2 = au * c2, e3 = au * c3, e4 = roe_fix_fast(ut + at, d, i2d) * c4;
  double fl0 = l1, fl1 = fma(l1, ul, pl), fl2 = l1 * vl, fl3 = (l3 + pl) * ul;
  double fr0 = r1, fr1 = fma(r1, ur, pr), fr2 = r1 * vr, fr3 = (r3 + pr) * ur;
  double ua = ut * at;
  return {0.5 * (fl0 + fr0) - 0.5 * (e1 + e2 + e4),
          0.5 * (fl1 + fr1) - 0.5 * fma(e1, ut - at, fma(e2, ut, e4 * (ut + at))),
          0.5 * (fl2 + fr2) - 0.5 * fma(e1 + e2 + e4, vt, e3),
          0.5 * (fl3 + fr3) - 0.5 * fma(e1, Ht - ua, fma(e2, 0.5 * q2, fma(e3, vt, e4 * (Ht + ua))))};
}
// common flux of the marching kernels, selected at compile time (FLUX = FRB_FLUX_HLL / LF / ROE)
template <int FLUX>
__device__ __forceinline__ Flux4 riemann4_fast(double l0, double l1, double l2, double l3, double r0, double r1,
                                               double r2, double r3, double gamma, double gm1) {
  if (FLUX == 1) return lf4_fast(l0, l1, l2, l3, r0, r1, r2, r3, gamma, gm1);
  if (FLUX == 2) return roe4_fast(l0, l1, l2, l3, r0, r1, r2, r3, gamma, gm1);
  return hll4_fast(l0, l1, l2, l3, r0, r1, r2, r3, gamma, gm1);
}
template <int FLUX>
__device__ __forceinline__ Flux4 riemann4_y_fast(double l0, double l1, double l2, double l3, double r0, double r1,
                                                 double r2, double r3, double gamma, double gm1) {
  Flux4 f = riemann4_fast<FLUX>(l0, l2, -l1, l3, r0, r2, -r1, r3, gamma, gm1);
  return {f.f0, -f.f2, f.f1, f.f3};
}

}  // namespace frb
