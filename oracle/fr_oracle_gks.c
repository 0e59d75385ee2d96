/* C restatement of the 2-D gas-kinetic Navier-Stokes RHS of example/ns_cavity.jl
 * (dudt! :147-287, boundary! :289-344, the in-script flux_gks! overloads :49-145)
 * -- TEST INFRASTRUCTURE ONLY (see fr_oracle.c header).  PARITY UNPINNED.
 *
 * [KB] marks KitBase.jl 0.9 closures restated from their published form:
 * gauss_moments, moments_conserve, moments_conserve_slope, pdf_slope,
 * vhs_collision_time, conserve_prim/prim_conserve, local_frame/global_frame.
 *
 * Layout: u[4, ns, nr, ny+2, nx+2]  (variable fastest; ns <-> s(y) index l, nr <-> r(x)
 * index k; ns_cavity.jl:33).  Stale indexing of the script (`ps.J[i,j][1]` on a 2x2,
 * :173,183,213,272) is read as Jx = dx/2, `[2]` as Jy = dy/2.
 * Quirks kept on purpose: the left cell's slope uses `dll`, the right cell's `dlr`
 * (:213-214, :244-245); slopes are NOT rotated into the local frame on y faces (:260);
 * `Mv1` instead of `Mv2` in the right state's time slope (:102).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NSPMAX 8

static inline void conserve_prim4(const double *W, double g, double *prim) {
  prim[0] = W[0];
  prim[1] = W[1] / W[0];
  prim[2] = W[2] / W[0];
  prim[3] = 0.5 * W[0] / (g - 1.0) / (W[3] - 0.5 * (W[1] * W[1] + W[2] * W[2]) / W[0]);
}
static inline void prim_conserve4(const double *prim, double g, double *W) {
  W[0] = prim[0];
  W[1] = prim[0] * prim[1];
  W[2] = prim[0] * prim[2];
  W[3] = 0.5 * prim[0] / prim[3] / (g - 1.0) + 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2]);
}

/* [KB] gauss_moments(prim, inK): Mu, Mv [0..6], Mxi [0..2], MuL, MuR [0..6] */
static void gauss_moments(const double *prim, double K, double *Mu, double *Mv, double *Mxi,
                          double *MuL, double *MuR) {
  const double U = prim[1], V = prim[2], lam = prim[3];
  MuL[0] = 0.5 * erfc(-sqrt(lam) * U);
  MuL[1] = U * MuL[0] + 0.5 * exp(-lam * U * U) / sqrt(M_PI * lam);
  MuR[0] = 0.5 * erfc(sqrt(lam) * U);
  MuR[1] = U * MuR[0] - 0.5 * exp(-lam * U * U) / sqrt(M_PI * lam);
  for (int i = 2; i <= 6; ++i) {
    MuL[i] = U * MuL[i - 1] + 0.5 * (i - 1) * MuL[i - 2] / lam;
    MuR[i] = U * MuR[i - 1] + 0.5 * (i - 1) * MuR[i - 2] / lam;
  }
  for (int i = 0; i <= 6; ++i) Mu[i] = MuL[i] + MuR[i];
  Mv[0] = 1.0;
  Mv[1] = V;
  for (int i = 2; i <= 6; ++i) Mv[i] = V * Mv[i - 1] + 0.5 * (i - 1) * Mv[i - 2] / lam;
  Mxi[0] = 1.0;
  Mxi[1] = 0.5 * K / lam;
  Mxi[2] = (K * K + 2.0 * K) / (4.0 * lam * lam);
}
/* [KB] moments_conserve(Mu, Mv, Mw, alpha, beta, delta) */
static void moments_conserve(const double *Mu, const double *Mv, const double *Mw, int a, int b, int d,
                             double *uv) {
  uv[0] = Mu[a] * Mv[b] * Mw[d / 2];
  uv[1] = Mu[a + 1] * Mv[b] * Mw[d / 2];
  uv[2] = Mu[a] * Mv[b + 1] * Mw[d / 2];
  uv[3] = 0.5 * (Mu[a + 2] * Mv[b] * Mw[d / 2] + Mu[a] * Mv[b + 2] * Mw[d / 2] +
                 Mu[a] * Mv[b] * Mw[(d + 2) / 2]);
}
/* [KB] moments_conserve_slope(a, Mu, Mv, Mw, alpha, beta) */
static void moments_conserve_slope(const double *sl, const double *Mu, const double *Mv, const double *Mw,
                                   int a, int b, double *au) {
  double t0[4], t1[4], t2[4], t3[4], t4[4], t5[4];
  moments_conserve(Mu, Mv, Mw, a + 0, b + 0, 0, t0);
  moments_conserve(Mu, Mv, Mw, a + 1, b + 0, 0, t1);
  moments_conserve(Mu, Mv, Mw, a + 0, b + 1, 0, t2);
  moments_conserve(Mu, Mv, Mw, a + 2, b + 0, 0, t3);
  moments_conserve(Mu, Mv, Mw, a + 0, b + 2, 0, t4);
  moments_conserve(Mu, Mv, Mw, a + 0, b + 0, 2, t5);
  for (int m = 0; m < 4; ++m)
    au[m] = sl[0] * t0[m] + sl[1] * t1[m] + sl[2] * t2[m] + 0.5 * sl[3] * t3[m] + 0.5 * sl[3] * t4[m] +
            0.5 * sl[3] * t5[m];
}
/* [KB] pdf_slope(prim, sw, inK) */
static void pdf_slope(const double *prim, const double *sw, double K, double *sl) {
  const double rho = prim[0], U = prim[1], V = prim[2], lam = prim[3];
  sl[3] = 4.0 * lam * lam / (K + 2.0) / rho *
          (2.0 * sw[3] - 2.0 * U * sw[1] - 2.0 * V * sw[2] + sw[0] * (U * U + V * V - 0.5 * (K + 2.0) / lam));
  sl[2] = 2.0 * lam / rho * (sw[2] - V * sw[0]) - V * sl[3];
  sl[1] = 2.0 * lam / rho * (sw[1] - U * sw[0]) - U * sl[3];
  sl[0] = sw[0] / rho - U * sl[1] - V * sl[2] - 0.5 * (U * U + V * V + 0.5 * (K + 2.0) / lam) * sl[3];
}
/* [KB] vhs_collision_time(prim, muRef, omega) */
static inline double vhs_collision_time(const double *prim, double mu, double omega) {
  return mu * 2.0 * pow(prim[3], 1.0 - omega) / prim[0];
}

/* ns_cavity.jl:49-73 with sw = zeros(4): fw = rho * (Muv - tau*Mau - tau*Mtu), Mau = Mtu = 0 */
static void flux_gks_point(double *fw, const double *w, double K, double g, double mu, double omega) {
  double prim[4], Mu[7], Mv[7], Mxi[3], MuL[7], MuR[7];
  conserve_prim4(w, g, prim);
  gauss_moments(prim, K, Mu, Mv, Mxi, MuL, MuR);
  double tau = vhs_collision_time(prim, mu, omega);
  double sw[4] = {0, 0, 0, 0}, a[4], dft[4], A[4], Muv[4], Mau[4], Mtu[4];
  pdf_slope(prim, sw, K, a);
  moments_conserve_slope(a, Mu, Mv, Mxi, 1, 0, dft);
  for (int m = 0; m < 4; ++m) dft[m] = -prim[0] * dft[m];
  pdf_slope(prim, dft, K, A);
  moments_conserve(Mu, Mv, Mxi, 1, 0, 0, Muv);
  moments_conserve_slope(a, Mu, Mv, Mxi, 2, 0, Mau);
  moments_conserve_slope(A, Mu, Mv, Mxi, 1, 0, Mtu);
  for (int m = 0; m < 4; ++m) fw[m] = prim[0] * (Muv[m] - tau * Mau[m] - tau * Mtu[m]);
}

/* ns_cavity.jl:75-145 */
static void flux_gks_face(double *fw, const double *wL, const double *wR, double K, double g, double mu,
                          double omega, double dt, const double *swL, const double *swR) {
  double pL[4], pR[4], prim[4], w[4];
  double Mu1[7], Mv1[7], Mxi1[3], MuL1[7], MuR1[7], Mu2[7], Mv2[7], Mxi2[3], MuL2[7], MuR2[7];
  conserve_prim4(wL, g, pL);
  conserve_prim4(wR, g, pR);
  gauss_moments(pL, K, Mu1, Mv1, Mxi1, MuL1, MuR1);
  gauss_moments(pR, K, Mu2, Mv2, Mxi2, MuL2, MuR2);
  double a0[4], b0[4];
  moments_conserve(MuL1, Mv1, Mxi1, 0, 0, 0, a0);
  moments_conserve(MuR2, Mv2, Mxi2, 0, 0, 0, b0);
  for (int m = 0; m < 4; ++m) w[m] = pL[0] * a0[m] + pR[0] * b0[m];
  conserve_prim4(w, g, prim);
  double tau = vhs_collision_time(prim, mu, omega) +
               2.0 * dt * fabs(pL[0] / pL[3] - pR[0] / pR[3]) / (pL[0] / pL[3] + pR[0] / pR[3]);
  double faL[4], faTL[4], faR[4], faTR[4], sw[4];
  pdf_slope(pL, swL, K, faL);
  moments_conserve_slope(faL, Mu1, Mv1, Mxi1, 1, 0, sw);
  for (int m = 0; m < 4; ++m) sw[m] = -pL[0] * sw[m];
  pdf_slope(pL, sw, K, faTL);
  pdf_slope(pR, swR, K, faR);
  moments_conserve_slope(faR, Mu2, Mv1, Mxi2, 1, 0, sw); /* Mv1: as written in :102 */
  for (int m = 0; m < 4; ++m) sw[m] = -pR[0] * sw[m];
  pdf_slope(pR, sw, K, faTR);
  double Mu[7], Mv[7], Mxi[3], MuL[7], MuR[7];
  gauss_moments(prim, K, Mu, Mv, Mxi, MuL, MuR);
  double Mt[5];
  Mt[3] = dt;
  Mt[4] = -tau * dt * exp(-dt / tau) + tau * Mt[3];
  Mt[0] = dt - Mt[3];
  Mt[1] = -tau * Mt[0] + Mt[4];
  Mt[2] = 0.5 * dt * dt - tau * Mt[0];
  (void)Mt[1]; (void)Mt[2];
  double Muv[4], MuvL[4], MauL[4], MauLT[4], MuvR[4], MauR[4], MauRT[4];
  moments_conserve(Mu, Mv, Mxi, 1, 0, 0, Muv);
  moments_conserve(MuL1, Mv1, Mxi1, 1, 0, 0, MuvL);
  moments_conserve_slope(faL, MuL1, Mv1, Mxi1, 2, 0, MauL);
  moments_conserve_slope(faTL, MuL1, Mv1, Mxi1, 1, 0, MauLT);
  moments_conserve(MuR2, Mv2, Mxi2, 1, 0, 0, MuvR);
  moments_conserve_slope(faR, MuR2, Mv2, Mxi2, 2, 0, MauR);
  moments_conserve_slope(faTR, MuR2, Mv2, Mxi2, 1, 0, MauRT);
  for (int m = 0; m < 4; ++m) {
    double f = Mt[0] * prim[0] * Muv[m];
    f += Mt[3] * pL[0] * MuvL[m] - tau * Mt[3] * pL[0] * MauL[m] - tau * Mt[3] * pL[0] * MauLT[m] +
         Mt[3] * pR[0] * MuvR[m] - tau * Mt[3] * pR[0] * MauR[m] - tau * Mt[3] * pR[0] * MauRT[m];
    fw[m] = f / dt;
  }
}

/* boundary!(u, p, lambda0): ns_cavity.jl:289-344; lid = pb[2] of :337 */
void fro_ns_boundary(double *u, int nx, int ny, int nsp, double g, double lam0, double lid) {
  const size_t e = (size_t)4 * nsp * nsp, NYG = ny + 2;
#define UN(m, l, k, j, i) u[(m) + 4 * ((l) + nsp * ((k) + nsp * ((size_t)(j) + NYG * (size_t)(i))))]
  (void)e;
  for (int side = 0; side < 4; ++side) {
    int n1 = side < 2 ? ny : nx;
    for (int a = 1; a <= n1; ++a)
      for (int k = 0; k < nsp; ++k)
        for (int l = 0; l < nsp; ++l) {
          int is, js, id, jd, kd, ld;
          if (side == 0) { is = 1; js = a; id = 0; jd = a; kd = nsp - 1 - k; ld = l; }
          else if (side == 1) { is = nx; js = a; id = nx + 1; jd = a; kd = nsp - 1 - k; ld = l; }
          else if (side == 2) { is = a; js = 1; id = a; jd = 0; kd = k; ld = nsp - 1 - l; }
          else { is = a; js = ny; id = a; jd = ny + 1; kd = k; ld = nsp - 1 - l; }
          double w[4] = {UN(0, l, k, js, is), UN(1, l, k, js, is), UN(2, l, k, js, is), UN(3, l, k, js, is)};
          double prim[4], pb[4], wb[4];
          conserve_prim4(w, g, prim);
          pb[3] = 2.0 * lam0 - prim[3];
          double tmp = (prim[3] - lam0) / lam0;
          pb[0] = (1.0 - tmp) / (1.0 + tmp) * prim[0];
          pb[1] = side == 3 ? lid : -prim[1];
          pb[2] = -prim[2];
          prim_conserve4(pb, g, wb);
          for (int m = 0; m < 4; ++m) UN(m, ld, kd, jd, id) = wb[m];
        }
  }
}

/* dudt!(du, u, p, t): ns_cavity.jl:147-287.  u is modified in place by boundary! like the
 * reference.  Work arrays are allocated per call (the script preallocates; timing is not the
 * point of this restatement). */
int fro_rhs_ns2d(double *u, double *du, int nx, int ny, int nsp, double Jx, double Jy, const double *ll,
                 const double *lr, const double *lpdm, const double *dhl, const double *dhr,
                 const double *dll, const double *dlr, double K, double g, double mu, double omega,
                 double dt, double lam0, double lid) {
  if (nsp > NSPMAX) return -1;
  const size_t NYG = ny + 2, NXG = nx + 2, n = (size_t)4 * nsp * nsp * NYG * NXG;
  fro_ns_boundary(u, nx, ny, nsp, g, lam0, lid);
  double *fx = (double *)calloc(n, sizeof(double)), *fy = (double *)calloc(n, sizeof(double));
  /* faces: [m, side, pt, j, i] */
  const size_t nf = (size_t)4 * 2 * nsp * NYG * NXG;
  double *uxf = (double *)calloc(nf, sizeof(double)), *uyf = (double *)calloc(nf, sizeof(double));
  double *fxf = (double *)calloc(nf, sizeof(double)), *fyf = (double *)calloc(nf, sizeof(double));
  double *fxi = (double *)calloc((size_t)4 * nsp * NYG * (NXG + 1), sizeof(double));
  double *fyi = (double *)calloc((size_t)4 * nsp * (NYG + 1) * NXG, sizeof(double));
  if (!fx || !fy || !uxf || !uyf || !fxf || !fyf || !fxi || !fyi) return -2;
#define A5(a, m, l, k, j, i) a[(m) + 4 * ((l) + nsp * ((k) + nsp * ((size_t)(j) + NYG * (size_t)(i))))]
#define FC(a, m, s, p, j, i) a[(m) + 4 * ((s) + 2 * ((p) + nsp * ((size_t)(j) + NYG * (size_t)(i))))]
#define FXI(m, l, j, i) fxi[(m) + 4 * ((l) + nsp * ((size_t)(j) + NYG * (size_t)(i)))]
#define FYI(m, k, j, i) fyi[(m) + 4 * ((k) + nsp * ((size_t)(j) + (NYG + 1) * (size_t)(i)))]
  /* point fluxes :169-187 */
#pragma omp parallel for schedule(static)
  for (int i = 1; i <= nx; ++i)
    for (int j = 1; j <= ny; ++j)
      for (int k = 0; k < nsp; ++k)
        for (int l = 0; l < nsp; ++l) {
          double w[4], fw[4], ul[4];
          for (int m = 0; m < 4; ++m) w[m] = A5(u, m, l, k, j, i);
          flux_gks_point(fw, w, K, g, mu, omega);
          for (int m = 0; m < 4; ++m) A5(fx, m, l, k, j, i) = fw[m] / Jx;
          ul[0] = w[0]; ul[1] = w[1] * 0.0 + w[2] * 1.0; ul[2] = w[2] * 0.0 - w[1] * 1.0; ul[3] = w[3];
          flux_gks_point(fw, ul, K, g, mu, omega);
          A5(fy, 0, l, k, j, i) = fw[0] / Jy;
          A5(fy, 1, l, k, j, i) = (fw[1] * 0.0 - fw[2] * 1.0) / Jy;
          A5(fy, 2, l, k, j, i) = (fw[1] * 1.0 + fw[2] * 0.0) / Jy;
          A5(fy, 3, l, k, j, i) = fw[3] / Jy;
        }
  /* traces :189-206 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i <= nx + 1; ++i)
    for (int j = 1; j <= ny; ++j)
      for (int l = 0; l < nsp; ++l)
        for (int m = 0; m < 4; ++m) {
          double a = 0, b = 0, c = 0, d = 0;
          for (int q = 0; q < nsp; ++q) {
            a += A5(u, m, l, q, j, i) * ll[q]; b += A5(u, m, l, q, j, i) * lr[q];
            c += A5(fx, m, l, q, j, i) * ll[q]; d += A5(fx, m, l, q, j, i) * lr[q];
          }
          FC(uxf, m, 0, l, j, i) = a; FC(uxf, m, 1, l, j, i) = b;
          FC(fxf, m, 0, l, j, i) = c; FC(fxf, m, 1, l, j, i) = d;
        }
#pragma omp parallel for schedule(static)
  for (int i = 1; i <= nx; ++i)
    for (int j = 0; j <= ny + 1; ++j)
      for (int k = 0; k < nsp; ++k)
        for (int m = 0; m < 4; ++m) {
          double a = 0, b = 0, c = 0, d = 0;
          for (int q = 0; q < nsp; ++q) {
            a += A5(u, m, q, k, j, i) * ll[q]; b += A5(u, m, q, k, j, i) * lr[q];
            c += A5(fy, m, q, k, j, i) * ll[q]; d += A5(fy, m, q, k, j, i) * lr[q];
          }
          FC(uyf, m, 0, k, j, i) = a; FC(uyf, m, 1, k, j, i) = b;
          FC(fyf, m, 0, k, j, i) = c; FC(fyf, m, 1, k, j, i) = d;
        }
  /* x interfaces :208-237 */
#pragma omp parallel for schedule(static)
  for (int i = 1; i <= nx + 1; ++i)
    for (int j = 1; j <= ny; ++j)
      for (int l = 0; l < nsp; ++l) {
        double swL[4], swR[4], wL[4], wR[4], fw[4];
        for (int m = 0; m < 4; ++m) {
          double a = 0, b = 0;
          for (int q = 0; q < nsp; ++q) { a += A5(u, m, l, q, j, i - 1) * dll[q]; b += A5(u, m, l, q, j, i) * dlr[q]; }
          swL[m] = a / Jx; swR[m] = b / Jx;
          wL[m] = FC(uxf, m, 1, l, j, i - 1); wR[m] = FC(uxf, m, 0, l, j, i);
        }
        flux_gks_face(fw, wL, wR, K, g, mu, omega, dt, swL, swR);
        for (int m = 0; m < 4; ++m) FXI(m, l, j, i) = fw[m];
      }
  /* y interfaces :239-262 */
#pragma omp parallel for schedule(static)
  for (int i = 1; i <= nx; ++i)
    for (int j = 1; j <= ny + 1; ++j)
      for (int k = 0; k < nsp; ++k) {
        double swL[4], swR[4], a4[4], b4[4], wL[4], wR[4], fw[4];
        for (int m = 0; m < 4; ++m) {
          double a = 0, b = 0;
          for (int q = 0; q < nsp; ++q) { a += A5(u, m, q, k, j - 1, i) * dll[q]; b += A5(u, m, q, k, j, i) * dlr[q]; }
          swL[m] = a / Jy; swR[m] = b / Jy;
          a4[m] = FC(uyf, m, 1, k, j - 1, i); b4[m] = FC(uyf, m, 0, k, j, i);
        }
        wL[0] = a4[0]; wL[1] = a4[1] * 0.0 + a4[2] * 1.0; wL[2] = a4[2] * 0.0 - a4[1] * 1.0; wL[3] = a4[3];
        wR[0] = b4[0]; wR[1] = b4[1] * 0.0 + b4[2] * 1.0; wR[2] = b4[2] * 0.0 - b4[1] * 1.0; wR[3] = b4[3];
        flux_gks_face(fw, wL, wR, K, g, mu, omega, dt, swL, swR);
        FYI(0, k, j, i) = fw[0];
        FYI(1, k, j, i) = fw[1] * 0.0 - fw[2] * 1.0;
        FYI(2, k, j, i) = fw[1] * 1.0 + fw[2] * 0.0;
        FYI(3, k, j, i) = fw[3];
      }
  /* derivative + correction :264-284 */
  memset(du, 0, sizeof(double) * n);
#pragma omp parallel for schedule(static)
  for (int i = 1; i <= nx; ++i)
    for (int j = 1; j <= ny; ++j)
      for (int k = 0; k < nsp; ++k)
        for (int l = 0; l < nsp; ++l)
          for (int m = 0; m < 4; ++m) {
            double r1 = 0, r2 = 0;
            for (int q = 0; q < nsp; ++q) {
              r1 += A5(fx, m, l, q, j, i) * lpdm[k * nsp + q];
              r2 += A5(fy, m, q, k, j, i) * lpdm[l * nsp + q];
            }
            A5(du, m, l, k, j, i) =
                -(r1 + r2 + (FXI(m, l, j, i) / Jx - FC(fxf, m, 0, l, j, i)) * dhl[k] +
                  (FXI(m, l, j, i + 1) / Jx - FC(fxf, m, 1, l, j, i)) * dhr[k] +
                  (FYI(m, k, j, i) / Jy - FC(fyf, m, 0, k, j, i)) * dhl[l] +
                  (FYI(m, k, j + 1, i) / Jy - FC(fyf, m, 1, k, j, i)) * dhr[l]);
          }
  free(fx); free(fy); free(uxf); free(uyf); free(fxf); free(fyf); free(fxi); free(fyi);
  return 0;
}

/* Euler-forward loop of ns_cavity.jl:380-384 */
int fro_integrate_ns2d(double *u, int nx, int ny, int nsp, double Jx, double Jy, const double *ll,
                       const double *lr, const double *lpdm, const double *dhl, const double *dhr,
                       const double *dll, const double *dlr, double K, double g, double mu,
                       double omega, double dt, double lam0, double lid, int nsteps) {
  const size_t n = (size_t)4 * nsp * nsp * (ny + 2) * (nx + 2);
  double *du = (double *)malloc(sizeof(double) * n);
  if (!du) return -2;
  for (int s = 0; s < nsteps; ++s) {
    int rc = fro_rhs_ns2d(u, du, nx, ny, nsp, Jx, Jy, ll, lr, lpdm, dhl, dhr, dll, dlr, K, g, mu, omega, dt,
                          lam0, lid);
    if (rc) { free(du); return rc; }
    for (size_t q = 0; q < n; ++q) u[q] = u[q] + dt * du[q];
  }
  free(du);
  return 0;
}

/* [KB] ref_vhs_vis(Kn, alpha, omega) */
double fro_ref_vhs_vis(double Kn, double alpha, double omega) {
  return 5.0 * (alpha + 1.0) * (alpha + 2.0) * sqrt(M_PI) /
         (4.0 * alpha * (5.0 - 2.0 * omega) * (7.0 - 2.0 * omega)) * Kn;
}

int fro_gks_available(void) { return 1; }
