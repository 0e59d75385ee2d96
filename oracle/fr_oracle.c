/* C restatement of the FluxReconstruction.jl hot path  --  TEST INFRASTRUCTURE ONLY.
 *
 * This is the oracle (and the timed CPU baseline): a loop-for-loop CPU restatement
 * of the reference's residual + explicit step.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (libfrb200.so) never links or calls it.
 *
 * PARITY UNPINNED: Julia and KitBase.jl are absent from this image and the
 * reference's tests pin nothing on this path (see oracle/fr_oracle.py header).
 * Functions marked [KB] restate KitBase.jl 0.9 closures from their published form.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off).
 * Citations are relative to /root/reference.  All arrays are Julia column-major
 * (first index fastest), 0-based here; ghost cells are explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FRO_MAXSP 8

int fro_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void fro_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* operators: src/Polynomial/poly_legendre.jl, poly_lagrange.jl               */
/* ------------------------------------------------------------------------- */

static void legendre_pd(int n, double x, double *P, double *dP) {
  /* P_n(x) and P_n'(x) by the three-term recurrence */
  double p0 = 1.0, p1 = x, d0 = 0.0, d1 = 1.0;
  if (n == 0) { *P = 1.0; *dP = 0.0; return; }
  for (int l = 1; l < n; ++l) {
    double p2 = ((2 * l + 1) * x * p1 - l * p0) / (l + 1);
    double d2 = d0 + (2 * l + 1) * p1;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
  *P = p1; *dP = d1;
}

/* Gauss-Legendre nodes/weights (FastGaussQuadrature.gausslegendre; poly_legendre.jl:6) */
void fro_gausslegendre(int n, double *x, double *w) {
  for (int i = 0; i < n; ++i) {
    double z = -cos(M_PI * (i + 0.75) / (n + 0.5));
    double P, dP;
    for (int it = 0; it < 100; ++it) {
      legendre_pd(n, z, &P, &dP);
      double dz = P / dP;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    legendre_pd(n, z, &P, &dP);
    x[i] = z;
    w[i] = 2.0 / ((1.0 - z * z) * dP * dP);
  }
  for (int i = 0; i < n / 2; ++i) { /* enforce exact symmetry like FastGaussQuadrature */
    double a = 0.5 * (x[n - 1 - i] - x[i]);
    x[i] = -a; x[n - 1 - i] = a;
    double b = 0.5 * (w[i] + w[n - 1 - i]);
    w[i] = b; w[n - 1 - i] = b;
  }
  if (n % 2) x[n / 2] = 0.0;
}

/* poly_lagrange.jl:6-21 */
static void lagrange_point(int nsp, const double *sp, double x, double *l) {
  for (int k = 0; k < nsp; ++k) {
    double tmp = 1.0;
    for (int j = 0; j < nsp; ++j)
      if (j != k) tmp *= (x - sp[j]) / (sp[k] - sp[j]);
    l[k] = tmp;
  }
}

/* poly_lagrange.jl:38-59: lpdm[m,k] (row-major here: lpdm[m*nsp+k]) */
static void dlagrange(int nsp, const double *sp, double *lpdm) {
  for (int k = 0; k < nsp; ++k)
    for (int m = 0; m < nsp; ++m) {
      double lsum = 0.0;
      for (int l = 0; l < nsp; ++l) {
        double tmp = 1.0;
        for (int j = 0; j < nsp; ++j)
          if (j != k && j != l) tmp *= (sp[m] - sp[j]) / (sp[k] - sp[j]);
        if (l != k) lsum += tmp / (sp[k] - sp[l]);
      }
      lpdm[m * nsp + k] = lsum;
    }
}

/* correction: 0 = radau (poly_legendre.jl:29-37), 1 = sd (:44-54), 2 = huynh (:61-71).
 * lpdm is row-major [m][k].  Returns 0 on success. */
int fro_operators(int deg, int correction, double *r, double *w, double *ll, double *lr,
                  double *lpdm, double *dgl, double *dgr) {
  int nsp = deg + 1;
  if (nsp > FRO_MAXSP || deg < 1) return -1;
  fro_gausslegendre(nsp, r, w);
  lagrange_point(nsp, r, -1.0, ll);
  lagrange_point(nsp, r, 1.0, lr);
  dlagrange(nsp, r, lpdm);
  double sgn = (deg % 2) ? -1.0 : 1.0;
  for (int i = 0; i < nsp; ++i) {
    double P, dm, d, dp;
    legendre_pd(deg - 1, r[i], &P, &dm);
    legendre_pd(deg, r[i], &P, &d);
    legendre_pd(deg + 1, r[i], &P, &dp);
    double y;
    if (correction == 0) y = dp;
    else if (correction == 1) y = (deg * dm + (deg + 1) * dp) / (2 * deg + 1);
    else y = ((deg + 1) * dm + deg * dp) / (2 * deg + 1);
    dgl[i] = sgn * 0.5 * (d - y);
    dgr[i] = 0.5 * (d + y);
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* [KB] pointwise physics                                                     */
/* ------------------------------------------------------------------------- */

static inline void conserve_prim3(const double *W, double g, double *prim) {
  prim[0] = W[0];
  prim[1] = W[1] / W[0];
  prim[2] = 0.5 * W[0] / (g - 1.0) / (W[2] - 0.5 * (W[1] * W[1]) / W[0]);
}
static inline void conserve_prim4(const double *W, double g, double *prim) {
  prim[0] = W[0];
  prim[1] = W[1] / W[0];
  prim[2] = W[2] / W[0];
  prim[3] = 0.5 * W[0] / (g - 1.0) / (W[3] - 0.5 * (W[1] * W[1] + W[2] * W[2]) / W[0]);
}
static inline void euler_flux3(const double *w, double g, double *F) {
  double prim[3];
  conserve_prim3(w, g, prim);
  double p = 0.5 * prim[0] / prim[2];
  F[0] = w[1];
  F[1] = (w[1] * w[1]) / w[0] + p;
  F[2] = (w[2] + p) * w[1] / w[0];
}
static inline void euler_flux4(const double *w, double g, double *F, double *G) {
  double prim[4];
  conserve_prim4(w, g, prim);
  double p = 0.5 * prim[0] / prim[3];
  F[0] = w[1];
  F[1] = (w[1] * w[1]) / w[0] + p;
  F[2] = w[1] * w[2] / w[0];
  F[3] = (w[3] + p) * w[1] / w[0];
  if (G) {
    G[0] = w[2];
    G[1] = w[2] * w[1] / w[0];
    G[2] = (w[2] * w[2]) / w[0] + p;
    G[3] = (w[3] + p) * w[2] / w[0];
  }
}
/* flux_hll!(fw, wL, wR, γ, dt) for 3 (1-D) or 4 (2-D) components */
static inline void flux_hll(int nv, double *fw, const double *wL, const double *wR, double g,
                            double dt) {
  double pL[4], pR[4], f1[4], f2[4];
  if (nv == 3) { conserve_prim3(wL, g, pL); conserve_prim3(wR, g, pR); }
  else { conserve_prim4(wL, g, pL); conserve_prim4(wR, g, pR); }
  double aL = sqrt(0.5 * g / pL[nv - 1]);
  double aR = sqrt(0.5 * g / pR[nv - 1]);
  double lmin = pL[1] - aL, lmax = pR[1] + aR;
  if (nv == 3) { euler_flux3(wL, g, f1); euler_flux3(wR, g, f2); }
  else { euler_flux4(wL, g, f1, 0); euler_flux4(wR, g, f2, 0); }
  if (lmin >= 0.0) {
    for (int m = 0; m < nv; ++m) fw[m] = f1[m];
  } else if (lmax <= 0.0) {
    for (int m = 0; m < nv; ++m) fw[m] = f2[m];
  } else {
    double factor = 1.0 / (lmax - lmin);
    for (int m = 0; m < nv; ++m)
      fw[m] = factor * (lmax * f1[m] - lmin * f2[m] + (lmax * lmin) * (wR[m] - wL[m]));
  }
  for (int m = 0; m < nv; ++m) fw[m] *= dt;
}

/* ------------------------------------------------------------------------- */
/* 1-D advection: eq_advection.jl:55-175, eq_scalar.jl:1-12,                  */
/* example/advection_lowlevel.jl:4-47 (variant 1)                             */
/* ------------------------------------------------------------------------- */
/* u, du [ncell, nsp]; J[ncell]; bc: 0 dirichlet, 1 period; variant: 0 packaged, 1 lowlevel */
int fro_rhs_adv1d(const double *u, double *du, int ncell, int nsp, const double *J,
                  const double *ll, const double *lr, const double *lpdm, const double *dgl,
                  const double *dgr, double a, int bc, int variant) {
  if (nsp > FRO_MAXSP) return -1;
  double *f = (double *)malloc(sizeof(double) * ncell * nsp);
  double *rhs1 = (double *)malloc(sizeof(double) * ncell * nsp);
  double *uf = (double *)malloc(sizeof(double) * ncell * 2);
  double *ff = (double *)malloc(sizeof(double) * ncell * 2);
  double *fi = (double *)calloc(ncell + 1, sizeof(double));
#define U2(i, p) u[(i) + (size_t)ncell * (p)]
#define F2(i, p) f[(i) + (size_t)ncell * (p)]
  for (int p = 0; p < nsp; ++p)
    for (int i = 0; i < ncell; ++i) F2(i, p) = a * U2(i, p) / J[i];
  for (int i = 0; i < ncell; ++i) {
    double ul = U2(i, 0) * ll[0], ur = U2(i, 0) * lr[0];
    double fl = F2(i, 0) * ll[0], fr = F2(i, 0) * lr[0];
    for (int q = 1; q < nsp; ++q) {
      ul = ul + U2(i, q) * ll[q]; ur = ur + U2(i, q) * lr[q];
      fl = fl + F2(i, q) * ll[q]; fr = fr + F2(i, q) * lr[q];
    }
    uf[i] = ul; uf[i + ncell] = ur; ff[i] = fl; ff[i + ncell] = fr;
  }
  for (int i = 1; i < ncell; ++i) {
    double au = (ff[i] - ff[i - 1 + ncell]) / (uf[i] - uf[i - 1 + ncell] + 1e-8);
    fi[i] = 0.5 * (ff[i] + ff[i - 1 + ncell]) - 0.5 * fabs(au) * (uf[i] - uf[i - 1 + ncell]);
  }
  for (int p = 0; p < nsp; ++p)
    for (int i = 0; i < ncell; ++i) {
      double acc = F2(i, 0) * lpdm[p * nsp + 0];
      for (int q = 1; q < nsp; ++q) acc = acc + F2(i, q) * lpdm[p * nsp + q];
      rhs1[i + (size_t)ncell * p] = acc;
    }
  int periodic = (bc == 1) || (variant == 1);
  int c0 = 1, c1 = ncell - 1;
  if (periodic) {
    double eps = variant == 1 ? 1e-8 : 1e-6;
    double au = (ff[0] - ff[ncell - 1 + ncell]) / (uf[0] - uf[ncell - 1 + ncell] + eps);
    fi[0] = 0.5 * (ff[ncell - 1 + ncell] + ff[0]) - 0.5 * fabs(au) * (uf[0] - uf[ncell - 1 + ncell]);
    fi[ncell] = fi[0];
    c0 = 0; c1 = ncell;
  }
  memset(du, 0, sizeof(double) * ncell * nsp);
  for (int p = 0; p < nsp; ++p)
    for (int i = c0; i < c1; ++i)
      du[i + (size_t)ncell * p] = -(rhs1[i + (size_t)ncell * p] + (fi[i] - ff[i]) * dgl[p] +
                                    (fi[i + 1] - ff[i + ncell]) * dgr[p]);
#undef U2
#undef F2
  free(f); free(rhs1); free(uf); free(ff); free(fi);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* 1-D Euler: eq_euler.jl:29-98                                               */
/* ------------------------------------------------------------------------- */
/* u, du [ncell, nsp, 3]; bc: 0 dirichlet, 1 period */
int fro_rhs_euler1d(const double *u, double *du, int ncell, int nsp, const double *J,
                    const double *ll, const double *lr, const double *lpdm, const double *dgl,
                    const double *dgr, double gamma, int bc) {
  if (nsp > FRO_MAXSP) return -1;
  size_t cs = (size_t)ncell, ps = cs * nsp;
  double *f = (double *)malloc(sizeof(double) * ps * 3);
  double *rhs1 = (double *)malloc(sizeof(double) * ps * 3);
  double *uf = (double *)malloc(sizeof(double) * cs * 2 * 3);
  double *ff = (double *)malloc(sizeof(double) * cs * 2 * 3);
  double *fi = (double *)calloc((cs + 1) * 3, sizeof(double));
#define U3(i, p, k) u[(i) + cs * (p) + ps * (k)]
#define F3(i, p, k) f[(i) + cs * (p) + ps * (k)]
#define UF(i, s, k) uf[(i) + cs * (s) + 2 * cs * (k)]
#define FF(i, s, k) ff[(i) + cs * (s) + 2 * cs * (k)]
#define FI(i, k) fi[(i) + (cs + 1) * (k)]
#pragma omp parallel for schedule(static)
  for (int i = 0; i < ncell; ++i)
    for (int p = 0; p < nsp; ++p) {
      double w[3] = {U3(i, p, 0), U3(i, p, 1), U3(i, p, 2)}, F[3];
      euler_flux3(w, gamma, F);
      for (int k = 0; k < 3; ++k) F3(i, p, k) = F[k] / J[i];
    }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < ncell; ++i)
    for (int k = 0; k < 3; ++k) {
      double ul = U3(i, 0, k) * ll[0], ur = U3(i, 0, k) * lr[0];
      double fl = F3(i, 0, k) * ll[0], fr = F3(i, 0, k) * lr[0];
      for (int q = 1; q < nsp; ++q) {
        ul = ul + U3(i, q, k) * ll[q]; ur = ur + U3(i, q, k) * lr[q];
        fl = fl + F3(i, q, k) * ll[q]; fr = fr + F3(i, q, k) * lr[q];
      }
      UF(i, 0, k) = ul; UF(i, 1, k) = ur; FF(i, 0, k) = fl; FF(i, 1, k) = fr;
    }
#pragma omp parallel for schedule(static)
  for (int i = 1; i < ncell; ++i) {
    double wL[3] = {UF(i - 1, 1, 0), UF(i - 1, 1, 1), UF(i - 1, 1, 2)};
    double wR[3] = {UF(i, 0, 0), UF(i, 0, 1), UF(i, 0, 2)}, fw[3];
    flux_hll(3, fw, wL, wR, gamma, 1.0);
    for (int k = 0; k < 3; ++k) FI(i, k) = fw[k];
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < ncell; ++i)
    for (int k = 0; k < 3; ++k)
      for (int p = 0; p < nsp; ++p) {
        double acc = F3(i, 0, k) * lpdm[p * nsp + 0];
        for (int q = 1; q < nsp; ++q) acc = acc + F3(i, q, k) * lpdm[p * nsp + q];
        rhs1[i + cs * p + ps * k] = acc;
      }
  int c0 = 1, c1 = ncell - 1;
  if (bc == 1) {
    double wL[3] = {UF(ncell - 1, 1, 0), UF(ncell - 1, 1, 1), UF(ncell - 1, 1, 2)};
    double wR[3] = {UF(0, 0, 0), UF(0, 0, 1), UF(0, 0, 2)}, fw[3];
    flux_hll(3, fw, wL, wR, gamma, 1.0);
    for (int k = 0; k < 3; ++k) { FI(0, k) = fw[k]; FI(ncell, k) = fw[k]; }
    c0 = 0; c1 = ncell;
  }
  memset(du, 0, sizeof(double) * ps * 3);
#pragma omp parallel for schedule(static)
  for (int i = c0; i < c1; ++i)
    for (int p = 0; p < nsp; ++p)
      for (int k = 0; k < 3; ++k)
        du[i + cs * p + ps * k] = -(rhs1[i + cs * p + ps * k] +
                                    (FI(i, k) / J[i] - FF(i, 0, k)) * dgl[p] +
                                    (FI(i + 1, k) / J[i] - FF(i, 1, k)) * dgr[p]);
#undef U3
#undef F3
#undef UF
#undef FF
#undef FI
  free(f); free(rhs1); free(uf); free(ff); free(fi);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* 2-D Euler: example/euler2d_wave.jl:35-107 in the preallocated, threaded form */
/* of example/shock-vortex.jl:26-118                                          */
/* ------------------------------------------------------------------------- */
typedef struct {
  int nx, ny, nsp;
  double *f, *u_face, *f_face, *fx, *fy, *rhs1, *rhs2;
} fro_work2d;

void *fro_work2d_create(int nx, int ny, int nsp) {
  fro_work2d *w = (fro_work2d *)calloc(1, sizeof(fro_work2d));
  size_t ne = (size_t)(nx + 2) * (ny + 2);
  w->nx = nx; w->ny = ny; w->nsp = nsp;
  w->f = (double *)malloc(sizeof(double) * ne * nsp * nsp * 4 * 2);
  w->u_face = (double *)malloc(sizeof(double) * ne * 4 * nsp * 4);
  w->f_face = (double *)malloc(sizeof(double) * ne * 4 * nsp * 4 * 2);
  w->fx = (double *)malloc(sizeof(double) * (size_t)(nx + 1) * ny * nsp * 4);
  w->fy = (double *)malloc(sizeof(double) * (size_t)nx * (ny + 1) * nsp * 4);
  w->rhs1 = (double *)malloc(sizeof(double) * (size_t)nx * ny * nsp * nsp * 4);
  w->rhs2 = (double *)malloc(sizeof(double) * (size_t)nx * ny * nsp * nsp * 4);
  if (!w->f || !w->u_face || !w->f_face || !w->fx || !w->fy || !w->rhs1 || !w->rhs2) {
    free(w->f); free(w->u_face); free(w->f_face); free(w->fx); free(w->fy);
    free(w->rhs1); free(w->rhs2); free(w);
    return 0;
  }
  return w;
}
void fro_work2d_destroy(void *p) {
  fro_work2d *w = (fro_work2d *)p;
  if (!w) return;
  free(w->f); free(w->u_face); free(w->f_face); free(w->fx); free(w->fy);
  free(w->rhs1); free(w->rhs2); free(w);
}

/* u, du [nx+2, ny+2, nsp, nsp, 4]; Jx = dx/2, Jy = dy/2 (J[i,j][k,l] = diag(Jx,Jy)). */
int fro_rhs_euler2d(const double *u, double *du, void *work, double Jx, double Jy,
                    const double *ll, const double *lr, const double *lpdm, const double *dhl,
                    const double *dhr, double gamma) {
  fro_work2d *w = (fro_work2d *)work;
  const int nx = w->nx, ny = w->ny, nsp = w->nsp;
  if (nsp > FRO_MAXSP) return -1;
  const size_t NXG = nx + 2, NYG = ny + 2, NE = NXG * NYG;
  const double iJ11 = 1.0 / Jx, iJ22 = 1.0 / Jy;
  double *f = w->f, *u_face = w->u_face, *f_face = w->f_face, *fx = w->fx, *fy = w->fy;
  double *rhs1 = w->rhs1, *rhs2 = w->rhs2;
#define U5(i, j, k, l, m) u[(i) + NXG * (j) + NE * ((k) + nsp * ((l) + nsp * (size_t)(m)))]
#define DU5(i, j, k, l, m) du[(i) + NXG * (j) + NE * ((k) + nsp * ((l) + nsp * (size_t)(m)))]
#define F6(i, j, k, l, m, n) f[(i) + NXG * (j) + NE * ((k) + nsp * ((l) + nsp * ((m) + 4 * (size_t)(n))))]
#define UFC(i, j, fc, l, m) u_face[(i) + NXG * (j) + NE * ((fc) + 4 * ((l) + nsp * (size_t)(m)))]
#define FFC(i, j, fc, l, m, n) f_face[(i) + NXG * (j) + NE * ((fc) + 4 * ((l) + nsp * ((m) + 4 * (size_t)(n))))]
#define FX(i, j, k, m) fx[(i) + (size_t)(nx + 1) * ((j) + (size_t)ny * ((k) + nsp * (size_t)(m)))]
#define FY(i, j, k, m) fy[(i) + (size_t)nx * ((j) + (size_t)(ny + 1) * ((k) + nsp * (size_t)(m)))]
#define R1(i, j, k, l, m) rhs1[(i) + (size_t)nx * ((j) + (size_t)ny * ((k) + nsp * ((l) + nsp * (size_t)(m))))]
#define R2(i, j, k, l, m) rhs2[(i) + (size_t)nx * ((j) + (size_t)ny * ((k) + nsp * ((l) + nsp * (size_t)(m))))]

  /* du .= 0 (:36) */
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < NE * nsp * nsp * 4; ++q) du[q] = 0.0;

  /* point fluxes, ghosts included (:45-50 / shock-vortex :48-55) */
#pragma omp parallel for collapse(2) schedule(static)
  for (int l = 0; l < nsp; ++l)
    for (int k = 0; k < nsp; ++k)
      for (size_t j = 0; j < NYG; ++j)
        for (size_t i = 0; i < NXG; ++i) {
          double wv[4] = {U5(i, j, k, l, 0), U5(i, j, k, l, 1), U5(i, j, k, l, 2), U5(i, j, k, l, 3)};
          double F[4], G[4];
          euler_flux4(wv, gamma, F, G);
          for (int s = 0; s < 4; ++s) {
            /* inv(J) * [F, G] with J = diag(Jx, Jy) */
            F6(i, j, k, l, s, 0) = iJ11 * F[s] + 0.0 * G[s];
            F6(i, j, k, l, s, 1) = 0.0 * F[s] + iJ22 * G[s];
          }
        }

  /* traces (:54-66 / shock-vortex :57-71) */
#pragma omp parallel for collapse(2) schedule(static)
  for (int m = 0; m < 4; ++m)
    for (int l = 0; l < nsp; ++l)
      for (size_t j = 0; j < NYG; ++j)
        for (size_t i = 0; i < NXG; ++i) {
          double a1 = U5(i, j, l, 0, m) * ll[0], a2 = U5(i, j, 0, l, m) * lr[0];
          double a3 = U5(i, j, l, 0, m) * lr[0], a4 = U5(i, j, 0, l, m) * ll[0];
          for (int q = 1; q < nsp; ++q) {
            a1 = a1 + U5(i, j, l, q, m) * ll[q];
            a2 = a2 + U5(i, j, q, l, m) * lr[q];
            a3 = a3 + U5(i, j, l, q, m) * lr[q];
            a4 = a4 + U5(i, j, q, l, m) * ll[q];
          }
          UFC(i, j, 0, l, m) = a1; UFC(i, j, 1, l, m) = a2;
          UFC(i, j, 2, l, m) = a3; UFC(i, j, 3, l, m) = a4;
          for (int n = 0; n < 2; ++n) {
            double b1 = F6(i, j, l, 0, m, n) * ll[0], b2 = F6(i, j, 0, l, m, n) * lr[0];
            double b3 = F6(i, j, l, 0, m, n) * lr[0], b4 = F6(i, j, 0, l, m, n) * ll[0];
            for (int q = 1; q < nsp; ++q) {
              b1 = b1 + F6(i, j, l, q, m, n) * ll[q];
              b2 = b2 + F6(i, j, q, l, m, n) * lr[q];
              b3 = b3 + F6(i, j, l, q, m, n) * lr[q];
              b4 = b4 + F6(i, j, q, l, m, n) * ll[q];
            }
            FFC(i, j, 0, l, m, n) = b1; FFC(i, j, 1, l, m, n) = b2;
            FFC(i, j, 2, l, m, n) = b3; FFC(i, j, 3, l, m, n) = b4;
          }
        }

  /* x faces (:68-74): reference i in 1:nx+1 -> fx index i-1 */
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nsp; ++k)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx + 1; ++i) {
        double uL[4], uR[4], fw[4];
        for (int m = 0; m < 4; ++m) { uL[m] = UFC(i - 1, j, 1, k, m); uR[m] = UFC(i, j, 3, k, m); }
        flux_hll(4, fw, uL, uR, gamma, 1.0);
        for (int m = 0; m < 4; ++m) FX(i - 1, j - 1, k, m) = fw[m];
      }
  /* y faces (:75-82): local_frame(.,0,1), HLL, global_frame(.,0,1) */
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nsp; ++k)
    for (int j = 1; j <= ny + 1; ++j)
      for (int i = 1; i <= nx; ++i) {
        double a[4], b[4], uL[4], uR[4], fw[4];
        for (int m = 0; m < 4; ++m) { a[m] = UFC(i, j - 1, 2, k, m); b[m] = UFC(i, j, 0, k, m); }
        const double c = 0.0, s = 1.0;
        uL[0] = a[0]; uL[1] = a[1] * c + a[2] * s; uL[2] = a[2] * c - a[1] * s; uL[3] = a[3];
        uR[0] = b[0]; uR[1] = b[1] * c + b[2] * s; uR[2] = b[2] * c - b[1] * s; uR[3] = b[3];
        flux_hll(4, fw, uL, uR, gamma, 1.0);
        FY(i - 1, j - 1, k, 0) = fw[0];
        FY(i - 1, j - 1, k, 1) = fw[1] * c - fw[2] * s;
        FY(i - 1, j - 1, k, 2) = fw[1] * s + fw[2] * c;
        FY(i - 1, j - 1, k, 3) = fw[3];
      }

  /* derivatives (:84-91) */
#pragma omp parallel for collapse(2) schedule(static)
  for (int m = 0; m < 4; ++m)
    for (int l = 0; l < nsp; ++l)
      for (int k = 0; k < nsp; ++k)
        for (int j = 1; j <= ny; ++j)
          for (int i = 1; i <= nx; ++i) {
            double a = F6(i, j, 0, l, m, 0) * lpdm[k * nsp + 0];
            double b = F6(i, j, k, 0, m, 1) * lpdm[l * nsp + 0];
            for (int q = 1; q < nsp; ++q) {
              a = a + F6(i, j, q, l, m, 0) * lpdm[k * nsp + q];
              b = b + F6(i, j, k, q, m, 1) * lpdm[l * nsp + q];
            }
            R1(i - 1, j - 1, k, l, m) = a;
            R2(i - 1, j - 1, k, l, m) = b;
          }

  /* correction (:93-104) */
#pragma omp parallel for collapse(2) schedule(static)
  for (int m = 0; m < 4; ++m)
    for (int l = 0; l < nsp; ++l)
      for (int k = 0; k < nsp; ++k)
        for (int j = 1; j <= ny; ++j)
          for (int i = 1; i <= nx; ++i) {
            DU5(i, j, k, l, m) =
                -(R1(i - 1, j - 1, k, l, m) + R2(i - 1, j - 1, k, l, m) +
                  (FX(i - 1, j - 1, l, m) * iJ11 - FFC(i, j, 3, l, m, 0)) * dhl[k] +
                  (FX(i, j - 1, l, m) * iJ11 - FFC(i, j, 1, l, m, 0)) * dhr[k] +
                  (FY(i - 1, j - 1, k, m) * iJ22 - FFC(i, j, 0, k, m, 1)) * dhl[l] +
                  (FY(i - 1, j, k, m) * iJ22 - FFC(i, j, 2, k, m, 1)) * dhr[l]);
          }
#undef F6
#undef UFC
#undef FFC
#undef FX
#undef FY
#undef R1
#undef R2
  return 0;
}

/* per-step ghost fill: mode 0 = euler2d_wave.jl:127-132 (x wave), 1 = :159-164 (y wave),
 * 2 = shock-vortex.jl:324-326 (copy) */
void fro_ghost_fill_euler2d(double *u, int nx, int ny, int nsp, int mode) {
  const size_t NXG = nx + 2, NYG = ny + 2, NE = NXG * NYG;
  const int npl = nsp * nsp * 4;
#define UP(i, j, p) u[(i) + NXG * (j) + NE * (size_t)(p)]
  for (int p = 0; p < npl; ++p) {
    int m = p / (nsp * nsp);
    if (mode == 0) {
      for (size_t j = 0; j < NYG; ++j) UP(0, j, p) = UP(nx, j, p);
      for (size_t j = 0; j < NYG; ++j) UP(nx + 1, j, p) = UP(1, j, p);
      double sg = (m == 2) ? -1.0 : 1.0;
      for (size_t i = 0; i < NXG; ++i) UP(i, 0, p) = sg * UP(i, ny, p);
      for (size_t i = 0; i < NXG; ++i) UP(i, ny + 1, p) = sg * UP(i, 1, p);
    } else if (mode == 1) {
      for (size_t i = 0; i < NXG; ++i) UP(i, 0, p) = UP(i, ny, p);
      for (size_t i = 0; i < NXG; ++i) UP(i, ny + 1, p) = UP(i, 1, p);
      double sg = (m == 1) ? -1.0 : 1.0;
      for (size_t j = 0; j < NYG; ++j) UP(0, j, p) = sg * UP(nx, j, p);
      for (size_t j = 0; j < NYG; ++j) UP(nx + 1, j, p) = sg * UP(1, j, p);
    } else {
      for (size_t i = 0; i < NXG; ++i) UP(i, 0, p) = UP(i, 1, p);
      for (size_t i = 0; i < NXG; ++i) UP(i, ny + 1, p) = UP(i, ny, p);
      for (size_t j = 0; j < NYG; ++j) UP(nx + 1, j, p) = UP(nx, j, p);
    }
  }
#undef UP
}

/* ------------------------------------------------------------------------- */
/* 1-D BGK: example/bgk_wave.jl:69-129                                        */
/* ------------------------------------------------------------------------- */
/* u, du [ncell, nu, nsp]; dx[ncell]; velo, weights [nu] */
int fro_rhs_bgk1d(const double *u, double *du, int ncell, int nu, int nsp, const double *dx,
                  const double *velo, const double *weights, const double *ll, const double *lr,
                  const double *lpdm, const double *dgl, const double *dgr, double tau) {
  if (nsp > FRO_MAXSP) return -1;
  const size_t cs = ncell, vs = cs * nu;
  double *M = (double *)malloc(sizeof(double) * vs * nsp);
  double *f = (double *)malloc(sizeof(double) * vs * nsp);
  double *ff = (double *)malloc(sizeof(double) * vs * 2);
  double *fi = (double *)malloc(sizeof(double) * (cs + 1) * nu);
#define U3(i, j, k) u[(i) + cs * (j) + vs * (k)]
#define F3(i, j, k) f[(i) + cs * (j) + vs * (k)]
#pragma omp parallel for schedule(static)
  for (int i = 0; i < ncell; ++i)
    for (int k = 0; k < nsp; ++k) {
      double w0 = 0, w1 = 0, w2 = 0;
      for (int j = 0; j < nu; ++j) {
        double fv = U3(i, j, k);
        w0 += weights[j] * fv;
        w1 += weights[j] * velo[j] * fv;
        w2 += weights[j] * (velo[j] * velo[j]) * fv;
      }
      double W[3] = {w0, w1, 0.5 * w2}, prim[3];
      conserve_prim3(W, 3.0, prim);
      for (int j = 0; j < nu; ++j) {
        double c = velo[j] - prim[1];
        M[i + cs * j + vs * k] = prim[0] * sqrt(prim[2] / M_PI) * exp(-prim[2] * (c * c));
      }
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < nu; ++j)
    for (int i = 0; i < ncell; ++i) {
      double J = 0.5 * dx[i];
      for (int k = 0; k < nsp; ++k) F3(i, j, k) = velo[j] * U3(i, j, k) / J;
      double a = F3(i, j, 0) * ll[0], b = F3(i, j, 0) * lr[0];
      for (int q = 1; q < nsp; ++q) { a = a + F3(i, j, q) * ll[q]; b = b + F3(i, j, q) * lr[q]; }
      ff[i + cs * j] = a; ff[i + cs * j + vs] = b;
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < nu; ++j) {
    double dl = velo[j] >= 0 ? 1.0 : 0.0;
    for (int fc = 0; fc <= ncell; ++fc) {
      int e1 = (fc == ncell) ? 0 : fc;             /* f2e[i,1] */
      int e2 = (fc == 0) ? ncell - 1 : fc - 1;     /* f2e[i,2] */
      fi[fc + (cs + 1) * j] = ff[e1 + cs * j] * (1.0 - dl) + ff[e2 + cs * j + vs] * dl;
    }
  }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < nu; ++j)
    for (int i = 0; i < ncell; ++i) {
      int fl = (i == 0) ? ncell : i;               /* e2f[i,2] */
      int fr = (i == ncell - 1) ? 0 : i + 1;       /* e2f[i,1] */
      for (int p = 0; p < nsp; ++p) {
        double r1 = F3(i, j, 0) * lpdm[p * nsp + 0];
        for (int q = 1; q < nsp; ++q) r1 = r1 + F3(i, j, q) * lpdm[p * nsp + q];
        du[i + cs * j + vs * p] =
            -(r1 + (fi[fl + (cs + 1) * j] - ff[i + cs * j]) * dgl[p] +
              (fi[fr + (cs + 1) * j] - ff[i + cs * j + vs]) * dgr[p]) +
            (M[i + cs * j + vs * p] - U3(i, j, p)) / tau;
      }
    }
#undef U3
#undef F3
  free(M); free(f); free(ff); free(fi);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* positivity limiter: dissipation.jl:61-123 (1-D), :125-206 (2-D), density branch */
/* (the energy branch of the reference throws at :116/:199; see fr_oracle.py)  */
/* ------------------------------------------------------------------------- */
/* returns the number of cells whose t1 fell outside (0,1] (the reference asserts) */
int fro_limiter_euler1d(double *u, int ncell, int nsp, double gamma, const double *weights,
                        const double *ll, const double *lr) {
  const size_t cs = ncell, ps = cs * nsp;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int i = 0; i < ncell; ++i) {
    double um[3], prim[3];
    for (int k = 0; k < 3; ++k) {
      double s = 0.0;
      for (int p = 0; p < nsp; ++p) s += u[i + cs * p + ps * k] * weights[p];
      um[k] = s;
    }
    conserve_prim3(um, gamma, prim);
    double t_mean = 1.0 / prim[2];
    double p_mean = 0.5 * um[0] * t_mean;
    double rl = u[i] * ll[0], rr = u[i] * lr[0];
    for (int q = 1; q < nsp; ++q) { rl = rl + u[i + cs * q] * ll[q]; rr = rr + u[i + cs * q] * lr[q]; }
    double eps = fmin(fmin(1e-13, um[0]), p_mean);
    double rmin = fmin(rl, rr);
    for (int p = 0; p < nsp; ++p) rmin = fmin(rmin, u[i + cs * p]);
    double t1 = fmin((um[0] - eps) / (um[0] - rmin + 1e-8), 1.0);
    if (!(t1 > 0 && t1 <= 1)) bad += 1;
    for (int p = 0; p < nsp; ++p) u[i + cs * p] = t1 * (u[i + cs * p] - um[0]) + um[0];
  }
  return bad;
}

int fro_limiter_euler2d(double *u, int nx, int ny, int nsp, double gamma, const double *weights,
                        const double *ll, const double *lr) {
  const size_t NXG = nx + 2, NYG = ny + 2, NE = NXG * NYG;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      double um[4], prim[4];
      for (int m = 0; m < 4; ++m) {
        double s = 0.0;
        for (int l = 0; l < nsp; ++l)
          for (int k = 0; k < nsp; ++k)
            s += U5(i, j, k, l, m) * weights[k + nsp * l];
        um[m] = s;
      }
      conserve_prim4(um, gamma, prim);
      double p_mean = 0.5 * um[0] * (1.0 / prim[3]);
      double eps = fmin(fmin(1e-13, um[0]), p_mean);
      double rmin = INFINITY;
      for (int a = 0; a < nsp; ++a) {
        double b1 = U5(i, j, a, 0, 0) * ll[0], b2 = U5(i, j, 0, a, 0) * lr[0];
        double b3 = U5(i, j, a, 0, 0) * lr[0], b4 = U5(i, j, 0, a, 0) * ll[0];
        for (int q = 1; q < nsp; ++q) {
          b1 = b1 + U5(i, j, a, q, 0) * ll[q]; b2 = b2 + U5(i, j, q, a, 0) * lr[q];
          b3 = b3 + U5(i, j, a, q, 0) * lr[q]; b4 = b4 + U5(i, j, q, a, 0) * ll[q];
        }
        rmin = fmin(rmin, fmin(fmin(b1, b2), fmin(b3, b4)));
        for (int q = 0; q < nsp; ++q) rmin = fmin(rmin, U5(i, j, a, q, 0));
      }
      double t1 = fmin((um[0] - eps) / (um[0] - rmin + 1e-8), 1.0);
      if (!(t1 > 0 && t1 <= 1)) bad += 1;
      for (int l = 0; l < nsp; ++l)
        for (int k = 0; k < nsp; ++k) {
          size_t q = (i) + NXG * (j) + NE * ((k) + nsp * ((l) + nsp * (size_t)0));
          u[q] = t1 * (u[q] - um[0]) + um[0];
        }
    }
  return bad;
}
#undef U5
#undef DU5

/* ------------------------------------------------------------------------- */
/* fixed-step integrators: scheme 0 Euler, 1 Midpoint, 2 SSPRK3   */
/* ------------------------------------------------------------------------- */
static void axpby(double *out, double a, const double *x, double b, const double *y, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) out[q] = a * x[q] + b * y[q];
}

/* One explicit step given an RHS callback.  tmp: 3 scratch arrays of n doubles. */
typedef int (*fro_rhs_fn)(const double *u, double *du, void *ctx);
static void step_generic(double *u, size_t n, double dt, int scheme, fro_rhs_fn rhs, void *ctx,
                         double *k, double *ua, double *ub) {
  if (scheme == 0) {
    rhs(u, k, ctx);
    axpby(u, 1.0, u, dt, k, n);
  } else if (scheme == 1) {
    rhs(u, k, ctx);
    axpby(ua, 1.0, u, 0.5 * dt, k, n);
    rhs(ua, k, ctx);
    axpby(u, 1.0, u, dt, k, n);
  } else {
    rhs(u, k, ctx);
    axpby(ua, 1.0, u, dt, k, n); /* u1 */
    rhs(ua, k, ctx);
    axpby(ub, 1.0, ua, dt, k, n);
    axpby(ub, 0.75, u, 0.25, ub, n); /* u2 */
    rhs(ub, k, ctx);
    axpby(ua, 1.0, ub, dt, k, n);
    axpby(u, 1.0 / 3.0, u, 2.0 / 3.0, ua, n);
  }
}

typedef struct {
  void *work; double Jx, Jy; const double *ll, *lr, *lpdm, *dhl, *dhr; double gamma;
} ctx2d;
static int rhs2d_cb(const double *u, double *du, void *c) {
  ctx2d *x = (ctx2d *)c;
  return fro_rhs_euler2d(u, du, x->work, x->Jx, x->Jy, x->ll, x->lr, x->lpdm, x->dhl, x->dhr, x->gamma);
}
/* The user loop of euler2d_wave.jl:125-135: ghost fill (ghost_mode >= 0), optional
 * positivity limiter (limiter_weights != NULL; shock-vortex.jl:298-303), then one step. */
int fro_integrate_euler2d(double *u, int nx, int ny, int nsp, double Jx, double Jy,
                          const double *ll, const double *lr, const double *lpdm,
                          const double *dhl, const double *dhr, double gamma, double dt,
                          int nsteps, int scheme, int ghost_mode, const double *limiter_weights) {
  size_t n = (size_t)(nx + 2) * (ny + 2) * nsp * nsp * 4;
  /* the caches (temporaries of dudt!, stage arrays) are preallocated once per mesh size and kept between
   * calls, like the reference's threaded twin does (shock-vortex.jl:26-45): a timed step-by-step loop must
   * not pay 16 GB of page faults per call.  Single caller at a time (tests and bench.py are). */
  static struct { int nx, ny, nsp; void *work; double *k, *ua, *ub; } cache = {0, 0, 0, 0, 0, 0, 0};
  if (cache.nx != nx || cache.ny != ny || cache.nsp != nsp || !cache.work) {
    if (cache.work) { fro_work2d_destroy(cache.work); free(cache.k); free(cache.ua); free(cache.ub); }
    cache.work = fro_work2d_create(nx, ny, nsp);
    cache.k = (double *)malloc(sizeof(double) * n);
    cache.ua = (double *)malloc(sizeof(double) * n);
    cache.ub = (double *)malloc(sizeof(double) * n);
    cache.nx = nx; cache.ny = ny; cache.nsp = nsp;
    if (!cache.work || !cache.k || !cache.ua || !cache.ub) {
      if (cache.work) fro_work2d_destroy(cache.work);
      free(cache.k); free(cache.ua); free(cache.ub);
      cache.work = 0; cache.k = cache.ua = cache.ub = 0; cache.nx = 0;
      return -2;
    }
  }
  ctx2d c = {cache.work, Jx, Jy, ll, lr, lpdm, dhl, dhr, gamma};
  for (int s = 0; s < nsteps; ++s) {
    if (limiter_weights) fro_limiter_euler2d(u, nx, ny, nsp, gamma, limiter_weights, ll, lr);
    if (ghost_mode >= 0) fro_ghost_fill_euler2d(u, nx, ny, nsp, ghost_mode);
    step_generic(u, n, dt, scheme, rhs2d_cb, &c, cache.k, cache.ua, cache.ub);
  }
  return 0;
}

typedef struct {
  int ncell, nsp; const double *J, *ll, *lr, *lpdm, *dgl, *dgr; double gamma; int bc;
} ctx1d;
static int rhs1d_cb(const double *u, double *du, void *c) {
  ctx1d *x = (ctx1d *)c;
  return fro_rhs_euler1d(u, du, x->ncell, x->nsp, x->J, x->ll, x->lr, x->lpdm, x->dgl, x->dgr, x->gamma, x->bc);
}
int fro_integrate_euler1d(double *u, int ncell, int nsp, const double *J, const double *ll,
                          const double *lr, const double *lpdm, const double *dgl,
                          const double *dgr, double gamma, int bc, double dt, int nsteps,
                          int scheme, const double *limiter_weights) {
  size_t n = (size_t)ncell * nsp * 3;
  double *k = (double *)malloc(sizeof(double) * n), *ua = (double *)malloc(sizeof(double) * n),
         *ub = (double *)malloc(sizeof(double) * n);
  ctx1d c = {ncell, nsp, J, ll, lr, lpdm, dgl, dgr, gamma, bc};
  for (int s = 0; s < nsteps; ++s) {
    if (limiter_weights) fro_limiter_euler1d(u, ncell, nsp, gamma, limiter_weights, ll, lr);
    step_generic(u, n, dt, scheme, rhs1d_cb, &c, k, ua, ub);
  }
  free(k); free(ua); free(ub);
  return 0;
}

typedef struct {
  int ncell, nsp; const double *J, *ll, *lr, *lpdm, *dgl, *dgr; double a; int bc, variant;
} ctxadv;
static int rhsadv_cb(const double *u, double *du, void *c) {
  ctxadv *x = (ctxadv *)c;
  return fro_rhs_adv1d(u, du, x->ncell, x->nsp, x->J, x->ll, x->lr, x->lpdm, x->dgl, x->dgr, x->a, x->bc, x->variant);
}
int fro_integrate_adv1d(double *u, int ncell, int nsp, const double *J, const double *ll,
                        const double *lr, const double *lpdm, const double *dgl, const double *dgr,
                        double a, int bc, int variant, double dt, int nsteps, int scheme) {
  size_t n = (size_t)ncell * nsp;
  double *k = (double *)malloc(sizeof(double) * n), *ua = (double *)malloc(sizeof(double) * n),
         *ub = (double *)malloc(sizeof(double) * n);
  ctxadv c = {ncell, nsp, J, ll, lr, lpdm, dgl, dgr, a, bc, variant};
  for (int s = 0; s < nsteps; ++s) step_generic(u, n, dt, scheme, rhsadv_cb, &c, k, ua, ub);
  free(k); free(ua); free(ub);
  return 0;
}

typedef struct {
  int ncell, nu, nsp; const double *dx, *velo, *weights, *ll, *lr, *lpdm, *dgl, *dgr; double tau;
} ctxbgk;
static int rhsbgk_cb(const double *u, double *du, void *c) {
  ctxbgk *x = (ctxbgk *)c;
  return fro_rhs_bgk1d(u, du, x->ncell, x->nu, x->nsp, x->dx, x->velo, x->weights, x->ll, x->lr, x->lpdm, x->dgl, x->dgr, x->tau);
}
int fro_integrate_bgk1d(double *u, int ncell, int nu, int nsp, const double *dx, const double *velo,
                        const double *weights, const double *ll, const double *lr,
                        const double *lpdm, const double *dgl, const double *dgr, double tau,
                        double dt, int nsteps, int scheme) {
  size_t n = (size_t)ncell * nu * nsp;
  double *k = (double *)malloc(sizeof(double) * n), *ua = (double *)malloc(sizeof(double) * n),
         *ub = (double *)malloc(sizeof(double) * n);
  ctxbgk c = {ncell, nu, nsp, dx, velo, weights, ll, lr, lpdm, dgl, dgr, tau};
  for (int s = 0; s < nsteps; ++s) step_generic(u, n, dt, scheme, rhsbgk_cb, &c, k, ua, ub);
  free(k); free(ua); free(ub);
  return 0;
}
