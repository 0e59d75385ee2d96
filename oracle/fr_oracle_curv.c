/* fr_oracle_curv.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; never linked by the product) for the
 * curvilinear structured-quadrilateral Euler residual, SURVEY 8f-2: a C / OpenMP restatement of
 *   dudt! of /root/reference/dev/parallelogram.jl:80-165  (correction factors (iJ[i,j][k,l] n)[c], :145-148)
 *   dudt! of /root/reference/dev/cylinder2.jl:52-164      (factors from the flux-point Ji, :155-158; mirror
 *                                                          wall on x face 1, :100-120)
 * with the scripts' arithmetic per point (the loops are reordered so that the element index streams), for sizes the
 * NumPy restatement (fr_oracle_curv.py) is too slow for.
 * PARITY UNPINNED like the rest of oracle/ (no Julia in the image); tests/test_oracle_curv.py holds it against
 * the NumPy restatement (1e-14) and through it against the invariants listed there.
 *
 * Layouts as at the C ABI (include/frb200.h, frb_euler2d_curv_create): Julia column-major with one ghost ring,
 *   u, du [nx+2, ny+2, nsp, nsp, 4];  iJ [nx+2, ny+2, nsp, nsp, 2, 2];  n1 [nx+1, ny, 2];  n2 [nx, ny+1, 2];
 *   fpc [nx, ny, nsp, 4] or NULL;  lpdm row-major [m][k];  flags: 1 = y common flux by the row index l (the
 *   scripts' literal form), 2 = mirror wall on x face 1.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void conserve_prim4c(const double *W, double g, double *prim) {
  prim[0] = W[0];
  prim[1] = W[1] / W[0];
  prim[2] = W[2] / W[0];
  prim[3] = 0.5 * W[0] / (g - 1.0) / (W[3] - 0.5 * (W[1] * W[1] + W[2] * W[2]) / W[0]);
}
static void prim_conserve4c(const double *prim, double g, double *W) {
  W[0] = prim[0];
  W[1] = prim[0] * prim[1];
  W[2] = prim[0] * prim[2];
  W[3] = 0.5 * prim[0] / prim[3] / (g - 1.0) + 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2]);
}
static void euler_flux4c(const double *w, double g, double *F, double *G) {
  double prim[4];
  conserve_prim4c(w, g, prim);
  const double p = 0.5 * prim[0] / prim[3];
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
/* [KB] flux_hll!(fw, wL, wR, gamma, 1.0) */
static void flux_hll4c(double *fw, const double *wL, const double *wR, double g) {
  double pL[4], pR[4], f1[4], f2[4];
  conserve_prim4c(wL, g, pL);
  conserve_prim4c(wR, g, pR);
  const double aL = sqrt(0.5 * g / pL[3]), aR = sqrt(0.5 * g / pR[3]);
  const double lmin = pL[1] - aL, lmax = pR[1] + aR;
  euler_flux4c(wL, g, f1, 0);
  euler_flux4c(wR, g, f2, 0);
  if (lmin >= 0.0) {
    memcpy(fw, f1, sizeof f1);
  } else if (lmax <= 0.0) {
    memcpy(fw, f2, sizeof f2);
  } else {
    const double factor = 1.0 / (lmax - lmin);
    for (int m = 0; m < 4; ++m) fw[m] = factor * (lmax * f1[m] - lmin * f2[m] + (lmax * lmin) * (wR[m] - wL[m]));
  }
}
static void local_frame(const double *w, double c, double s, double *o) {
  o[0] = w[0];
  o[1] = w[1] * c + w[2] * s;
  o[2] = w[2] * c - w[1] * s;
  o[3] = w[3];
}
static void global_frame(const double *w, double c, double s, double *o) {
  o[0] = w[0];
  o[1] = w[1] * c - w[2] * s;
  o[2] = w[1] * s + w[2] * c;
  o[3] = w[3];
}

int fro_rhs_euler2d_curv(const double *u, double *du, int nx, int ny, int nsp, const double *iJ, const double *n1,
                         const double *n2, const double *fpc, int flags, double gamma, const double *ll,
                         const double *lr, const double *lpdm, const double *dgl, const double *dgr) {
  const size_t NXG = nx + 2, NYG = ny + 2, NE = NXG * NYG, NP = (size_t)nsp * nsp;
  const int fy_row = flags & 1, wall = (flags & 2) != 0;
  /* f[k, l, m, n], u_face[face, p, m], f_face[face, p, m] (component 1 on faces 2, 4; 2 on faces 1, 3), each per element */
  double *f = (double *)malloc(sizeof(double) * NE * NP * 4 * 2);
  double *uf = (double *)malloc(sizeof(double) * NE * 4 * nsp * 4);
  double *ff = (double *)malloc(sizeof(double) * NE * 4 * nsp * 4);
  double *fx = (double *)malloc(sizeof(double) * (size_t)(nx + 1) * ny * nsp * 4);
  double *fy = (double *)malloc(sizeof(double) * (size_t)nx * (ny + 1) * nsp * 4);
  if (!f || !uf || !ff || !fx || !fy) {
    free(f); free(uf); free(ff); free(fx); free(fy);
    return -1;
  }
#define U(e, k, l, m) u[(e) + NE * ((k) + nsp * ((l) + (size_t)nsp * (m)))]
#define IJ(e, k, l, a, b) iJ[(e) + NE * ((k) + nsp * ((l) + (size_t)nsp * ((a) + 2 * (b))))]
  /* work arrays are plane-major like the state (element index fastest), so that the inner loops stream */
#define F_(e, k, l, m, n) f[(e) + NE * ((k) + nsp * ((l) + (size_t)nsp * ((m) + 4 * (n))))]
#define UF(e, face, p, m) uf[(e) + NE * ((m) + 4 * ((p) + (size_t)nsp * (face)))]
#define FF(e, face, p, m) ff[(e) + NE * ((m) + 4 * ((p) + (size_t)nsp * (face)))]
  /* point fluxes, parallelogram.jl:88-96 (ghosts included) */
  for (int l = 0; l < nsp; ++l)
    for (int k = 0; k < nsp; ++k) {
#pragma omp parallel for schedule(static)
      for (long e = 0; e < (long)NE; ++e) {
        double w[4], F[4], G[4];
        for (int m = 0; m < 4; ++m) w[m] = U(e, k, l, m);
        euler_flux4c(w, gamma, F, G);
        for (int m = 0; m < 4; ++m) {
          F_(e, k, l, m, 0) = IJ(e, k, l, 0, 0) * F[m] + IJ(e, k, l, 0, 1) * G[m];
          F_(e, k, l, m, 1) = IJ(e, k, l, 1, 0) * F[m] + IJ(e, k, l, 1, 1) * G[m];
        }
      }
    }
  /* traces, :98-112: face 1 dot(u[i,j,p,:,m], ll), 2 dot(u[i,j,:,p,m], lr), 3 dot(u[i,j,p,:,m], lr), 4 dot(u[i,j,:,p,m], ll) */
  for (int p = 0; p < nsp; ++p)
    for (int m = 0; m < 4; ++m) {
#pragma omp parallel for schedule(static)
      for (long e = 0; e < (long)NE; ++e) {
        double a1 = 0, a2 = 0, a3 = 0, a4 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0;
        for (int q = 0; q < nsp; ++q) {
          a1 += U(e, p, q, m) * ll[q];
          a2 += U(e, q, p, m) * lr[q];
          a3 += U(e, p, q, m) * lr[q];
          a4 += U(e, q, p, m) * ll[q];
          b1 += F_(e, p, q, m, 1) * ll[q];
          b2 += F_(e, q, p, m, 0) * lr[q];
          b3 += F_(e, p, q, m, 1) * lr[q];
          b4 += F_(e, q, p, m, 0) * ll[q];
        }
        UF(e, 0, p, m) = a1; UF(e, 1, p, m) = a2; UF(e, 2, p, m) = a3; UF(e, 3, p, m) = a4;
        FF(e, 0, p, m) = b1; FF(e, 1, p, m) = b2; FF(e, 2, p, m) = b3; FF(e, 3, p, m) = b4;
      }
    }
  /* x faces i = 1..nx+1, :114-125 (cylinder2.jl:100-120 on face 1 when wall) */
#pragma omp parallel for schedule(static) collapse(2)
  for (int p = 0; p < nsp; ++p)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx + 1; ++i) {
        const size_t e = i + NXG * j, fi = (size_t)(i - 1) + (size_t)(nx + 1) * (j - 1);
        const double c = n1[fi], s = n1[fi + (size_t)(nx + 1) * ny];
        double tL[4], tR[4], uL[4], uR[4], fw[4], fg[4];
        for (int m = 0; m < 4; ++m) { tL[m] = UF(e - 1, 1, p, m); tR[m] = UF(e, 3, p, m); }
        local_frame(tL, c, s, uL);
        local_frame(tR, c, s, uR);
        if (wall && i == 1) {
          double prim[4], pn[4];
          conserve_prim4c(uR, gamma, prim);
          pn[1] = -prim[1];
          pn[2] = prim[2];
          pn[3] = 2.0 - prim[3];
          const double tmp = prim[3] - 1.0;
          pn[0] = (1 - tmp) / (1 + tmp) * prim[0];
          prim_conserve4c(pn, gamma, uL);
        }
        flux_hll4c(fw, uL, uR, gamma);
        global_frame(fw, c, s, fg);
        for (int m = 0; m < 4; ++m) fx[fi + (size_t)(nx + 1) * ny * (p + (size_t)nsp * m)] = fg[m];
      }
  /* y faces j = 1..ny+1, :126-136 */
#pragma omp parallel for schedule(static) collapse(2)
  for (int p = 0; p < nsp; ++p)
    for (int j = 1; j <= ny + 1; ++j)
      for (int i = 1; i <= nx; ++i) {
        const size_t e = i + NXG * j, fi = (size_t)(i - 1) + (size_t)nx * (j - 1);
        const double c = n2[fi], s = n2[fi + (size_t)nx * (ny + 1)];
        double tL[4], tR[4], uL[4], uR[4], fw[4], fg[4];
        for (int m = 0; m < 4; ++m) { tL[m] = UF(e - NXG, 2, p, m); tR[m] = UF(e, 0, p, m); }
        local_frame(tL, c, s, uL);
        local_frame(tR, c, s, uR);
        flux_hll4c(fw, uL, uR, gamma);
        global_frame(fw, c, s, fg);
        for (int m = 0; m < 4; ++m) fy[fi + (size_t)nx * (ny + 1) * (p + (size_t)nsp * m)] = fg[m];
      }
  /* derivative + correction, :138-163; du = 0 in the ghosts */
  memset(du, 0, sizeof(double) * NE * NP * 4);
  const size_t s1 = (size_t)(nx + 1) * ny, s2 = (size_t)nx * (ny + 1), sp = (size_t)nx * ny;
#pragma omp parallel for schedule(static) collapse(2)
  for (int m = 0; m < 4; ++m)
    for (int j = 1; j <= ny; ++j)
      for (int l = 0; l < nsp; ++l)
        for (int k = 0; k < nsp; ++k) {
          const int yi = fy_row ? l : k;
          for (int i = 1; i <= nx; ++i) {
            const size_t e = i + NXG * j;
            const size_t f1i = (size_t)(i - 1) + (size_t)(nx + 1) * (j - 1);
            const size_t f2i = (size_t)(i - 1) + (size_t)nx * (j - 1);
            double cxL, cxR, cyL, cyR;
            if (fpc) {
              cxL = fpc[f2i + sp * (l + (size_t)nsp * 0)];
              cxR = fpc[f2i + sp * (l + (size_t)nsp * 1)];
              cyL = fpc[f2i + sp * (k + (size_t)nsp * 2)];
              cyR = fpc[f2i + sp * (k + (size_t)nsp * 3)];
            } else {
              cxL = IJ(e, k, l, 0, 0) * n1[f1i] + IJ(e, k, l, 0, 1) * n1[f1i + s1];
              cxR = IJ(e, k, l, 0, 0) * n1[f1i + 1] + IJ(e, k, l, 0, 1) * n1[f1i + 1 + s1];
              cyL = IJ(e, k, l, 1, 0) * n2[f2i] + IJ(e, k, l, 1, 1) * n2[f2i + s2];
              cyR = IJ(e, k, l, 1, 0) * n2[f2i + nx] + IJ(e, k, l, 1, 1) * n2[f2i + nx + s2];
            }
            double rhs1 = 0, rhs2 = 0;
            for (int q = 0; q < nsp; ++q) {
              rhs1 += F_(e, q, l, m, 0) * lpdm[k * nsp + q];
              rhs2 += F_(e, k, q, m, 1) * lpdm[l * nsp + q];
            }
            const double fxL = cxL * fx[f1i + s1 * (l + (size_t)nsp * m)];
            const double fxR = cxR * fx[f1i + 1 + s1 * (l + (size_t)nsp * m)];
            const double fyL = cyL * fy[f2i + s2 * (yi + (size_t)nsp * m)];
            const double fyR = cyR * fy[f2i + nx + s2 * (yi + (size_t)nsp * m)];
            du[e + NE * (k + nsp * (l + (size_t)nsp * m))] =
                -(rhs1 + rhs2 + (fxL - FF(e, 3, l, m)) * dgl[k] + (fxR - FF(e, 1, l, m)) * dgr[k] +
                  (fyL - FF(e, 0, k, m)) * dgl[l] + (fyR - FF(e, 2, k, m)) * dgr[l]);
          }
        }
#undef U
#undef IJ
#undef F_
#undef UF
#undef FF
  free(f); free(uf); free(ff); free(fx); free(fy);
  return 0;
}
