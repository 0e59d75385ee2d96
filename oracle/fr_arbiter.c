/* Extended-precision ARBITER for the floating-point parity tests -- TEST INFRASTRUCTURE ONLY.
 *
 * The FP64 CUDA kernels and the FP64 C/NumPy oracle evaluate the same formulas in different
 * operation orders (FMA chains, folded operators, MUFU-seeded reciprocals on the GPU; the
 * reference's literal order in the oracle), so at BASELINE size they differ by rounding noise that
 * the conditioning of the residual amplifies (du is a small difference of terms ~200 x its size on the
 * smooth wave).  This file evaluates the residual of SELECTED cells in x87 `long double` (64-bit
 * mantissa, eps = 1.1e-19: ~2000 x finer than double) straight from the double inputs, cell-locally,
 * with no shared temporaries -- a third evaluation, independent in structure of both others -- so that
 * the tests can state  |gpu - exact| <= c * |oracle - exact|  instead of a hand-picked tolerance.
 *
 * Formulas (same citations as fr_oracle.c; [KB] = KitBase 0.9 closures restated from recall,
 * PARITY UNPINNED like the rest of the oracle):
 *   2-D Euler  dudt! of example/euler2d_wave.jl:35-107 (== example/shock-vortex.jl:26-118)
 *   1-D BGK    mol!  of example/bgk_wave.jl:69-129
 * Inputs stay double (state, operators, Jacobians: they ARE the problem data); every intermediate
 * is long double; the result is rounded to double once, at the end.
 */
#include <math.h>
#include <stddef.h>

typedef long double R;
#define ARB_MAXSP 8

/* [KB] euler_flux(w, gamma) -> F, G with p = (gamma-1)(E - 1/2 rho |v|^2) */
static void arb_flux4(const R *w, R g, R *F, R *G) {
  R p = (g - 1.0L) * (w[3] - 0.5L * (w[1] * w[1] + w[2] * w[2]) / w[0]);
  F[0] = w[1];
  F[1] = w[1] * w[1] / w[0] + p;
  F[2] = w[1] * w[2] / w[0];
  F[3] = (w[3] + p) * w[1] / w[0];
  if (G) {
    G[0] = w[2];
    G[1] = w[1] * w[2] / w[0];
    G[2] = w[2] * w[2] / w[0] + p;
    G[3] = (w[3] + p) * w[2] / w[0];
  }
}

/* [KB] flux_hll!(fw, wL, wR, gamma, 1.0): lambda = rho / (2 p), a = sqrt(gamma / (2 lambda)) */
static void arb_hll4(const R *wL, const R *wR, R g, R *fw) {
  R pL = (g - 1.0L) * (wL[3] - 0.5L * (wL[1] * wL[1] + wL[2] * wL[2]) / wL[0]);
  R pR = (g - 1.0L) * (wR[3] - 0.5L * (wR[1] * wR[1] + wR[2] * wR[2]) / wR[0]);
  R aL = sqrtl(g * pL / wL[0]), aR = sqrtl(g * pR / wR[0]);
  R lmin = wL[1] / wL[0] - aL, lmax = wR[1] / wR[0] + aR;
  R f1[4], f2[4];
  arb_flux4(wL, g, f1, 0);
  arb_flux4(wR, g, f2, 0);
  for (int m = 0; m < 4; ++m) {
    if (lmin >= 0.0L) fw[m] = f1[m];
    else if (lmax <= 0.0L) fw[m] = f2[m];
    else fw[m] = (lmax * f1[m] - lmin * f2[m] + lmax * lmin * (wR[m] - wL[m])) / (lmax - lmin);
  }
}

/* rotate into the frame of the y faces and back: local_frame(w, 0, 1), global_frame(f, 0, 1) */
static void arb_to_y(const R *w, R *o) { o[0] = w[0]; o[1] = w[2]; o[2] = -w[1]; o[3] = w[3]; }
static void arb_from_y(const R *f, R *o) { o[0] = f[0]; o[1] = -f[2]; o[2] = f[1]; o[3] = f[3]; }

/* u[i, j, k, l, m] (i fastest, one ghost ring), cells = ncells (i, j) pairs with 1 <= i <= nx,
 * 1 <= j <= ny; out[c][k + nsp (l + nsp m)] = du of that cell.  lpdm row-major [m][k]. */
int fra_rhs_euler2d_cells(const double *u, int nx, int ny, int nsp, double Jx, double Jy, const double *ll,
                          const double *lr, const double *lpdm, const double *dhl, const double *dhr,
                          double gamma, const int *cells, int ncells, double *out) {
  if (nsp > ARB_MAXSP) return -1;
  const size_t NXG = nx + 2, NE = NXG * (size_t)(ny + 2);
  const R g = gamma, iJx = 1.0L / (R)Jx, iJy = 1.0L / (R)Jy;
#define U5(i, j, k, l, m) ((R)u[(i) + NXG * (size_t)(j) + NE * ((k) + nsp * ((l) + nsp * (size_t)(m)))])
#pragma omp parallel for schedule(dynamic, 16)
  for (int c = 0; c < ncells; ++c) {
    const int i = cells[2 * c], j = cells[2 * c + 1];
    R F[ARB_MAXSP][ARB_MAXSP][4], G[ARB_MAXSP][ARB_MAXSP][4]; /* [k][l][m], already times inv(J) */
    for (int l = 0; l < nsp; ++l)
      for (int k = 0; k < nsp; ++k) {
        R w[4] = {U5(i, j, k, l, 0), U5(i, j, k, l, 1), U5(i, j, k, l, 2), U5(i, j, k, l, 3)};
        R f[4], gg[4];
        arb_flux4(w, g, f, gg);
        for (int m = 0; m < 4; ++m) { F[k][l][m] = iJx * f[m]; G[k][l][m] = iJy * gg[m]; }
      }
    /* common fluxes on the four faces of the cell, per flux point */
    R fxL[ARB_MAXSP][4], fxR[ARB_MAXSP][4], fyB[ARB_MAXSP][4], fyT[ARB_MAXSP][4];
    for (int q = 0; q < nsp; ++q) {
      R a[4], b[4], ra[4], rb[4], h[4];
      /* x face left of the cell: right trace of (i-1, j) row q | left trace of (i, j) row q */
      for (int m = 0; m < 4; ++m) {
        a[m] = 0; b[m] = 0;
        for (int k = 0; k < nsp; ++k) { a[m] += U5(i - 1, j, k, q, m) * (R)lr[k]; b[m] += U5(i, j, k, q, m) * (R)ll[k]; }
      }
      arb_hll4(a, b, g, fxL[q]);
      for (int m = 0; m < 4; ++m) {
        a[m] = 0; b[m] = 0;
        for (int k = 0; k < nsp; ++k) { a[m] += U5(i, j, k, q, m) * (R)lr[k]; b[m] += U5(i + 1, j, k, q, m) * (R)ll[k]; }
      }
      arb_hll4(a, b, g, fxR[q]);
      /* y face below the cell: top trace of (i, j-1) column q | bottom trace of (i, j) column q */
      for (int m = 0; m < 4; ++m) {
        a[m] = 0; b[m] = 0;
        for (int l = 0; l < nsp; ++l) { a[m] += U5(i, j - 1, q, l, m) * (R)lr[l]; b[m] += U5(i, j, q, l, m) * (R)ll[l]; }
      }
      arb_to_y(a, ra); arb_to_y(b, rb); arb_hll4(ra, rb, g, h); arb_from_y(h, fyB[q]);
      for (int m = 0; m < 4; ++m) {
        a[m] = 0; b[m] = 0;
        for (int l = 0; l < nsp; ++l) { a[m] += U5(i, j, q, l, m) * (R)lr[l]; b[m] += U5(i, j + 1, q, l, m) * (R)ll[l]; }
      }
      arb_to_y(a, ra); arb_to_y(b, rb); arb_hll4(ra, rb, g, h); arb_from_y(h, fyT[q]);
    }
    for (int m = 0; m < 4; ++m)
      for (int l = 0; l < nsp; ++l)
        for (int k = 0; k < nsp; ++k) {
          R r1 = 0, r2 = 0, fl = 0, fr = 0, gb = 0, gt = 0;
          for (int q = 0; q < nsp; ++q) {
            r1 += F[q][l][m] * (R)lpdm[k * nsp + q];
            r2 += G[k][q][m] * (R)lpdm[l * nsp + q];
            fl += F[q][l][m] * (R)ll[q];
            fr += F[q][l][m] * (R)lr[q];
            gb += G[k][q][m] * (R)ll[q];
            gt += G[k][q][m] * (R)lr[q];
          }
          R d = -(r1 + r2 + (fxL[l][m] * iJx - fl) * (R)dhl[k] + (fxR[l][m] * iJx - fr) * (R)dhr[k] +
                  (fyB[k][m] * iJy - gb) * (R)dhl[l] + (fyT[k][m] * iJy - gt) * (R)dhr[l]);
          out[(size_t)c * nsp * nsp * 4 + k + nsp * (l + nsp * m)] = (double)d;
        }
  }
#undef U5
  return 0;
}

/* u[cell, velocity, sp] (cell fastest), periodic; cells = ncells cell indices (0-based);
 * out[c][j + nu p] = du[cell, j, p].  gamma of the moments is 3 (bgk_wave.jl:79, one velocity dimension,
 * no internal degrees of freedom): lambda = rho / (4 (E - 1/2 rho U^2)). */
int fra_rhs_bgk1d_cells(const double *u, int ncell, int nu, int nsp, const double *dx, const double *velo,
                        const double *weights, const double *ll, const double *lr, const double *lpdm,
                        const double *dgl, const double *dgr, double tau, const int *cells, int ncells,
                        double *out) {
  if (nsp > ARB_MAXSP) return -1;
  const size_t cs = ncell, vs = cs * nu;
  const R pi = 3.141592653589793238462643383279502884L;
#define U3(i, j, k) ((R)u[(i) + cs * (size_t)(j) + vs * (size_t)(k)])
#pragma omp parallel for schedule(dynamic, 4)
  for (int c = 0; c < ncells; ++c) {
    const int i = cells[c];
    const int il = i == 0 ? ncell - 1 : i - 1, ir = i == ncell - 1 ? 0 : i + 1;
    const R iJ = 1.0L / (0.5L * (R)dx[i]), iJl = 1.0L / (0.5L * (R)dx[il]), iJr = 1.0L / (0.5L * (R)dx[ir]);
    R rho[ARB_MAXSP], U[ARB_MAXSP], lam[ARB_MAXSP];
    for (int k = 0; k < nsp; ++k) {
      R w0 = 0, w1 = 0, w2 = 0;
      for (int j = 0; j < nu; ++j) {
        R f = U3(i, j, k), v = velo[j], w = weights[j];
        w0 += w * f; w1 += w * v * f; w2 += w * v * v * f;
      }
      w2 *= 0.5L;
      rho[k] = w0; U[k] = w1 / w0;
      lam[k] = 0.5L * w0 / (3.0L - 1.0L) / (w2 - 0.5L * w1 * w1 / w0);
    }
    for (int j = 0; j < nu; ++j) {
      const R v = velo[j];
      R f[ARB_MAXSP], fL = 0, fR = 0, nbL = 0, nbR = 0;
      for (int k = 0; k < nsp; ++k) {
        f[k] = v * U3(i, j, k) * iJ;
        fL += f[k] * (R)ll[k];
        fR += f[k] * (R)lr[k];
        nbL += v * U3(il, j, k) * iJl * (R)lr[k]; /* right trace of the left neighbour */
        nbR += v * U3(ir, j, k) * iJr * (R)ll[k]; /* left trace of the right neighbour */
      }
      /* upwind (bgk_wave.jl:103-107): delta = heaviside(v) picks the left cell's right trace */
      const R hatL = v >= 0 ? nbL : fL, hatR = v >= 0 ? fR : nbR;
      for (int p = 0; p < nsp; ++p) {
        R r1 = 0;
        for (int q = 0; q < nsp; ++q) r1 += f[q] * (R)lpdm[p * nsp + q];
        R cc = v - U[p];
        R M = rho[p] * sqrtl(lam[p] / pi) * expl(-lam[p] * cc * cc);
        R d = -(r1 + (hatL - fL) * (R)dgl[p] + (hatR - fR) * (R)dgr[p]) + (M - U3(i, j, p)) / (R)tau;
        out[(size_t)c * nu * nsp + j + (size_t)nu * p] = (double)d;
      }
    }
  }
#undef U3
  return 0;
}
