"""NumPy restatement of the FluxReconstruction.jl hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: a CPU restatement of the reference's semi-discrete
flux-reconstruction residual and explicit time step.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``fluxreconstruction.jl_b200``) never does.

PARITY UNPINNED: Julia and KitBase.jl (compat 0.9, Project.toml:9,18; no Manifest)
are absent from this image, and the reference's own tests never evaluate an RHS,
a flux or a time step (test/runtests.jl:1-43).  The physics closures that live in
KitBase are restated from its published formulas (marked [KB]); what pins this
oracle is (i) the reference's own two-way operator check
(example/vandermonde_lagrange.jl:14-15,25), (ii) analytic invariants
(free-stream preservation, hll(w,w)=F(w), conservation, design-order convergence,
travelling waves returning to the IC) and (iii) bit-level agreement with the
independent C restatement in ``oracle/fr_oracle.c``.

All citations are relative to /root/reference.  Arrays keep the reference's Julia
index order as NumPy axes (0-based, ghost cells included explicitly) and are
allocated Fortran-ordered so that the memory image equals Julia's column-major one.
"""
from __future__ import annotations

import math
import numpy as np
from numpy.polynomial import legendre as _npleg

# ----------------------------------------------------------------------------
# L1: polynomial operators  (src/Polynomial/*.jl)
# ----------------------------------------------------------------------------


def legendre_point(p: int) -> np.ndarray:
    """Gauss-Legendre nodes, poly_legendre.jl:6 (gausslegendre(p+1)[1])."""
    return _npleg.leggauss(p + 1)[0].astype(np.float64)


def gausslegendre(n: int):
    """FastGaussQuadrature.gausslegendre(n): nodes, weights (struct.jl:47)."""
    x, w = _npleg.leggauss(n)
    return x.astype(np.float64), w.astype(np.float64)


def lagrange_point(sp: np.ndarray, x: float) -> np.ndarray:
    """l_k(x) for the nodal basis on ``sp``; poly_lagrange.jl:6-21."""
    nsp = len(sp)
    l = np.empty(nsp)
    for k in range(nsp):
        tmp = 1.0
        for j in range(nsp):
            if j != k:
                tmp *= (x - sp[j]) / (sp[k] - sp[j])
        l[k] = tmp
    return l


def dlagrange(sp: np.ndarray) -> np.ndarray:
    """lpdm[m,k] = l_k'(sp[m]); poly_lagrange.jl:38-59."""
    nsp = len(sp)
    lpdm = np.empty((nsp, nsp))
    for k in range(nsp):
        for m in range(nsp):
            lsum = 0.0
            for l in range(nsp):
                tmp = 1.0
                for j in range(nsp):
                    if j != k and j != l:
                        tmp *= (sp[m] - sp[j]) / (sp[k] - sp[j])
                if l != k:
                    lsum += tmp / (sp[k] - sp[l])
            lpdm[m, k] = lsum
    return lpdm


def standard_lagrange(x):
    """poly_lagrange.jl:99-105."""
    return lagrange_point(x, -1.0), lagrange_point(x, 1.0), dlagrange(x)


def dlegendre(p: int, x) -> np.ndarray:
    """P_p'(x) (GSL sf_legendre_Pl_deriv_array in poly_legendre.jl:13-22).

    Evaluated with the three-term recurrence GSL uses: P'_{l+1} = ((2l+1)(P_l + x P'_l)
    ... ) expressed through P_l; values agree with GSL to a few ulp.
    """
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if p < 0:
        return np.zeros_like(x)
    P0 = np.ones_like(x)
    d0 = np.zeros_like(x)
    if p == 0:
        return d0
    P1 = x.copy()
    d1 = np.ones_like(x)
    for l in range(1, p):
        P2 = ((2 * l + 1) * x * P1 - l * P0) / (l + 1)
        d2 = d0 + (2 * l + 1) * P1  # P'_{l+1} = P'_{l-1} + (2l+1) P_l
        P0, P1 = P1, P2
        d0, d1 = d1, d2
    return d1


def dradau(p: int, x):
    """poly_legendre.jl:29-37."""
    d = dlegendre(p, x)
    dp = dlegendre(p + 1, x)
    dgl = (-1.0) ** p * 0.5 * (d - dp)
    dgr = 0.5 * (d + dp)
    return dgl, dgr


def dsd(p: int, x):
    """poly_legendre.jl:44-54."""
    dm = dlegendre(p - 1, x)
    d = dlegendre(p, x)
    dp = dlegendre(p + 1, x)
    y = (p * dm + (p + 1) * dp) / (2 * p + 1)
    return (-1.0) ** p * 0.5 * (d - y), 0.5 * (d + y)


def dhuynh(p: int, x):
    """poly_legendre.jl:61-71."""
    dm = dlegendre(p - 1, x)
    d = dlegendre(p, x)
    dp = dlegendre(p + 1, x)
    y = ((p + 1) * dm + p * dp) / (2 * p + 1)
    return (-1.0) ** p * 0.5 * (d - y), 0.5 * (d + y)


_CORRECTION = {"radau": dradau, "sd": dsd, "huynh": dhuynh}


def jacobi_p(x, alpha, beta, N):
    """Orthonormal Jacobi polynomial, poly_jacobi.jl:47-86."""
    xp = np.atleast_1d(np.asarray(x, dtype=np.float64))
    PL = np.zeros((N + 1, len(xp)))
    g = math.gamma
    gamma0 = 2.0 ** (alpha + beta + 1) / (alpha + beta + 1) * g(alpha + 1) * g(beta + 1) / g(alpha + beta + 1)
    PL[0, :] = 1.0 / math.sqrt(gamma0)
    if N == 0:
        return PL[0].copy()
    gamma1 = (alpha + 1) * (beta + 1) / (alpha + beta + 3) * gamma0
    PL[1, :] = ((alpha + beta + 2) * xp / 2 + (alpha - beta) / 2) / math.sqrt(gamma1)
    if N == 1:
        return PL[1].copy()
    aold = 2 / (2 + alpha + beta) * math.sqrt((alpha + 1) * (beta + 1) / (alpha + beta + 3))
    for i in range(1, N):
        h1 = 2 * i + alpha + beta
        anew = 2 / (h1 + 2) * math.sqrt(
            (i + 1) * (i + 1 + alpha + beta) * (i + 1 + alpha) * (i + 1 + beta) / (h1 + 1) / (h1 + 3)
        )
        bnew = -(alpha**2 - beta**2) / h1 / (h1 + 2)
        PL[i + 1, :] = 1 / anew * (-aold * PL[i - 1, :] + (xp - bnew) * PL[i, :])
        aold = anew
    return PL[N].copy()


def djacobi_p(r, alpha, beta, N):
    """poly_jacobi.jl:100-107."""
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    if N == 0:
        return np.zeros_like(r)
    return math.sqrt(N * (N + alpha + beta + 1)) * jacobi_p(r, alpha + 1, beta + 1, N - 1)


def vandermonde_matrix(N, r):
    """transform.jl:18-26."""
    r = np.asarray(r, dtype=np.float64)
    V = np.zeros((len(r), N + 1))
    for j in range(N + 1):
        V[:, j] = jacobi_p(r, 0, 0, j)
    return V


def dvandermonde_matrix(N, r):
    """transform.jl:76-84."""
    r = np.asarray(r, dtype=np.float64)
    Vr = np.zeros((len(r), N + 1))
    for i in range(N + 1):
        Vr[:, i] = djacobi_p(r, 0, 0, i)
    return Vr


# ----------------------------------------------------------------------------
# L1: spaces  (src/struct.jl, src/Geometry/*)
# ----------------------------------------------------------------------------


def r_x(r, vl, vr):
    """geo_transform.jl:7."""
    return ((1.0 - r) / 2.0) * vl + ((1.0 + r) / 2.0) * vr


class FRPSpace1D:
    """struct.jl:13-88.  Arrays over cells carry ``ng`` ghost cells on each side;
    index ``i`` of the reference (1-based, ghosts at <=0) is ``i - 1 + ng`` here."""

    def __init__(self, x0, x1, nx, deg, ng=0, correction="radau"):
        self.x0, self.x1, self.nx, self.deg, self.ng = float(x0), float(x1), int(nx), int(deg), int(ng)
        dx = (self.x1 - self.x0) / nx
        ntot = nx + 2 * ng
        idx = np.arange(1 - ng, nx + ng + 1)
        self.x = self.x0 + (idx - 0.5) * dx  # [KB] PSpace1D cell centres
        self.dx = np.full(ntot, dx)
        self.J = self.dx / 2  # struct.jl:42
        self.np = deg + 1
        r = legendre_point(deg)
        self.xpl = r
        xi = np.append(self.x - 0.5 * self.dx, self.x[-1] + 0.5 * self.dx[-1])  # struct.jl:45
        self.xpg = np.empty((ntot, deg + 1), order="F")
        for j in range(deg + 1):
            self.xpg[:, j] = r_x(r[j], xi[:-1], xi[1:])  # geo_points.jl:8-15
        self.wp = gausslegendre(deg + 1)[1]
        self.ll = lagrange_point(r, -1.0)
        self.lr = lagrange_point(r, 1.0)
        self.dl = dlagrange(r)
        V = vandermonde_matrix(deg, r)
        self.V = V
        self.iV = np.linalg.inv(V)
        dVf = dvandermonde_matrix(deg, np.array([-1.0, 1.0]))
        dlf = np.zeros((2, deg + 1))
        for i in range(2):
            dlf[i, :] = np.linalg.solve(V.T, dVf[i, :])  # struct.jl:55-59
        self.dll, self.dlr = dlf[0].copy(), dlf[1].copy()
        self.dhl, self.dhr = _CORRECTION[correction](deg, r)


class FRPSpace2D:
    """struct.jl:99-245 for a uniform rectangular PSpace2D.  ``J[i,j][k,l]`` is
    diag(dx/2, dy/2) on such a mesh (geo_jacobi.jl:77-108); it is kept as the two
    scalars Jx, Jy.  Axis 0/1 of cell arrays include ``ngx``/``ngy`` ghosts."""

    def __init__(self, x0, x1, nx, y0, y1, ny, deg, ngx=0, ngy=0):
        self.x0, self.x1, self.nx = float(x0), float(x1), int(nx)
        self.y0, self.y1, self.ny = float(y0), float(y1), int(ny)
        self.deg, self.ngx, self.ngy = int(deg), int(ngx), int(ngy)
        dx = (self.x1 - self.x0) / nx
        dy = (self.y1 - self.y0) / ny
        self.dx, self.dy = dx, dy
        self.Jx, self.Jy = dx / 2, dy / 2
        nsp = deg + 1
        self.np = nsp * nsp
        r = legendre_point(deg)
        self.xpl = r
        ii = np.arange(1 - ngx, nx + ngx + 1)
        jj = np.arange(1 - ngy, ny + ngy + 1)
        xc = self.x0 + (ii - 0.5) * dx
        yc = self.y0 + (jj - 0.5) * dy
        self.x = np.repeat(xc[:, None], len(jj), axis=1)
        self.y = np.repeat(yc[None, :], len(ii), axis=0)
        # vertices CCW from bottom-left [KB]; bilinear map of struct.jl:169-175
        xl, xr = xc - 0.5 * dx, xc + 0.5 * dx
        yl, yr = yc - 0.5 * dy, yc + 0.5 * dy
        self.xpg = np.empty((len(ii), len(jj), nsp, nsp, 2), order="F")
        for k in range(nsp):
            for l in range(nsp):
                a1 = (r[k] - 1.0) * (r[l] - 1.0) / 4
                a2 = (r[k] + 1.0) * (1.0 - r[l]) / 4
                a3 = (r[k] + 1.0) * (r[l] + 1.0) / 4
                a4 = (1.0 - r[k]) * (r[l] + 1.0) / 4
                self.xpg[:, :, k, l, 0] = (a1 * xl + a2 * xr + a3 * xr + a4 * xl)[:, None]
                self.xpg[:, :, k, l, 1] = (a1 * yl + a2 * yl + a3 * yr + a4 * yr)[None, :]
        w = gausslegendre(nsp)[1]
        self.wp = np.outer(w, w)
        self.ll, self.lr, self.dl = standard_lagrange(r)
        V = vandermonde_matrix(deg, r)
        dVf = dvandermonde_matrix(deg, np.array([-1.0, 1.0]))
        dlf = np.zeros((2, nsp))
        for i in range(2):
            dlf[i, :] = np.linalg.solve(V.T, dVf[i, :])
        self.dll, self.dlr = dlf[0].copy(), dlf[1].copy()
        self.dhl, self.dhr = dradau(deg, r)  # struct.jl:193 (radau only)


# ----------------------------------------------------------------------------
# L0: KitBase physics [KB]  (vectorised over leading axes; last axis = variable)
# ----------------------------------------------------------------------------


def prim_conserve(prim, gamma):
    prim = np.asarray(prim, dtype=np.float64)
    W = np.empty_like(prim)
    if prim.shape[-1] == 3:
        W[..., 0] = prim[..., 0]
        W[..., 1] = prim[..., 0] * prim[..., 1]
        W[..., 2] = 0.5 * prim[..., 0] / prim[..., 2] / (gamma - 1.0) + 0.5 * prim[..., 0] * prim[..., 1] ** 2
    else:
        W[..., 0] = prim[..., 0]
        W[..., 1] = prim[..., 0] * prim[..., 1]
        W[..., 2] = prim[..., 0] * prim[..., 2]
        W[..., 3] = 0.5 * prim[..., 0] / prim[..., 3] / (gamma - 1.0) + 0.5 * prim[..., 0] * (
            prim[..., 1] ** 2 + prim[..., 2] ** 2
        )
    return W


def conserve_prim(W, gamma):
    W = np.asarray(W, dtype=np.float64)
    prim = np.empty_like(W)
    if W.shape[-1] == 3:
        prim[..., 0] = W[..., 0]
        prim[..., 1] = W[..., 1] / W[..., 0]
        prim[..., 2] = 0.5 * W[..., 0] / (gamma - 1.0) / (W[..., 2] - 0.5 * W[..., 1] ** 2 / W[..., 0])
    else:
        prim[..., 0] = W[..., 0]
        prim[..., 1] = W[..., 1] / W[..., 0]
        prim[..., 2] = W[..., 2] / W[..., 0]
        prim[..., 3] = (
            0.5 * W[..., 0] / (gamma - 1.0) / (W[..., 3] - 0.5 * (W[..., 1] ** 2 + W[..., 2] ** 2) / W[..., 0])
        )
    return prim


def sound_speed(prim, gamma):
    return np.sqrt(0.5 * gamma / prim[..., -1])


def euler_flux(w, gamma):
    """[KB] euler_flux(w, γ) -> (F,) in 1-D, (F, G) in 2-D."""
    w = np.asarray(w, dtype=np.float64)
    prim = conserve_prim(w, gamma)
    p = 0.5 * prim[..., 0] / prim[..., -1]
    F = np.empty_like(w)
    if w.shape[-1] == 3:
        F[..., 0] = w[..., 1]
        F[..., 1] = w[..., 1] ** 2 / w[..., 0] + p
        F[..., 2] = (w[..., 2] + p) * w[..., 1] / w[..., 0]
        return (F,)
    G = np.empty_like(w)
    F[..., 0] = w[..., 1]
    F[..., 1] = w[..., 1] ** 2 / w[..., 0] + p
    F[..., 2] = w[..., 1] * w[..., 2] / w[..., 0]
    F[..., 3] = (w[..., 3] + p) * w[..., 1] / w[..., 0]
    G[..., 0] = w[..., 2]
    G[..., 1] = w[..., 2] * w[..., 1] / w[..., 0]
    G[..., 2] = w[..., 2] ** 2 / w[..., 0] + p
    G[..., 3] = (w[..., 3] + p) * w[..., 2] / w[..., 0]
    return F, G


def flux_hll(wL, wR, gamma, dt=1.0, length=1.0):
    """[KB] flux_hll!(fw, wL, wR, γ, dt, len)."""
    wL = np.asarray(wL, dtype=np.float64)
    wR = np.asarray(wR, dtype=np.float64)
    primL = conserve_prim(wL, gamma)
    primR = conserve_prim(wR, gamma)
    aL = sound_speed(primL, gamma)
    aR = sound_speed(primR, gamma)
    lmin = primL[..., 1] - aL
    lmax = primR[..., 1] + aR
    f1 = euler_flux(wL, gamma)[0]
    f2 = euler_flux(wR, gamma)[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        factor = 1.0 / (lmax - lmin)
        mid = factor[..., None] * (
            lmax[..., None] * f1 - lmin[..., None] * f2 + (lmax * lmin)[..., None] * (wR - wL)
        )
    fw = np.where((lmin >= 0.0)[..., None], f1, np.where((lmax <= 0.0)[..., None], f2, mid))
    return fw * (dt * length)


def flux_lf(wL, wR, gamma, dt=1.0, length=1.0):
    """Local Lax-Friedrichs (Rusanov) flux normal to the face (component 1 = normal momentum):
    0.5 (F_L + F_R) - 0.5 alpha (w_R - w_L), alpha = max(|u_L| + a_L, |u_R| + a_R).
    Oracle-defined: the reference only sketches an LF-like flux (dev/cylinder.jl:86-88, with the
    time step as dissipation coefficient); SURVEY 0.1 makes LF/Roe extras without a reference."""
    wL = np.asarray(wL, dtype=np.float64)
    wR = np.asarray(wR, dtype=np.float64)
    primL, primR = conserve_prim(wL, gamma), conserve_prim(wR, gamma)
    alpha = np.maximum(np.abs(primL[..., 1]) + sound_speed(primL, gamma),
                       np.abs(primR[..., 1]) + sound_speed(primR, gamma))
    f1, f2 = euler_flux(wL, gamma)[0], euler_flux(wR, gamma)[0]
    return (0.5 * (f1 + f2) - 0.5 * alpha[..., None] * (wR - wL)) * (dt * length)


def flux_roe(wL, wR, gamma, dt=1.0, length=1.0):
    """Roe flux normal to the face with Harten's entropy fix on the acoustic waves
    (|lam| -> (lam^2 + d^2) / (2 d) below d = 0.1 a~).  3 (1-D) or 4 (2-D) components.
    Oracle-defined (the reference names flux_roe! only in dev/gks.jl:176-177)."""
    wL = np.asarray(wL, dtype=np.float64)
    wR = np.asarray(wR, dtype=np.float64)
    nv = wL.shape[-1]
    gm1 = gamma - 1.0

    def prim(w):
        rho = w[..., 0]
        u = w[..., 1] / rho
        v = w[..., 2] / rho if nv == 4 else np.zeros_like(rho)
        E = w[..., -1]
        p = gm1 * (E - 0.5 * rho * (u * u + v * v))
        return rho, u, v, p, (E + p) / rho

    rL, uL, vL, pL, HL = prim(wL)
    rR, uR, vR, pR, HR = prim(wR)
    R = np.sqrt(rR / rL)
    ut = (uL + R * uR) / (1.0 + R)
    vt = (vL + R * vR) / (1.0 + R)
    Ht = (HL + R * HR) / (1.0 + R)
    q2 = ut * ut + vt * vt
    a2 = gm1 * (Ht - 0.5 * q2)
    at = np.sqrt(a2)
    rt = R * rL
    dr, du, dv, dp = rR - rL, uR - uL, vR - vL, pR - pL
    a1 = (dp - rt * at * du) / (2.0 * a2)
    a2w = dr - dp / a2
    a3 = rt * dv
    a4 = (dp + rt * at * du) / (2.0 * a2)
    d = 0.1 * at

    def fix(lam):
        al = np.abs(lam)
        return np.where(al < d, (lam * lam + d * d) / (2.0 * d), al)

    l1, l2, l4 = fix(ut - at), np.abs(ut), fix(ut + at)
    one, zero = np.ones_like(ut), np.zeros_like(ut)
    if nv == 4:
        K1 = np.stack([one, ut - at, vt, Ht - ut * at], axis=-1)
        K2 = np.stack([one, ut, vt, 0.5 * q2], axis=-1)
        K3 = np.stack([zero, zero, one, vt], axis=-1)
        K4 = np.stack([one, ut + at, vt, Ht + ut * at], axis=-1)
        diss = ((l1 * a1)[..., None] * K1 + (l2 * a2w)[..., None] * K2 + (l2 * a3)[..., None] * K3
                + (l4 * a4)[..., None] * K4)
    else:
        K1 = np.stack([one, ut - at, Ht - ut * at], axis=-1)
        K2 = np.stack([one, ut, 0.5 * q2], axis=-1)
        K4 = np.stack([one, ut + at, Ht + ut * at], axis=-1)
        diss = (l1 * a1)[..., None] * K1 + (l2 * a2w)[..., None] * K2 + (l4 * a4)[..., None] * K4
    f1, f2 = euler_flux(wL, gamma)[0], euler_flux(wR, gamma)[0]
    return (0.5 * (f1 + f2) - 0.5 * diss) * (dt * length)


RIEMANN = {"hll": flux_hll, "lf": flux_lf, "roe": flux_roe}


def local_frame(w, c, s):
    w = np.asarray(w, dtype=np.float64)
    out = np.empty_like(w)
    out[..., 0] = w[..., 0]
    out[..., 1] = w[..., 1] * c + w[..., 2] * s
    out[..., 2] = w[..., 2] * c - w[..., 1] * s
    out[..., 3] = w[..., 3]
    return out


def global_frame(w, c, s):
    w = np.asarray(w, dtype=np.float64)
    out = np.empty_like(w)
    out[..., 0] = w[..., 0]
    out[..., 1] = w[..., 1] * c - w[..., 2] * s
    out[..., 2] = w[..., 1] * s + w[..., 2] * c
    out[..., 3] = w[..., 3]
    return out


def vspace1d(u0, u1, nu):
    """[KB] VSpace1D(u0,u1,nu): midpoint nodes and uniform weights."""
    du = (u1 - u0) / nu
    u = u0 + (np.arange(1, nu + 1) - 0.5) * du
    return u.astype(np.float64), np.full(nu, du)


def maxwellian(v, prim):
    """[KB] maxwellian(u, prim) 1-D1V: ρ sqrt(λ/π) exp(-λ (u-U)^2); prim[..., 3]."""
    rho, U, lam = prim[..., 0:1], prim[..., 1:2], prim[..., 2:3]
    return rho * np.sqrt(lam / np.pi) * np.exp(-lam * (v - U) ** 2)


def moments_conserve_1v(f, v, w):
    """[KB] moments_conserve(f,u,ω) = [Σωf, Σωuf, ½Σωu²f]; f[..., nu]."""
    out = np.empty(f.shape[:-1] + (3,))
    out[..., 0] = np.sum(w * f, axis=-1)
    out[..., 1] = np.sum(v * w * f, axis=-1)
    out[..., 2] = 0.5 * np.sum(v**2 * w * f, axis=-1)
    return out


def heaviside(x):
    return (np.asarray(x) >= 0).astype(np.float64)


# ----------------------------------------------------------------------------
# L3: RHS restatements
# ----------------------------------------------------------------------------


def _dot_last(a, l):
    """sequential dot over the last axis (Julia generic dot order)."""
    acc = a[..., 0] * l[0]
    for q in range(1, len(l)):
        acc = acc + a[..., q] * l[q]
    return acc


def rhs_advection1d(u, ps: FRPSpace1D, a, bc="period", variant="packaged"):
    """1-D advection RHS.

    variant="packaged": src/Equation/eq_advection.jl:55-77 + eq_scalar.jl:1-12 +
        period_advection! :158-175 (seam epsilon 1e-6) / dirichlet_advection! :153-156.
    variant="lowlevel": example/advection_lowlevel.jl:4-47 (periodic, seam epsilon 1e-8).
    u[ncell, nsp]  (no ghosts).
    """
    ncell, nsp = u.shape
    J = ps.J[ps.ng : ps.ng + ncell]
    f = a * u / J[:, None]  # advection_dflux! :103-111
    u_face = np.stack([_dot_last(u, ps.ll), _dot_last(u, ps.lr)], axis=1)
    f_face = np.stack([_dot_last(f, ps.ll), _dot_last(f, ps.lr)], axis=1)
    fi = np.zeros(ncell + 1)
    # faces 2:ncell (0-based 1..ncell-1): eq_advection.jl:128-137
    au = (f_face[1:, 0] - f_face[:-1, 1]) / (u_face[1:, 0] - u_face[:-1, 1] + 1e-8)
    fi[1:ncell] = 0.5 * (f_face[1:, 0] + f_face[:-1, 1]) - 0.5 * np.abs(au) * (u_face[1:, 0] - u_face[:-1, 1])
    rhs1 = np.zeros_like(u)
    for p in range(nsp):
        rhs1[:, p] = _dot_last(f, ps.dl[p, :])  # f * lpdm'
    du = np.zeros_like(u)
    periodic = bc == "period" or variant == "lowlevel"
    if periodic:
        eps = 1e-8 if variant == "lowlevel" else 1e-6
        au0 = (f_face[0, 0] - f_face[-1, 1]) / (u_face[0, 0] - u_face[-1, 1] + eps)
        fi[0] = 0.5 * (f_face[-1, 1] + f_face[0, 0]) - 0.5 * abs(au0) * (u_face[0, 0] - u_face[-1, 1])
        fi[-1] = fi[0]
        cells = slice(0, ncell)
    else:
        cells = slice(1, ncell - 1)
    for p in range(nsp):
        du[cells, p] = -(
            rhs1[cells, p]
            + (fi[:-1][cells] - f_face[cells, 0]) * ps.dhl[p]
            + (fi[1:][cells] - f_face[cells, 1]) * ps.dhr[p]
        )
    return du


def rhs_euler1d(u, ps: FRPSpace1D, gamma, bc="dirichlet", flux="hll"):
    """src/Equation/eq_euler.jl:29-98.  u[ncell, nsp, 3].  flux: the common flux (reference: hll)."""
    flux_hll = RIEMANN[flux]  # noqa: F811 (the reference's call sites keep their name)
    ncell, nsp, _ = u.shape
    J = ps.J[ps.ng : ps.ng + ncell]
    f = euler_flux(u, gamma)[0] / J[:, None, None]  # :35-39
    # traces :41-49   (u[:, :, j] * ll)
    uL = np.stack([_dot_last(u[:, :, j], ps.ll) for j in range(3)], axis=-1)
    uR = np.stack([_dot_last(u[:, :, j], ps.lr) for j in range(3)], axis=-1)
    fL = np.stack([_dot_last(f[:, :, j], ps.ll) for j in range(3)], axis=-1)
    fR = np.stack([_dot_last(f[:, :, j], ps.lr) for j in range(3)], axis=-1)
    fi = np.zeros((ncell + 1, 3))
    fi[1:ncell] = flux_hll(uR[:-1], uL[1:], gamma, 1.0)  # :51-54
    rhs1 = np.zeros_like(u)
    for k in range(3):
        for p in range(nsp):
            rhs1[:, p, k] = _dot_last(f[:, :, k], ps.dl[p, :])  # :56-58
    du = np.zeros_like(u)
    if bc == "period":
        fi[0] = flux_hll(uR[-1], uL[0], gamma, 1.0)  # :86-89
        fi[-1] = fi[0]
        cells = slice(0, ncell)
    else:
        cells = slice(1, ncell - 1)
    for p in range(nsp):
        for k in range(3):
            du[cells, p, k] = -(
                rhs1[cells, p, k]
                + (fi[:-1, k][cells] / J[cells] - fL[cells, k]) * ps.dhl[p]
                + (fi[1:, k][cells] / J[cells] - fR[cells, k]) * ps.dhr[p]
            )
    return du


def rhs_euler2d(u, ps: FRPSpace2D, gamma, flux="hll"):
    """example/euler2d_wave.jl:35-107.  flux: the common flux (reference: hll).  u[nx+2, ny+2, nsp, nsp, 4] with one ghost
    ring (axis index = reference index).  du is zero in ghosts (:36)."""
    flux_hll = RIEMANN[flux]  # noqa: F811
    nxg, nyg, nsp, _, _ = u.shape
    nx, ny = nxg - 2, nyg - 2
    ll, lr, dhl, dhr, lpdm = ps.ll, ps.lr, ps.dhl, ps.dhr, ps.dl
    iJx, iJy = 1.0 / ps.Jx, 1.0 / ps.Jy
    F, G = euler_flux(u, gamma)  # :45-50, ghosts included
    f1 = F * iJx
    f2 = G * iJy
    # traces :54-66.  u[i,j,k,l,m]: k <-> r(x), l <-> s(y)
    # face 1 (s=-1): dot(u[i,j,l,:,m], ll) -> point index = first tensor index
    # face 2 (r=+1): dot(u[i,j,:,l,m], lr) -> point index = second tensor index
    def tr_s(a, l):  # contract second tensor index (axis 3)
        acc = a[:, :, :, 0, :] * l[0]
        for q in range(1, nsp):
            acc = acc + a[:, :, :, q, :] * l[q]
        return acc

    def tr_r(a, l):  # contract first tensor index (axis 2)
        acc = a[:, :, 0, :, :] * l[0]
        for q in range(1, nsp):
            acc = acc + a[:, :, q, :, :] * l[q]
        return acc

    u1, u2, u3, u4 = tr_s(u, ll), tr_r(u, lr), tr_s(u, lr), tr_r(u, ll)
    f_face1_2 = tr_s(f2, ll)  # f_face[.,.,1,k,m,2]
    f_face2_1 = tr_r(f1, lr)
    f_face3_2 = tr_s(f2, lr)
    f_face4_1 = tr_r(f1, ll)
    # x faces :68-74  (i in 1:nx+1, j in 1:ny)
    fx = flux_hll(u2[0 : nx + 1, 1 : ny + 1], u4[1 : nx + 2, 1 : ny + 1], gamma, 1.0)
    # y faces :75-82
    uLy = local_frame(u3[1 : nx + 1, 0 : ny + 1], 0.0, 1.0)
    uRy = local_frame(u1[1 : nx + 1, 1 : ny + 2], 0.0, 1.0)
    fy = global_frame(flux_hll(uLy, uRy, gamma, 1.0), 0.0, 1.0)
    du = np.zeros_like(u)
    I = slice(1, nx + 1)
    Jn = slice(1, ny + 1)
    for k in range(nsp):
        for l in range(nsp):
            rhs1 = f1[I, Jn, 0, l, :] * lpdm[k, 0]
            rhs2 = f2[I, Jn, k, 0, :] * lpdm[l, 0]
            for q in range(1, nsp):
                rhs1 = rhs1 + f1[I, Jn, q, l, :] * lpdm[k, q]
                rhs2 = rhs2 + f2[I, Jn, k, q, :] * lpdm[l, q]
            du[I, Jn, k, l, :] = -(
                rhs1
                + rhs2
                + (fx[0:nx, :, l, :] * iJx - f_face4_1[I, Jn, l, :]) * dhl[k]
                + (fx[1 : nx + 1, :, l, :] * iJx - f_face2_1[I, Jn, l, :]) * dhr[k]
                + (fy[:, 0:ny, k, :] * iJy - f_face1_2[I, Jn, k, :]) * dhl[l]
                + (fy[:, 1 : ny + 1, k, :] * iJy - f_face3_2[I, Jn, k, :]) * dhr[l]
            )
    return du


def ghost_fill_euler2d(u, mode="wave_x"):
    """Per-step ghost fill of example/euler2d_wave.jl:127-132 (mode 'wave_x') and
    :159-164 ('wave_y'); 'copy' is shock-vortex.jl:324-326 (zero-gradient)."""
    nx, ny = u.shape[0] - 2, u.shape[1] - 2
    if mode == "wave_x":
        u[0] = u[nx]
        u[nx + 1] = u[1]
        u[:, 0] = u[:, ny]
        u[:, 0, :, :, 2] *= -1
        u[:, ny + 1] = u[:, 1]
        u[:, ny + 1, :, :, 2] *= -1
    elif mode == "wave_y":
        u[:, 0] = u[:, ny]
        u[:, ny + 1] = u[:, 1]
        u[0] = u[nx]
        u[0, :, :, :, 1] *= -1
        u[nx + 1] = u[1]
        u[nx + 1, :, :, :, 1] *= -1
    elif mode == "copy":
        u[:, 0] = u[:, 1]
        u[:, ny + 1] = u[:, ny]
        u[nx + 1] = u[nx]
    else:
        raise ValueError(mode)
    return u


def rhs_bgk1d(u, dx, velo, weights, ll, lr, lpdm, dgl, dgr, tau=1e-2, model="bgk", a=1.0):
    """example/bgk_wave.jl:69-129 with the periodic e2f/f2e tables of :42-67.
    u[ncell, nu, nsp].  model="advection": mol! of example/advection_kinetic.jl:73-128 -- identical but for the
    Maxwellian, built from rho = sum(u .* weights) and prim = [rho, a, 1.0] (:80-88; tau = 2e-3 there)."""
    ncell, nu, nsp = u.shape
    delta = heaviside(velo)
    M = np.empty_like(u)
    for k in range(nsp):
        w = moments_conserve_1v(u[:, :, k], velo, weights)
        if model == "advection":
            prim = np.stack([w[..., 0], np.full_like(w[..., 0], a), np.ones_like(w[..., 0])], axis=-1)
        else:
            prim = conserve_prim(w, 3.0)
        M[:, :, k] = maxwellian(velo[None, :], prim)
    J = 0.5 * np.asarray(dx)
    f = velo[None, :, None] * u / J[:, None, None]
    f_face = np.stack([_dot_last(f, ll), _dot_last(f, lr)], axis=-1)  # [:,:,0]=ll, [:,:,1]=lr
    nface = ncell + 1
    # f2e[i,1] = element right of the face (wraps), f2e[i,2] = element left (wraps)
    f2e1 = np.arange(nface) % ncell
    f2e1[-1] = 0
    f2e2 = (np.arange(nface) - 1) % ncell
    fi = f_face[f2e1, :, 0] * (1.0 - delta)[None, :] + f_face[f2e2, :, 1] * delta[None, :]
    rhs1 = np.stack([_dot_last(f, lpdm[p, :]) for p in range(nsp)], axis=-1)
    # e2f[i,2] = left face, e2f[i,1] = right face (with the wraps of :42-55)
    e2f2 = np.arange(ncell)
    e2f2[0] = nface - 1
    e2f1 = np.arange(ncell) + 1
    e2f1[-1] = 0
    du = np.empty_like(u)
    for p in range(nsp):
        du[:, :, p] = (
            -(
                rhs1[:, :, p]
                + (fi[e2f2, :] - f_face[:, :, 0]) * dgl[p]
                + (fi[e2f1, :] - f_face[:, :, 1]) * dgr[p]
            )
            + (M[:, :, p] - u[:, :, p]) / tau
        )
    return du


# ----------------------------------------------------------------------------
# positivity limiter, src/dissipation.jl:61-123 (1-D Euler) and :125-206 (2-D)
# ----------------------------------------------------------------------------


def positive_limiter_euler1d(u, gamma, weights, ll, lr):
    """dissipation.jl:61-123 applied to every cell of u[ncell, nsp, 3] in place.
    Density branch exactly as written (:70-88).  The energy branch of the reference
    ends in ``minimum(tj, t0)`` (:116), which is a MethodError for a Vector ``tj``:
    the reference *throws* whenever a flux or solution point has pressure < eps.
    The oracle therefore applies the density branch and returns the number of cells
    in which the reference would have thrown (0 on every BASELINE config)."""
    ncell, nsp, _ = u.shape
    um = np.stack([np.sum(u[:, :, j] * weights, axis=1) for j in range(3)], axis=-1)  # :71 (not normalised)
    t_mean = 1.0 / conserve_prim(um, gamma)[..., -1]
    p_mean = 0.5 * um[:, 0] * t_mean
    rb = np.stack([_dot_last(u[:, :, 0], ll), _dot_last(u[:, :, 0], lr)], axis=1)
    eps = np.minimum(np.minimum(1e-13, um[:, 0]), p_mean)
    rho_min = np.minimum(rb.min(axis=1), u[:, :, 0].min(axis=1))
    t1 = np.minimum((um[:, 0] - eps) / (um[:, 0] - rho_min + 1e-8), 1.0)
    if not np.all((t1 > 0) & (t1 <= 1)):
        raise AssertionError("incorrect range of limiter parameter t")
    u[:, :, 0] = t1[:, None] * (u[:, :, 0] - um[:, 0:1]) + um[:, 0:1]
    # energy corrector trigger (:92-113), evaluated on the density-limited state
    mb = np.stack([_dot_last(u[:, :, 1], ll), _dot_last(u[:, :, 1], lr)], axis=1)
    eb = np.stack([_dot_last(u[:, :, 2], ll), _dot_last(u[:, :, 2], lr)], axis=1)
    with np.errstate(all="ignore"):
        lam_b = conserve_prim(np.stack([rb, mb, eb], axis=-1), gamma)[..., -1]
        lam_p = conserve_prim(u, gamma)[..., -1]
    would_throw = (lam_b < eps[:, None]).any(axis=1) | (lam_p < eps[:, None]).any(axis=1)
    return int(would_throw.sum())


def tj_equation(t, w, um, gamma, eps):
    """dissipation.jl:208-215 (kept for completeness; unreachable without a throw)."""
    ut = t * (w - um) + um
    prim = conserve_prim(ut, gamma)
    return 0.5 * prim[0] / prim[-1] - eps


def positive_limiter_euler2d(u, gamma, weights, ll, lr):
    """dissipation.jl:125-206, density branch, for every interior cell of
    u[nx+2, ny+2, nsp, nsp, 4] in place (shock-vortex.jl:298-303 loop).  The mean
    follows the *intent* of :135 (one mean per conserved variable); the literal
    ``for j in axes(u, 2)`` coincides with it when nsp == 4."""
    nsp = u.shape[2]
    I = slice(1, u.shape[0] - 1)
    Jn = slice(1, u.shape[1] - 1)
    ui = u[I, Jn]
    um = np.stack([np.sum(ui[..., m] * weights[None, None], axis=(2, 3)) for m in range(4)], axis=-1)
    t_mean = 1.0 / conserve_prim(um, gamma)[..., -1]
    p_mean = 0.5 * um[..., 0] * t_mean
    rho = ui[..., 0]
    # :149-152 density traces on the four faces
    rb1 = sum(rho[:, :, :, q] * ll[q] for q in range(nsp))
    rb2 = sum(rho[:, :, q, :] * lr[q] for q in range(nsp))
    rb3 = sum(rho[:, :, :, q] * lr[q] for q in range(nsp))
    rb4 = sum(rho[:, :, q, :] * ll[q] for q in range(nsp))
    eps = np.minimum(np.minimum(1e-13, um[..., 0]), p_mean)
    rmin = np.minimum.reduce([rb1.min(-1), rb2.min(-1), rb3.min(-1), rb4.min(-1), rho.min(axis=(2, 3))])
    t1 = np.minimum((um[..., 0] - eps) / (um[..., 0] - rmin + 1e-8), 1.0)
    if not np.all((t1 > 0) & (t1 <= 1)):
        raise AssertionError("incorrect range of limiter parameter t")
    u[I, Jn, :, :, 0] = t1[..., None, None] * (rho - um[..., 0][..., None, None]) + um[..., 0][..., None, None]
    return u


# ----------------------------------------------------------------------------
# L4: fixed-step integrators (OrdinaryDiffEq semantics)
# ----------------------------------------------------------------------------

SCHEMES = ("euler", "midpoint", "ssprk3")


# ------------------------------------------------------------------ shock sensor + modal filter
def shock_detector(Se, deg, S0=None, kappa=4.0):
    """src/dissipation.jl:13-23."""
    if S0 is None:
        S0 = -3.0 * np.log10(deg)
    if Se < S0 - kappa:
        sigma = 1.0
    elif S0 - kappa <= Se < S0 + kappa:
        sigma = 0.5 * (1.0 - np.sin(0.5 * np.pi * (Se - S0) / kappa))
    else:
        sigma = 0.0
    return sigma < 0.99


def modal_filter_l2(u_hat, lam):
    """[KB] KitBase.modal_filter!(u, lam; filter=:l2) on a vector of modes: mode k >= 1 is divided
    by 1 + lam k^2 (k+1)^2, mode 0 is kept."""
    out = np.array(u_hat, dtype=np.float64)
    for k in range(1, len(out)):
        out[k] /= 1.0 + lam * (k + 1) ** 2 * k**2
    return out


def filter_pass_1d(u, V, iV, deg, lam, eps=1e-6, S0=None, kappa=4.0):
    """example/euler_highlevel.jl:37-52 on u[ncell, nsp, 3]; returns the number of filtered cells."""
    n = 0
    for i in range(u.shape[0]):
        uh = iV @ u[i, :, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            su = uh[-1] ** 2 / (np.sum(uh**2) + eps)
            Se = np.log10(su)
        if shock_detector(Se, deg, S0, kappa):
            n += 1
            for s in range(u.shape[2]):
                u[i, :, s] = V @ modal_filter_l2(iV @ u[i, :, s], lam)
    return n


def filter_pass_2d(u, V, iV, deg, lam, eps=0.0, S0=None, kappa=9.0, ghosts=True):
    """example/shock-vortex.jl:308-321 on u[nx+2, ny+2, nsp, nsp, 4]: the modes of the element
    block in [:] order (first tensor index fastest) go through the *vector* l2 filter."""
    nsp = u.shape[2]
    n = 0
    I = range(u.shape[0]) if ghosts else range(1, u.shape[0] - 1)
    Jr = range(u.shape[1]) if ghosts else range(1, u.shape[1] - 1)
    for i in I:
        for j in Jr:
            uh = iV @ u[i, j, :, :, 0].reshape(-1, order="F")
            with np.errstate(divide="ignore", invalid="ignore"):
                su = uh[-1] ** 2 / (np.sum(uh**2) + eps)
                Se = np.log10(su)
            if shock_detector(Se, deg, S0, kappa):
                n += 1
                for s in range(4):
                    uh = iV @ u[i, j, :, :, s].reshape(-1, order="F")
                    u[i, j, :, :, s] = (V @ modal_filter_l2(uh, lam)).reshape(nsp, nsp, order="F")
    return n


def step(u, dt, rhs, scheme="midpoint"):
    """One fixed step.  Euler: u+dt L(u).  Midpoint: k1=L(u), k2=L(u+dt/2 k1),
    u+dt k2.  SSPRK3: Shu-Osher convex form (not in the reference; north_star)."""
    if scheme == "euler":
        return u + dt * rhs(u)
    if scheme == "midpoint":
        k1 = rhs(u)
        return u + dt * rhs(u + (0.5 * dt) * k1)
    if scheme == "ssprk3":
        u1 = u + dt * rhs(u)
        u2 = 0.75 * u + 0.25 * (u1 + dt * rhs(u1))
        return (1.0 / 3.0) * u + (2.0 / 3.0) * (u2 + dt * rhs(u2))
    if scheme in RK_TABLEAUS:
        A, b = RK_TABLEAUS[scheme]
        k = []
        for i in range(len(b)):
            ui = u
            if any(A[i][j] != 0.0 for j in range(i)):
                ui = u + dt * sum(A[i][j] * k[j] for j in range(i) if A[i][j] != 0.0)
            k.append(rhs(ui))
        return u + dt * sum(b[j] * k[j] for j in range(len(b)) if b[j] != 0.0)
    raise ValueError(scheme)


# Explicit tableaus for the schemes the reference's scripts use besides Euler / Midpoint:
# Tsit5 with adaptive=false (advection_highlevel.jl:26, euler1d_convergence.jl:133) -- the six
# evaluated stages of Tsitouras (2011) in OrdinaryDiffEq's u = uprev + dt * (a_i1 k1 + ...) form -- and RK4.
RK_TABLEAUS = {
    "rk4": ([[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 1.0, 0]], [1 / 6, 1 / 3, 1 / 3, 1 / 6]),
    "tsit5": (
        [
            [0, 0, 0, 0, 0, 0],
            [0.161, 0, 0, 0, 0, 0],
            [-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0],
            [2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0],
            [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0],
            [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0],
        ],
        [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
    ),
}


def integrate(u, dt, nsteps, rhs, scheme="midpoint", before_step=None):
    """The user loop of euler2d_wave.jl:125-135: optional in-place hook, then step."""
    u = u.copy(order="F")
    for _ in range(nsteps):
        if before_step is not None:
            before_step(u)
        u = np.asfortranarray(step(u, dt, rhs, scheme))
    return u


# ----------------------------------------------------------------------------
# initial conditions of the BASELINE configs
# ----------------------------------------------------------------------------


def ic_advection1d(ps: FRPSpace1D):
    """advection_lowlevel.jl:55-58."""
    return np.asfortranarray(np.sin(np.pi * ps.xpg[ps.ng : ps.ng + ps.nx]))


def ic_sod1d(ps: FRPSpace1D, gamma=5.0 / 3.0):
    """euler_lowlevel.jl:18-28."""
    xc = ps.x[ps.ng : ps.ng + ps.nx]
    prim = np.where((xc <= 0.5)[:, None], np.array([1.0, 0.0, 0.5]), np.array([0.3, 0.0, 0.625]))
    W = prim_conserve(prim, gamma)
    return np.asfortranarray(np.repeat(W[:, None, :], ps.deg + 1, axis=1))


def ic_wave1d(ps: FRPSpace1D, gamma=5.0 / 3.0, amp=0.1):
    """smooth variant commented in euler_lowlevel.jl:26 (density wave, unit velocity)."""
    x = ps.xpg[ps.ng : ps.ng + ps.nx]
    prim = np.stack([1 + amp * np.sin(2 * np.pi * x), np.ones_like(x), np.ones_like(x)], axis=-1)
    return np.asfortranarray(prim_conserve(prim, gamma))


def ic_wave2d(ps: FRPSpace2D, gamma=5.0 / 3.0, direction="x"):
    """euler2d_wave.jl:115-120 / :146-151."""
    coord = ps.xpg[..., 0 if direction == "x" else 1]
    rho = 1.0 + 0.1 * np.sin(2 * np.pi * coord)
    prim = np.empty(coord.shape + (4,))
    prim[..., 0] = rho
    prim[..., 1] = 1.0 if direction == "x" else 0.0
    prim[..., 2] = 0.0 if direction == "x" else 1.0
    prim[..., 3] = rho
    return np.asfortranarray(prim_conserve(prim, gamma))


def ic_bgk1d(ps: FRPSpace1D, velo):
    """bgk_wave.jl:33-40."""
    x = ps.xpg[ps.ng : ps.ng + ps.nx]
    rho = 1.0 + 0.1 * np.sin(2.0 * np.pi * x)
    T = 2 * 0.5 / rho
    prim = np.stack([rho, np.ones_like(rho), 1.0 / T], axis=-1)  # [ncell, nsp, 3]
    f0 = np.empty((ps.nx, len(velo), ps.deg + 1), order="F")
    for k in range(ps.deg + 1):
        f0[:, :, k] = maxwellian(velo[None, :], prim[:, k, :])
    return f0


def ic_kinetic_advection1d(ps: FRPSpace1D, velo, a=1.0):
    """advection_kinetic.jl:37-45: u = 1 - sin(pi x), f = maxwellian(v, [u, a, 1]) ([KB-recall]: the scalar
    conserve_prim(u, a) the script calls is the triple the mol! spells out at :84)."""
    x = ps.xpg[ps.ng : ps.ng + ps.nx]
    rho = 1.0 - np.sin(np.pi * x)
    prim = np.stack([rho, np.full_like(rho, a), np.ones_like(rho)], axis=-1)
    f0 = np.empty((ps.nx, len(velo), ps.deg + 1), order="F")
    for k in range(ps.deg + 1):
        f0[:, :, k] = maxwellian(velo[None, :], prim[:, k, :])
    return f0


# ----------------------------------------------------------------------------
# 2-D gas-kinetic Navier-Stokes (config 5): example/ns_cavity.jl
# ----------------------------------------------------------------------------
# Layout u[4, ns, nr, ny+2, nx+2] (variable fastest; ns_cavity.jl:33).  [KB] closures restated
# from KitBase 0.9.  Quirks kept: see oracle/fr_oracle_gks.c header.


def ref_vhs_vis(Kn, alpha, omega):
    """[KB] ref_vhs_vis."""
    return 5.0 * (alpha + 1.0) * (alpha + 2.0) * math.sqrt(math.pi) / (
        4.0 * alpha * (5.0 - 2.0 * omega) * (7.0 - 2.0 * omega)) * Kn


def _erfc(x):
    from scipy.special import erfc

    return erfc(x)


def gauss_moments(prim, K):
    """[KB] gauss_moments(prim, inK) for prim[..., 4] -> Mu, Mv, Mxi, MuL, MuR (moment index first)."""
    U, V, lam = prim[..., 1], prim[..., 2], prim[..., 3]
    MuL = np.empty((7,) + U.shape)
    MuR = np.empty_like(MuL)
    MuL[0] = 0.5 * _erfc(-np.sqrt(lam) * U)
    MuL[1] = U * MuL[0] + 0.5 * np.exp(-lam * U * U) / np.sqrt(np.pi * lam)
    MuR[0] = 0.5 * _erfc(np.sqrt(lam) * U)
    MuR[1] = U * MuR[0] - 0.5 * np.exp(-lam * U * U) / np.sqrt(np.pi * lam)
    for i in range(2, 7):
        MuL[i] = U * MuL[i - 1] + 0.5 * (i - 1) * MuL[i - 2] / lam
        MuR[i] = U * MuR[i - 1] + 0.5 * (i - 1) * MuR[i - 2] / lam
    Mu = MuL + MuR
    Mv = np.empty_like(MuL)
    Mv[0] = 1.0
    Mv[1] = V
    for i in range(2, 7):
        Mv[i] = V * Mv[i - 1] + 0.5 * (i - 1) * Mv[i - 2] / lam
    Mxi = np.empty((3,) + U.shape)
    Mxi[0] = 1.0
    Mxi[1] = 0.5 * K / lam
    Mxi[2] = (K * K + 2.0 * K) / (4.0 * lam * lam)
    return Mu, Mv, Mxi, MuL, MuR


def moments_conserve_2d(Mu, Mv, Mw, a, b, d):
    """[KB] moments_conserve(Mu, Mv, Mw, alpha, beta, delta) -> [..., 4]."""
    uv = np.empty(Mu.shape[1:] + (4,))
    uv[..., 0] = Mu[a] * Mv[b] * Mw[d // 2]
    uv[..., 1] = Mu[a + 1] * Mv[b] * Mw[d // 2]
    uv[..., 2] = Mu[a] * Mv[b + 1] * Mw[d // 2]
    uv[..., 3] = 0.5 * (Mu[a + 2] * Mv[b] * Mw[d // 2] + Mu[a] * Mv[b + 2] * Mw[d // 2]
                        + Mu[a] * Mv[b] * Mw[(d + 2) // 2])
    return uv


def moments_conserve_slope_2d(sl, Mu, Mv, Mw, a, b):
    """[KB] moments_conserve_slope."""
    s = [sl[..., q:q + 1] for q in range(4)]
    return (s[0] * moments_conserve_2d(Mu, Mv, Mw, a, b, 0) + s[1] * moments_conserve_2d(Mu, Mv, Mw, a + 1, b, 0)
            + s[2] * moments_conserve_2d(Mu, Mv, Mw, a, b + 1, 0)
            + 0.5 * s[3] * moments_conserve_2d(Mu, Mv, Mw, a + 2, b, 0)
            + 0.5 * s[3] * moments_conserve_2d(Mu, Mv, Mw, a, b + 2, 0)
            + 0.5 * s[3] * moments_conserve_2d(Mu, Mv, Mw, a, b, 2))


def pdf_slope_2d(prim, sw, K):
    """[KB] pdf_slope."""
    rho, U, V, lam = (prim[..., q] for q in range(4))
    sl = np.empty_like(sw)
    sl[..., 3] = 4.0 * lam * lam / (K + 2.0) / rho * (
        2.0 * sw[..., 3] - 2.0 * U * sw[..., 1] - 2.0 * V * sw[..., 2]
        + sw[..., 0] * (U * U + V * V - 0.5 * (K + 2.0) / lam))
    sl[..., 2] = 2.0 * lam / rho * (sw[..., 2] - V * sw[..., 0]) - V * sl[..., 3]
    sl[..., 1] = 2.0 * lam / rho * (sw[..., 1] - U * sw[..., 0]) - U * sl[..., 3]
    sl[..., 0] = sw[..., 0] / rho - U * sl[..., 1] - V * sl[..., 2] - 0.5 * (
        U * U + V * V + 0.5 * (K + 2.0) / lam) * sl[..., 3]
    return sl


def vhs_collision_time(prim, mu, omega):
    return mu * 2.0 * prim[..., 3] ** (1.0 - omega) / prim[..., 0]


def flux_gks_point(w, K, gamma, mu, omega):
    """ns_cavity.jl:49-73 with sw = 0."""
    prim = conserve_prim(w, gamma)
    Mu, Mv, Mxi, _, _ = gauss_moments(prim, K)
    tau = vhs_collision_time(prim, mu, omega)[..., None]
    a = pdf_slope_2d(prim, np.zeros_like(w), K)
    dft = -prim[..., 0:1] * moments_conserve_slope_2d(a, Mu, Mv, Mxi, 1, 0)
    A = pdf_slope_2d(prim, dft, K)
    Muv = moments_conserve_2d(Mu, Mv, Mxi, 1, 0, 0)
    Mau = moments_conserve_slope_2d(a, Mu, Mv, Mxi, 2, 0)
    Mtu = moments_conserve_slope_2d(A, Mu, Mv, Mxi, 1, 0)
    return prim[..., 0:1] * (Muv - tau * Mau - tau * Mtu)


def flux_gks_face(wL, wR, K, gamma, mu, omega, dt, swL, swR):
    """ns_cavity.jl:75-145."""
    pL, pR = conserve_prim(wL, gamma), conserve_prim(wR, gamma)
    Mu1, Mv1, Mxi1, MuL1, _ = gauss_moments(pL, K)
    Mu2, Mv2, Mxi2, _, MuR2 = gauss_moments(pR, K)
    w = pL[..., 0:1] * moments_conserve_2d(MuL1, Mv1, Mxi1, 0, 0, 0) + pR[..., 0:1] * moments_conserve_2d(
        MuR2, Mv2, Mxi2, 0, 0, 0)
    prim = conserve_prim(w, gamma)
    tau = vhs_collision_time(prim, mu, omega) + 2.0 * dt * np.abs(
        pL[..., 0] / pL[..., 3] - pR[..., 0] / pR[..., 3]) / (pL[..., 0] / pL[..., 3] + pR[..., 0] / pR[..., 3])
    tau = tau[..., None]
    faL = pdf_slope_2d(pL, swL, K)
    faTL = pdf_slope_2d(pL, -pL[..., 0:1] * moments_conserve_slope_2d(faL, Mu1, Mv1, Mxi1, 1, 0), K)
    faR = pdf_slope_2d(pR, swR, K)
    faTR = pdf_slope_2d(pR, -pR[..., 0:1] * moments_conserve_slope_2d(faR, Mu2, Mv1, Mxi2, 1, 0), K)  # Mv1 (:102)
    Mu, Mv, Mxi, _, _ = gauss_moments(prim, K)
    Mt4 = dt
    Mt1 = dt - Mt4
    fw = Mt1 * prim[..., 0:1] * moments_conserve_2d(Mu, Mv, Mxi, 1, 0, 0)
    fw = fw + (
        Mt4 * pL[..., 0:1] * moments_conserve_2d(MuL1, Mv1, Mxi1, 1, 0, 0)
        - tau * Mt4 * pL[..., 0:1] * moments_conserve_slope_2d(faL, MuL1, Mv1, Mxi1, 2, 0)
        - tau * Mt4 * pL[..., 0:1] * moments_conserve_slope_2d(faTL, MuL1, Mv1, Mxi1, 1, 0)
        + Mt4 * pR[..., 0:1] * moments_conserve_2d(MuR2, Mv2, Mxi2, 1, 0, 0)
        - tau * Mt4 * pR[..., 0:1] * moments_conserve_slope_2d(faR, MuR2, Mv2, Mxi2, 2, 0)
        - tau * Mt4 * pR[..., 0:1] * moments_conserve_slope_2d(faTR, MuR2, Mv2, Mxi2, 1, 0))
    return fw / dt


def ns_boundary(u, gamma, lam0=1.0, lid=0.15):
    """boundary!(u, p, 1.0): ns_cavity.jl:289-344, in place.  u[4, ns, nr, ny+2, nx+2]."""
    _, ns, nr, nyg, nxg = u.shape
    nx, ny = nxg - 2, nyg - 2

    def mirror(w, top):
        prim = conserve_prim(w, gamma)
        pb = np.empty_like(prim)
        pb[..., 3] = 2 * lam0 - prim[..., 3]
        tmp = (prim[..., 3] - lam0) / lam0
        pb[..., 0] = (1.0 - tmp) / (1.0 + tmp) * prim[..., 0]
        pb[..., 1] = lid if top else -prim[..., 1]
        pb[..., 2] = -prim[..., 2]
        return prim_conserve(pb, gamma)

    # arrays as [..., 4] views: move variable axis last
    ul = np.moveaxis(u, 0, -1)  # [ns, nr, nyg, nxg, 4]
    ul[:, ::-1, 1:ny + 1, 0, :] = mirror(ul[:, :, 1:ny + 1, 1, :], False)
    ul[:, ::-1, 1:ny + 1, nx + 1, :] = mirror(ul[:, :, 1:ny + 1, nx, :], False)
    ul[::-1, :, 0, 1:nx + 1, :] = mirror(ul[:, :, 1, 1:nx + 1, :], False)
    ul[::-1, :, ny + 1, 1:nx + 1, :] = mirror(ul[:, :, ny, 1:nx + 1, :], True)
    return u


def rhs_ns2d(u, ps: FRPSpace2D, K, gamma, mu, omega, dt, lam0=1.0, lid=0.15):
    """dudt! of ns_cavity.jl:147-287 (u's ghosts are rewritten in place, as in the reference)."""
    ns_boundary(u, gamma, lam0, lid)
    _, ns, nr, nyg, nxg = u.shape
    nx, ny = nxg - 2, nyg - 2
    Jx, Jy = ps.Jx, ps.Jy
    ul = np.moveaxis(u, 0, -1)  # [l, k, j, i, m]
    I, Jn = slice(1, nx + 1), slice(1, ny + 1)
    fx = np.zeros_like(ul)
    fy = np.zeros_like(ul)
    fx[:, :, Jn, I] = flux_gks_point(ul[:, :, Jn, I], K, gamma, mu, omega) / Jx
    fy[:, :, Jn, I] = global_frame(flux_gks_point(local_frame(ul[:, :, Jn, I], 0.0, 1.0), K, gamma, mu, omega),
                                   0.0, 1.0) / Jy

    def con_k(a, l):  # contract the r index (axis 1)
        return sum(a[:, q] * l[q] for q in range(nr))

    def con_l(a, l):  # contract the s index (axis 0)
        return sum(a[q] * l[q] for q in range(ns))

    uxL, uxR = con_k(ul, ps.ll), con_k(ul, ps.lr)  # [l, j, i, m]
    fxL, fxR = con_k(fx, ps.ll), con_k(fx, ps.lr)
    uyB, uyT = con_l(ul, ps.ll), con_l(ul, ps.lr)  # [k, j, i, m]
    fyB, fyT = con_l(fy, ps.ll), con_l(fy, ps.lr)
    # x interfaces i = 1..nx+1 (left cell i-1, right cell i), rows j = 1..ny
    swL = con_k(ul, ps.dll)[:, Jn, 0:nx + 1] / Jx
    swR = con_k(ul, ps.dlr)[:, Jn, 1:nx + 2] / Jx
    fxi = flux_gks_face(uxR[:, Jn, 0:nx + 1], uxL[:, Jn, 1:nx + 2], K, gamma, mu, omega, dt, swL, swR)  # [l, j, iface, m]
    swL = con_l(ul, ps.dll)[:, 0:ny + 1, I] / Jy
    swR = con_l(ul, ps.dlr)[:, 1:ny + 2, I] / Jy
    fyi = global_frame(flux_gks_face(local_frame(uyT[:, 0:ny + 1, I], 0.0, 1.0), local_frame(uyB[:, 1:ny + 2, I], 0.0, 1.0),
                                     K, gamma, mu, omega, dt, swL, swR), 0.0, 1.0)  # [k, jface, i, m]
    du = np.zeros_like(ul)
    for k in range(nr):
        for l in range(ns):
            r1 = sum(fx[l, q, Jn, I] * ps.dl[k, q] for q in range(nr))
            r2 = sum(fy[q, k, Jn, I] * ps.dl[l, q] for q in range(ns))
            du[l, k, Jn, I] = -(
                r1 + r2
                + (fxi[l, :, 0:nx] / Jx - fxL[l, Jn, I]) * ps.dhl[k]
                + (fxi[l, :, 1:nx + 1] / Jx - fxR[l, Jn, I]) * ps.dhr[k]
                + (fyi[k, 0:ny] / Jy - fyB[k, Jn, I]) * ps.dhl[l]
                + (fyi[k, 1:ny + 1] / Jy - fyT[k, Jn, I]) * ps.dhr[l])
    return np.asfortranarray(np.moveaxis(du, -1, 0))


def ic_cavity(ps: FRPSpace2D, gamma=5.0 / 3.0):
    """ns_cavity.jl:33-36: prim = [1, 0, 0, 1] everywhere."""
    nsp = ps.deg + 1
    u = np.empty((4, nsp, nsp, ps.ny + 2, ps.nx + 2), order="F")
    u[...] = prim_conserve(np.array([1.0, 0.0, 0.0, 1.0]), gamma)[:, None, None, None, None]
    return u
