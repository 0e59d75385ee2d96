"""Host-side mirror of the reference's FR spaces and operator builders.

Setup-time only (the reference builds these once per space on the CPU as well):
``FRPSpace1D`` / ``FRPSpace2D`` (src/struct.jl:13-88, 99-245), ``legendre_point``
(src/Polynomial/poly_legendre.jl:6), ``lagrange_point`` / ``∂lagrange`` /
``standard_lagrange`` (poly_lagrange.jl:6-59, 99-105), ``∂radau`` / ``∂sd`` / ``∂huynh``
(poly_legendre.jl:29-71), ``global_sp`` (src/Geometry/geo_points.jl:8-41) and
``rs_jacobi`` for rectangles (geo_jacobi.jl:77-108).  The arrays produced here are what
the C ABI takes as ``frb_operators``; nothing in this file runs per stage.

The constructions are deliberately *not* the reference's product loops: Lagrange values
and the differentiation matrix come from barycentric weights, Legendre derivatives from
``numpy.polynomial``.  tests/ check them against the oracle's literal restatement the way
example/vandermonde_lagrange.jl:14-15,25 checks ``ll``/``lr``/``dl`` two ways.

Index convention: a reference array with axes ``0:nx+1`` (OffsetArray, one ghost layer)
is the plain array here, so NumPy index == Julia index when ng == 1 and
NumPy index == Julia index - 1 when ng == 0.  All arrays are Fortran-ordered.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import legendre as L

__all__ = [
    "legendre_point", "gausslegendre", "lagrange_point", "dlagrange", "standard_lagrange", "dlegendre",
    "dradau", "dsd", "dhuynh", "vandermonde_matrix", "dvandermonde_matrix", "global_sp", "r_x",
    "FRPSpace1D", "FRPSpace2D",
]


def gausslegendre(n: int):
    x, w = L.leggauss(n)
    return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)


def legendre_point(p: int) -> np.ndarray:
    return gausslegendre(p + 1)[0]


def _bary_weights(sp):
    sp = np.asarray(sp, dtype=np.float64)
    d = sp[:, None] - sp[None, :]
    np.fill_diagonal(d, 1.0)
    return 1.0 / d.prod(axis=1)


def lagrange_point(sp, x):
    """Values l_k(x) of the Lagrange basis on ``sp``; x scalar -> (nsp,), vector -> (len(x), nsp)."""
    sp = np.asarray(sp, dtype=np.float64)
    if np.ndim(x) > 0:
        return np.stack([lagrange_point(sp, float(xi)) for xi in x], axis=0)
    hit = np.isclose(sp, x, rtol=0, atol=0)
    if hit.any():
        return hit.astype(np.float64)
    w = _bary_weights(sp)
    t = w / (x - sp)
    return t / t.sum()


def dlagrange(sp):
    """lpdm[m, k] = l_k'(sp[m]) from the barycentric differentiation matrix."""
    sp = np.asarray(sp, dtype=np.float64)
    w = _bary_weights(sp)
    n = len(sp)
    D = np.zeros((n, n))
    for m in range(n):
        for k in range(n):
            if m != k:
                D[m, k] = (w[k] / w[m]) / (sp[m] - sp[k])
        D[m, m] = -np.sum(np.delete(D[m], m))
    return D


def standard_lagrange(x):
    return lagrange_point(x, -1.0), lagrange_point(x, 1.0), dlagrange(x)


def dlegendre(p: int, x):
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if p <= 0:
        return np.zeros_like(x)
    return L.Legendre.basis(p).deriv()(x)


def dradau(p: int, x):
    d, dp = dlegendre(p, x), dlegendre(p + 1, x)
    return (-1.0) ** p * 0.5 * (d - dp), 0.5 * (d + dp)


def dsd(p: int, x):
    dm, d, dp = dlegendre(p - 1, x), dlegendre(p, x), dlegendre(p + 1, x)
    y = (p * dm + (p + 1) * dp) / (2 * p + 1)
    return (-1.0) ** p * 0.5 * (d - y), 0.5 * (d + y)


def dhuynh(p: int, x):
    dm, d, dp = dlegendre(p - 1, x), dlegendre(p, x), dlegendre(p + 1, x)
    y = ((p + 1) * dm + p * dp) / (2 * p + 1)
    return (-1.0) ** p * 0.5 * (d - y), 0.5 * (d + y)


_CORRECTION = {"radau": dradau, "sd": dsd, "huynh": dhuynh}


def vandermonde_matrix(N: int, r):
    """Orthonormal-Legendre Vandermonde matrix (src/Transform/transform.jl:18-26)."""
    r = np.asarray(r, dtype=np.float64)
    return np.stack([np.sqrt((2 * j + 1) / 2.0) * L.Legendre.basis(j)(r) for j in range(N + 1)], axis=1)


def dvandermonde_matrix(N: int, r):
    r = np.asarray(r, dtype=np.float64)
    return np.stack([np.sqrt((2 * j + 1) / 2.0) * dlegendre(j, r) for j in range(N + 1)], axis=1)


def r_x(r, vl, vr):
    return ((1.0 - r) / 2.0) * vl + ((1.0 + r) / 2.0) * vr


def global_sp(xi, r):
    xi = np.asarray(xi, dtype=np.float64)
    out = np.empty((len(xi) - 1, len(r)), order="F")
    for j, rj in enumerate(r):
        out[:, j] = r_x(rj, xi[:-1], xi[1:])
    return out


def _edge_slopes(deg, r):
    V = vandermonde_matrix(deg, r)
    dVf = dvandermonde_matrix(deg, np.array([-1.0, 1.0]))
    return np.linalg.solve(V.T, dVf[0]), np.linalg.solve(V.T, dVf[1]), V


class FRPSpace1D:
    """FRPSpace1D(x0, x1, nx, deg, ng=0, correction=:radau)  (struct.jl:40-88)."""

    def __init__(self, x0, x1, nx, deg, ng=0, correction="radau", **_):
        self.x0, self.x1, self.nx, self.deg, self.ng = float(x0), float(x1), int(nx), int(deg), int(ng)
        correction = str(correction).lstrip(":")
        dx = (self.x1 - self.x0) / self.nx
        idx = np.arange(1 - ng, nx + ng + 1)
        self.x = self.x0 + (idx - 0.5) * dx
        self.dx = np.full(idx.shape, dx)
        self.J = self.dx / 2.0
        self.np = deg + 1
        r = legendre_point(deg)
        self.xpl = r
        xi = np.append(self.x - 0.5 * self.dx, self.x[-1] + 0.5 * self.dx[-1])
        self.xpg = global_sp(xi, r)
        self.wp = gausslegendre(deg + 1)[1]
        self.ll, self.lr, self.dl = standard_lagrange(r)
        self.dll, self.dlr, self.V = _edge_slopes(deg, r)
        self.iV = np.linalg.inv(self.V)
        self.dhl, self.dhr = _CORRECTION[correction](deg, r)

    # interior views (reference index 1:nx)
    def interior(self, a):
        return a[self.ng : self.ng + self.nx]


class FRPSpace2D:
    """FRPSpace2D(x0, x1, nx, y0, y1, ny, deg, ngx, ngy) on the uniform rectangular PSpace2D
    (struct.jl:130-245).  ``J[i,j][k,l]`` is diag(dx/2, dy/2) here and kept as Jx, Jy."""

    def __init__(self, x0, x1, nx, y0, y1, ny, deg, ngx=0, ngy=0, **_):
        self.x0, self.x1, self.nx = float(x0), float(x1), int(nx)
        self.y0, self.y1, self.ny = float(y0), float(y1), int(ny)
        self.deg, self.ngx, self.ngy = int(deg), int(ngx), int(ngy)
        self.dx = (self.x1 - self.x0) / self.nx
        self.dy = (self.y1 - self.y0) / self.ny
        self.Jx, self.Jy = self.dx / 2.0, self.dy / 2.0
        nsp = deg + 1
        self.np = nsp * nsp
        r = legendre_point(deg)
        self.xpl = r
        ii = np.arange(1 - ngx, nx + ngx + 1)
        jj = np.arange(1 - ngy, ny + ngy + 1)
        xc = self.x0 + (ii - 0.5) * self.dx
        yc = self.y0 + (jj - 0.5) * self.dy
        self.x = np.asfortranarray(np.repeat(xc[:, None], len(jj), axis=1))
        self.y = np.asfortranarray(np.repeat(yc[None, :], len(ii), axis=0))
        xs = global_sp(np.append(xc - 0.5 * self.dx, xc[-1] + 0.5 * self.dx), r)  # [i, k]
        ys = global_sp(np.append(yc - 0.5 * self.dy, yc[-1] + 0.5 * self.dy), r)  # [j, l]
        self.xpg = np.empty((len(ii), len(jj), nsp, nsp, 2), order="F")
        self.xpg[..., 0] = xs[:, None, :, None]
        self.xpg[..., 1] = ys[None, :, None, :]
        w = gausslegendre(nsp)[1]
        self.wp = np.asfortranarray(np.outer(w, w))
        self.ll, self.lr, self.dl = standard_lagrange(r)
        self.dll, self.dlr, _ = _edge_slopes(deg, r)
        self.dhl, self.dhr = dradau(deg, r)
        # 2-D Vandermonde (struct.jl:196-205; vandermonde_matrix(Quad, ...), transform.jl:55-69):
        # rows = points in [:] order of u[i,j,:,:,m] (k fastest), columns = modes (i, j), j fastest
        V1 = vandermonde_matrix(deg, r)  # [point, mode]
        self.V = np.asfortranarray(np.einsum("ka,lb->lkab", V1, V1).reshape(nsp * nsp, nsp * nsp))
        self.iV = np.asfortranarray(np.linalg.inv(self.V))

    @property
    def J(self):
        """ps.J[i,j][k,l] of the reference, materialised lazily: (nxg, nyg, nsp, nsp, 2, 2)."""
        nsp = self.deg + 1
        J = np.zeros((self.nx + 2 * self.ngx, self.ny + 2 * self.ngy, nsp, nsp, 2, 2), order="F")
        J[..., 0, 0] = self.Jx
        J[..., 1, 1] = self.Jy
        return J
