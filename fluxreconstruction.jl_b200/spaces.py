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
    "FRPSpace1D", "FRPSpace2D", "PSpace2D", "CSpace2D", "quad_jacobians", "face_normals",
    "correction_factors_fp", "embed_ghostless_x",
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


def vandermonde_matrix(N, r, *rest):
    """Orthonormal-Legendre Vandermonde matrix (src/Transform/transform.jl:18-26), or -- with the reference's
    shape tag first -- ``vandermonde_matrix(Line, N, r)``, ``vandermonde_matrix(Tri, N, r, s)`` (:36-51) and
    ``vandermonde_matrix(Quad, N, r, s)`` (:55-69, modes (i, j), j fastest)."""
    if isinstance(N, type):
        from . import unstruct as un

        shape, N, r, rest = N, r, rest[0], rest[1:]
        if shape is un.Line:
            return vandermonde_matrix(N, r)
        if shape is un.Tri:
            return un.simplex_vandermonde(N, r, rest[0])
        if shape is un.Quad:
            Vr, Vs = vandermonde_matrix(N, r), vandermonde_matrix(N, rest[0])
            return np.einsum("pa,pb->pab", Vr, Vs).reshape(len(np.atleast_1d(r)), (N + 1) ** 2)
        raise NotImplementedError(f"vandermonde_matrix for {shape.__name__}")
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


class PSpace2D:
    """[KB-recall] KitBase.PSpace2D: the structured physical space the 2-D scripts build, either uniform
    ``PSpace2D(x0, x1, nx, y0, y1, ny, ngx, ngy)`` or from explicit cell data
    ``PSpace2D(x0, x1, nx, y0, y1, ny, x, y, dx, dy, vertices)`` (dev/parallelogram.jl:71).  Arrays cover
    the ghost cells (index 0 = reference index 1 - ng); ``vertices[i, j, 4, 2]`` run counter-clockwise from
    the lower left corner (geo_jacobi.jl:50-66)."""

    def __init__(self, x0, x1, nx, y0, y1, ny, *rest):
        self.x0, self.x1, self.nx = float(x0), float(x1), int(nx)
        self.y0, self.y1, self.ny = float(y0), float(y1), int(ny)
        if len(rest) == 5:
            self.x, self.y, self.dx, self.dy, self.vertices = (np.asfortranarray(a, dtype=np.float64) for a in rest)
            self.ngx = (self.vertices.shape[0] - self.nx) // 2
            self.ngy = (self.vertices.shape[1] - self.ny) // 2
            return
        ngx, ngy = (list(rest) + [0, 0])[:2]
        self.ngx, self.ngy = int(ngx), int(ngy)
        dx, dy = (self.x1 - self.x0) / self.nx, (self.y1 - self.y0) / self.ny
        xc = self.x0 + (np.arange(1 - self.ngx, self.nx + self.ngx + 1) - 0.5) * dx
        yc = self.y0 + (np.arange(1 - self.ngy, self.ny + self.ngy + 1) - 0.5) * dy
        self.x, self.y = (np.asfortranarray(a) for a in np.meshgrid(xc, yc, indexing="ij"))
        self.dx, self.dy = np.full_like(self.x, dx), np.full_like(self.x, dy)
        sx, sy = np.array([-0.5, 0.5, 0.5, -0.5]), np.array([-0.5, -0.5, 0.5, 0.5])
        self.vertices = np.asfortranarray(np.stack(
            [self.x[..., None] + sx * dx, self.y[..., None] + sy * dy], axis=-1))


class CSpace2D:
    """[KB-recall] KitBase.CSpace2D(r0, r1, nr, th0, th1, nth, ngr, ngth) (dev/cylinder2.jl:22): polar
    cells, centre (r0 + (i - 1/2) dr, th0 + (j - 1/2) dth), vertices at (r -+ dr/2, th -+ dth/2)."""

    def __init__(self, r0, r1, nr, th0, th1, nth, ngr=0, ngth=0):
        self.r0, self.r1, self.nr = float(r0), float(r1), int(nr)
        self.th0, self.th1, self.nth = float(th0), float(th1), int(nth)
        self.ngr, self.ngth = int(ngr), int(ngth)
        self.nx, self.ny, self.ngx, self.ngy = self.nr, self.nth, self.ngr, self.ngth
        dr, dth = (self.r1 - self.r0) / self.nr, (self.th1 - self.th0) / self.nth
        rc = self.r0 + (np.arange(1 - ngr, nr + ngr + 1) - 0.5) * dr
        tc = self.th0 + (np.arange(1 - ngth, nth + ngth + 1) - 0.5) * dth
        self.r, self.theta = (np.asfortranarray(a) for a in np.meshgrid(rc, tc, indexing="ij"))
        self.dr, self.dtheta = np.full_like(self.r, dr), np.full_like(self.r, dth)
        self.x, self.y = self.r * np.cos(self.theta), self.r * np.sin(self.theta)
        sr, st = np.array([-0.5, 0.5, 0.5, -0.5]), np.array([-0.5, -0.5, 0.5, 0.5])
        rv, tv = self.r[..., None] + sr * dr, self.theta[..., None] + st * dth
        self.vertices = np.asfortranarray(np.stack([rv * np.cos(tv), rv * np.sin(tv)], axis=-1))


def embed_ghostless_x(base):
    """The cylinder scripts hold every array over i = 1:nr without radial ghosts (CSpace2D(..., 0, 1),
    dev/cylinder2.jl:22,33) while the ABI wants a full ghost ring.  Cell nr is never updated there
    (cylinder2.jl:154 runs i over 1:nx-1) and only feeds face nr: it *is* the outer ghost column.  This
    returns the same cells as a PSpace2D with nx = nr - 1 interior columns, column nr as the upper ghost and a
    dummy copy of column 1 as the lower ghost (never read when x face 1 is the mirror wall).  States are
    embedded the same way: ``np.concatenate([u[:1], u])``."""
    if base.ngx != 0 or base.ngy != 1:
        raise ValueError("expects a space without ghosts in x and with one ghost layer in y")
    cat = lambda a: np.asfortranarray(np.concatenate([a[:1], a], axis=0))  # noqa: E731
    dx = getattr(base, "dx", getattr(base, "dr", None))
    dy = getattr(base, "dy", getattr(base, "dtheta", None))
    lo = getattr(base, "x0", getattr(base, "r0", 0.0)), getattr(base, "y0", getattr(base, "th0", 0.0))
    hi = getattr(base, "x1", getattr(base, "r1", 0.0)), getattr(base, "y1", getattr(base, "th1", 0.0))
    return PSpace2D(lo[0], hi[0], base.nx - 1, lo[1], hi[1], base.ny, cat(base.x), cat(base.y), cat(dx), cat(dy),
                    cat(base.vertices))


# derivatives of the bilinear shape functions lambda^1..4 (geo_jacobi.jl:55-66) as tables in (r, s)
def _shape_derivatives(r, s):
    r, s = np.asarray(r, dtype=np.float64), np.asarray(s, dtype=np.float64)
    d_r = np.stack([s - 1.0, 1.0 - s, s + 1.0, -(s + 1.0)], axis=-1) / 4.0
    d_s = np.stack([r - 1.0, -(r + 1.0), r + 1.0, 1.0 - r], axis=-1) / 4.0
    return d_r, d_s


def quad_jacobians(vertices, r, s):
    """J[i, j, ..., (x,y), (r,s)] of the bilinear map of every cell at the points (r[...], s[...])."""
    d_r, d_s = _shape_derivatives(r, s)
    D = np.stack([d_r, d_s], axis=-1)  # [..., vertex, (r, s)]
    return np.einsum("ijvc,...vd->ij...cd", np.asarray(vertices, dtype=np.float64), D)


def _inv2(J):
    det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    adj = np.stack([np.stack([J[..., 1, 1], -J[..., 0, 1]], -1), np.stack([-J[..., 1, 0], J[..., 0, 0]], -1)], -2)
    return adj / det[..., None, None]


def face_normals(vertices, ng=1):
    """Unit normals of the interior-and-boundary faces of a structured quad mesh, in the layout of the
    scripts' hand-built tables: n1[nx+1, ny, 2] (x faces, pointing towards +r) and n2[nx, ny+1, 2]
    (y faces, towards +s).  For the parallelogram and the polar mesh this reproduces
    dev/parallelogram.jl:176-186 and dev/cylinder2.jl:39-49 (theta0 = 0)."""
    v = np.asarray(vertices, dtype=np.float64)
    nx, ny = v.shape[0] - 2 * ng, v.shape[1] - 2 * ng
    I, Jn = slice(ng, ng + nx), slice(ng, ng + ny)
    # x face i = left edge (vertex 1 -> 4) of cell i; the last one = right edge (2 -> 3) of cell nx
    t = np.concatenate([v[I, Jn, 3] - v[I, Jn, 0], (v[ng + nx - 1, Jn, 2] - v[ng + nx - 1, Jn, 1])[None]], axis=0)
    n1 = np.stack([t[..., 1], -t[..., 0]], axis=-1)
    n1 /= np.linalg.norm(n1, axis=-1, keepdims=True)
    # y face j = bottom edge (1 -> 2) of cell j; the last one = top edge (4 -> 3) of cell ny
    t = np.concatenate([v[I, Jn, 1] - v[I, Jn, 0], (v[I, ng + ny - 1, 2] - v[I, ng + ny - 1, 3])[:, None]], axis=1)
    n2 = np.stack([-t[..., 1], t[..., 0]], axis=-1)
    n2 /= np.linalg.norm(n2, axis=-1, keepdims=True)
    return np.asfortranarray(n1), np.asfortranarray(n2)


def correction_factors_fp(Ji, n1, n2, ng=1):
    """The flux-point correction factors of dev/cylinder2.jl:155-158 for the ABI's ``fpc[nx, ny, nsp, 4]``:
    (inv(Ji[i,j][4,l]) n1[i,j])[1], (inv(Ji[i,j][2,l]) n1[i+1,j])[1], (inv(Ji[i,j][1,k]) n2[i,j])[2],
    (inv(Ji[i,j][3,k]) n2[i,j+1])[2]."""
    nx, ny = n1.shape[0] - 1, n2.shape[1] - 1
    iJi = _inv2(Ji[ng : ng + nx, ng : ng + ny])  # [nx, ny, face, pt, 2, 2]
    row = lambda face, c, n: np.einsum("ijpd,ijd->ijp", iJi[:, :, face, :, c, :], n)  # noqa: E731
    return np.asfortranarray(np.stack(
        [row(3, 0, n1[:-1]), row(1, 0, n1[1:]), row(0, 1, n2[:, :-1]), row(2, 1, n2[:, 1:])], axis=-1))


class FRPSpace2D:
    """FRPSpace2D(x0, x1, nx, y0, y1, ny, deg, ngx, ngy) on the uniform rectangular PSpace2D
    (struct.jl:130-245).  ``J[i,j][k,l]`` is diag(dx/2, dy/2) there and kept as Jx, Jy.

    ``FRPSpace2D(base, deg)`` (struct.jl:130) with a base space that carries ``vertices`` (PSpace2D from
    explicit cell data, CSpace2D) builds the point-wise metric instead: ``J``, ``iJ`` as
    [nxg, nyg, nsp, nsp, 2, 2] and ``Ji`` [nxg, nyg, 4, nsp, 2, 2].  ``literal_Ji=True`` (default) keeps the
    reference's flux-point tables as they are (struct.jl:145-158 puts 0.0 where -1.0 is meant and
    geo_jacobi.jl:93-94 takes one Jacobian per face); ``False`` evaluates at the real flux points."""

    def __init__(self, x0, x1=None, nx=None, y0=None, y1=None, ny=None, deg=None, ngx=0, ngy=0, literal_Ji=True, **_):
        if hasattr(x0, "vertices"):
            self._init_from_base(x0, int(x1), literal_Ji)
            return
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

    def _init_from_base(self, base, deg, literal_Ji):
        self.base, self.deg = base, deg
        self.nx, self.ny, self.ngx, self.ngy = base.nx, base.ny, base.ngx, base.ngy
        self.x, self.y, self.vertices = base.x, base.y, np.asarray(base.vertices, dtype=np.float64)
        nsp = deg + 1
        self.np = nsp * nsp
        r = legendre_point(deg)
        self.xpl = r
        rr, ss = np.meshgrid(r, r, indexing="ij")  # [k, l]: k <-> r, l <-> s (struct.jl:169-174)
        self._J = np.asfortranarray(quad_jacobians(self.vertices, rr, ss))
        self._iJ = np.asfortranarray(_inv2(self._J))
        if literal_Ji:  # one point per face: (ri[face, 1], si[face, 1]) of the tables at struct.jl:145-155
            fr = np.repeat(np.array([r[0], 1.0, r[-1], 0.0])[:, None], nsp, axis=1)
            fs = np.repeat(np.array([0.0, r[0], 1.0, r[-1]])[:, None], nsp, axis=1)
        else:  # faces 1..4: s = -1, r = +1, s = +1, r = -1, points in trace order
            one = np.ones(nsp)
            fr, fs = np.stack([r, one, r, -one]), np.stack([-one, r, one, r])
        self._Ji = np.asfortranarray(quad_jacobians(self.vertices, fr, fs))
        # bilinear map of the solution points (struct.jl:161-175)
        N = np.stack([(rr - 1) * (ss - 1), (rr + 1) * (1 - ss), (rr + 1) * (ss + 1), (1 - rr) * (ss + 1)], -1) / 4.0
        self.xpg = np.asfortranarray(np.einsum("ijvc,klv->ijklc", self.vertices, N))
        w = gausslegendre(nsp)[1]
        self.wp = np.asfortranarray(np.outer(w, w))
        self.ll, self.lr, self.dl = standard_lagrange(r)
        self.dll, self.dlr, _ = _edge_slopes(deg, r)
        self.dhl, self.dhr = dradau(deg, r)
        V1 = vandermonde_matrix(deg, r)
        self.V = np.asfortranarray(np.einsum("ka,lb->lkab", V1, V1).reshape(nsp * nsp, nsp * nsp))
        self.iV = np.asfortranarray(np.linalg.inv(self.V))

    @property
    def J(self):
        """ps.J[i,j][k,l] of the reference, materialised lazily: (nxg, nyg, nsp, nsp, 2, 2)."""
        if hasattr(self, "_J"):
            return self._J
        nsp = self.deg + 1
        J = np.zeros((self.nx + 2 * self.ngx, self.ny + 2 * self.ngy, nsp, nsp, 2, 2), order="F")
        J[..., 0, 0] = self.Jx
        J[..., 1, 1] = self.Jy
        return J

    @property
    def vertices(self):
        """ps.vertices[i, j, 4, 2] (forwarded from the base space, tools.jl:5-22), counter-clockwise from the
        lower left corner."""
        if "_vertices" not in self.__dict__:
            self._vertices = PSpace2D(self.x0, self.x1, self.nx, self.y0, self.y1, self.ny, self.ngx, self.ngy).vertices
        return self._vertices

    @vertices.setter
    def vertices(self, v):
        self._vertices = v

    @property
    def iJ(self):
        """ps.iJ[i,j][k,l] (struct.jl:137-142): the point-wise inverse, (nxg, nyg, nsp, nsp, 2, 2)."""
        if hasattr(self, "_iJ"):
            return self._iJ
        iJ = np.zeros_like(self.J)
        iJ[..., 0, 0] = 1.0 / self.Jx
        iJ[..., 1, 1] = 1.0 / self.Jy
        return iJ

    @property
    def Ji(self):
        """ps.Ji[i,j][face, pt] (struct.jl:145-158): Jacobians at the flux points, (nxg, nyg, 4, nsp, 2, 2)."""
        if hasattr(self, "_Ji"):
            return self._Ji
        nsp = self.deg + 1
        Ji = np.zeros((self.nx + 2 * self.ngx, self.ny + 2 * self.ngy, 4, nsp, 2, 2), order="F")
        Ji[..., 0, 0] = self.Jx
        Ji[..., 1, 1] = self.Jy
        return Ji
