"""Host-side mirror of the reference's triangle space: ``TriFRPSpace(file, deg)``
(src/struct.jl:305-352) with everything it pulls in --

* a native Gmsh reader (MSH 4.1 and 2.2, ASCII) in place of KitBase -> PyCall -> meshio
  (``UnstructPSpace(file)``; the meshes of the reference are assets/*.msh, all 4.1 ASCII);
* the mesh connectivity KitBase derives (cellNeighbors, cellFaces, facePoints, faceCells, centres,
  areas, outward unit normals, cell / face types);
* the Williams-Shunn-Jameson solution points (src/Quadrature/quadrature.jl:14-71, which calls the
  reference's qpmin.py through PyCall), the face Gauss points (:78-101), the orthonormal simplex basis
  and its gradient (src/Polynomial/poly_triangle.jl), ``∂lagrange`` / ``lf`` / the correction field
  ``ϕ`` (src/Polynomial/poly_correction.jl) and the flux-point connectivity ``fpn``
  (src/Geometry/geo_neighbor.jl:8-60).

Setup-time only; the arrays built here are what ``frb_tri_euler_create`` takes.  The constructions
are not the reference's: the basis is evaluated as  sqrt((2i+1)(i+j+1)/2) * Q_i(xi, t) * P_j^(2i+1,0)(s)
with the scaled Legendre polynomials Q_i(xi, t) = P_i(xi/t) t^i (xi = r + (1+s)/2, t = (1-s)/2), which
has no collapsed-coordinate singularity at the top vertex; the WSJ points are expanded from their
symmetry orbits; the connectivity comes from sorted edge keys, and neighbouring flux points are
matched by edge orientation instead of comparing coordinates with ``==``.  tests/test_unstruct.py
checks every array against the reference's own golden tables (dev/check_phi.jl:60-89) and against the
oracle's literal restatement.

Index convention: everything is 0-based here (``-1`` = no neighbour);
``TriEulerProblem.from_space`` converts to the 1-based tuples of the reference at the C ABI.
"""
from __future__ import annotations

import numpy as np

from .spaces import gausslegendre

__all__ = [
    "read_msh", "read_su2", "read_mesh", "UnstructPSpace", "UnstructFRPSpace", "TriFRPSpace", "tri_quadrature", "triface_quadrature",
    "simplex_vandermonde", "dsimplex_vandermonde", "rs_xy", "wsj_points",
    "AbstractElementShape", "Line", "Quad", "Tri", "Hex", "Wed", "Pyr", "Tet",
    "JacobiP", "dJacobiP", "simplex_basis", "dsimplex_basis", "correction_field", "global_fp", "global_sp_tri",
    "neighbor_fpidx",
]

# ------------------------------------------------------------------------------------------------
# Williams-Shunn-Jameson symmetric points (J. Comput. Appl. Math. 266 (2014) 18-38), as orbit
# generators in barycentric coordinates.  ("c", w): centroid; ("s21", a, w): (a, a, 1-2a) and its
# 3 rotations; ("s111", a, b, w): the 6 permutations of (a, b, 1-a-b).  The expansion order below
# reproduces the point order the reference gets from qpmin (schemes 1..5 = deg 0..4).
_WSJ = {
    0: [("c", 1.0)],
    1: [("s21", 1.0 / 6.0, 1.0 / 3.0)],
    2: [("s21", 0.09157621350977073, 0.1099517436553219), ("s21", 0.4459484909159649, 0.2233815896780114)],
    3: [("c", 0.2015429868248577), ("s21", 0.05556405873493273, 0.04195551878545865),
        ("s111", 0.2955337173474313, 0.6342107415763597, 0.1120984094697944)],
    4: [("s21", 0.03587089730389267, 0.01791546811194674), ("s21", 0.2417293971024203, 0.1277121938457533),
        ("s21", 0.4743087877649925, 0.07620605385370617),
        ("s111", 0.2015039089284195, 0.751183601617799, 0.0557498087609636)],
}


def wsj_points(deg: int):
    """Barycentric coordinates [np, 3] and weights [np] (sum 1) of the degree-``deg`` WSJ points."""
    if deg not in _WSJ:
        raise ValueError(f"Williams-Shunn-Jameson points are tabulated for deg 0..4, got {deg}")
    orbits = _WSJ[deg]
    pts, w = [], []
    for o in orbits:
        if o[0] == "c":
            pts.append((1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0))
            w.append(o[1])
    s21 = [o for o in orbits if o[0] == "s21"]
    for rot in range(3):  # the s21 orbits are interleaved: rotation-major
        for _, a, wt in s21:
            p = [a, a, a]
            p[2 - rot] = 1.0 - 2.0 * a
            pts.append(tuple(p))
            w.append(wt)
    for o in orbits:
        if o[0] == "s111":
            a, b = o[1], o[2]
            c = 1.0 - a - b
            pts += [(a, b, c), (c, a, b), (b, c, a), (b, a, c), (c, b, a), (a, c, b)]
            w += [o[3]] * 6
    return np.array(pts, dtype=np.float64), np.array(w, dtype=np.float64)


def tri_quadrature(deg: int, vertices=None, transform=True):
    """Solution points and weights (quadrature.jl:14-71).  Default: the points in the right reference
    triangle (-1,-1), (1,-1), (-1,1) -- the barycentric point (x, y, z) sits at x*v1 + y*v2 + z*v3.
    With ``vertices`` = three corner points the reference's route is taken: trilinear -> Cartesian
    coordinates in that triangle, then (``transform``) the equilateral -> right-triangle map ``xy_rs``."""
    lam, w = wsj_points(deg)
    if vertices is None and transform:
        r = -lam[:, 0] + lam[:, 1] - lam[:, 2]
        s = -lam[:, 0] - lam[:, 1] + lam[:, 2]
        return np.stack([r, s], axis=1), w
    if vertices is None:
        vertices = ((-1.0, -1.0 / np.sqrt(3.0)), (1.0, -1.0 / np.sqrt(3.0)), (0.0, 2.0 / np.sqrt(3.0)))
    p1, p2, p3 = (np.asarray(v, dtype=np.float64) for v in vertices)
    side = np.array([np.linalg.norm(p2 - p3), np.linalg.norm(p3 - p1), np.linalg.norm(p2 - p1)])
    t = lam * side  # trilinear coordinates weighted by the opposite side lengths
    pts = (t[:, :1] * p1 + t[:, 1:2] * p2 + t[:, 2:] * p3) / t.sum(axis=1, keepdims=True)
    if transform:
        from .tools import xy_rs

        r, s = xy_rs(pts[:, 0], pts[:, 1])
        pts = np.stack([r, s], axis=1)
    return pts, w


def triface_quadrature(deg: int):
    """Gauss points along the three faces, each walked from vertex j to vertex j+1, and their weights
    scaled by the face length / 2 (quadrature.jl:78-101)."""
    g, w = gausslegendre(deg + 1)
    pf = np.zeros((3, deg + 1, 2))
    pf[0, :, 0], pf[0, :, 1] = g, -1.0
    pf[1, :, 0], pf[1, :, 1] = -g, g
    pf[2, :, 0], pf[2, :, 1] = -1.0, -g
    wf = np.stack([w, np.sqrt(2.0) * w, w])
    return pf, wf


# ------------------------------------------------------------------------------------------------
# orthonormal basis of the triangle
def _jacobi(n: int, alpha: float, beta: float, x):
    """Classical Jacobi polynomial P_n^(alpha,beta)(x), three-term recurrence."""
    x = np.asarray(x, dtype=np.float64)
    p0 = np.ones_like(x)
    if n == 0:
        return p0
    p1 = 0.5 * (alpha - beta + (alpha + beta + 2.0) * x)
    for k in range(1, n):
        c = 2.0 * k + alpha + beta
        a1 = 2.0 * (k + 1.0) * (k + alpha + beta + 1.0) * c
        a2 = (c + 1.0) * (alpha * alpha - beta * beta)
        a3 = c * (c + 1.0) * (c + 2.0)
        a4 = 2.0 * (k + alpha) * (k + beta) * (c + 2.0)
        p0, p1 = p1, ((a2 + a3 * x) * p1 - a4 * p0) / a1
    return p1


def _djacobi(n: int, alpha: float, beta: float, x):
    if n == 0:
        return np.zeros_like(np.asarray(x, dtype=np.float64))
    return 0.5 * (n + alpha + beta + 1.0) * _jacobi(n - 1, alpha + 1.0, beta + 1.0, x)


def _scaled_legendre(n: int, xi, t):
    """Q_k(xi, t) = P_k(xi / t) t^k for k = 0..n with its partial derivatives (lists of arrays)."""
    Q = [np.ones_like(xi)]
    Qx = [np.zeros_like(xi)]
    Qt = [np.zeros_like(xi)]
    if n >= 1:
        Q.append(xi.copy())
        Qx.append(np.ones_like(xi))
        Qt.append(np.zeros_like(xi))
    for k in range(1, n):
        a, b = (2.0 * k + 1.0) / (k + 1.0), k / (k + 1.0)
        Q.append(a * xi * Q[k] - b * t * t * Q[k - 1])
        Qx.append(a * (Q[k] + xi * Qx[k]) - b * t * t * Qx[k - 1])
        Qt.append(a * xi * Qt[k] - b * (2.0 * t * Q[k - 1] + t * t * Qt[k - 1]))
    return Q, Qx, Qt


def _simplex(deg: int, r, s, grad: bool):
    r, s = np.asarray(r, dtype=np.float64), np.asarray(s, dtype=np.float64)
    xi, t = r + 0.5 * (1.0 + s), 0.5 * (1.0 - s)
    Q, Qx, Qt = _scaled_legendre(deg, xi, t)
    V, Vr, Vs = [], [], []
    for i in range(deg + 1):
        for j in range(deg + 1 - i):
            c = np.sqrt((2.0 * i + 1.0) * (i + j + 1.0) / 2.0)
            Pj = _jacobi(j, 2.0 * i + 1.0, 0.0, s)
            V.append(c * Q[i] * Pj)
            if grad:
                dPj = _djacobi(j, 2.0 * i + 1.0, 0.0, s)
                Vr.append(c * Qx[i] * Pj)
                Vs.append(c * ((0.5 * Qx[i] - 0.5 * Qt[i]) * Pj + Q[i] * dPj))
    if grad:
        return np.stack(Vr, axis=-1), np.stack(Vs, axis=-1)
    return np.stack(V, axis=-1)


def simplex_vandermonde(deg: int, r, s):
    """vandermonde_matrix(Tri, N, r, s): V[k, m] = psi_m(r_k, s_k), modes ordered (i, j), j fastest."""
    return _simplex(deg, r, s, False)


def dsimplex_vandermonde(deg: int, r, s):
    """∂vandermonde_matrix(Tri, N, r, s) -> (Vr, Vs)."""
    return _simplex(deg, r, s, True)


# ------------------------------------------------------------------------------------------------
# the reference's exported names for the same objects (src/FluxReconstruction.jl:19-51), on top of the
# constructions above: setup-time helpers, nothing here runs per stage
class AbstractElementShape:
    """src/data.jl:4-11: tags the reference dispatches ``vandermonde_matrix(Tri, N, r, s)`` etc. on."""


class Line(AbstractElementShape):
    pass


class Quad(AbstractElementShape):
    pass


class Tri(AbstractElementShape):
    pass


class Hex(AbstractElementShape):
    pass


class Wed(AbstractElementShape):
    pass


class Pyr(AbstractElementShape):
    pass


class Tet(AbstractElementShape):
    pass


def _jacobi_norm(n: int, alpha: float, beta: float) -> float:
    """gamma_n = int_{-1}^{1} (1-x)^alpha (1+x)^beta P_n^2 dx of the classical polynomial."""
    from math import exp, lgamma, log

    ab = alpha + beta
    lg = (ab + 1.0) * log(2.0) - log(2.0 * n + ab + 1.0) + lgamma(n + alpha + 1.0) + lgamma(n + beta + 1.0) \
        - lgamma(n + ab + 1.0) - lgamma(n + 1.0)
    return exp(lg)


def JacobiP(x, alpha, beta, N):
    """JacobiP(x, alpha, beta, N) (src/Polynomial/poly_jacobi.jl:7-84): the *orthonormal* Jacobi polynomial
    of degree N at x (scalar or array) -- the classical one divided by sqrt(gamma_N)."""
    return _jacobi(int(N), float(alpha), float(beta), x) / np.sqrt(_jacobi_norm(int(N), float(alpha), float(beta)))


def dJacobiP(r, alpha, beta, N):
    """∂JacobiP (poly_jacobi.jl:88-110): sqrt(N (N + alpha + beta + 1)) JacobiP(r, alpha+1, beta+1, N-1)."""
    r = np.asarray(r, dtype=np.float64)
    if N == 0:
        return np.zeros_like(r)
    return np.sqrt(N * (N + alpha + beta + 1.0)) * JacobiP(r, alpha + 1.0, beta + 1.0, N - 1)


def simplex_basis(a, b, i, j):
    """simplex_basis(a, b, i, j) (src/Transform/transform_triangle.jl:8-20): mode (i, j) of the orthonormal
    triangle basis in collapsed coordinates (a, b) = rs_ab(r, s)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.sqrt(2.0) * JacobiP(a, 0, 0, i) * JacobiP(b, 2 * i + 1, 0, j) * (1.0 - b) ** i


def dsimplex_basis(a, b, i, j):
    """∂simplex_basis(a, b, id, jd) -> (d/dr, d/ds) (transform_triangle.jl:22-69)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fa, dfa = JacobiP(a, 0, 0, i), dJacobiP(a, 0, 0, i)
    gb, dgb = JacobiP(b, 2 * i + 1, 0, j), dJacobiP(b, 2 * i + 1, 0, j)
    h = 0.5 * (1.0 - b)
    hp = h ** (i - 1) if i > 0 else 1.0
    dr = dfa * gb * hp
    ds = dfa * gb * 0.5 * (1.0 + a) * hp
    tmp = dgb * h**i
    if i > 0:
        tmp = tmp - 0.5 * i * gb * h ** (i - 1)
    ds = ds + fa * tmp
    return 2.0 ** (i + 0.5) * dr, 2.0 ** (i + 0.5) * ds


def correction_field(N, V=None):
    """correction_field(N, V) (src/Polynomial/poly_triangle.jl:1-28): phi[f, j, i] = sum_m wf[f, j]
    psi_m(x_fj) psi_m(x_i) on the WSJ solution points (the argument V is recomputed there as well)."""
    pl, _ = tri_quadrature(N)
    pf, wf = triface_quadrature(N)
    psif = simplex_vandermonde(N, pf[:, :, 0], pf[:, :, 1])
    Vl = simplex_vandermonde(N, pl[:, 0], pl[:, 1])
    return np.einsum("fj,fjm,im->fji", wf, psif, Vl)


def global_sp_tri(points, cellid, N):
    """global_sp(points, cellid, N) for triangles (src/Geometry/geo_points.jl:56-70): [ncell, np, 2];
    cellid 0-based."""
    pl, _ = tri_quadrature(N)
    xy, cid = np.asarray(points, dtype=np.float64)[:, :2], np.asarray(cellid)
    v1, v2, v3 = (xy[cid[:, k]][:, None, :] for k in range(3))
    return rs_xy(pl[None, :, 0], pl[None, :, 1], v1, v2, v3)


def global_fp(points, cellid, N):
    """global_fp(points, cellid, N) (geo_points.jl:101-115): flux points [ncell, 3, N+1, 2]; cellid 0-based."""
    pf, _ = triface_quadrature(N)
    xy, cid = np.asarray(points, dtype=np.float64)[:, :2], np.asarray(cellid)
    v1, v2, v3 = (xy[cid[:, k]][:, None, None, :] for k in range(3))
    return rs_xy(pf[None, :, :, 0], pf[None, :, :, 1], v1, v2, v3)


def neighbor_fpidx(IDs, ps, fpg=None):
    """neighbor_fpidx((cell, face, point), ps, fpg) (src/Geometry/geo_neighbor.jl:8-60): the flux point of the
    neighbouring cell that coincides with mine, as (cell, face, point) -- 0-based here, (-1, -1, -1) on a
    boundary face (the reference returns (neighbor_cid <= 0, -1, -1)).  The reference searches the neighbour's
    flux-point coordinates for an exact match; the space has the answer from the edge orientation already
    (``ps.fpn``), which is what the kernels consume; ``fpg`` is accepted for signature parity."""
    c, f, k = IDs
    return tuple(int(v) for v in ps.fpn[c, f, k])


def rs_xy(r, s, v1, v2, v3):
    """Reference -> physical triangle (geo_transform.jl:16-31); broadcasts over leading axes."""
    r, s = np.asarray(r)[..., None], np.asarray(s)[..., None]
    return -0.5 * (r + s) * v1 + 0.5 * (r + 1.0) * v2 + 0.5 * (s + 1.0) * v3


# ------------------------------------------------------------------------------------------------
# Gmsh reader
_NODES_OF = {1: 2, 2: 3, 3: 4, 15: 1}  # line, triangle, quadrangle, point (first-order elements)
_NAME_OF = {1: "line", 2: "triangle", 3: "quad", 15: "vertex"}


def _sections(path):
    sec, name, buf = {}, None, []
    with open(path, "r") as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line.startswith("$End"):
                sec[name] = buf
                name, buf = None, []
            elif line.startswith("$"):
                name, buf = line[1:], []
            elif name is not None:
                buf.append(line)
    if name is not None:
        raise ValueError(f"{path}: section ${name} is not closed")
    return sec


def read_msh(path):
    """Reads a Gmsh .msh file (format 4.1 or 2.2, ASCII, first-order elements).

    Returns ``points`` [npoint, 3] in file order and ``cells``: {"triangle": [n, 3], "line": [n, 2],
    "quad": [n, 4], "vertex": [n, 1]} with 0-based indices into ``points`` -- the two things
    KitBase's ``read_mesh`` takes from meshio."""
    sec = _sections(path)
    if "MeshFormat" not in sec or "Nodes" not in sec or "Elements" not in sec:
        raise ValueError(f"{path}: not a Gmsh mesh ($MeshFormat / $Nodes / $Elements missing)")
    ver, ftype = sec["MeshFormat"][0].split()[:2]
    if int(ftype) != 0:
        raise ValueError(f"{path}: binary .msh files are not supported")
    major = int(float(ver))
    nodes, elems = sec["Nodes"], sec["Elements"]
    tags, xyz, found = [], [], {}
    if major == 4:
        nblk, nnode = (int(x) for x in nodes[0].split()[:2])
        pos = 1
        for _ in range(nblk):
            _, _, parametric, nb = (int(x) for x in nodes[pos].split())
            pos += 1
            tags += [int(x) for x in nodes[pos:pos + nb]]
            pos += nb
            xyz += [[float(x) for x in ln.split()[:3]] for ln in nodes[pos:pos + nb]]
            pos += nb
        if len(tags) != nnode:
            raise ValueError(f"{path}: $Nodes announces {nnode} nodes, holds {len(tags)}")
        index = {t: k for k, t in enumerate(tags)}
        nblk = int(elems[0].split()[0])
        pos = 1
        for _ in range(nblk):
            _, _, etype, nb = (int(x) for x in elems[pos].split())
            pos += 1
            if etype not in _NODES_OF:
                raise ValueError(f"{path}: element type {etype} is not a first-order point/line/triangle/quad")
            nn = _NODES_OF[etype]
            rows = [[index[int(x)] for x in ln.split()[1:1 + nn]] for ln in elems[pos:pos + nb]]
            found.setdefault(_NAME_OF[etype], []).extend(rows)
            pos += nb
    elif major == 2:
        nnode = int(nodes[0])
        for ln in nodes[1:1 + nnode]:
            f = ln.split()
            tags.append(int(f[0]))
            xyz.append([float(x) for x in f[1:4]])
        index = {t: k for k, t in enumerate(tags)}
        for ln in elems[1:1 + int(elems[0])]:
            f = [int(x) for x in ln.split()]
            etype, ntag = f[1], f[2]
            if etype not in _NODES_OF:
                raise ValueError(f"{path}: element type {etype} is not a first-order point/line/triangle/quad")
            found.setdefault(_NAME_OF[etype], []).append([index[t] for t in f[3 + ntag:3 + ntag + _NODES_OF[etype]]])
    else:
        raise ValueError(f"{path}: unsupported MSH version {ver}")
    points = np.array(xyz, dtype=np.float64).reshape(-1, 3)
    cells = {k: np.array(v, dtype=np.int64) for k, v in found.items()}
    return points, cells


# SU2 native mesh format (the reference ships assets/linesource.su2; KitBase reads it through meshio):
# NDIME / NELEM + connectivity rows "vtk_type n0 n1 ... [index]" / NPOIN + coordinate rows "x y [z] [index]" /
# NMARK, then per marker MARKER_TAG, MARKER_ELEMS and its boundary elements.  0-based node indices.
_SU2_NODES = {3: ("line", 2), 5: ("triangle", 3), 9: ("quad", 4)}


def read_su2(path):
    """Reads an ASCII .su2 mesh with first-order lines / triangles / quadrilaterals; same return value as
    ``read_msh`` (the marker elements of every MARKER_TAG are the ``line`` cells)."""
    with open(path, "r") as fh:
        lines = [ln.split("%")[0].strip() for ln in fh]
    lines = [ln for ln in lines if ln]
    ndim, pos, found, xyz = None, 0, {}, None

    def key(ln):
        return ln.split("=")[0].strip().upper() if "=" in ln else None

    def val(ln):
        return ln.split("=")[1].strip()

    def rows(start, n, sink):
        for ln in lines[start:start + n]:
            f = ln.split()
            et = int(f[0])
            if et not in _SU2_NODES:
                raise ValueError(f"{path}: element type {et} is not a first-order line/triangle/quad")
            name, nn = _SU2_NODES[et]
            sink.setdefault(name, []).append([int(x) for x in f[1:1 + nn]])

    while pos < len(lines):
        k = key(lines[pos])
        if k == "NDIME":
            ndim = int(val(lines[pos]))
            pos += 1
        elif k == "NELEM":
            n = int(val(lines[pos]))
            rows(pos + 1, n, found)
            pos += 1 + n
        elif k == "NPOIN":
            n = int(val(lines[pos]).split()[0])
            if ndim is None:
                raise ValueError(f"{path}: NPOIN before NDIME")
            xyz = np.zeros((n, 3))
            for q, ln in enumerate(lines[pos + 1:pos + 1 + n]):
                xyz[q, :ndim] = [float(x) for x in ln.split()[:ndim]]
            pos += 1 + n
        elif k == "MARKER_ELEMS":
            n = int(val(lines[pos]))
            rows(pos + 1, n, found)
            pos += 1 + n
        else:  # NMARK, MARKER_TAG and anything this reader has no use for
            pos += 1
    if xyz is None or not found:
        raise ValueError(f"{path}: not an SU2 mesh (NPOIN / NELEM missing)")
    return xyz, {k: np.array(v, dtype=np.int64) for k, v in found.items()}


def read_mesh(path):
    """KitBase.read_mesh(file): by extension, ``.msh`` (Gmsh 4.1 / 2.2) or ``.su2``."""
    return read_su2(path) if str(path).lower().endswith(".su2") else read_msh(path)


# ------------------------------------------------------------------------------------------------
class UnstructPSpace:
    """The mesh fields of KitBase's ``UnstructPSpace`` the FR path reads (struct.jl:270-286), for a
    triangle mesh: ``UnstructPSpace(file)`` or ``UnstructPSpace(points, cellid)``.

    points [npoint, 2|3]; cellid [ncell, 3]; cellNeighbors / cellFaces [ncell, 3] (local face j joins
    vertices j and j+1); facePoints / faceCells [nface, 2]; cellCenter, cellArea, cellNormals
    [ncell, 3, 2] (outward, unit); faceCenter, faceArea (edge length); cellType / faceType
    (0 interior, 1 boundary)."""

    def __init__(self, points, cellid=None):
        if cellid is None:
            points, cells = read_mesh(points)
            if "triangle" not in cells:
                raise ValueError("the mesh holds no triangles")
            cellid = cells["triangle"]
            self.cells = cells
        else:
            self.cells = {"triangle": np.asarray(cellid, dtype=np.int64)}
        self.points = np.asarray(points, dtype=np.float64)
        self.cellid = np.asarray(cellid, dtype=np.int64)
        nc = self.cellid.shape[0]
        if self.cellid.ndim != 2 or self.cellid.shape[1] != 3:
            raise ValueError("cellid must be [ncell, 3]")
        xy = self.points[:, :2]
        v = xy[self.cellid]  # [ncell, 3, 2]

        # edges: local face j of a cell joins vertices j, j+1; one global face per unordered pair
        a, b = self.cellid, np.roll(self.cellid, -1, axis=1)
        lo, hi = np.minimum(a, b).ravel(), np.maximum(a, b).ravel()
        key = lo * (self.points.shape[0] + 1) + hi
        uniq, first, inv, cnt = np.unique(key, return_index=True, return_inverse=True, return_counts=True)
        if cnt.max() > 2:
            raise ValueError("non-manifold mesh: an edge is shared by more than two triangles")
        nf = uniq.shape[0]
        self.cellFaces = inv.reshape(nc, 3)
        self.facePoints = np.stack([lo[first], hi[first]], axis=1)
        owner = np.repeat(np.arange(nc), 3)
        order = np.argsort(inv, kind="stable")  # half-edges grouped by face, cell order kept
        start = np.searchsorted(inv[order], np.arange(nf))
        self.faceCells = -np.ones((nf, 2), dtype=np.int64)
        self.faceCells[:, 0] = owner[order[start]]
        two = cnt == 2
        self.faceCells[two, 1] = owner[order[start[two] + 1]]
        fc = self.faceCells[self.cellFaces]  # [ncell, 3, 2]
        me = np.arange(nc)[:, None]
        self.cellNeighbors = np.where(fc[:, :, 0] == me, fc[:, :, 1], fc[:, :, 0])
        # a face both of whose sides are the same cell cannot happen for triangles

        self.cellCenter = v.mean(axis=1)
        e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
        self.cellArea = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
        edge = np.roll(v, -1, axis=1) - v
        n = np.stack([edge[:, :, 1], -edge[:, :, 0]], axis=2)
        mid = 0.5 * (np.roll(v, -1, axis=1) + v)
        outward = np.sum(n * (mid - self.cellCenter[:, None, :]), axis=2) >= 0.0
        n = np.where(outward[:, :, None], n, -n)
        self.cellNormals = n / np.linalg.norm(n, axis=2, keepdims=True)
        pf = xy[self.facePoints]
        self.faceCenter = pf.mean(axis=1)
        self.faceArea = np.linalg.norm(pf[:, 1] - pf[:, 0], axis=1)
        self.faceType = (self.faceCells[:, 1] < 0).astype(np.int64)
        self.cellType = (self.cellNeighbors < 0).any(axis=1).astype(np.int64)


class UnstructFRPSpace:
    """``TriFRPSpace(file, deg)`` (struct.jl:305-352).  ``file`` is a Gmsh mesh, or pass an
    ``UnstructPSpace`` / ``(points, cellid)``.  Field names follow the reference with ASCII spellings:
    ``dl`` = ∂l [np, np, 2], ``phi`` = ϕ [3, deg+1, np], ``psif`` = ψf, ``lf`` [3, deg+1, np],
    ``J`` [ncell, 2, 2] = [xr xs; yr ys] per cell, ``xpg`` [ncell, np, 2], ``xfg`` [ncell, 3, deg+1, 2],
    ``fpn`` [ncell, 3, deg+1, 3] = (cell, face, point) of the coincident flux point, 0-based, -1 on
    boundary faces.  The mesh fields (``points``, ``cellid``, ``cellType``, ``cellNormals``, ...) are
    reachable directly, like the reference's property forwarding to ``base``."""

    def __init__(self, mesh, deg: int, cellid=None):
        if isinstance(mesh, UnstructPSpace):
            base = mesh
        elif cellid is not None:
            base = UnstructPSpace(mesh, cellid)
        elif isinstance(mesh, (tuple, list)) and len(mesh) == 2:
            base = UnstructPSpace(mesh[0], mesh[1])
        else:
            base = UnstructPSpace(mesh)
        deg = int(deg)
        self.base, self.deg = base, deg
        self.np = (deg + 1) * (deg + 2) // 2
        self.xpl, self.wp = tri_quadrature(deg)
        self.V = simplex_vandermonde(deg, self.xpl[:, 0], self.xpl[:, 1])
        self.Vr, self.Vs = dsimplex_vandermonde(deg, self.xpl[:, 0], self.xpl[:, 1])
        # l_k(x) = sum_m (V^-1)[m, k] psi_m(x)  =>  values of all l_k at points with basis rows B: B V^-1
        iV = np.linalg.inv(self.V)
        self.dl = np.stack([self.Vr @ iV, self.Vs @ iV], axis=2)  # ∂l[i, k, d] = ∂_d l_k at point i
        self.xfl, self.wf = triface_quadrature(deg)
        self.psif = simplex_vandermonde(deg, self.xfl[:, :, 0], self.xfl[:, :, 1])  # [3, deg+1, np]
        self.lf = self.psif @ iV
        # correction field: phi[f, j, i] = sum_m wf[f, j] psi_m(x_fj) psi_m(x_i)
        self.phi = np.einsum("fj,fjm,im->fji", self.wf, self.psif, self.V)

        cid, xy = base.cellid, base.points[:, :2]
        v1, v2, v3 = (xy[cid[:, k]] for k in range(3))
        self.J = np.stack([0.5 * (v2 - v1), 0.5 * (v3 - v1)], axis=2)
        b = lambda a: a[:, None, :]  # noqa: E731
        self.xpg = rs_xy(self.xpl[None, :, 0], self.xpl[None, :, 1], b(v1), b(v2), b(v3))
        bb = lambda a: a[:, None, None, :]  # noqa: E731
        self.xfg = rs_xy(self.xfl[None, :, :, 0], self.xfl[None, :, :, 1], bb(v1), bb(v2), bb(v3))

        # flux-point connectivity: my face j is walked from vertex j to j+1; the neighbour holds the
        # same edge as its face nj, walked the other way round unless its orientation differs
        nc = cid.shape[0]
        nb = base.cellNeighbors
        has = nb >= 0
        nbs = np.where(has, nb, 0)
        myface = base.cellFaces
        nface = (base.cellFaces[nbs] == myface[:, :, None]).argmax(axis=2)  # [ncell, 3]: its local face
        first = cid  # first vertex of my face j
        nfirst = np.take_along_axis(cid[nbs], nface[:, :, None], axis=2)[:, :, 0]
        same_dir = nfirst == first
        k = np.arange(deg + 1)[None, None, :]
        nk = np.where(same_dir[:, :, None], k, deg - k)
        fpn = -np.ones((nc, 3, deg + 1, 3), dtype=np.int64)
        fpn[..., 0] = np.where(has[:, :, None], nb[:, :, None], -1)
        fpn[..., 1] = np.where(has[:, :, None], nface[:, :, None], -1)
        fpn[..., 2] = np.where(has[:, :, None], nk, -1)
        self.fpn = fpn

    def __getattr__(self, name):  # ps.cellid, ps.cellType, ... forward to the mesh
        base = self.__dict__.get("base")
        if base is not None and hasattr(base, name):
            return getattr(base, name)
        raise AttributeError(name)


TriFRPSpace = UnstructFRPSpace
