"""CPU restatement of the reference's TRIANGLE operator builders -- test infrastructure, like the
rest of oracle/ (only tests/ import it).

    rs_ab, xy_rs                     src/Geometry/geo_transform.jl:65-125
    simplex_basis, dsimplex_basis    src/Transform/transform_triangle.jl:8-68   (Dubiner basis)
    vandermonde_tri, dvandermonde_tri  src/Transform/transform.jl:36-49, 94-110
    tri_quadrature                   src/Quadrature/quadrature.jl:14-71  (Williams-Shunn-Jameson points
                                     from the reference's own qpmin.py; here from the committed fixture
                                     tests/golden/tri_golden.npz, made by tests/golden/make_tri_golden.py)
    triface_quadrature               src/Quadrature/quadrature.jl:86-106
    dlagrange_tri                    src/Polynomial/poly_lagrange.jl:82-92
    correction_field                 src/Polynomial/poly_triangle.jl:1-28
    tri_operators                    what TriFRPSpace precomputes, src/struct.jl:305-352 (V, Vr, Vs, dl, lf, phi)

PINNED: unlike the rest of the oracle this part has reference golden vectors --
dev/check_phi.jl:60-89 (py_V, py_Vf, phifj_ref for the degree-2 triangle, asserted by the reference at
:110, :117, :125).  tests/test_oracle_tri.py checks this module against them.
"""
from __future__ import annotations

import os

import numpy as np

from fr_oracle import djacobi_p, gausslegendre, jacobi_p

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "tri_golden.npz")


def rs_ab(r, s):
    r, s = np.asarray(r, dtype=np.float64), np.asarray(s, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(s != 1.0, 2.0 * (1.0 + r) / (1.0 - s) - 1.0, -1.0)
    return a, 1.0 * s


def xy_rs(x, y):
    """equilateral -> right triangle"""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    L1 = (np.sqrt(3.0) * y + 1.0) / 3.0
    L2 = (-3.0 * x - np.sqrt(3.0) * y + 2.0) / 6.0
    L3 = (3.0 * x - np.sqrt(3.0) * y + 2.0) / 6.0
    return -L2 + L3 - L1, -L2 - L3 + L1


def simplex_basis(a, b, i, j):
    h1 = jacobi_p(a, 0, 0, i)
    h2 = jacobi_p(b, 2 * i + 1, 0, j)
    return np.sqrt(2.0) * h1 * h2 * (1 - np.asarray(b)) ** i


def dsimplex_basis(a, b, i, j):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fa, dfa = jacobi_p(a, 0, 0, i), djacobi_p(a, 0, 0, i)
    gb, dgb = jacobi_p(b, 2 * i + 1, 0, j), djacobi_p(b, 2 * i + 1, 0, j)
    dr = dfa * gb
    if i > 0:
        dr = dr * (0.5 * (1.0 - b)) ** (i - 1)
    ds = dfa * (gb * (0.5 * (1.0 + a)))
    if i > 0:
        ds = ds * (0.5 * (1.0 - b)) ** (i - 1)
    tmp = dgb * (0.5 * (1.0 - b)) ** i
    if i > 0:
        tmp = tmp - 0.5 * i * gb * (0.5 * (1.0 - b)) ** (i - 1)
    ds = ds + fa * tmp
    return dr * 2.0 ** (i + 0.5), ds * 2.0 ** (i + 0.5)


def vandermonde_tri(N, r, s):
    a, b = rs_ab(r, s)
    cols = [simplex_basis(a, b, i, j) for i in range(N + 1) for j in range(N + 1 - i)]
    return np.stack(cols, axis=1)


def dvandermonde_tri(N, r, s):
    a, b = rs_ab(r, s)
    d = [dsimplex_basis(a, b, i, j) for i in range(N + 1) for j in range(N + 1 - i)]
    return np.stack([x[0] for x in d], axis=1), np.stack([x[1] for x in d], axis=1)


def tri_quadrature(deg, vertices=((-1.0, -1 / np.sqrt(3)), (1.0, -1 / np.sqrt(3)), (0.0, 2 / np.sqrt(3))),
                   transform=True):
    g = np.load(_GOLDEN)
    pts0, w = g[f"wsj{deg + 1}_points"], g[f"wsj{deg + 1}_weights"]
    p1, p2, p3 = (np.asarray(v, dtype=np.float64) for v in vertices)
    a, b, c = np.linalg.norm(p2 - p3), np.linalg.norm(p3 - p1), np.linalg.norm(p2 - p1)
    pts = np.zeros((pts0.shape[1], 2))
    for i in range(pts.shape[0]):
        x, y, z = pts0[:, i]  # trilinear -> cartesian
        pts[i] = (a * x * p1 + b * y * p2 + c * z * p3) / (a * x + b * y + c * z)
    if transform:
        r, s = xy_rs(pts[:, 0], pts[:, 1])
        pts = np.stack([r, s], axis=1)
    return pts, w.copy()


def triface_quadrature(N):
    p0, w0 = gausslegendre(N + 1)
    pf = np.zeros((3, N + 1, 2))
    pf[0, :, 0], pf[1, :, 0], pf[2, :, 0] = p0, p0[::-1], -1.0
    pf[0, :, 1], pf[1, :, 1], pf[2, :, 1] = -1.0, p0, p0[::-1]
    wf = np.stack([w0 * 1.0, w0 * np.sqrt(2.0), w0 * 1.0])
    return pf, wf


def dlagrange_tri(V, Vr, Vs):
    Np = V.shape[0]
    dl = np.zeros((Np, Np, 2))
    for i in range(Np):
        dl[i, :, 0] = np.linalg.solve(V.T, Vr[i])
        dl[i, :, 1] = np.linalg.solve(V.T, Vs[i])
    return dl


def correction_field(N, V=None):
    pl, _ = tri_quadrature(N)
    pf, wf = triface_quadrature(N)
    psi = np.stack([vandermonde_tri(N, pf[i, :, 0], pf[i, :, 1]) for i in range(3)])  # [3, N+1, Np]
    sigma = wf[:, :, None] * psi
    V = vandermonde_tri(N, pl[:, 0], pl[:, 1])
    return np.einsum("fjk,ik->fji", sigma, V)  # phi[f, j, i] = sum_k sigma[f,j,k] V[i,k]


def tri_operators(deg):
    """The constant arrays of TriFRPSpace (struct.jl:305-352)."""
    xpl, wp = tri_quadrature(deg)
    V = vandermonde_tri(deg, xpl[:, 0], xpl[:, 1])
    Vr, Vs = dvandermonde_tri(deg, xpl[:, 0], xpl[:, 1])
    xfl, wf = triface_quadrature(deg)
    psi = np.stack([vandermonde_tri(deg, xfl[i, :, 0], xfl[i, :, 1]) for i in range(3)])
    lf = np.stack([[np.linalg.solve(V.T, psi[i, j]) for j in range(deg + 1)] for i in range(3)])
    return {"xpl": xpl, "wp": wp, "V": V, "Vr": Vr, "Vs": Vs, "dl": dlagrange_tri(V, Vr, Vs), "xfl": xfl, "wf": wf,
            "psif": psi, "lf": lf, "phi": correction_field(deg, V)}


# ---------------------------------------------------------------- meshes, spaces, the Euler residual
def rs_xy(r, s, v1, v2, v3):
    """src/Geometry/geo_transform.jl:16-31 (triangle)."""
    return -(r + s) / 2 * np.asarray(v1) + (r + 1) / 2 * np.asarray(v2) + (s + 1) / 2 * np.asarray(v3)


def tri_mesh_rect(nx, ny, x0=0.0, x1=1.0, y0=0.0, y1=1.0, jitter=0.0, seed=0):
    """A rectangle cut into 2*nx*ny counter-clockwise triangles (test meshes; the reference reads
    Gmsh files through KitBase/meshio).  Returns points [npt, 2], cellid [ncell, 3] (0-based)."""
    xs, ys = np.linspace(x0, x1, nx + 1), np.linspace(y0, y1, ny + 1)
    pts = np.array([[x, y] for y in ys for x in xs])
    if jitter > 0:
        rng = np.random.default_rng(seed)
        inner = [(j * (nx + 1) + i) for j in range(1, ny) for i in range(1, nx)]
        pts[inner] += jitter * min((x1 - x0) / nx, (y1 - y0) / ny) * (rng.random((len(inner), 2)) - 0.5)
    cells = []
    for j in range(ny):
        for i in range(nx):
            a, b = j * (nx + 1) + i, j * (nx + 1) + i + 1
            c, d = (j + 1) * (nx + 1) + i + 1, (j + 1) * (nx + 1) + i
            cells += [[a, b, c], [a, c, d]] if (i + j) % 2 == 0 else [[a, b, d], [b, c, d]]
    return pts, np.array(cells, dtype=np.int64)


def tri_space(points, cellid, deg):
    """What TriFRPSpace holds (struct.jl:305-352) plus the KitBase mesh fields dev/sod.jl uses:
    J (geo_jacobi.jl:32-43), xpg / xfg (geo_points.jl global_sp / global_fp), fpn
    (geo_neighbor.jl:8-60, 0-based here, -1 = no neighbour), cell normals ([KB] outward unit normal of
    face j = vertices j -> j+1) and cellType (0 interior, 1 boundary)."""
    ops = tri_operators(deg)
    ncell = cellid.shape[0]
    v = points[cellid]  # [ncell, 3, 2]
    J = np.zeros((ncell, 2, 2))
    J[:, :, 0] = (v[:, 1] - v[:, 0]) / 2  # [xr; yr]
    J[:, :, 1] = (v[:, 2] - v[:, 0]) / 2  # [xs; ys]
    xpg = np.stack([[rs_xy(r, s, v[i, 0], v[i, 1], v[i, 2]) for r, s in ops["xpl"]] for i in range(ncell)])
    xfg = np.stack([[[rs_xy(ops["xfl"][j, k, 0], ops["xfl"][j, k, 1], v[i, 0], v[i, 1], v[i, 2])
                      for k in range(deg + 1)] for j in range(3)] for i in range(ncell)])
    normals = np.zeros((ncell, 3, 2))
    for j in range(3):
        e = v[:, (j + 1) % 3] - v[:, j]
        normals[:, j, 0], normals[:, j, 1] = e[:, 1], -e[:, 0]
    normals /= np.linalg.norm(normals, axis=2, keepdims=True)
    edge = {}
    for i in range(ncell):
        for j in range(3):
            edge.setdefault(tuple(sorted((cellid[i, j], cellid[i, (j + 1) % 3]))), []).append((i, j))
    fpn = -np.ones((ncell, 3, deg + 1, 3), dtype=np.int64)
    for pair in edge.values():
        if len(pair) != 2:
            continue
        for (i, j), (ni, nj) in (pair, pair[::-1]):
            for k in range(deg + 1):
                d = np.abs(xfg[ni, nj] - xfg[i, j, k]).sum(axis=1)
                nk = int(np.argmin(d))  # the reference matches coordinates with ==
                assert d[nk] < 1e-12
                fpn[i, j, k] = (ni, nj, nk)
    cell_type = (fpn[:, :, 0, 0] < 0).any(axis=1).astype(np.int64)
    ops.update({"points": points, "cellid": cellid, "deg": deg, "np": ops["V"].shape[0], "J": J, "xpg": xpg,
                "xfg": xfg, "normals": normals, "fpn": fpn, "cellType": cell_type})
    return ops


_NREF = np.array([[0.0, -1.0], [1 / np.sqrt(2), 1 / np.sqrt(2)], [-1.0, 0.0]])


def rhs_tri_euler(u, sp, gamma):
    """dudt! of dev/sod.jl:31-123 (the same loop is dev/euler.jl, dev/euler_naca.jl).  u[ncell, np, 4]."""
    from fr_oracle import euler_flux, flux_hll, global_frame, local_frame

    ncell, nsp, _ = u.shape
    deg = sp["deg"]
    iJ = np.linalg.inv(sp["J"])
    F, G = euler_flux(u, gamma)
    f = np.einsum("iab,ipkb->ipka", iJ, np.stack([F, G], axis=-1))  # f[i,p,k,:] = inv(J[i]) * [F_k, G_k]
    lf, dl, phi, nrm = sp["lf"], sp["dl"], sp["phi"], sp["normals"]
    u_face = np.einsum("ipl,jkp->ijkl", u, lf)
    f_face = np.einsum("ipla,jkp->ijkla", f, lf)
    fn_face = np.einsum("ijkla,ja->ijkl", f_face, _NREF)
    fn_int = np.zeros((ncell, 3, deg + 1, 4))
    for i in range(ncell):
        for j in range(3):
            c, s = nrm[i, j]
            for k in range(deg + 1):
                uL = local_frame(u_face[i, j, k], c, s)
                ni, nj, nk = sp["fpn"][i, j, k]
                fl = np.zeros(4)
                if ni >= 0:
                    fl = flux_hll(uL, local_frame(u_face[ni, nj, nk], c, s), gamma, 1.0)
                elif sp["cellType"][i] == 2:
                    w = u_face[i, j, k].copy()
                    w[2] = -w[2]
                    fl = flux_hll(uL, local_frame(w, c, s), gamma, 1.0)
                fg = global_frame(fl, c, s)
                fxy = fg[:, None] * nrm[i, j][None, :]
                frs = fxy @ iJ[i].T  # inv(J[i]) * fwn_xy[idx, :]
                fn_int[i, j, k] = frs @ _NREF[j]
    rhs1 = -np.einsum("iqk,pq->ipk", f[..., 0], dl[:, :, 0]) - np.einsum("iqk,pq->ipk", f[..., 1], dl[:, :, 1])
    rhs2 = -np.einsum("ijkl,jkp->ipl", fn_int - fn_face, phi)
    du = np.zeros_like(u)
    act = np.isin(sp["cellType"], (0, 2))
    du[act] = rhs1[act] + rhs2[act]
    return du
