"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the curvilinear structured-quadrilateral Euler residual --
SURVEY 8f-2.  Only tests/ may import this file; nothing in the product path does.

What is restated, loop for loop, from /root/reference (citations relative to it):

* ``FRPSpace2D(base, deg)`` on a mesh given by its cell vertices: ``J = rs_jacobi(r, base.vertices)``
  (src/struct.jl:135, src/Geometry/geo_jacobi.jl:77-108), the point-wise inverse ``iJ`` (struct.jl:137-142)
  and the flux-point Jacobians ``Ji`` (struct.jl:145-158);
* the meshes of the two scratch scripts that use those metrics: the 45-degree parallelogram of
  dev/parallelogram.jl:36-75 and ``KitBase.CSpace2D(r0, r1, nr, th0, th1, nth, ngr, ngth)`` of
  dev/cylinder2.jl:22 ([KB-recall]: KitBase 0.9 is not under /root/reference);
* the face normals ``n1`` / ``n2`` the scripts build by hand (parallelogram.jl:176-186,
  cylinder2.jl:39-49);
* ``dudt!`` of dev/parallelogram.jl:80-165 (correction factors from the solution-point ``iJ``) and of
  dev/cylinder2.jl:52-164 (correction factors from the flux-point ``Ji``, mirror wall on the inner face).

PARITY UNPINNED: the reference cannot run here (no Julia, no KitBase) and has no test on this path.
What pins the restatement (tests/test_oracle_curv.py): it reduces to the rectangular residual
(example/euler2d_wave.jl:35-107, fr_oracle.rhs_euler2d) on a rectangular mesh, it is equivariant under a
rigid rotation of the mesh and the velocity field, a uniform state is preserved on the parallelogram, and
it converges to the analytic flux divergence at the design order.

Reference quirks that are restated literally and can be switched off:

* both scripts index the y common flux with the *row* index in the correction step,
  ``fy_interaction[i, j, l, m]`` (parallelogram.jl:147-148, cylinder2.jl:157-158), where the rectangular
  scripts use the flux-point index ``k`` (euler2d_wave.jl:100-103) -- ``fy_index="l"`` is the scripts'
  form, ``"k"`` the consistent one.  With ``"l"`` the scheme is not conservative (sum of wp det(J) du over a
  periodic sheared mesh is O(0.1) instead of 1e-16, tests/test_oracle_curv.py) -- parallelogram.jl:12:
  "Instability is somehow detected for order larger than 2".
* ``Ji``: ``rs_jacobi(ri, si, vertices)`` with matrix arguments evaluates ``rs_jacobi(r[i], s[i], ...)``
  with the *linear* index ``i`` over the 4 faces for every flux point ``j`` (geo_jacobi.jl:93-94), and the
  face tables put 0.0 where -1.0 is meant (struct.jl:149,152).  ``flux_point_jacobi(..., literal=True)``
  reproduces that (one Jacobian per face, taken at ``(ri[face, 1], si[face, 1])``); ``literal=False``
  evaluates the Jacobian at the actual flux points.  Elements whose opposite sides are parallel
  (parallelograms) have a constant Jacobian, so the quirk is invisible there.
"""
from __future__ import annotations

import numpy as np

import fr_oracle as o

# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------


def rs_jacobi_point(r, s, v):
    """geo_jacobi.jl:77-88.  v[..., 4, 2] vertices (CCW from bottom-left) -> J[..., 2, 2] = [xr xs; yr ys]."""
    v = np.asarray(v, dtype=np.float64)
    Xr = (s - 1.0) * v[..., 0, :] / 4 + (1.0 - s) * v[..., 1, :] / 4 + (s + 1.0) * v[..., 2, :] / 4 - (s + 1.0) * v[..., 3, :] / 4
    Xs = (r - 1.0) * v[..., 0, :] / 4 - (r + 1.0) * v[..., 1, :] / 4 + (r + 1.0) * v[..., 2, :] / 4 + (1.0 - r) * v[..., 3, :] / 4
    J = np.empty(v.shape[:-2] + (2, 2))
    J[..., 0, 0] = Xr[..., 0]
    J[..., 0, 1] = Xs[..., 0]
    J[..., 1, 0] = Xr[..., 1]
    J[..., 1, 1] = Xs[..., 1]
    return J


def solution_point_jacobi(r, vertices):
    """struct.jl:135: J[i, j][k, l] = rs_jacobi(r[k], r[l], vertices[i, j]) (geo_jacobi.jl:102-108)."""
    nsp = len(r)
    J = np.empty(vertices.shape[:2] + (nsp, nsp, 2, 2))
    for k in range(nsp):
        for l in range(nsp):
            J[:, :, k, l] = rs_jacobi_point(r[k], r[l], vertices)
    return J


def inverse_jacobi(J):
    """struct.jl:137-142: iJ[i, j][k, l] = inv(J[i, j][k, l])."""
    det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    iJ = np.empty_like(J)
    iJ[..., 0, 0] = J[..., 1, 1] / det
    iJ[..., 0, 1] = -J[..., 0, 1] / det
    iJ[..., 1, 0] = -J[..., 1, 0] / det
    iJ[..., 1, 1] = J[..., 0, 0] / det
    return iJ


def flux_point_jacobi(r, vertices, literal=True):
    """struct.jl:145-158: Ji[i, j][face, pt].  literal=True restates the reference including its two
    quirks (module docstring); literal=False evaluates at the flux points of faces 1..4
    (s=-1, r=+1, s=+1, r=-1) in the point order of the traces (u_face[., ., face, pt, .])."""
    nsp = len(r)
    Ji = np.empty(vertices.shape[:2] + (4, nsp, 2, 2))
    if literal:
        ri = np.zeros((4, nsp))
        si = np.zeros((4, nsp))
        ri[0, :] = r
        ri[1, :] = 1.0
        ri[2, :] = r[::-1]
        ri[3, :] = 0.0
        si[0, :] = 0.0
        si[1, :] = r
        si[2, :] = 1.0
        si[3, :] = r[::-1]
        for face in range(4):
            # r[i], s[i] with the linear index i = face (column-major): element [face, 1]
            Jf = rs_jacobi_point(ri[face, 0], si[face, 0], vertices)
            for pt in range(nsp):
                Ji[:, :, face, pt] = Jf
    else:
        for pt in range(nsp):
            Ji[:, :, 0, pt] = rs_jacobi_point(r[pt], -1.0, vertices)
            Ji[:, :, 1, pt] = rs_jacobi_point(1.0, r[pt], vertices)
            Ji[:, :, 2, pt] = rs_jacobi_point(r[pt], 1.0, vertices)
            Ji[:, :, 3, pt] = rs_jacobi_point(-1.0, r[pt], vertices)
    return Ji


def global_sp(r, vertices):
    """struct.jl:161-175: bilinear map of the solution points, xpg[i, j, k, l, 1:2]."""
    nsp = len(r)
    xpg = np.empty(vertices.shape[:2] + (nsp, nsp, 2))
    for k in range(nsp):
        for l in range(nsp):
            xpg[:, :, k, l, :] = (
                (r[k] - 1.0) * (r[l] - 1.0) / 4 * vertices[:, :, 0, :]
                + (r[k] + 1.0) * (1.0 - r[l]) / 4 * vertices[:, :, 1, :]
                + (r[k] + 1.0) * (r[l] + 1.0) / 4 * vertices[:, :, 2, :]
                + (1.0 - r[k]) * (r[l] + 1.0) / 4 * vertices[:, :, 3, :]
            )
    return xpg


class CurvSpace2D:
    """FRPSpace2D(base, deg) for a base space that carries ``vertices[i, j, 4, 2]`` with one ghost ring
    (array index == reference index 0:nx+1).  Operators as in fr_oracle.FRPSpace2D."""

    def __init__(self, vertices, deg, literal_Ji=True):
        vertices = np.asarray(vertices, dtype=np.float64)
        self.vertices = vertices
        self.nx, self.ny = vertices.shape[0] - 2, vertices.shape[1] - 2
        self.deg = int(deg)
        r = o.legendre_point(deg)
        self.xpl = r
        self.J = solution_point_jacobi(r, vertices)
        self.iJ = inverse_jacobi(self.J)
        self.Ji = flux_point_jacobi(r, vertices, literal_Ji)
        self.xpg = global_sp(r, vertices)
        w = o.gausslegendre(deg + 1)[1]
        self.wp = np.outer(w, w)
        self.ll, self.lr, self.dl = o.standard_lagrange(r)
        self.dhl, self.dhr = o.dradau(deg, r)


def rect_vertices(x0, x1, nx, y0, y1, ny):
    """[KB-recall] PSpace2D(x0, x1, nx, y0, y1, ny, 1, 1).vertices: CCW from bottom-left."""
    dx, dy = (x1 - x0) / nx, (y1 - y0) / ny
    i = np.arange(0, nx + 2)[:, None]
    j = np.arange(0, ny + 2)[None, :]
    xl = x0 + (i - 1) * dx + 0 * j
    yl = y0 + (j - 1) * dy + 0 * i
    v = np.empty((nx + 2, ny + 2, 4, 2))
    v[:, :, 0, 0], v[:, :, 0, 1] = xl, yl
    v[:, :, 1, 0], v[:, :, 1, 1] = xl + dx, yl
    v[:, :, 2, 0], v[:, :, 2, 1] = xl + dx, yl + dy
    v[:, :, 3, 0], v[:, :, 3, 1] = xl, yl + dy
    return v


def rotate_vertices(v, angle):
    c, s = np.cos(angle), np.sin(angle)
    out = np.empty_like(v)
    out[..., 0] = c * v[..., 0] - s * v[..., 1]
    out[..., 1] = s * v[..., 0] + c * v[..., 1]
    return out


def parallelogram_vertices(nx=30, ny=15, lx=1.0, ly=0.5, x0=0.0, y0=0.0):
    """dev/parallelogram.jl:36-61, literally (including the 0.5 factors of :51,:56 that make
    neighbouring rows overlap: every cell is the same 45-degree sheared dx x dy cell, which is all the
    residual sees)."""
    dx, dy = lx / nx, ly / ny
    v = np.empty((nx + 2, ny + 2, 4, 2))
    for j in range(ny + 2):
        for i in range(nx + 2):
            v[i, j, 0, 0] = x0 + 0.5 * (j - 1) / ny + (i - 1) * dx
            v[i, j, 1, 0] = v[i, j, 0, 0] + dx
            v[i, j, 2, 0] = v[i, j, 1, 0] + dy
            v[i, j, 3, 0] = v[i, j, 2, 0] - dx
            v[i, j, 0, 1] = y0 + 0.5 * (j - 1) * dy
            v[i, j, 1, 1] = v[i, j, 0, 1]
            v[i, j, 2, 1] = v[i, j, 0, 1] + dy
            v[i, j, 3, 1] = v[i, j, 2, 1]
    return v


def parallelogram_normals(nx, ny):
    """dev/parallelogram.jl:176-186."""
    n1 = np.empty((nx + 1, ny, 2))
    n1[..., 0], n1[..., 1] = np.cos(-np.pi / 4), np.sin(-np.pi / 4)
    n2 = np.empty((nx, ny + 1, 2))
    n2[..., 0], n2[..., 1] = 0.0, 1.0
    return n1, n2


def cspace2d_vertices(r0, r1, nr, th0, th1, nth, ngr=0, ngth=1):
    """[KB-recall] KitBase.CSpace2D(r0, r1, nr, th0, th1, nth, ngr, ngth): polar cell centres
    r = r0 + (i - 1/2) dr, th = th0 + (j - 1/2) dth, vertices at (r -+ dr/2, th -+ dth/2) CCW from
    (inner, lower).  Returns (vertices[i, j, 4, 2], dtheta[j]) with i = 1-ngr..nr+ngr, j = 1-ngth..nth+ngth."""
    dr, dth = (r1 - r0) / nr, (th1 - th0) / nth
    ii = np.arange(1 - ngr, nr + ngr + 1)
    jj = np.arange(1 - ngth, nth + ngth + 1)
    rc = (r0 + (ii - 0.5) * dr)[:, None]
    tc = (th0 + (jj - 0.5) * dth)[None, :]
    v = np.empty((len(ii), len(jj), 4, 2))
    for q, (sr, st) in enumerate(((-1, -1), (1, -1), (1, 1), (-1, 1))):
        rr, tt = rc + 0.5 * sr * dr, tc + 0.5 * st * dth
        v[:, :, q, 0] = rr * np.cos(tt)
        v[:, :, q, 1] = rr * np.sin(tt)
    return v, np.full(len(jj), dth)


def cylinder_normals(nr, nth, dth):
    """dev/cylinder2.jl:39-49 with a uniform d-theta: n1[i, j] points along the radius through the middle
    of cell row j, n2[i, j] along the tangent at the lower edge of row j."""
    j = np.arange(1, nth + 1)
    a1 = (j - 1) * dth + 0.5 * dth
    n1 = np.empty((nr + 1, nth, 2))
    n1[..., 0], n1[..., 1] = np.cos(a1)[None, :], np.sin(a1)[None, :]
    j = np.arange(1, nth + 2)
    a2 = np.pi / 2 + (j - 1) * dth
    n2 = np.empty((nr, nth + 1, 2))
    n2[..., 0], n2[..., 1] = np.cos(a2)[None, :], np.sin(a2)[None, :]
    return n1, n2


def embed_cylinder(a_ref):
    """The cylinder scripts hold arrays over i = 1:nr (no radial ghost); the ABI wants a full ghost ring.
    Cell nr is never updated there (cylinder2.jl:154: i in 1:nx-1) and only feeds face nr, i.e. it *is* the
    outer ghost column: interior nx = nr - 1, column 0 is a dummy copy of column 1 (never read when the
    inner face is the mirror wall)."""
    return np.concatenate([a_ref[:1], a_ref], axis=0)


# ----------------------------------------------------------------------------------------------
# correction factors
# ----------------------------------------------------------------------------------------------


def corr_factors_sp(iJ, n1, n2):
    """dev/parallelogram.jl:145-148: (iJ[i,j][k,l] * n)[component] at every solution point.
    Returns c[nx, ny, nsp, nsp, 4] = (xL, xR, yL, yR)."""
    nx, ny = n1.shape[0] - 1, n2.shape[1] - 1
    I, Jn = slice(1, nx + 1), slice(1, ny + 1)
    A = iJ[I, Jn]
    c = np.empty(A.shape[:4] + (4,))
    e = (slice(None), slice(None), None, None)
    c[..., 0] = A[..., 0, 0] * n1[:-1, :, 0][e] + A[..., 0, 1] * n1[:-1, :, 1][e]
    c[..., 1] = A[..., 0, 0] * n1[1:, :, 0][e] + A[..., 0, 1] * n1[1:, :, 1][e]
    c[..., 2] = A[..., 1, 0] * n2[:, :-1, 0][e] + A[..., 1, 1] * n2[:, :-1, 1][e]
    c[..., 3] = A[..., 1, 0] * n2[:, 1:, 0][e] + A[..., 1, 1] * n2[:, 1:, 1][e]
    return c


def corr_factors_fp(Ji, n1, n2):
    """dev/cylinder2.jl:155-158: (inv(Ji[i,j][face, pt]) * n)[component] at the flux points.
    Returns c[nx, ny, nsp, 4] = (xL by l via face 4, xR by l via face 2, yL by k via face 1, yR by k via face 3)."""
    nx, ny = n1.shape[0] - 1, n2.shape[1] - 1
    I, Jn = slice(1, nx + 1), slice(1, ny + 1)
    iJi = inverse_jacobi(Ji[I, Jn])  # [nx, ny, 4, nsp, 2, 2]
    c = np.empty(iJi.shape[:2] + (iJi.shape[3], 4))
    e = (slice(None), slice(None), None)
    c[..., 0] = iJi[:, :, 3, :, 0, 0] * n1[:-1, :, 0][e] + iJi[:, :, 3, :, 0, 1] * n1[:-1, :, 1][e]
    c[..., 1] = iJi[:, :, 1, :, 0, 0] * n1[1:, :, 0][e] + iJi[:, :, 1, :, 0, 1] * n1[1:, :, 1][e]
    c[..., 2] = iJi[:, :, 0, :, 1, 0] * n2[:, :-1, 0][e] + iJi[:, :, 0, :, 1, 1] * n2[:, :-1, 1][e]
    c[..., 3] = iJi[:, :, 2, :, 1, 0] * n2[:, 1:, 0][e] + iJi[:, :, 2, :, 1, 1] * n2[:, 1:, 1][e]
    return c


# ----------------------------------------------------------------------------------------------
# the residual
# ----------------------------------------------------------------------------------------------


def wall_state(ul, gamma):
    """dev/cylinder2.jl:103-114: the mirror state behind the inner wall, in the face frame."""
    prim = o.conserve_prim(ul, gamma)
    pn = np.empty_like(prim)
    pn[..., 1] = -prim[..., 1]
    pn[..., 2] = prim[..., 2]
    pn[..., 3] = 2.0 - prim[..., 3]
    tmp = prim[..., 3] - 1.0
    pn[..., 0] = (1 - tmp) / (1 + tmp) * prim[..., 0]
    return o.prim_conserve(pn, gamma)


def rhs_euler2d_curv(u, ps, n1, n2, gamma, corr="sp", fpc=None, fy_index="l", wall_xlo=False, flux="hll"):
    """dudt! of dev/parallelogram.jl:80-165 (corr="sp") / dev/cylinder2.jl:52-164 (corr="fp",
    wall_xlo=True).  u[nx+2, ny+2, nsp, nsp, 4] with one ghost ring; n1[nx+1, ny, 2], n2[nx, ny+1, 2];
    fpc: flux-point correction factors (corr_factors_fp) when corr == "fp".  du = 0 in the ghosts.
    flux: the common flux in the face frame -- "hll" is what the scripts call; "lf" / "roe" are the north
    star's extras as defined in fr_oracle.flux_lf / flux_roe."""
    riemann = o.RIEMANN[flux]
    nxg, nyg, nsp, _, _ = u.shape
    nx, ny = nxg - 2, nyg - 2
    ll, lr, dhl, dhr, lpdm = ps.ll, ps.lr, ps.dhl, ps.dhr, ps.dl
    iJ = ps.iJ
    F, G = o.euler_flux(u, gamma)
    # f[i,j,k,l,m,:] = iJ[i,j][k,l] * [F_m, G_m]   (parallelogram.jl:88-96)
    f1 = iJ[..., 0, 0, None] * F + iJ[..., 0, 1, None] * G
    f2 = iJ[..., 1, 0, None] * F + iJ[..., 1, 1, None] * G

    def tr_s(a, l):  # dot(a[i,j,p,:,m], l): contracts the second tensor index
        acc = a[:, :, :, 0, :] * l[0]
        for q in range(1, nsp):
            acc = acc + a[:, :, :, q, :] * l[q]
        return acc

    def tr_r(a, l):  # dot(a[i,j,:,p,m], l)
        acc = a[:, :, 0, :, :] * l[0]
        for q in range(1, nsp):
            acc = acc + a[:, :, q, :, :] * l[q]
        return acc

    u1, u2, u3, u4 = tr_s(u, ll), tr_r(u, lr), tr_s(u, lr), tr_r(u, ll)  # :100-105
    f_face1_2, f_face3_2 = tr_s(f2, ll), tr_s(f2, lr)
    f_face2_1, f_face4_1 = tr_r(f1, lr), tr_r(f1, ll)

    # x faces i = 1..nx+1 (:115-125)
    c, s = n1[:, :, None, 0], n1[:, :, None, 1]
    uL = o.local_frame(u2[0 : nx + 1, 1 : ny + 1], c, s)
    uR = o.local_frame(u4[1 : nx + 2, 1 : ny + 1], c, s)
    if wall_xlo:  # cylinder2.jl:100-120: face 1 is flux_hll!(fw, ub, ul)
        uL[0] = wall_state(uR[0], gamma)
    fx = o.global_frame(riemann(uL, uR, gamma, 1.0), c, s)
    # y faces j = 1..ny+1 (:126-136)
    c, s = n2[:, :, None, 0], n2[:, :, None, 1]
    uL = o.local_frame(u3[1 : nx + 1, 0 : ny + 1], c, s)
    uR = o.local_frame(u1[1 : nx + 1, 1 : ny + 2], c, s)
    fy = o.global_frame(riemann(uL, uR, gamma, 1.0), c, s)

    if corr == "sp":
        cf = corr_factors_sp(iJ, n1, n2)
    elif corr == "fp":
        assert fpc is not None
    else:
        raise ValueError(corr)

    du = np.zeros_like(u)
    I, Jn = slice(1, nx + 1), slice(1, ny + 1)
    for k in range(nsp):
        for l in range(nsp):
            rhs1 = f1[I, Jn, 0, l, :] * lpdm[k, 0]
            rhs2 = f2[I, Jn, k, 0, :] * lpdm[l, 0]
            for q in range(1, nsp):
                rhs1 = rhs1 + f1[I, Jn, q, l, :] * lpdm[k, q]
                rhs2 = rhs2 + f2[I, Jn, k, q, :] * lpdm[l, q]
            if corr == "sp":
                cxL, cxR, cyL, cyR = (cf[:, :, k, l, q, None] for q in range(4))
            else:
                cxL, cxR = fpc[:, :, l, 0, None], fpc[:, :, l, 1, None]
                cyL, cyR = fpc[:, :, k, 2, None], fpc[:, :, k, 3, None]
            yi = l if fy_index == "l" else k
            fxL = cxL * fx[0:nx, :, l, :]
            fxR = cxR * fx[1 : nx + 1, :, l, :]
            fyL = cyL * fy[:, 0:ny, yi, :]
            fyR = cyR * fy[:, 1 : ny + 1, yi, :]
            du[I, Jn, k, l, :] = -(
                rhs1
                + rhs2
                + (fxL - f_face4_1[I, Jn, l, :]) * dhl[k]
                + (fxR - f_face2_1[I, Jn, l, :]) * dhr[k]
                + (fyL - f_face1_2[I, Jn, k, :]) * dhl[l]
                + (fyR - f_face3_2[I, Jn, k, :]) * dhr[l]
            )
    return du


def ghost_fill_periodic(u):
    """dev/parallelogram.jl:201-205."""
    nx, ny = u.shape[0] - 2, u.shape[1] - 2
    u[0] = u[nx]
    u[nx + 1] = u[1]
    u[:, 0] = u[:, ny]
    u[:, ny + 1] = u[:, 1]
    return u


def ghost_fill_cylinder(u, nsp):
    """dev/cylinder2.jl:176-187 on the embedded array (embed_cylinder): theta ghosts = point-reversed
    mirror images with the y momentum flipped (the script's ``4 - k`` is ``nsp + 1 - k`` at deg 2, the only
    degree it runs at); the outer column copies its inner neighbour over the first half of the rows."""
    nx, ny = u.shape[0] - 2, u.shape[1] - 2  # nx = nr - 1
    flip = np.array([1.0, 1.0, -1.0, 1.0])
    src1 = u[1:, 1].copy()
    src2 = u[1:, ny].copy()
    u[1:, 0] = src1[:, ::-1, ::-1, :] * flip
    u[1:, ny + 1] = src2[:, ::-1, ::-1, :] * flip
    u[nx + 1, 1 : ny // 2 + 1] = u[nx, 1 : ny // 2 + 1]
    return u
