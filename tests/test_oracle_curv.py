"""CPU tests of the curvilinear-quadrilateral path: the oracle restatement of
dev/parallelogram.jl / dev/cylinder2.jl (oracle/fr_oracle_curv.py), the host mirror of the metric
(FRPSpace2D(base, deg), PSpace2D, CSpace2D, face_normals) and -- through tests/harness -- the very
__host__ __device__ routines the CUDA kernels are made of, run in CPU loops.

The reference pins nothing here (no Julia, no test on this path): the oracle is held by invariants.
"""
import ctypes

import numpy as np
import pytest

import fr_oracle as o
import fr_oracle_curv as c

GAMMA = 5.0 / 3.0


def rand_state(shape, seed):
    rng = np.random.default_rng(seed)
    prim = np.empty(shape + (4,))
    prim[..., 0] = 1.0 + 0.2 * rng.random(shape)
    prim[..., 1] = 0.3 + 0.1 * rng.standard_normal(shape)
    prim[..., 2] = 0.3 + 0.1 * rng.standard_normal(shape)
    prim[..., 3] = 1.0 + 0.2 * rng.random(shape)
    return np.asfortranarray(o.prim_conserve(prim, GAMMA))


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def axis_normals(nx, ny, th=0.0):
    n1 = np.zeros((nx + 1, ny, 2))
    n2 = np.zeros((nx, ny + 1, 2))
    n1[..., 0], n1[..., 1] = np.cos(th), np.sin(th)
    n2[..., 0], n2[..., 1] = -np.sin(th), np.cos(th)
    return n1, n2


def sheared_vertices(nx, ny, lx=1.0, ly=0.5, shear=1.0):
    """A consistent tiling by parallelograms: x -> x + shear * y."""
    v = c.rect_vertices(0.0, lx, nx, 0.0, ly, ny)
    out = v.copy()
    out[..., 0] = v[..., 0] + shear * v[..., 1]
    return out


# ------------------------------------------------------------------------------- oracle invariants
@pytest.mark.parametrize("deg", [1, 2, 3])
@pytest.mark.parametrize("corr", ["sp", "fp"])
def test_reduces_to_rectangular_residual(deg, corr):
    nx, ny = 6, 5
    ps_r = o.FRPSpace2D(0, 1, nx, 0, 0.5, ny, deg, 1, 1)
    ps = c.CurvSpace2D(c.rect_vertices(0, 1, nx, 0, 0.5, ny), deg)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 1)
    n1, n2 = axis_normals(nx, ny)
    fpc = c.corr_factors_fp(ps.Ji, n1, n2) if corr == "fp" else None
    a = o.rhs_euler2d(u, ps_r, GAMMA)  # example/euler2d_wave.jl:35-107
    b = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr=corr, fpc=fpc, fy_index="k")
    assert rel(b, a) < 1e-14
    # the scripts' row index differs from the rectangular form as soon as the state varies along x
    b2 = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr=corr, fpc=fpc, fy_index="l")
    assert rel(b2, a) > 1e-3


@pytest.mark.parametrize("deg", [1, 2, 3])
@pytest.mark.parametrize("corr,fy", [("sp", "k"), ("sp", "l"), ("fp", "k"), ("fp", "l")])
def test_rotation_equivariance(deg, corr, fy):
    """Rotating the mesh, the normals and the momentum by the same angle rotates du."""
    nx, ny, th = 6, 5, 0.37
    v0 = c.rect_vertices(0, 1, nx, 0, 0.5, ny)
    ps0, psr = c.CurvSpace2D(v0, deg), c.CurvSpace2D(c.rotate_vertices(v0, th), deg)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 2)
    ur = o.global_frame(u, np.cos(th), np.sin(th))
    n1, n2 = axis_normals(nx, ny)
    n1r, n2r = axis_normals(nx, ny, th)
    kw0, kwr = dict(corr=corr, fy_index=fy), dict(corr=corr, fy_index=fy)
    if corr == "fp":
        kw0["fpc"] = c.corr_factors_fp(ps0.Ji, n1, n2)
        kwr["fpc"] = c.corr_factors_fp(psr.Ji, n1r, n2r)
    d0 = c.rhs_euler2d_curv(u, ps0, n1, n2, GAMMA, **kw0)
    dr = c.rhs_euler2d_curv(ur, psr, n1r, n2r, GAMMA, **kwr)
    assert rel(dr, o.global_frame(d0, np.cos(th), np.sin(th))) < 1e-13


@pytest.mark.parametrize("deg", [1, 2, 3])
@pytest.mark.parametrize("fy", ["k", "l"])
def test_free_stream_on_the_parallelogram(deg, fy):
    nx, ny = 8, 4
    ps = c.CurvSpace2D(c.parallelogram_vertices(nx, ny), deg)
    n1, n2 = c.parallelogram_normals(nx, ny)
    u = np.empty((nx + 2, ny + 2, deg + 1, deg + 1, 4))
    u[...] = o.prim_conserve(np.array([1.0, 0.7, 0.3, 1.2]), GAMMA)
    du = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, fy_index=fy)
    assert np.abs(du).max() < 1e-12


@pytest.mark.parametrize("deg", [1, 2, 3])
def test_conservation_on_a_periodic_sheared_mesh(deg):
    """sum over the mesh of wp det(J) du vanishes for every variable when the y common flux is indexed by the
    flux point (k); the scripts' row index (l, dev/parallelogram.jl:147-148) loses conservation at O(1) -- what
    the script's header reports as "instability is somehow detected for order larger than 2"."""
    nx, ny = 6, 5
    ps = c.CurvSpace2D(sheared_vertices(nx, ny, 1.0, 0.5, 1.0), deg)
    n1 = np.empty((nx + 1, ny, 2))
    n1[..., 0], n1[..., 1] = np.cos(-np.pi / 4), np.sin(-np.pi / 4)
    n2 = np.zeros((nx, ny + 1, 2))
    n2[..., 1] = 1.0
    u = c.ghost_fill_periodic(rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 5))
    det = ps.J[..., 0, 0] * ps.J[..., 1, 1] - ps.J[..., 0, 1] * ps.J[..., 1, 0]
    total = {}
    for fy in "kl":
        du = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, fy_index=fy)
        total[fy] = np.einsum("ijklm,kl,ijkl->m", du[1:-1, 1:-1], ps.wp, det[1:-1, 1:-1])
    assert np.abs(total["k"]).max() < 1e-13
    assert np.abs(total["l"]).max() > 1e-2


def test_converges_to_the_flux_divergence_on_a_sheared_mesh():
    """rho_t = -(U rho_x + V rho_y) for a density wave carried by a uniform flow at constant pressure."""
    U, V, deg = 0.8, 0.5, 2
    errs = []
    for n in (8, 16, 32):
        v = sheared_vertices(n, n, 1.0, 1.0, 1.0)
        ps = c.CurvSpace2D(v, deg)
        n1 = np.empty((n + 1, n, 2))
        n1[..., 0], n1[..., 1] = np.cos(-np.pi / 4), np.sin(-np.pi / 4)
        n2 = np.zeros((n, n + 1, 2))
        n2[..., 1] = 1.0
        x, y = ps.xpg[..., 0], ps.xpg[..., 1]
        ph = 2 * np.pi * (x + 0.5 * y)
        rho = 1.0 + 0.1 * np.sin(ph)
        prim = np.stack([rho, np.full_like(rho, U), np.full_like(rho, V), rho], axis=-1)  # p = rho / (2 lambda) = 1/2
        u = o.prim_conserve(prim, GAMMA)
        du = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, fy_index="k")
        exact = -(U + 0.5 * V) * 0.1 * 2 * np.pi * np.cos(ph)
        errs.append(np.abs(du[1:-1, 1:-1, :, :, 0] - exact[1:-1, 1:-1]).max())
    order = np.log2(errs[0] / errs[1]), np.log2(errs[1] / errs[2])
    assert min(order) > deg - 0.3, (errs, order)


def test_literal_flux_point_jacobians_hide_on_parallelograms_only():
    r = o.legendre_point(2)
    v = c.parallelogram_vertices(5, 4)
    assert np.abs(c.flux_point_jacobi(r, v, True) - c.flux_point_jacobi(r, v, False)).max() < 1e-15
    vc, _ = c.cspace2d_vertices(1.0, 6.0, 6, 0.0, np.pi, 8, 0, 1)
    assert np.abs(c.flux_point_jacobi(r, vc, True) - c.flux_point_jacobi(r, vc, False)).max() > 1e-3


def test_wall_state_mirrors_the_normal_velocity():
    w = rand_state((5,), 3)
    prim, pw = o.conserve_prim(w, GAMMA), o.conserve_prim(c.wall_state(w, GAMMA), GAMMA)
    assert np.allclose(pw[:, 1], -prim[:, 1]) and np.allclose(pw[:, 2], prim[:, 2])
    assert np.allclose(pw[:, 3], 2.0 - prim[:, 3])


def test_cylinder_ghost_fill():
    nr, nth, nsp = 5, 6, 3
    u = rand_state((nr + 1, nth + 2, nsp, nsp), 4)
    ref = u.copy()
    c.ghost_fill_cylinder(u, nsp)
    for k in range(nsp):
        for l in range(nsp):  # cylinder2.jl:177-185 with 4 - k -> nsp + 1 - k
            src = ref[1:, 1, nsp - 1 - k, nsp - 1 - l]
            assert np.array_equal(u[1:, 0, k, l], src * np.array([1.0, 1.0, -1.0, 1.0]))
    assert np.array_equal(u[nr, 1 : nth // 2 + 1], ref[nr - 1, 1 : nth // 2 + 1])
    assert np.array_equal(u[nr, nth // 2 + 1 : nth + 1], ref[nr, nth // 2 + 1 : nth + 1])


@pytest.mark.parametrize("deg", [1, 2, 3])
def test_c_restatement_equals_the_numpy_one(coracle, deg):
    """oracle/fr_oracle_curv.c (the fast checker for large meshes) against oracle/fr_oracle_curv.py."""
    nx, ny = 7, 5
    ps = c.CurvSpace2D(c.parallelogram_vertices(nx, ny), deg)
    n1, n2 = c.parallelogram_normals(nx, ny)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 21)
    for fy in "kl":
        a = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="sp", fy_index=fy)
        assert rel(coracle.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="sp", fy_index=fy), a) < 1e-14
    nr, nth = 6, 8
    vv, dth = c.cspace2d_vertices(1.0, 6.0, nr, 0.0, np.pi, nth, 0, 1)
    ps = c.CurvSpace2D(c.embed_cylinder(vv), deg)
    n1, n2 = c.cylinder_normals(nr, nth, dth[0])
    n1, n2 = n1[:nr], n2[: nr - 1]
    fpc = c.corr_factors_fp(ps.Ji, n1, n2)
    u = rand_state((nr + 1, nth + 2, deg + 1, deg + 1), 22)
    for fy in "kl":
        kw = dict(corr="fp", fpc=fpc, fy_index=fy, wall_xlo=True)
        assert rel(coracle.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, **kw), c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, **kw)) < 1e-14


# ------------------------------------------------------------------------------- host mirror
@pytest.mark.parametrize("deg", [1, 2, 3])
def test_host_space_matches_the_oracle(FR, deg):
    nx, ny = 30, 15
    v = c.parallelogram_vertices(nx, ny)
    z = np.zeros((nx + 2, ny + 2))
    ps = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, nx, 0.0, 0.5, ny, z, z, z, z, v), deg)
    po = c.CurvSpace2D(v, deg)
    for name in ("J", "iJ", "Ji", "xpg"):
        assert np.abs(getattr(ps, name) - getattr(po, name)).max() < 1e-13, name
    n1, n2 = FR.face_normals(v)
    m1, m2 = c.parallelogram_normals(nx, ny)  # dev/parallelogram.jl:176-186
    assert np.abs(n1 - m1).max() < 1e-14 and np.abs(n2 - m2).max() < 1e-14

    cs = FR.CSpace2D(1.0, 6.0, 8, 0.0, np.pi, 10, 0, 1)
    vv, dth = c.cspace2d_vertices(1.0, 6.0, 8, 0.0, np.pi, 10, 0, 1)
    assert np.abs(cs.vertices - vv).max() < 1e-14
    for literal in (True, False):
        ps = FR.FRPSpace2D(FR.embed_ghostless_x(cs), deg, literal_Ji=literal)
        po = c.CurvSpace2D(c.embed_cylinder(vv), deg, literal_Ji=literal)
        assert (ps.nx, ps.ny) == (7, 10)
        assert np.abs(ps.J - po.J).max() < 1e-13 and np.abs(ps.Ji - po.Ji).max() < 1e-13
    n1, n2 = FR.face_normals(ps.vertices)
    m1, m2 = c.cylinder_normals(8, 10, dth[0])  # dev/cylinder2.jl:39-49
    assert np.abs(n1 - m1[:8]).max() < 1e-14 and np.abs(n2 - m2[:7]).max() < 1e-14
    ps = FR.FRPSpace2D(FR.embed_ghostless_x(cs), deg)
    po = c.CurvSpace2D(c.embed_cylinder(vv), deg)
    assert np.abs(FR.correction_factors_fp(ps.Ji, n1, n2) - c.corr_factors_fp(po.Ji, m1[:8], m2[:7])).max() < 1e-13


def test_uniform_pspace2d_is_the_rectangular_space(FR):
    base = FR.PSpace2D(0.0, 1.0, 6, 0.0, 0.5, 4, 1, 1)
    assert np.abs(base.vertices - c.rect_vertices(0.0, 1.0, 6, 0.0, 0.5, 4)).max() < 1e-15
    ps, pr = FR.FRPSpace2D(base, 2), FR.FRPSpace2D(0.0, 1.0, 6, 0.0, 0.5, 4, 2, 1, 1)
    assert np.abs(ps.J - pr.J).max() < 1e-15 and np.abs(ps.xpg - pr.xpg).max() < 1e-14
    assert np.array_equal(ps.V, pr.V) and np.array_equal(ps.dl, pr.dl)


# ------------------------------------------------------------------------------- device routines on the CPU
@pytest.fixture(scope="module")
def host_stage():
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "harness"))
    import build as hb

    lib = ctypes.CDLL(hb.build())
    D = ctypes.POINTER(ctypes.c_double)
    lib.curv_host_stage.argtypes = ([ctypes.c_int] * 3 + [D] * 11 + [ctypes.c_int, ctypes.c_double] + [D] * 5
                                    + [ctypes.c_double] * 3 + [ctypes.c_int] * 2)

    def P(a):
        return None if a is None else a.ctypes.data_as(D)

    def run(u, ps, n1, n2, fpc, flags, stage=(0.0, 0.0, 1.0, 0, 1), ua=None, vertices=False):
        nx, ny, nsp = u.shape[0] - 2, u.shape[1] - 2, u.shape[2]
        out = np.zeros_like(u, order="F")
        keep = [np.asfortranarray(ps.iJ), np.asfortranarray(n1), np.asfortranarray(n2),
                None if fpc is None else np.asfortranarray(fpc), np.zeros((nx + 1) * ny * nsp * 4),
                np.zeros(nx * (ny + 1) * nsp * 4), np.asfortranarray(ps.dl),
                np.asfortranarray(ps.vertices) if vertices else None, np.ascontiguousarray(ps.xpl)]
        ops = [np.ascontiguousarray(a) for a in (ps.ll, ps.lr)] + [keep[6]] + [np.ascontiguousarray(a) for a in (ps.dhl, ps.dhr)]
        rc = lib.curv_host_stage(nx, ny, nsp, P(u), P(ua), P(out), P(keep[4]), P(keep[5]), P(keep[0]), P(keep[1]),
                                 P(keep[2]), P(keep[3]), P(keep[7]), P(keep[8]), flags, GAMMA, *[P(a) for a in ops],
                                 float(stage[0]), float(stage[1]), float(stage[2]), int(stage[3]), int(stage[4]))
        assert rc == 0
        return out

    return run


@pytest.mark.parametrize("fast", [0, 1 << 10])
@pytest.mark.parametrize("deg", [1, 2, 3])
def test_device_routines_on_the_cpu(host_stage, deg, fast):
    """face_xy / row_xpass / row_ypass of csrc/frb_euler2d_curv_elem.cuh, the code the kernels run: the literal
    instantiation (size_t indices, the correction written as in the scripts) and the one the kernels launch
    (fast: 32-bit indices, the flux traces folded into the derivative matrix -- rounding differs in the last bits)."""
    tol = 1e-13 if fast else 1e-14
    nx, ny = 7, 5
    ps = c.CurvSpace2D(c.parallelogram_vertices(nx, ny), deg)
    n1, n2 = c.parallelogram_normals(nx, ny)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 5)
    for fy, flags in (("l", 1), ("k", 0)):
        ref = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="sp", fy_index=fy)
        assert rel(host_stage(u, ps, n1, n2, None, flags | fast), ref) < tol
    nr, nth = 6, 8
    vv, dth = c.cspace2d_vertices(1.0, 6.0, nr, 0.0, np.pi, nth, 0, 1)
    ps = c.CurvSpace2D(c.embed_cylinder(vv), deg)
    n1, n2 = c.cylinder_normals(nr, nth, dth[0])
    n1, n2 = n1[:nr], n2[: nr - 1]
    fpc = c.corr_factors_fp(ps.Ji, n1, n2)
    u = rand_state((nr + 1, nth + 2, deg + 1, deg + 1), 6)
    for fy, flags in (("l", 3), ("k", 2)):
        ref = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index=fy, wall_xlo=True)
        assert rel(host_stage(u, ps, n1, n2, fpc, flags | fast), ref) < tol
    # the metric evaluated from the vertices on the fly (frb_euler2d_curv_set_vertices) instead of the stored iJ
    ref = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index="k", wall_xlo=True)
    assert rel(host_stage(u, ps, n1, n2, fpc, 2 | fast, vertices=True), ref) < 1e-13
    for kind, name in ((1, "lf"), (2, "roe")):  # the extra common fluxes, in the face frame
        ref = c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index="k", wall_xlo=True, flux=name)
        assert rel(host_stage(u, ps, n1, n2, fpc, 2 | (kind << 8) | fast), ref) < 1e-13
    ua = rand_state(u.shape[:-1], 7)
    ref = 0.75 * ua + 0.25 * u + 0.25e-3 * c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index="k",
                                                               wall_xlo=True)
    got = host_stage(u, ps, n1, n2, fpc, 2 | fast, stage=(0.75, 0.25, 0.25e-3, 1, 0), ua=ua)
    assert np.abs(got[1:-1, 1:-1] - ref[1:-1, 1:-1]).max() < 1e-14


# ------------------------------------------------------------------------------- committed fixtures
def _golden():
    import os

    return np.load(os.path.join(os.path.dirname(__file__), "golden", "curv_golden.npz"))


def test_oracle_reproduces_the_golden_vectors(host_stage):
    """tests/golden/curv_golden.npz (make_curv_golden.py): the oracle and the device routines on the CPU."""
    g, deg = _golden(), 2
    nx, ny = 6, 4
    ps = c.CurvSpace2D(c.parallelogram_vertices(nx, ny), deg)
    n1, n2 = c.parallelogram_normals(nx, ny)
    u = np.asfortranarray(g["para_u"])
    for fy, flags in (("k", 0), ("l", 1)):
        assert rel(c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="sp", fy_index=fy), g[f"para_du_{fy}"]) < 1e-14
        assert rel(host_stage(u, ps, n1, n2, None, flags), g[f"para_du_{fy}"]) < 1e-13
    nr, nth = 5, 6
    vv, dth = c.cspace2d_vertices(1.0, 6.0, nr, 0.0, np.pi, nth, 0, 1)
    ps = c.CurvSpace2D(c.embed_cylinder(vv), deg)
    n1, n2 = c.cylinder_normals(nr, nth, dth[0])
    n1, n2 = n1[:nr], n2[: nr - 1]
    fpc = c.corr_factors_fp(ps.Ji, n1, n2)
    assert np.abs(fpc - g["cyl_fpc"]).max() < 1e-14
    u = np.asfortranarray(g["cyl_u"])
    for fy, flags in (("k", 2), ("l", 3)):
        ref = g[f"cyl_du_{fy}"]
        assert rel(c.rhs_euler2d_curv(u, ps, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index=fy, wall_xlo=True), ref) < 1e-14
        assert rel(host_stage(u, ps, n1, n2, fpc, flags), ref) < 1e-13
    assert np.array_equal(c.ghost_fill_cylinder(u.copy(), deg + 1), g["cyl_ghost"])


# ------------------------------------------------------------------------------- row slabs of a curvilinear mesh
@pytest.mark.parametrize("world", [2, 3, 5])
def test_slab_space_carries_the_global_metric(FR, world):
    """What tests/dist/check_dist_curv.py hands to DistributedEuler2DCurv on every rank: the space built from the
    slab's vertices has the global mesh's metric of those rows bit for bit, the normal tables slice by rows / faces."""
    nx, nyg, deg = 9, 11, 2
    v = c.parallelogram_vertices(nx, nyg)
    pog = c.CurvSpace2D(v, deg)
    n1g, n2g = c.parallelogram_normals(nx, nyg)
    covered = []
    for rank in range(world):
        sl = FR.partition.slab(nyg, world, rank)
        rows = slice(sl.start - 1, sl.stop + 2)
        zl = np.zeros((nx + 2, sl.count + 2))
        psl = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, nx, 0.0, 0.5, sl.count, zl, zl, zl, zl, v[:, rows]), deg)
        assert (psl.nx, psl.ny, psl.ngx, psl.ngy) == (nx, sl.count, 1, 1)
        assert np.array_equal(psl.iJ, pog.iJ[:, rows])
        n1l, n2l = n1g[:, sl.start - 1: sl.stop], n2g[:, sl.start - 1: sl.stop + 1]
        assert n1l.shape == (nx + 1, sl.count, 2) and n2l.shape == (nx, sl.count + 1, 2)
        f1, f2 = FR.face_normals(psl.vertices)
        assert np.abs(f1 - n1l).max() < 1e-14 and np.abs(f2 - n2l).max() < 1e-14
        covered += list(range(sl.start, sl.stop + 1))
    assert covered == list(range(1, nyg + 1))
