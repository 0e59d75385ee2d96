"""GPU parity of the curvilinear-quadrilateral Euler path (frb_euler2d_curv_create) against
the NumPy restatement of dev/parallelogram.jl:80-165 and dev/cylinder2.jl:52-187 (oracle/fr_oracle_curv.py),
through the C ABI.  Tolerances as in test_gpu_parity.py: 1e-12 relative per RHS, 1e-9 after the steps.
"""
import numpy as np
import pytest

import fr_oracle as o
import fr_oracle_curv as c

pytestmark = pytest.mark.gpu

GAMMA = 5.0 / 3.0
RTOL_RHS = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def rand_state(shape, seed):
    rng = np.random.default_rng(seed)
    prim = np.empty(shape + (4,))
    prim[..., 0] = 1.0 + 0.2 * rng.random(shape)
    prim[..., 1] = 0.3 + 0.1 * rng.standard_normal(shape)
    prim[..., 2] = 0.3 + 0.1 * rng.standard_normal(shape)
    prim[..., 3] = 1.0 + 0.2 * rng.random(shape)
    return np.asfortranarray(o.prim_conserve(prim, GAMMA))


def parallelogram(FR, nx, ny, deg):
    v = c.parallelogram_vertices(nx, ny)
    z = np.zeros((nx + 2, ny + 2))
    ps = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, nx, 0.0, 0.5, ny, z, z, z, z, v), deg)
    return ps, c.CurvSpace2D(v, deg), c.parallelogram_normals(nx, ny)


def cylinder(FR, nr, nth, deg):
    cs = FR.CSpace2D(1.0, 6.0, nr, 0.0, np.pi, nth, 0, 1)
    ps = FR.FRPSpace2D(FR.embed_ghostless_x(cs), deg)
    vv, dth = c.cspace2d_vertices(1.0, 6.0, nr, 0.0, np.pi, nth, 0, 1)
    po = c.CurvSpace2D(c.embed_cylinder(vv), deg)
    n1, n2 = c.cylinder_normals(nr, nth, dth[0])
    return ps, po, (n1[:nr], n2[: nr - 1])


# auto: one launch, every block evaluates the fluxes of its own faces; generic: face kernel + element kernel;
# curv_march: the one-launch row-marching kernel
KERNELS = ["auto", "generic", "curv_march"]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("deg", [1, 2, 3])
@pytest.mark.parametrize("fy", ["k", "l"])
@pytest.mark.parametrize("nx,ny", [(30, 15), (33, 7), (1, 1), (31, 2), (61, 3)])
def test_parallelogram_rhs(FR, deg, fy, nx, ny, kernel):
    ps, po, (n1, n2) = parallelogram(FR, nx, ny, deg)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 11)
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, n1, n2, corr="sp", fy_index=fy)
    prob.set_kernel(kernel)
    du = np.full_like(u, np.nan, order="F")
    prob.f(du, u, None, 0.0)
    ref = c.rhs_euler2d_curv(u, po, n1, n2, GAMMA, corr="sp", fy_index=fy)
    assert rel(du, ref) <= RTOL_RHS
    assert np.all(du[0] == 0) and np.all(du[-1] == 0) and np.all(du[:, 0] == 0) and np.all(du[:, -1] == 0)
    prob.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("deg", [1, 2, 3])
@pytest.mark.parametrize("fy", ["k", "l"])
def test_cylinder_rhs(FR, deg, fy, kernel):
    """dev/cylinder2.jl:52-164: flux-point factors from the literal Ji, mirror wall on the inner face."""
    nr, nth = 30, 40
    ps, po, (n1, n2) = cylinder(FR, nr, nth, deg)
    u = rand_state((nr + 1, nth + 2, deg + 1, deg + 1), 12)
    u[0] = np.nan  # the dummy column behind the wall is never read
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, None, None, corr="fp", fy_index=fy, wall_xlo=True)
    prob.set_kernel(kernel)
    du = np.zeros_like(u, order="F")
    prob.f(du, u, None, 0.0)
    u0 = u.copy()
    u0[0] = u0[1]
    fpc = c.corr_factors_fp(po.Ji, n1, n2)
    ref = c.rhs_euler2d_curv(u0, po, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index=fy, wall_xlo=True)
    assert rel(du, ref) <= RTOL_RHS
    prob.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("flux", ["lf", "roe"])
def test_extra_fluxes_in_the_face_frame(FR, flux, kernel):
    nr, nth, deg = 12, 16, 2
    ps, po, (n1, n2) = cylinder(FR, nr, nth, deg)
    u = rand_state((nr + 1, nth + 2, deg + 1, deg + 1), 18)
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, corr="fp", fy_index="k", wall_xlo=True)
    prob.set_kernel(kernel)
    prob.set_flux(flux)
    du = np.zeros_like(u, order="F")
    prob.f(du, u, None, 0.0)
    fpc = c.corr_factors_fp(po.Ji, n1, n2)
    ref = c.rhs_euler2d_curv(u, po, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index="k", wall_xlo=True, flux=flux)
    assert rel(du, ref) <= RTOL_RHS
    prob.close()


@pytest.mark.parametrize("deg", [1, 2, 3])
def test_metric_from_vertices(FR, deg):
    """frb_euler2d_curv_set_vertices: iJ evaluated on the fly from the cell vertices (trapezoids: J varies)."""
    nr, nth = 20, 24
    ps, po, (n1, n2) = cylinder(FR, nr, nth, deg)
    u = rand_state((nr + 1, nth + 2, deg + 1, deg + 1), 19)
    fpc = c.corr_factors_fp(po.Ji, n1, n2)
    for corr in ("sp", "fp"):
        prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, corr=corr, wall_xlo=True, metric="vertices")
        du = np.zeros_like(u, order="F")
        prob.f(du, u, None, 0.0)
        ref = c.rhs_euler2d_curv(u, po, n1, n2, GAMMA, corr=corr, fpc=fpc, fy_index="k", wall_xlo=True)
        assert rel(du, ref) <= RTOL_RHS
        prob.set_metric("stored")
        du2 = np.zeros_like(u, order="F")
        prob.f(du2, u, None, 0.0)
        assert rel(du2, ref) <= RTOL_RHS
        prob.close()


@pytest.mark.parametrize("rows", [1, 2, 4, 5, 37, 1000])
@pytest.mark.parametrize("deg", [1, 2, 3])
def test_marching_kernel_segment_lengths(FR, deg, rows, monkeypatch):
    """The one-launch kernel with every split of the rows into segments (FRB_CURV_ROWS), several strips, a ragged
    last strip and a ragged last segment: the residual and three SSPRK3 steps against the two-kernel form."""
    nx, ny = 95, 37
    ps, po, (n1, n2) = parallelogram(FR, nx, ny, deg)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 21)
    ref = c.rhs_euler2d_curv(u, po, n1, n2, GAMMA, corr="sp", fy_index="k")
    monkeypatch.setenv("FRB_CURV_ROWS", str(rows))
    got = {}
    for kernel in KERNELS:
        prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, n1, n2, corr="sp", fy_index="k")
        prob.set_kernel(kernel)
        du = np.full_like(u, np.nan, order="F")
        prob.f(du, u, None, 0.0)
        assert rel(du, ref) <= RTOL_RHS, kernel
        _, n = prob.last_timing()
        assert n == (2 if kernel == "generic" else 1)
        itg = FR.init(prob, FR.SSPRK33(), dt=2e-4)
        itg.set_hooks(ghost="periodic")
        FR.step_(itg, 3)
        got[kernel] = itg.u.copy()
        prob.close()
    for kernel in KERNELS[1:]:
        assert np.isfinite(got[kernel]).all()
        assert rel(got[kernel], got["auto"]) <= 1e-13


def test_rectangular_mesh_equals_the_rectangular_problem(FR):
    nx, ny, deg = 40, 24, 3
    ps = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, nx, 0.0, 0.5, ny, 1, 1), deg)
    pr = FR.FRPSpace2D(0.0, 1.0, nx, 0.0, 0.5, ny, deg, 1, 1)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 13)
    a, b = np.zeros_like(u, order="F"), np.zeros_like(u, order="F")
    p1 = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA)
    p2 = FR.Euler2DProblem(u, (0.0, 1.0), pr, GAMMA)
    p1.f(a, u, None, 0.0)
    p2.f(b, u, None, 0.0)
    assert rel(a, b) <= RTOL_RHS
    p1.close()
    p2.close()


@pytest.mark.parametrize("scheme", ["euler", "midpoint", "ssprk3", "rk4"])
def test_parallelogram_steps(FR, scheme):
    """The user loop of dev/parallelogram.jl:188-208: periodic ghost copies, then step!(itg) (Euler there)."""
    nx, ny, deg, dt, nsteps = 30, 15, 1, 0.001, 100
    ps, po, (n1, n2) = parallelogram(FR, nx, ny, deg)
    u0 = np.empty((nx + 2, ny + 2, deg + 1, deg + 1, 4), order="F")
    for i in range(nx + 2):  # parallelogram.jl:167-174
        rho = 1.0 + 0.1 * np.sin(2 * np.pi * i / nx)
        u0[i] = o.prim_conserve(np.array([rho, 1.0, 0.0, rho]), GAMMA)
    alg = {"euler": FR.Euler, "midpoint": FR.Midpoint, "ssprk3": FR.SSPRK33, "rk4": FR.RK4}[scheme]
    prob = FR.Euler2DCurvProblem(u0, (0.0, 1.0), ps, GAMMA, n1, n2, corr="sp", fy_index="l")
    itg = FR.init(prob, alg(), dt=dt)
    itg.set_hooks(ghost="periodic")
    FR.step_(itg, nsteps)
    rhs = lambda w: c.rhs_euler2d_curv(w, po, n1, n2, GAMMA, corr="sp", fy_index="l")  # noqa: E731
    ref = o.integrate(u0, dt, nsteps, rhs, scheme, before_step=c.ghost_fill_periodic)
    assert rel(itg.u, ref) <= 1e-10
    prob.close()


def test_parallelogram_1000_steps(FR):
    """The north star's second tolerance on this path: <= 1e-9 relative after 1000 steps (the script's Euler forward,
    dt = 0.001, to t = 1).  With the flux-point index k the run is stable (a 1e-15 perturbation of the data grows 12 x);
    the scripts' row index l turns non-finite before t = 1 in the oracle as well, so there is nothing to compare."""
    nx, ny, deg, dt, nsteps = 30, 15, 1, 0.001, 1000
    ps, po, (n1, n2) = parallelogram(FR, nx, ny, deg)
    u0 = np.empty((nx + 2, ny + 2, deg + 1, deg + 1, 4), order="F")
    for i in range(nx + 2):
        rho = 1.0 + 0.1 * np.sin(2 * np.pi * i / nx)
        u0[i] = o.prim_conserve(np.array([rho, 1.0, 0.0, rho]), GAMMA)
    prob = FR.Euler2DCurvProblem(u0, (0.0, 1.0), ps, GAMMA, n1, n2, corr="sp", fy_index="k")
    itg = FR.init(prob, FR.Euler(), dt=dt)
    itg.set_hooks(ghost="periodic")
    FR.step_(itg, nsteps)
    rhs = lambda w: c.rhs_euler2d_curv(w, po, n1, n2, GAMMA, corr="sp", fy_index="k")  # noqa: E731
    ref = o.integrate(u0, dt, nsteps, rhs, "euler", before_step=c.ghost_fill_periodic)
    assert np.isfinite(ref).all()
    assert rel(itg.u, ref) <= 1e-9
    prob.close()


def test_cylinder_steps(FR):
    """dev/cylinder2.jl:170-190: mirror rows in theta, outflow copy on half of the outer column, Euler steps.
    The script's own case (uniform Mach-1 flow against the wall) loses positivity next to the wall at its 17th
    step in the oracle too (dev/cylinder3.jl:2: "the cylinder flow blows up somehow"), so 10 steps here."""
    nr, nth, deg, dt, nsteps = 30, 40, 2, 0.0005, 10
    ps, po, (n1, n2) = cylinder(FR, nr, nth, deg)
    u_ref = np.empty((nr, nth + 2, deg + 1, deg + 1, 4))
    u_ref[...] = o.prim_conserve(np.array([1.0, 1.0, 0.0, 1.0]), GAMMA)  # cylinder2.jl:33-37
    u0 = np.asfortranarray(c.embed_cylinder(u_ref))
    prob = FR.Euler2DCurvProblem(u0, (0.0, 1.0), ps, GAMMA, corr="fp", fy_index="l", wall_xlo=True)
    itg = FR.init(prob, FR.Euler(), dt=dt)
    itg.set_hooks(ghost="cylinder")
    FR.step_(itg, nsteps)
    fpc = c.corr_factors_fp(po.Ji, n1, n2)
    rhs = lambda w: c.rhs_euler2d_curv(w, po, n1, n2, GAMMA, corr="fp", fpc=fpc, fy_index="l", wall_xlo=True)  # noqa: E731
    ref = o.integrate(u0, dt, nsteps, rhs, "euler", before_step=lambda w: c.ghost_fill_cylinder(w, deg + 1))
    got = itg.u
    assert np.isfinite(got[1:]).all()
    assert rel(got[1:], ref[1:]) <= 1e-10
    prob.close()


def test_cylinder_ghost_fill_call(FR):
    nr, nth, deg = 9, 10, 2
    ps, _, _ = cylinder(FR, nr, nth, deg)
    u = rand_state((nr + 1, nth + 2, deg + 1, deg + 1), 14)
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, corr="fp", wall_xlo=True)
    prob.ghost_fill("cylinder")
    ref = c.ghost_fill_cylinder(u.copy(), deg + 1)
    assert np.array_equal(prob.download()[1:], ref[1:])
    prob.close()


def test_periodic_ghost_mode_on_the_rectangular_problem(FR):
    """FRB_GHOST_PERIODIC on both layouts of the rectangular path (reference image and row-chunk)."""
    nx, ny, deg = 64, 20, 3
    pr = FR.FRPSpace2D(0.0, 1.0, nx, 0.0, 1.0, ny, deg, 1, 1)
    po = o.FRPSpace2D(0.0, 1.0, nx, 0.0, 1.0, ny, deg, 1, 1)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 15)
    ref = o.integrate(u, 1e-4, 5, lambda w: o.rhs_euler2d(w, po, GAMMA), "ssprk3", before_step=c.ghost_fill_periodic)
    for kernel in ("auto", "generic"):
        prob = FR.Euler2DProblem(u, (0.0, 1.0), pr, GAMMA, kernel=kernel)
        itg = FR.init(prob, FR.SSPRK33(), dt=1e-4)
        itg.set_hooks(ghost="periodic")
        FR.step_(itg, 5)
        assert rel(itg.u, ref) <= 1e-11, kernel
        prob.close()


def test_limiter_and_filter_apply(FR):
    """Every euler2d call applies to the curvilinear problem: the limiter leaves an admissible state alone."""
    nx, ny, deg = 12, 9, 2
    ps, _, (n1, n2) = parallelogram(FR, nx, ny, deg)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 16)
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, n1, n2)
    assert prob.limiter(ps.wp / 4.0) == 0
    ref = o.positive_limiter_euler2d(u.copy(), GAMMA, ps.wp / 4.0, ps.ll, ps.lr)
    assert rel(prob.download()[1:-1, 1:-1], ref[1:-1, 1:-1]) <= 1e-13
    prob.close()


def test_unsupported_combinations_fail_loudly(FR):
    nx, ny, deg = 8, 6, 2
    ps, _, (n1, n2) = parallelogram(FR, nx, ny, deg)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 17)
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, n1, n2)
    with pytest.raises(FR.FRBError):
        prob.set_kernel("march")
    prob.set_kernel("curv_march")
    prob.set_metric("vertices")  # the marching kernel reads the stored metric only
    with pytest.raises(FR.FRBError):
        prob.f(np.zeros_like(u, order="F"), u, None, 0.0)
    prob.close()
    pr = FR.FRPSpace2D(0.0, 1.0, nx, 0.0, 1.0, ny, deg, 1, 1)
    p2 = FR.Euler2DProblem(u, (0.0, 1.0), pr, GAMMA)
    with pytest.raises(FR.FRBError):
        p2.set_hooks(ghost="cylinder")
    with pytest.raises(FR.FRBError):
        p2.set_kernel("curv_march")
    p2.close()


def test_large_mesh(FR):
    """512 x 256 sheared elements, p3: one residual against the oracle and the launch count."""
    nx, ny, deg = 512, 256, 3
    ps, po, (n1, n2) = parallelogram(FR, nx, ny, deg)
    x = po.xpg[..., 0]
    rho = 1.0 + 0.1 * np.sin(2 * np.pi * x)
    prim = np.stack([rho, np.ones_like(rho), 0.2 * np.ones_like(rho), rho], axis=-1)
    u = np.asfortranarray(o.prim_conserve(prim, GAMMA))
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, n1, n2)
    du = np.zeros_like(u, order="F")
    prob.f(du, u, None, 0.0)
    ref = c.rhs_euler2d_curv(u, po, n1, n2, GAMMA, fy_index="k")
    # a smooth state on a fine mesh: du = O(1) is the sum of terms of size |iJ| |F| |lpdm| = O(1e4), and that is
    # the scale rounding differences (FMA contraction, summation order) live on
    F, G = o.euler_flux(u, GAMMA)
    terms = np.abs(po.iJ).max() * max(np.abs(F).max(), np.abs(G).max()) * np.abs(po.dl).max()
    assert np.abs(du - ref).max() <= RTOL_RHS * terms
    assert rel(du, ref) <= 1e-9
    ms, n = prob.last_timing()
    assert n == 1  # one launch
    for kernel, launches in (("generic", 2), ("curv_march", 1)):
        prob.set_kernel(kernel)
        du2 = np.zeros_like(u, order="F")
        prob.f(du2, u, None, 0.0)
        assert np.abs(du2 - ref).max() <= RTOL_RHS * terms, kernel
        assert prob.last_timing()[1] == launches
    prob.close()


def test_golden_vectors(FR):
    """The committed fixtures of tests/golden/curv_golden.npz (make_curv_golden.py) through the C ABI."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "curv_golden.npz"))
    deg = 2
    ps, _, (n1, n2) = parallelogram(FR, 6, 4, deg)
    u = np.asfortranarray(g["para_u"])
    for fy in "kl":
        prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, n1, n2, corr="sp", fy_index=fy)
        du = np.zeros_like(u, order="F")
        prob.f(du, u, None, 0.0)
        assert rel(du, g[f"para_du_{fy}"]) <= RTOL_RHS
        if fy == "l":
            itg = FR.init(prob, FR.Euler(), dt=0.001)
            itg.set_hooks(ghost="periodic")
            FR.step_(itg, 5)
            assert rel(itg.u, g["para_u5"]) <= 1e-12
        prob.close()
    ps, _, _ = cylinder(FR, 5, 6, deg)
    u = np.asfortranarray(g["cyl_u"])
    for fy in "kl":
        prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, GAMMA, corr="fp", fy_index=fy, wall_xlo=True)
        du = np.zeros_like(u, order="F")
        prob.f(du, u, None, 0.0)
        assert rel(du, g[f"cyl_du_{fy}"]) <= RTOL_RHS
        prob.ghost_fill("cylinder")
        assert np.array_equal(prob.download()[1:], g["cyl_ghost"][1:])
        prob.close()
