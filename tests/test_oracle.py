"""CPU tests of the oracle itself: known answers, analytic invariants, C == NumPy restatement,
and the committed golden vectors.  (PARITY UNPINNED: no reference runtime exists in this image;
see oracle/fr_oracle.py.)"""
import os

import numpy as np
import pytest

G = 5.0 / 3.0
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_known_answer_operators(oracle):
    """SURVEY.md 8(a4) values (reference formulas poly_lagrange.jl:6-59, poly_legendre.jl:29-37)."""
    ps = oracle.FRPSpace1D(0, 1, 100, 2)
    assert np.allclose(ps.ll, [1.478830557701236, -0.6666666666666665, 0.1878361089654305], rtol=0, atol=1e-15)
    assert np.allclose(ps.dhl, [-2.6618950038622256, 0.75, -0.33810499613777534], rtol=0, atol=2e-15)
    assert np.allclose(ps.dl[0], [-1.9364916731037085, 2.581988897471611, -0.6454972243679028], rtol=0, atol=2e-15)
    ps = oracle.FRPSpace1D(0, 1, 100, 3)
    assert np.allclose(ps.ll, [1.5267881254572664, -0.813632449486927, 0.40076152031165035, -0.11391719628198993],
                       rtol=0, atol=2e-15)
    assert np.allclose(ps.dhl, [-4.3891529665310856, 1.247624770988934, -0.6145280959667937, 0.3274848629375171],
                       rtol=0, atol=5e-15)
    assert np.allclose(ps.dhr, -ps.dhl[::-1], atol=5e-15) and np.allclose(ps.lr, ps.ll[::-1], atol=2e-15)


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5, 7])
def test_vandermonde_two_way_check(oracle, deg):
    """The reference's own assertion, example/vandermonde_lagrange.jl:14-15,25."""
    ps = oracle.FRPSpace1D(0, 1, 100, deg)
    V = oracle.vandermonde_matrix(deg, ps.xpl)
    psi = oracle.vandermonde_matrix(deg, np.array([-1.0, 1.0]))
    Vr = oracle.dvandermonde_matrix(deg, ps.xpl)
    assert np.allclose(np.linalg.solve(V.T, psi[0]), ps.ll, atol=1e-12)
    assert np.allclose(np.linalg.solve(V.T, psi[1]), ps.lr, atol=1e-12)
    dl = np.array([np.linalg.solve(V.T, Vr[i]) for i in range(deg + 1)])
    assert np.allclose(dl, ps.dl, atol=1e-11)
    assert np.allclose(ps.dl.sum(axis=1), 0.0, atol=1e-12)  # derivative of a constant


def test_physics_identities(oracle):
    rng = np.random.default_rng(0)
    prim = np.stack([1 + rng.random(50), rng.standard_normal(50), rng.standard_normal(50), 0.5 + rng.random(50)], -1)
    w = oracle.prim_conserve(prim, G)
    assert np.allclose(oracle.conserve_prim(w, G), prim, rtol=1e-13)
    assert np.allclose(oracle.flux_hll(w, w, G), oracle.euler_flux(w, G)[0], rtol=1e-13, atol=1e-14)
    assert np.allclose(oracle.global_frame(oracle.local_frame(w, 0.0, 1.0), 0.0, 1.0), w)
    # supersonic to the right / left: pure upwinding
    wl = oracle.prim_conserve(np.array([1.0, 5.0, 0.0, 1.0]), G)
    wr = oracle.prim_conserve(np.array([0.8, 5.0, 0.0, 1.0]), G)
    assert np.array_equal(oracle.flux_hll(wl, wr, G), oracle.euler_flux(wl, G)[0])
    wl[1] *= -1
    wr[1] *= -1
    assert np.array_equal(oracle.flux_hll(wl, wr, G), oracle.euler_flux(wr, G)[0])


def test_advection_rhs_is_consistent(oracle):
    errs = []
    for n in (25, 50, 100):
        ps = oracle.FRPSpace1D(-1, 1, n, 2)
        du = oracle.rhs_advection1d(oracle.ic_advection1d(ps), ps, 1.0, "period", "lowlevel")
        errs.append(np.abs(du + np.pi * np.cos(np.pi * ps.xpg)).max())
    assert errs[0] / errs[1] > 3.5 and errs[1] / errs[2] > 3.5  # order >= p = 2 for the derivative


def test_euler1d_freestream_convergence_conservation(oracle):
    ps = oracle.FRPSpace1D(0, 1, 32, 3)
    u = np.empty((32, 4, 3), order="F")
    u[...] = oracle.prim_conserve(np.array([1.0, 0.3, 0.8]), G)
    assert np.abs(oracle.rhs_euler1d(u, ps, G, "period")).max() < 1e-12
    errs = []
    for n in (8, 16, 32):
        ps = oracle.FRPSpace1D(0, 1, n, 3)
        rho = 1 + 0.2 * np.sin(2 * np.pi * ps.xpg)
        u = oracle.prim_conserve(np.stack([rho, np.ones_like(rho), rho], -1), G)
        du = oracle.rhs_euler1d(u, ps, G, "period")
        errs.append(np.abs(du[:, :, 0] + 0.2 * 2 * np.pi * np.cos(2 * np.pi * ps.xpg)).max())
        tot = np.einsum("ipk,p->k", du, ps.wp) * ps.J[0]
        assert np.abs(tot).max() < 1e-12  # periodic FR is conservative
    assert errs[0] / errs[1] > 6 and errs[1] / errs[2] > 6  # order ~3


def test_euler2d_freestream_and_directions(oracle):
    ps = oracle.FRPSpace2D(0, 1, 6, 0, 2, 5, 3, 1, 1)
    u = np.empty((8, 7, 4, 4, 4), order="F")
    u[...] = oracle.prim_conserve(np.array([1.0, 0.3, -0.2, 0.8]), G)
    assert np.abs(oracle.rhs_euler2d(u, ps, G)).max() < 1e-11
    for d, mode in (("x", "wave_x"), ("y", "wave_y")):
        errs = []
        for nx, ny in ((4, 6), (8, 12)):
            ps = oracle.FRPSpace2D(0, 1, nx, 0, 1, ny, 3, 1, 1)
            u = oracle.ic_wave2d(ps, G, d)
            oracle.ghost_fill_euler2d(u, mode)
            du = oracle.rhs_euler2d(u, ps, G)
            ex = -0.1 * 2 * np.pi * np.cos(2 * np.pi * ps.xpg[..., 0 if d == "x" else 1])
            errs.append(np.abs(du[1:-1, 1:-1, :, :, 0] - ex[1:-1, 1:-1]).max())
        assert errs[0] / errs[1] > 5


def test_euler2d_wave_returns_to_ic(oracle, coracle):
    """Analytic check of the stepping semantics: the density wave of euler2d_wave.jl travels one
    period (velocity 1, domain 1) and comes back."""
    ps = oracle.FRPSpace2D(0, 1, 10, 0, 1, 4, 3, 1, 1)
    u0 = oracle.ic_wave2d(ps, G, "x")
    errs = []
    for dt, n in ((0.002, 500), (0.001, 1000)):
        u1 = coracle.integrate_euler2d(u0, ps, G, dt, n, "midpoint", "wave_x")
        errs.append(np.abs(u1[1:-1, 1:-1] - u0[1:-1, 1:-1]).max())
    # the ghosts are refreshed once per step and frozen through the midpoint stage
    # (euler2d_wave.jl:125-135), so the periodic seam carries an O(dt) error: ~5.3 * dt here
    assert errs[0] < 1.2e-2 and 1.8 < errs[0] / errs[1] < 2.2


def test_c_equals_numpy(oracle, coracle):
    rng = np.random.default_rng(3)
    ps = oracle.FRPSpace1D(-1, 1, 40, 2)
    u = np.asfortranarray(oracle.ic_advection1d(ps) + 0.1 * rng.standard_normal((40, 3)))
    for bc, var in (("period", "packaged"), ("period", "lowlevel"), ("dirichlet", "packaged")):
        assert np.array_equal(coracle.rhs_advection1d(u, ps, 1.0, bc, var), oracle.rhs_advection1d(u, ps, 1.0, bc, var))
    ps = oracle.FRPSpace1D(0, 1, 40, 3)
    u = np.asfortranarray(oracle.ic_sod1d(ps, G) * (1 + 0.02 * rng.standard_normal((40, 4, 3))))
    for bc in ("period", "dirichlet"):
        assert np.array_equal(coracle.rhs_euler1d(u, ps, G, bc), oracle.rhs_euler1d(u, ps, G, bc))
    ps = oracle.FRPSpace2D(0, 1, 9, 0, 1, 7, 3, 1, 1)
    u = np.asfortranarray(oracle.ic_wave2d(ps, G, "x") * (1 + 0.02 * rng.standard_normal((11, 9, 4, 4, 4))))
    assert np.array_equal(coracle.rhs_euler2d(u, ps, G), oracle.rhs_euler2d(u, ps, G))
    a, b = u.copy(order="F"), u.copy(order="F")
    coracle.limiter_euler2d(a, G, ps.wp / 4, ps.ll, ps.lr)
    oracle.positive_limiter_euler2d(b, G, ps.wp / 4, ps.ll, ps.lr)
    assert np.allclose(a, b, rtol=1e-14, atol=1e-15)
    ps = oracle.FRPSpace1D(0, 1, 12, 2)
    v, w = oracle.vspace1d(-5, 5, 24)
    f0 = oracle.ic_bgk1d(ps, v)
    args = (ps.dx, v, w, ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr)
    assert np.allclose(coracle.rhs_bgk1d(f0, *args), oracle.rhs_bgk1d(f0, *args), rtol=1e-12, atol=1e-13)
    ops = coracle.operators(3)
    ps = oracle.FRPSpace1D(0, 1, 4, 3)
    for k, a in (("ll", ps.ll), ("lr", ps.lr), ("lpdm", ps.dl), ("dgl", ps.dhl), ("dgr", ps.dhr), ("w", ps.wp)):
        assert np.allclose(ops[k], a, atol=5e-15)


def test_limiter_keeps_mean_and_constant_states(oracle):
    """test/runtests.jl:25-26 calls the limiter on constant states; it must leave them alone."""
    ps = oracle.FRPSpace1D(0, 1, 20, 5)
    u = np.ones((20, 6, 3), order="F")
    u[..., 2] = 2.0
    before = u.copy()
    assert oracle.positive_limiter_euler1d(u, G, np.full(6, 1 / 6), ps.ll, ps.lr) == 0
    assert np.allclose(u, before)
    rng = np.random.default_rng(1)
    ps = oracle.FRPSpace1D(0, 1, 30, 3)
    u = np.asfortranarray(oracle.ic_sod1d(ps, G) * (1 + 0.3 * rng.standard_normal((30, 4, 3))))
    m0 = (u[..., 0] * ps.wp / 2).sum(1)
    oracle.positive_limiter_euler1d(u, G, ps.wp / 2, ps.ll, ps.lr)
    assert np.allclose((u[..., 0] * ps.wp / 2).sum(1), m0, rtol=1e-13)
    assert (u[..., 0] > 0).all()


def test_golden_vectors(oracle, coracle):
    g = np.load(os.path.join(GOLD, "operators.npz"))
    for deg in range(1, 6):
        ps = oracle.FRPSpace1D(0.0, 1.0, 4, deg)
        for k in ("xpl", "wp", "ll", "lr", "dl", "dhl", "dhr", "dll", "dlr"):
            assert np.allclose(getattr(ps, k), g[f"ops_deg{deg}_{k}"], rtol=0, atol=1e-13)
    g = np.load(os.path.join(GOLD, "cfg1_advection.npz"))
    ps = oracle.FRPSpace1D(-1.0, 1.0, 100, 2)
    u = np.asfortranarray(g["u"])
    assert np.allclose(coracle.rhs_advection1d(u, ps, 1.0, "period", "lowlevel"), g["du_lowlevel"], rtol=1e-13, atol=1e-13)
    assert np.allclose(coracle.rhs_advection1d(u, ps, 1.0, "period", "packaged"), g["du_packaged"], rtol=1e-13, atol=1e-13)
    assert np.allclose(coracle.rhs_advection1d(u, ps, 1.0, "dirichlet", "packaged"), g["du_dirichlet"], rtol=1e-13, atol=1e-13)
    assert np.allclose(coracle.integrate_advection1d(u, ps, 1.0, "period", "lowlevel", 1e-3, 10, "midpoint"), g["u_mid10"],
                       rtol=1e-12, atol=1e-13)
    g = np.load(os.path.join(GOLD, "cfg2_euler1d.npz"))
    ps = oracle.FRPSpace1D(0.0, 1.0, 128, 3)
    u = np.asfortranarray(g["u"])
    assert np.allclose(coracle.rhs_euler1d(u, ps, G, "dirichlet"), g["du_dirichlet"], rtol=1e-12, atol=1e-10)
    assert np.allclose(coracle.rhs_euler1d(u, ps, G, "period"), g["du_period"], rtol=1e-12, atol=1e-10)
    ul = u.copy(order="F")
    coracle.limiter_euler1d(ul, G, ps.wp / 2, ps.ll, ps.lr)
    assert np.allclose(ul, g["u_limited"], rtol=1e-13)
    g = np.load(os.path.join(GOLD, "cfg3_euler2d.npz"))
    ps = oracle.FRPSpace2D(0.0, 1.0, 12, 0.0, 1.0, 10, 3, 1, 1)
    u = np.asfortranarray(g["u"])
    assert np.allclose(coracle.rhs_euler2d(u, ps, G), g["du"], rtol=1e-12, atol=1e-10)
    assert np.allclose(coracle.integrate_euler2d(u, ps, G, 1e-3, 5, "ssprk3", "wave_x"), g["u_ssprk3_5"], rtol=1e-12)
    g = np.load(os.path.join(GOLD, "cfg4_bgk.npz"))
    ps = oracle.FRPSpace1D(0.0, 1.0, 16, 2)
    du = coracle.rhs_bgk1d(np.asfortranarray(g["f0"]), ps.dx, g["velo"], g["weights"], ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr, 1e-2)
    assert np.allclose(du, g["du"], rtol=1e-11, atol=1e-12)


def test_long_double_arbiter_agrees_with_the_fp64_oracle():
    """oracle/fr_arbiter.c evaluates selected cells in x87 long double, cell-locally, with its own code for the
    fluxes: a third, structurally independent evaluation.  On well-conditioned data it must agree with the FP64
    oracle to rounding (a few ulp of the summed terms), for the 2-D Euler residual and the BGK residual."""
    import c_oracle
    import fr_oracle as o

    g = 5.0 / 3.0
    ps = o.FRPSpace2D(0.0, 1.0, 20, 0.0, 1.0, 30, 3, 1, 1)
    rng = np.random.default_rng(1)
    u0 = o.ic_wave2d(ps, g, "x")
    u0 = np.asfortranarray(u0 * (1.0 + 0.02 * rng.standard_normal(u0.shape)))
    u0[..., 2] += 0.1 * u0[..., 0]
    o.ghost_fill_euler2d(u0, "wave_x")
    ref = c_oracle.rhs_euler2d(u0, ps, g)
    cells = [(i, j) for i in (1, 2, 10, 20) for j in (1, 15, 30)]
    ex = c_oracle.arbiter_euler2d_cells(u0, ps, g, cells)
    for c, (i, j) in enumerate(cells):
        assert np.abs(ex[c] - ref[i, j]).max() <= 1e-14 * np.abs(ref).max()
    # supersonic states: the upwind branches of HLL
    prim = np.empty(ps.xpg.shape[:-1] + (4,))
    prim[..., 0] = 1.0 + 0.1 * np.sin(2 * np.pi * ps.xpg[..., 0]) * np.cos(2 * np.pi * ps.xpg[..., 1])
    prim[..., 1], prim[..., 2], prim[..., 3] = -3.0, 2.5, 1.0
    us = np.asfortranarray(o.prim_conserve(prim, g))
    ref = c_oracle.rhs_euler2d(us, ps, g)
    ex = c_oracle.arbiter_euler2d_cells(us, ps, g, cells)
    term = np.abs((us[..., 3] + 0.5) * 3.0).max() / min(ps.Jx, ps.Jy)  # the residual is a difference of terms this size
    for c, (i, j) in enumerate(cells):
        assert np.abs(ex[c] - ref[i, j]).max() <= 1e-14 * term
    ps1 = o.FRPSpace1D(0.0, 1.0, 64, 2)
    velo, w = o.vspace1d(-5.0, 5.0, 64)
    f0 = o.ic_bgk1d(ps1, velo) * (1.0 + 0.05 * rng.standard_normal((64, 64, 3)))
    dx = np.full(64, 1.0 / 64)
    args = (dx, velo, w, ps1.ll, ps1.lr, ps1.dl, ps1.dhl, ps1.dhr, 1e-2)
    du = c_oracle.rhs_bgk1d(f0, *args)
    ex = c_oracle.arbiter_bgk1d_cells(f0, *args, [0, 1, 31, 63])
    for c, i in enumerate([0, 1, 31, 63]):
        assert np.abs(ex[c] - du[i]).max() <= 1e-13 * np.abs(du).max()
