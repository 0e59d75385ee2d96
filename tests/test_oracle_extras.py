"""CPU checks of the oracle pieces added for the widened scope: LF / Roe common fluxes (defined in
DESIGN.md section 2), the shock sensor + modal filter pass, the triangle Euler residual."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fr_oracle as o  # noqa: E402
import fr_oracle_tri as T  # noqa: E402

G = 5.0 / 3.0


@pytest.mark.parametrize("flux", ["hll", "lf", "roe"])
def test_common_fluxes_are_consistent_and_mirror_symmetric(flux):
    rng = np.random.default_rng(1)
    prim = np.stack([1 + 0.3 * rng.random(50), rng.normal(0, 0.8, 50), rng.normal(0, 0.5, 50), 0.6 + 0.4 * rng.random(50)], axis=-1)
    wL = o.prim_conserve(prim, G)
    wR = o.prim_conserve(prim[::-1] * np.array([1.1, 0.9, 1.0, 1.05]), G)
    f = o.RIEMANN[flux]
    assert np.abs(f(wL, wL, G) - o.euler_flux(wL, G)[0]).max() <= 1e-14  # f(w, w) = F(w)
    # mirror the x axis: states swap sides, normal momentum flips -> mass/energy/tangential fluxes flip
    m = np.array([1.0, -1.0, 1.0, 1.0])
    a, b = f(wL, wR, G), f(wR * m, wL * m, G)
    assert np.abs(a + b * m).max() <= 1e-13
    # 1-D = 2-D with zero tangential momentum
    w3L, w3R = wL[:, [0, 1, 3]].copy(), wR[:, [0, 1, 3]].copy()
    w4L, w4R = wL.copy(), wR.copy()
    for w4, w3 in ((w4L, w3L), (w4R, w3R)):
        w3[:, 2] = w4[:, 3] - 0.5 * w4[:, 2] ** 2 / w4[:, 0]  # remove the tangential kinetic energy
        w4[:, 3] = w3[:, 2]
        w4[:, 2] = 0.0
    assert np.abs(f(w3L, w3R, G) - f(w4L, w4R, G)[:, [0, 1, 3]]).max() <= 1e-13


def test_roe_and_lf_upwind_supersonic_states():
    wl = o.prim_conserve(np.array([1.0, 3.0, 0.1, 0.8]), G)
    wr = o.prim_conserve(np.array([0.9, 3.2, -0.1, 0.7]), G)
    assert np.abs(o.flux_roe(wl, wr, G) - o.euler_flux(wl, G)[0]).max() <= 1e-13
    assert np.abs(o.flux_hll(wl, wr, G) - o.euler_flux(wl, G)[0]).max() <= 1e-13


@pytest.mark.parametrize("flux", ["lf", "roe"])
def test_euler2d_other_fluxes_freestream_and_conservation(flux):
    ps = o.FRPSpace2D(0.0, 1.0, 8, 0.0, 1.0, 6, 3, 1, 1)
    u = np.empty((10, 8, 4, 4, 4), order="F")
    u[...] = o.prim_conserve(np.array([1.0, 0.3, -0.2, 0.8]), G)
    assert np.abs(o.rhs_euler2d(u, ps, G, flux=flux)).max() <= 1e-11
    u = o.ic_wave2d(ps, G, "x")
    o.ghost_fill_euler2d(u, "wave_x")
    du = o.rhs_euler2d(u, ps, G, flux=flux)
    for m in (0, 1, 3):
        assert abs(np.einsum("ijkl,kl->", du[1:-1, 1:-1, :, :, m], ps.wp)) <= 1e-11 * np.abs(du).max() * 48


def test_shock_detector_branches():
    S0 = -3.0 * np.log10(3)
    assert not o.shock_detector(S0 - 5.0, 3)          # smooth: sigma = 1
    assert o.shock_detector(S0 + 5.0, 3)              # rough: sigma = 0
    assert o.shock_detector(S0, 3)                    # sigma = 0.5
    assert not o.shock_detector(-np.inf, 3)           # su = 0
    assert o.shock_detector(np.nan, 3)                # 0/0: falls into the else branch of dissipation.jl:19-21


def test_modal_filter_pass_keeps_means_and_smooth_cells():
    ps = o.FRPSpace1D(0.0, 1.0, 40, 3)
    u = o.ic_wave1d(ps, G, amp=1e-6)
    u[ps.xpg[ps.ng: ps.ng + 40] > 0.5031] *= 0.4
    ref = u.copy()
    n = o.filter_pass_1d(u, ps.V, ps.iV, ps.deg, 1e-2)
    assert 0 < n < 5
    changed = np.abs(u - ref).max(axis=(1, 2)) > 0
    assert changed.sum() == n
    w = ps.wp / 2  # the l2 filter keeps mode 0, i.e. the cell mean
    assert np.abs(np.einsum("ipk,p->ik", u - ref, w)).max() <= 1e-14


def test_triangle_residual_freestream_and_convergence():
    errs = []
    for n in (4, 8):
        pts, cells = T.tri_mesh_rect(n, n, jitter=0.1)
        sp = T.tri_space(pts, cells, 2)
        if n == 4:
            u = np.zeros((len(cells), sp["np"], 4))
            u[:] = o.prim_conserve(np.array([1.0, 0.4, -0.3, 0.8]), G)
            assert np.abs(T.rhs_tri_euler(u, sp, G)).max() <= 1e-12
            # every interior flux point has exactly one partner and the pairing is an involution
            fpn = sp["fpn"]
            for i, j, k in zip(*np.nonzero(fpn[..., 0] >= 0)):
                ni, nj, nk = fpn[i, j, k]
                assert tuple(fpn[ni, nj, nk]) == (i, j, k)
        x, y = sp["xpg"][..., 0], sp["xpg"][..., 1]
        rho = 1 + 0.1 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
        prim = np.stack([rho, 0.5 * np.ones_like(rho), 0.25 * np.ones_like(rho), rho], axis=-1)  # p = 1/2
        du = T.rhs_tri_euler(o.prim_conserve(prim, G), sp, G)
        exact = -(0.5 * 0.2 * np.pi * np.cos(2 * np.pi * x) * np.cos(2 * np.pi * y)
                  - 0.25 * 0.2 * np.pi * np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y))
        act = sp["cellType"] == 0
        errs.append(np.abs(du[act, :, 0] - exact[act]).max())
    assert errs[1] < 0.7 * errs[0]  # the derivative of a degree-2 reconstruction converges


# ------------------------------------------------------------------ explicit RK tableaus (Tsit5, RK4)
def _order_conditions(A, b):
    A, b = np.array(A, dtype=np.float64), np.array(b, dtype=np.float64)
    c = A.sum(axis=1)
    Ac = A @ c
    return [
        (b.sum(), 1), (b @ c, 1 / 2), (b @ c**2, 1 / 3), (b @ Ac, 1 / 6),
        (b @ c**3, 1 / 4), (b @ (c * Ac), 1 / 8), (b @ (A @ c**2), 1 / 12), (b @ (A @ Ac), 1 / 24),
        (b @ c**4, 1 / 5), (b @ (c**2 * Ac), 1 / 10), (b @ (Ac * Ac), 1 / 20), (b @ (c * (A @ c**2)), 1 / 15),
        (b @ (A @ c**3), 1 / 20), (b @ (c * (A @ Ac)), 1 / 30), (b @ (A @ (c * Ac)), 1 / 40),
        (b @ (A @ (A @ c**2)), 1 / 60), (b @ (A @ (A @ Ac)), 1 / 120),
    ]


def test_tsit5_tableau_satisfies_all_order_conditions_up_to_5():
    """No Julia here to print OrdinaryDiffEq's table: the 17 rooted-tree conditions pin the digits."""
    import fr_oracle as O
    import frb200 as FR

    for A, b in (O.RK_TABLEAUS["tsit5"], FR.Tsit5().tableau):
        for got, want in _order_conditions(A, b):
            assert abs(got - want) < 3e-14  # coefficients of size 12 cancel
        c = np.array(A).sum(axis=1)
        assert np.allclose(c, [0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0], atol=1e-15)
    for got, want in _order_conditions(*O.RK_TABLEAUS["rk4"])[:8]:
        assert abs(got - want) < 1e-15
    assert np.array_equal(FR.RK4().A, np.array(O.RK_TABLEAUS["rk4"][0]))


def test_tableau_stepping_converges_at_design_order():
    import fr_oracle as O

    rhs = lambda u: np.array([u[1], -u[0]])  # noqa: E731  harmonic oscillator
    exact = np.array([np.sin(1.0), np.cos(1.0)])
    for scheme, order in (("rk4", 4), ("tsit5", 5)):
        err = []
        for n in (8, 16, 32):
            u = O.integrate(np.array([0.0, 1.0]), 1.0 / n, n, rhs, scheme)
            err.append(np.abs(u - exact).max())
        rates = np.log2(np.array(err[:-1]) / np.array(err[1:]))
        assert (rates > order - 0.3).all(), (scheme, err, rates)


# ---------------------------------------------------------------- example/advection_kinetic.jl
def test_kinetic_advection_model(oracle):
    """mol! of example/advection_kinetic.jl:73-128 = the BGK residual with the Maxwellian of [rho, a, 1]:
    the relaxation term vanishes on that Maxwellian's own moments, mass is conserved, and the transport part
    equals the bgk_wave residual's (the two differ by the relaxation term only)."""
    ps = oracle.FRPSpace1D(-1.0, 1.0, 40, 2)
    velo, wts = oracle.vspace1d(-5.0, 5.0, 28)  # advection_kinetic.jl:13-16,24
    f0 = oracle.ic_kinetic_advection1d(ps, velo, 1.0)
    args = (ps.dx, velo, wts, ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr)
    du = oracle.rhs_bgk1d(f0, *args, 2e-3, model="advection", a=1.0)
    du_slow = oracle.rhs_bgk1d(f0, *args, 2e3, model="advection", a=1.0)
    # the midpoint rule on 28 nodes reproduces rho to ~1e-10, so the stiff term is small against 1/tau
    assert np.abs(du - du_slow).max() < 1e-6 / 2e-3
    # mass of the transport part (periodic, conservative interface flux); the relaxation term conserves mass only
    # up to the quadrature error of the discrete Maxwellian (28 midpoint nodes), which 1/tau = 500 amplifies
    mass = np.einsum("ijk,j,k->", du_slow, wts, ps.wp) * ps.J[0]
    assert abs(mass) < 1e-9
    assert abs(np.einsum("ijk,j,k->", du, wts, ps.wp) * ps.J[0]) < 1e-4
    # against the bgk_wave model on the same state only the Maxwellian changes
    d_bgk = oracle.rhs_bgk1d(f0, *args, 2e-3)
    k = 1
    w = oracle.moments_conserve_1v(f0[:, :, k], velo, wts)
    M_adv = oracle.maxwellian(velo[None, :], np.stack([w[:, 0], np.ones(40), np.ones(40)], axis=-1))
    M_bgk = oracle.maxwellian(velo[None, :], oracle.conserve_prim(w, 3.0))
    assert np.allclose((du - d_bgk)[:, :, k], (M_adv - M_bgk) / 2e-3, rtol=1e-9, atol=1e-9)


def test_factored_slope_moments_equal_the_literal_form():
    """csrc/frb_ns2d.cu evaluates moments_conserve_slope(sl, Mu, Mv, Mxi, a, 0) in a factored form (all six calls of
    ns_cavity.jl:75-145 have beta = 0): E_i = (m_(i+2) + A0 m_i)/2, F = (V1 m_2 + B0 m_0)/2,
    Q = (m_4 + 2 A0 m_2 + C0 m_0)/4 with A0 = V2 + X1, B0 = V3 + V1 X1, C0 = V4 + X2 + 2 V2 X1.  The same algebra
    in NumPy against the literal six-term sum of the oracle, incl. the mixed (Mv of one state, Mxi of the other) call."""
    import numpy as np
    import fr_oracle as o

    rng = np.random.default_rng(0)
    n = 500
    mk = lambda: np.stack([1 + 0.2 * rng.random(n), 0.3 * rng.standard_normal(n), 0.3 * rng.standard_normal(n),  # noqa: E731
                           1 + 0.3 * rng.random(n)], -1)
    Mu, Mv, Mxi, MuL, _ = o.gauss_moments(mk(), 1.0)
    Mu2, Mv2, Mxi2, _, MuR2 = o.gauss_moments(mk(), 1.0)
    sl = rng.standard_normal((n, 4))

    def fast(sl, m, a, V, X):
        A0, B0, C0 = V[2] + X[1], V[3] + V[1] * X[1], V[4] + X[2] + 2 * V[2] * X[1]
        E = lambda i: 0.5 * (m[a + i + 2] + A0 * m[a + i])  # noqa: E731
        F = 0.5 * (V[1] * m[a + 2] + B0 * m[a])
        Q = 0.25 * (m[a + 4] + 2 * A0 * m[a + 2] + C0 * m[a])
        s0, s1, s2, s3 = (sl[:, q] for q in range(4))
        k = s0 + s2 * V[1]
        return np.stack([k * m[a] + s1 * m[a + 1] + s3 * E(0), k * m[a + 1] + s1 * m[a + 2] + s3 * E(1),
                         (s0 * V[1] + s2 * V[2]) * m[a] + s1 * V[1] * m[a + 1] + s3 * F,
                         s0 * E(0) + s1 * E(1) + s2 * F + s3 * Q], -1)

    for M, V, X, a in ((Mu, Mv, Mxi, 1), (Mu2, Mv, Mxi2, 1), (MuL, Mv, Mxi, 2), (MuR2, Mv2, Mxi2, 2), (MuR2, Mv2, Mxi2, 1)):
        ref = o.moments_conserve_slope_2d(sl, M, V, X, a, 0)
        assert np.abs(fast(sl, M, a, V, X) - ref).max() <= 4e-15 * np.abs(ref).max()
