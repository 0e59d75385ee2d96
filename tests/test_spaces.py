"""Host-side FR spaces (fluxreconstruction.jl_b200/spaces.py) against the oracle's literal
restatement of src/struct.jl and src/Polynomial/* -- two independent constructions, the way
example/vandermonde_lagrange.jl:14-15,25 checks ll/lr/dl."""
import numpy as np
import pytest


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize("ng", [0, 1])
def test_frpspace1d_matches_oracle(FR, oracle, deg, ng):
    a = FR.FRPSpace1D(-1.0, 2.0, 17, deg, ng)
    b = oracle.FRPSpace1D(-1.0, 2.0, 17, deg, ng)
    for k in ("x", "dx", "J", "xpl", "xpg", "wp", "ll", "lr", "dl", "dll", "dlr", "dhl", "dhr", "V", "iV"):
        x, y = getattr(a, k), getattr(b, k)
        assert x.shape == y.shape, k
        assert np.allclose(x, y, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(y).max())), k
    assert a.np == deg + 1


@pytest.mark.parametrize("correction", ["radau", "sd", "huynh", ":radau"])
def test_correction_functions(FR, oracle, correction):
    a = FR.FRPSpace1D(0.0, 1.0, 5, 3, 0, correction)
    b = oracle.FRPSpace1D(0.0, 1.0, 5, 3, 0, correction.lstrip(":"))
    assert np.allclose(a.dhl, b.dhl, atol=1e-13) and np.allclose(a.dhr, b.dhr, atol=1e-13)


def test_frpspace2d_matches_oracle(FR, oracle):
    a = FR.FRPSpace2D(0.0, 1.0, 20, 0.0, 1.0, 30, 2, 1, 1)  # example/euler2d_wave.jl:25-27
    b = oracle.FRPSpace2D(0.0, 1.0, 20, 0.0, 1.0, 30, 2, 1, 1)
    assert a.xpg.shape == (22, 32, 3, 3, 2) and a.xpg.flags.f_contiguous
    for k in ("xpg", "wp", "ll", "lr", "dl", "dhl", "dhr", "dll", "dlr", "x", "y"):
        assert np.allclose(getattr(a, k), getattr(b, k), rtol=1e-13, atol=1e-13), k
    assert a.Jx == b.Jx and a.Jy == b.Jy and a.np == 9
    J = a.J
    assert J.shape == (22, 32, 3, 3, 2, 2) and J[3, 4, 1, 2, 0, 0] == a.Jx and J[3, 4, 1, 2, 0, 1] == 0.0


def test_keyword_constructors(FR):
    cfg = dict(x0=-1, x1=1, nx=100, deg=2, cfl=0.05, t=0.0, a=1.0)  # example/advection_lowlevel.jl:49-53
    ps = FR.FRPSpace1D(**cfg)
    assert ps.nx == 100 and ps.deg == 2 and ps.xpg.shape == (100, 3)
    ps2 = FR.FRPSpace2D(x0=0, x1=1, nx=4, y0=0, y1=1, ny=6, deg=3, ngx=1, ngy=1)
    assert ps2.xpg.shape == (6, 8, 4, 4, 2)


def test_triangle_space_needs_a_readable_mesh(FR):
    with pytest.raises(OSError):
        FR.TriFRPSpace("../assets/does_not_exist.msh", 2)  # test/runtests.jl:22 builds one from assets/


def test_interp_and_derivative_identities(FR):
    """ll/lr interpolate and dl differentiates polynomials of degree <= deg exactly."""
    ps = FR.FRPSpace1D(0.0, 1.0, 4, 4)
    r = ps.xpl
    for q in range(5):
        assert abs(r**q @ ps.ll - (-1.0) ** q) < 1e-13 and abs(r**q @ ps.lr - 1.0) < 1e-13
        d = ps.dl @ r**q
        assert np.allclose(d, q * r ** max(q - 1, 0) if q else 0 * r, atol=1e-12)


def test_rectangular_space_carries_the_pointwise_metric(FR):
    """ps.iJ, ps.Ji and ps.vertices exist on the rectangular FRPSpace2D too (struct.jl:99-128; example/shock-vortex.jl
    reads ps.iJ) and equal what FRPSpace2D(base, deg) computes from the vertices."""
    ps = FR.FRPSpace2D(0.0, 1.0, 6, 0.0, 0.5, 4, 2, 1, 1)
    pb = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, 6, 0.0, 0.5, 4, 1, 1), 2)
    assert ps.iJ.shape == (8, 6, 3, 3, 2, 2) and ps.Ji.shape == (8, 6, 4, 3, 2, 2)
    for name in ("J", "iJ", "Ji", "vertices", "xpg"):
        a, b = getattr(ps, name), getattr(pb, name)
        assert np.abs(a - b).max() <= 1e-14 * np.abs(b).max(), name
    n1, n2 = FR.face_normals(ps.vertices)
    assert np.allclose(n1, [1.0, 0.0]) and np.allclose(n2, [0.0, 1.0])
