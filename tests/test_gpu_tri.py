"""2-D Euler on triangles (UnstructFRPSpace path, dev/sod.jl:31-130) against the oracle restatement,
whose operator builders are pinned by the reference's golden tables (tests/test_oracle_tri.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu
GAMMA = 5.0 / 3.0


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _case(T, oracle, nx, ny, deg, walls=False, seed=3):
    pts, cells = T.tri_mesh_rect(nx, ny, jitter=0.2, seed=seed)
    sp = T.tri_space(pts, cells, deg)
    if walls:  # dev/sod.jl:7-16: boundary cells at the bottom / top become mirror walls
        yc = pts[cells].mean(axis=1)[:, 1]
        sp["cellType"][(sp["cellType"] == 1) & ((yc < 1.0 / ny) | (yc > 1.0 - 1.0 / ny))] = 2
    x, y = sp["xpg"][..., 0], sp["xpg"][..., 1]
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
    prim = np.stack([rho, 0.5 + 0.1 * y, -0.25 + 0.1 * x, 1.0 / (0.8 + 0.2 * rho)], axis=-1)
    u = np.asfortranarray(oracle.prim_conserve(prim, GAMMA))
    return sp, u


def _problem(FR, sp, u):
    return FR.TriEulerProblem(u, (0.0, 0.1), sp["cellType"], sp["J"], sp["lf"], sp["normals"], sp["fpn"], sp["dl"],
                              sp["phi"], GAMMA, fpn_base=0)


@pytest.mark.parametrize("deg", [1, 2, 3])
@pytest.mark.parametrize("walls", [False, True])
def test_tri_euler_rhs(FR, oracle, deg, walls):
    import fr_oracle_tri as T

    sp, u = _case(T, oracle, 9, 7, deg, walls)
    prob = _problem(FR, sp, u)
    du = np.zeros_like(u, order="F")
    prob.f(du, u, None, 0.0)
    ref = T.rhs_tri_euler(u, sp, GAMMA)
    assert np.abs(ref).max() > 0.1
    assert rel(du, ref) <= 1e-12
    assert np.abs(du[sp["cellType"] == 1]).max() == 0.0  # frozen boundary cells
    prob.close()


def test_tri_euler_freestream(FR, oracle):
    import fr_oracle_tri as T

    sp, u = _case(T, oracle, 12, 10, 2)
    u[...] = oracle.prim_conserve(np.array([1.0, 0.4, -0.3, 0.8]), GAMMA)
    prob = _problem(FR, sp, u)
    du = np.zeros_like(u, order="F")
    prob.f(du, u, None, 0.0)
    assert np.abs(du).max() <= 1e-11
    prob.close()


@pytest.mark.parametrize("scheme", ["euler", "ssprk3"])
def test_tri_euler_steps(FR, oracle, scheme):
    """dev/sod.jl:126-136 steps with Euler(); SSPRK3 as well."""
    import fr_oracle_tri as T

    sp, u = _case(T, oracle, 8, 8, 2, walls=True)
    prob = _problem(FR, sp, u)
    alg = {"euler": FR.Euler, "ssprk3": FR.SSPRK33}[scheme]
    itg = FR.init(prob, alg(), dt=5e-4)
    FR.step_(itg, 100)
    ref = oracle.integrate(u.copy(order="F"), 5e-4, 100, lambda v: T.rhs_tri_euler(v, sp, GAMMA), scheme)
    assert np.isfinite(itg.u).all()
    assert rel(itg.u, ref) <= 1e-10
    prob.close()


def test_sod_from_a_mesh_file_through_the_package_space(FR, oracle, tmp_path):
    """The whole dev/sod.jl set-up through this package: TriFRPSpace(file, 2), wall retagging (:7-16),
    the Sod initial state (:19-30), ODEProblem + Euler() steps (:124-136) -- against the oracle running on
    ITS OWN space for the same mesh."""
    import fr_oracle_tri as T
    from test_unstruct import write_msh41

    pts, cells = T.tri_mesh_rect(24, 6, 0.0, 1.0, 0.0, 0.1, jitter=0.2, seed=5)
    f = str(tmp_path / "sod_like.msh")
    write_msh41(f, pts, cells)
    ps = FR.TriFRPSpace(f, 2)
    ct = ps.cellType.copy()
    ct[(ct == 1) & ((ps.cellCenter[:, 1] < 0.014) | (ps.cellCenter[:, 1] > 0.0835))] = 2
    left = ps.cellCenter[:, 0] < 0.5
    prim = np.where(left[:, None, None], np.array([1.0, 0.0, 0.0, 0.5]), np.array([0.3, 0.0, 0.0, 0.625]))
    u0 = np.asfortranarray(oracle.prim_conserve(np.broadcast_to(prim, (ct.size, ps.np, 4)), GAMMA))
    prob = FR.TriEulerProblem.from_space(ps, u0, (0.0, 0.1), GAMMA, cell_type=ct)
    itg = FR.init(prob, FR.Euler(), dt=5e-4)
    FR.step_(itg, 40)

    sp = T.tri_space(pts, cells, 2)
    sp["cellType"] = ct
    ref = oracle.integrate(u0.copy(order="F"), 5e-4, 40, lambda v: T.rhs_tri_euler(v, sp, GAMMA), "euler")
    assert np.isfinite(itg.u).all() and np.abs(itg.u - u0).max() > 1e-2
    assert rel(itg.u, ref) <= 1e-10
    prob.close()
