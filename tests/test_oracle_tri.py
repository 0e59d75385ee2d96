"""The triangle operator builders of the oracle against the reference's OWN golden tables
(dev/check_phi.jl:60-89, asserted there at :110, :117, :125) and analytic properties."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fr_oracle_tri as T  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "tri_golden.npz"))
PERM = [0, 2, 4, 1, 3, 5]  # the point reordering check_phi.jl applies (:107-108, :121-123)


def test_vandermonde_at_solution_points_matches_check_phi():
    ops = T.tri_operators(2)
    assert np.allclose(ops["V"][PERM], G["py_V"], rtol=0, atol=5e-8)  # 8 printed digits


def test_vandermonde_at_flux_points_matches_check_phi():
    ops = T.tri_operators(2)
    Vf = ops["psif"].reshape(9, 6)  # [(face, point), mode], check_phi.jl:112-116
    assert np.allclose(Vf, G["py_Vf"], rtol=0, atol=5e-8)


def test_correction_field_matches_check_phi():
    phi = T.correction_field(2)  # [3, 3, 6]
    phi1 = phi.reshape(9, 6).T  # phi1[:, (i-1)*3+j] = phi[i, j, :], check_phi.jl:119-122
    assert np.allclose(phi1[PERM], G["phifj_ref"], rtol=0, atol=5e-7)


@pytest.mark.parametrize("deg", [1, 2, 3, 4])
def test_tri_operator_identities(deg):
    ops = T.tri_operators(deg)
    np_ = (deg + 1) * (deg + 2) // 2
    assert ops["V"].shape == (np_, np_) and abs(np.linalg.det(ops["V"])) > 1e-6
    # quadrature: weights sum to the area factor the reference uses, points inside the right triangle
    assert np.isclose(ops["wp"].sum(), 1.0)
    r, s = ops["xpl"][:, 0], ops["xpl"][:, 1]
    assert (r > -1).all() and (s > -1).all() and (r + s < 0).all()
    # nodal derivative matrices differentiate polynomials of degree <= deg exactly
    f = (1 + r) ** deg + 0.5 * s ** max(deg - 1, 0) * r
    fr = deg * (1 + r) ** (deg - 1) + 0.5 * s ** max(deg - 1, 0)
    fs = 0.5 * max(deg - 1, 0) * s ** max(deg - 2, 0) * r if deg >= 2 else np.zeros_like(r)
    assert np.allclose(ops["dl"][:, :, 0] @ f, fr, atol=1e-11)
    assert np.allclose(ops["dl"][:, :, 1] @ f, fs, atol=1e-11)
    # face interpolation reproduces the polynomial at the face Gauss points; partition of unity
    for face in range(3):
        rf, sf = ops["xfl"][face, :, 0], ops["xfl"][face, :, 1]
        assert np.allclose(ops["lf"][face] @ f, (1 + rf) ** deg + 0.5 * sf ** max(deg - 1, 0) * rf, atol=1e-11)
        assert np.allclose(ops["lf"][face].sum(axis=1), 1.0, atol=1e-12)
    # the default quadrature equals the one on explicit equilateral vertices (test/test_triangle.jl:15-16)
    p2, w2 = T.tri_quadrature(deg, vertices=((-1.0, -1 / np.sqrt(3)), (1.0, -1 / np.sqrt(3)), (0.0, 2 / np.sqrt(3))))
    assert np.array_equal(p2, ops["xpl"]) and np.array_equal(w2, ops["wp"])
