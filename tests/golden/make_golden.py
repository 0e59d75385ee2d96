"""Generates tests/golden/*.npz with the NumPy oracle (oracle/fr_oracle.py).

The reference is Julia + KitBase.jl and cannot run in this image, so these are NOT outputs of the
reference itself: they pin the oracle (and through it the CUDA path) against regressions, and hold
the known-answer operator values of SURVEY.md 8(a4).  Inputs are seeded; rerun with

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import fr_oracle as o  # noqa: E402

G = 5.0 / 3.0


def noisy(u, amp, seed):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(u * (1.0 + amp * rng.standard_normal(u.shape)))


def main():
    out = {}
    # operators (struct.jl:40-88) for deg 1..5
    for deg in range(1, 6):
        ps = o.FRPSpace1D(0.0, 1.0, 4, deg)
        for k in ("xpl", "wp", "ll", "lr", "dl", "dhl", "dhr", "dll", "dlr"):
            out[f"ops_deg{deg}_{k}"] = getattr(ps, k)
    np.savez_compressed(os.path.join(HERE, "operators.npz"), **out)

    # cfg1: 1-D advection, deg 2, 100 periodic cells
    ps = o.FRPSpace1D(-1.0, 1.0, 100, 2)
    u = o.ic_advection1d(ps)
    np.savez_compressed(
        os.path.join(HERE, "cfg1_advection.npz"), u=u,
        du_lowlevel=o.rhs_advection1d(u, ps, 1.0, "period", "lowlevel"),
        du_packaged=o.rhs_advection1d(u, ps, 1.0, "period", "packaged"),
        du_dirichlet=o.rhs_advection1d(u, ps, 1.0, "dirichlet", "packaged"),
        u_mid10=o.integrate(u, 1e-3, 10, lambda w: o.rhs_advection1d(w, ps, 1.0, "period", "lowlevel"), "midpoint"),
    )
    # cfg2 (shrunk): 1-D Euler Sod, deg 3, 128 cells
    ps = o.FRPSpace1D(0.0, 1.0, 128, 3)
    u = noisy(o.ic_sod1d(ps, G), 0.01, 21)
    ul = u.copy(order="F")
    o.positive_limiter_euler1d(ul, G, ps.wp / 2, ps.ll, ps.lr)
    np.savez_compressed(
        os.path.join(HERE, "cfg2_euler1d.npz"), u=u, du_dirichlet=o.rhs_euler1d(u, ps, G, "dirichlet"),
        du_period=o.rhs_euler1d(u, ps, G, "period"), u_limited=ul,
    )
    # cfg3 (shrunk): 2-D Euler wave, deg 3, 12 x 10
    ps = o.FRPSpace2D(0.0, 1.0, 12, 0.0, 1.0, 10, 3, 1, 1)
    u = noisy(o.ic_wave2d(ps, G, "x"), 0.01, 22)
    u[..., 2] += 0.1 * u[..., 0]
    o.ghost_fill_euler2d(u, "wave_x")
    u5 = o.integrate(u, 1e-3, 5, lambda w: o.rhs_euler2d(w, ps, G), "ssprk3", lambda w: o.ghost_fill_euler2d(w, "wave_x"))
    np.savez_compressed(os.path.join(HERE, "cfg3_euler2d.npz"), u=u, du=o.rhs_euler2d(u, ps, G), u_ssprk3_5=u5)
    # cfg4 (shrunk): BGK, deg 2, 16 cells x 32 velocities
    ps = o.FRPSpace1D(0.0, 1.0, 16, 2)
    v, w = o.vspace1d(-5.0, 5.0, 32)
    f0 = noisy(o.ic_bgk1d(ps, v), 0.01, 23)
    np.savez_compressed(os.path.join(HERE, "cfg4_bgk.npz"), f0=f0, velo=v, weights=w,
                        du=o.rhs_bgk1d(f0, ps.dx, v, w, ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr, 1e-2))


if __name__ == "__main__":
    main()
