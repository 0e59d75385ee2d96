"""Generates tests/golden/curv_golden.npz with the NumPy restatement of the curvilinear scratch scripts
(oracle/fr_oracle_curv.py: dev/parallelogram.jl:80-165, dev/cylinder2.jl:52-187).

Like the other fixtures of this directory these are NOT outputs of the reference (Julia + KitBase.jl do not run
in this image): they pin the oracle -- and through it the CUDA path -- against regressions.  Seeded inputs; rerun

    python tests/golden/make_curv_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import fr_oracle as o  # noqa: E402
import fr_oracle_curv as c  # noqa: E402

G = 5.0 / 3.0


def rand_state(shape, seed):
    rng = np.random.default_rng(seed)
    prim = np.empty(shape + (4,))
    prim[..., 0] = 1.0 + 0.2 * rng.random(shape)
    prim[..., 1] = 0.3 + 0.1 * rng.standard_normal(shape)
    prim[..., 2] = 0.3 + 0.1 * rng.standard_normal(shape)
    prim[..., 3] = 1.0 + 0.2 * rng.random(shape)
    return np.asfortranarray(o.prim_conserve(prim, G))


def main():
    out = {}
    deg = 2
    # dev/parallelogram.jl: 6 x 4 cells of the 45-degree mesh, factors from the solution-point iJ
    nx, ny = 6, 4
    ps = c.CurvSpace2D(c.parallelogram_vertices(nx, ny), deg)
    n1, n2 = c.parallelogram_normals(nx, ny)
    u = rand_state((nx + 2, ny + 2, deg + 1, deg + 1), 20261017)
    out["para_u"] = u
    for fy in "kl":
        out[f"para_du_{fy}"] = c.rhs_euler2d_curv(u, ps, n1, n2, G, corr="sp", fy_index=fy)
    w = u.copy()
    for _ in range(5):  # parallelogram.jl:200-208 (Euler, dt = 0.001)
        c.ghost_fill_periodic(w)
        w = w + 0.001 * c.rhs_euler2d_curv(w, ps, n1, n2, G, corr="sp", fy_index="l")
    out["para_u5"] = w
    # dev/cylinder2.jl: 5 x 6 polar cells, flux-point factors from the literal Ji, mirror wall
    nr, nth = 5, 6
    vv, dth = c.cspace2d_vertices(1.0, 6.0, nr, 0.0, np.pi, nth, 0, 1)
    ps = c.CurvSpace2D(c.embed_cylinder(vv), deg)
    n1, n2 = c.cylinder_normals(nr, nth, dth[0])
    n1, n2 = n1[:nr], n2[: nr - 1]
    fpc = c.corr_factors_fp(ps.Ji, n1, n2)
    u = rand_state((nr + 1, nth + 2, deg + 1, deg + 1), 20261018)
    out["cyl_u"] = u
    out["cyl_fpc"] = fpc
    for fy in "kl":
        out[f"cyl_du_{fy}"] = c.rhs_euler2d_curv(u, ps, n1, n2, G, corr="fp", fpc=fpc, fy_index=fy, wall_xlo=True)
    out["cyl_ghost"] = c.ghost_fill_cylinder(u.copy(), deg + 1)
    np.savez_compressed(os.path.join(HERE, "curv_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
