"""Generates tests/golden/tri_golden.npz from the reference tree (run in the build container only):

* the Williams-Shunn-Jameson points / weights the reference obtains through PyCall from its own
  src/Quadrature/qpmin.py (tri_quadrature, quadrature.jl:14-71) for schemes 1..5 (deg 0..4), in the
  trilinear coordinates qpmin returns;
* the 8-digit golden tables of dev/check_phi.jl:60-89 (py_V 6x6, py_Vf 9x6, phifj_ref 6x9), which
  the reference asserts with `@test ... ≈` at :110,117,125 for the degree-2 triangle.

    python tests/golden/make_tri_golden.py [/root/reference]
"""
import os
import re
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, os.path.join(ref, "src", "Quadrature"))
sys.dont_write_bytecode = True
import qpmin  # noqa: E402  (the reference's own trimmed quadpy)

out = {}
for n in range(1, 6):
    sch = getattr(qpmin, f"williams_shunn_jameson_{n}")()
    out[f"wsj{n}_points"] = np.asarray(sch.points, dtype=np.float64)
    out[f"wsj{n}_weights"] = np.asarray(sch.weights, dtype=np.float64)

txt = open(os.path.join(ref, "dev", "check_phi.jl")).read()
for name, shape in (("py_V", (6, 6)), ("py_Vf", (9, 6)), ("phifj_ref", (6, 9))):
    m = re.search(name + r"\s*=\s*\[(.*?)\]", txt, re.S)
    vals = [float(x) for x in m.group(1).split()]
    out[name] = np.array(vals).reshape(shape)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tri_golden.npz")
np.savez(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items()})
