"""Golden for the layer-absorption epilogue: the reference's real Model.run + _calc_absorption
(recipe B), run in its own process because recipes A and B install different
`sys.modules["crt1d"]` objects.  Called by make_golden.py."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = os.environ.get("CRT1D_REFERENCE", "/root/reference")

for m in ("xarray", "matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
v = types.ModuleType("crt1d._version")
v.version = "0+ref"
sys.modules["crt1d._version"] = v
sys.path.insert(0, REF)
import crt1d  # noqa: E402
import crt1d.model  # noqa: E402

from crt1d_b200 import cases  # noqa: E402


def _case(nlayers):
    p = cases.load_default_case(nlayers)
    p.pop("leaf_angle")
    return p


crt1d.model.load_default_case = _case
out = {}
for scheme in ("2s", "bf", "zq"):
    m = crt1d.Model(scheme, nlayers=60).run().calc_absorption()
    for k, a in m.absorption.items():
        out[f"{scheme}__{k}"] = a
    for k, a in m.out.items():
        out[f"{scheme}__out_{k}"] = a
# Bonan case through Model (as the reference's test does), per-leaf-area sunlit/shaded absorbed PAR/NIR
pb = cases.load_bonan_sp1403_case()
for k in ("leaf_angle", "mla", "green", "orient"):
    pb.pop(k)
m = crt1d.Model(scheme="n79", mla=60.0, orient=1.0, **pb).run(tau_d_method="9sky").calc_absorption()
for k, a in m.absorption.items():
    out[f"bonan_n79__{k}"] = a
for k, a in m.out.items():
    out[f"bonan_n79__out_{k}"] = a
path = os.path.join(HERE, "ref_absorption.npz")
np.savez_compressed(path, **out)
print(f"  wrote ref_absorption.npz: {os.path.getsize(path)/1024:.0f} KB")
