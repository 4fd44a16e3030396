"""ref_leaf_area.npz: the UNMODIFIED reference's non-beta LAI-profile generators (leaf_area.py:156-641) and its solvers on
the NON-UNIFORM cumulative-LAI axes they produce (weibull_z: unequal steps and zero-thickness layers below the crown
base; gamma: unequal first and last steps).  Run HERE:   python tests/golden/make_golden_leaf_area.py"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _refimport import import_reference_solvers  # noqa: E402
from _refimport import tight_4s_solver  # noqa: E402

import util  # noqa: E402


def main():
    S, _, LAREA = import_reference_solvers()
    tight = tight_4s_solver()
    out = {}
    warnings.simplefilter("ignore")
    t0 = time.time()
    # ---- generators
    for tag, fn, args, kw in util.LEAF_AREA_CASES:
        res = getattr(LAREA, fn)(*args, **kw)
        for k in ("lai", "lad", "z"):
            v = getattr(res, k)
            if v is not None:
                out[f"gen__{tag}__{k}"] = np.asarray(v, dtype=np.float64)
        print(tag, f"{time.time() - t0:.1f} s", flush=True)
    # ---- solvers on the non-uniform axes (12-band subset of the default case, like ref_variants.npz)
    for tag in util.NONUNIFORM_AXES:
        q = util.nonuniform_case(tag)
        for scheme in ("2s", "bf", "bl", "g77", "n79", "zq", "zq_pa", "4s_tight"):
            sd = S.AVAILABLE_SCHEMES["4s" if scheme == "4s_tight" else scheme]
            kw = {k: q[k] for k in sd["args"]}
            try:
                sol = tight(**kw) if scheme == "4s_tight" else sd["solver"](**kw)
            except Exception as e:  # noqa: BLE001
                out[f"sol__{tag}__{scheme}__raises"] = np.array(type(e).__name__)
                print(f"  {tag} {scheme}: reference raises {type(e).__name__}: {e}")
                continue
            for k, v in sol.items():
                out[f"sol__{tag}__{scheme}__{k}"] = v
            print(tag, scheme, f"{time.time() - t0:.1f} s", flush=True)
    path = os.path.join(HERE, "ref_leaf_area.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
