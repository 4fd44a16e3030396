"""ref_4s_random.npz: the UNMODIFIED reference's solve_4s (solve_bvp forced to tol = 1e-11, same closures) on the
seeded random scenarios of tests/util.py::random_4s_batch, both mu_s.  Run HERE:
    python tests/golden/make_golden_4s_random.py"""
import os
import sys
import time

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
    S, _, _ = import_reference_solvers()
    tight = tight_4s_solver()
    args = S.AVAILABLE_SCHEMES["4s"]["args"]
    b = util.random_4s_batch()
    out = {}
    t0 = time.time()
    for mu_s in util.EDGE_4S_MU_S:
        for s in range(b.n_scen):
            q = b.scenario_params(s)
            sol = tight(**{k: q[k] for k in args}, mu_s=mu_s)
            for k, v in sol.items():
                out[f"mus{int(round(mu_s * 1000))}__s{s}__{k}"] = v
            print(mu_s, s, f"{time.time() - t0:.0f} s", flush=True)
    path = os.path.join(HERE, "ref_4s_random.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
