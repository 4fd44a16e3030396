"""Golden fixtures for the 4s edge cases (omega -> 1, kappa = lambda_k) from the UNMODIFIED reference.

Run HERE (build container; needs /root/reference):   python tests/golden/make_golden_4s_edge.py
Writes ref_4s_edge.npz: for every (psi, mu_s) of tests/util.py::EDGE_4S_*, the reference's solve_4s as
shipped (tol = 1e-6) and with solve_bvp forced to tol = 1e-11 (same closures; _refimport.tight_4s_solver),
on the bands of util.edge_4s_case (inputs stored alongside so the GPU box needs neither scipy.optimize
nor the reference)."""
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
    out = {}
    t0 = time.time()
    for psi_deg in util.EDGE_4S_PSI_DEG:
        for mu_s in util.EDGE_4S_MU_S:
            q = util.edge_4s_case(psi_deg, mu_s)
            tag = f"psi{int(psi_deg)}_mus{int(round(mu_s * 1000))}"
            for k in ("psi", "lai", "leaf_r", "leaf_t", "soil_r", "I_dr0_all", "I_df0_all"):
                out[f"{tag}__in__{k}"] = np.asarray(q[k])
            kw = {k: q[k] for k in args}
            t1 = time.time()
            sol = S.AVAILABLE_SCHEMES["4s"]["solver"](**kw, mu_s=mu_s)
            t2 = time.time()
            solt = tight(**kw, mu_s=mu_s)
            print(f"{tag}: {q['leaf_r'].size} bands, shipped {t2 - t1:.1f} s, tight {time.time() - t2:.1f} s", flush=True)
            for k, v in sol.items():
                out[f"{tag}__shipped__{k}"] = v
            for k, v in solt.items():
                out[f"{tag}__tight__{k}"] = v
    path = os.path.join(HERE, "ref_4s_edge.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KB in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
