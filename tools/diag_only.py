"""Reduced-diagnostic mode timing: canopy-absorbed PAR/NIR of the 10^6-scenario sweep with NO profile output
(crt1d_out profile pointers NULL).  The row-sweep kernels then stop after the coefficient phase: the absorbed
reduction telescopes to the ground/top levels.  Run on the GPU box: python tools/diag_only.py [scheme ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from crt1d_b200 import sweep  # noqa: E402


def main(schemes):
    spec = sweep.synthetic_sweep_spec(seed=0)
    rows = []
    for scheme in schemes:
        full = sweep.SweepRunner(spec.slice(0, 66304), scheme, chunk=4144).upload()
        full.step()
        torch.cuda.synchronize()
        ref = full.absorbed.clone()
        r = sweep.SweepRunner(spec, scheme, chunk=16576, profiles=False).upload()
        for _ in range(3):
            r.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        same = float(((r.absorbed[:66304] - ref).abs() / ref.abs().clamp_min(1e-300)).max())
        rows.append(dict(scheme=scheme, ms_per_1e6_scenarios=ms, scenarios_per_s=spec.n_scen / ms * 1e3,
                         band_columns_per_s=spec.n_scen * spec.n_wl / ms * 1e3, max_rel_diff_vs_full_profile_run=same))
        print(rows[-1])
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "diag_only.json"), "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:] or ["2s", "4s", "bl", "bf", "g77"])
