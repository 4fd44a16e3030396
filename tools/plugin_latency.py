"""Latency of the single-scenario plugin path (BASELINE.json configs[0]: default case, n_z = 60, n_wl = 107,
and the same at 2100 bands) next to the CPU oracle port.  Run on the GPU box:  python tools/plugin_latency.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import crt_oracle as oracle  # noqa: E402

import crt1d_b200 as crt  # noqa: E402
from crt1d_b200 import sweep  # noqa: E402


def best(f, n=20):
    f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def main():
    rows = []
    spec = sweep.synthetic_sweep_spec(seed=0)
    big = spec.scenario_params(424242)
    for scheme in ("2s", "4s", "bf", "bl", "g77", "n79", "zq", "zq_pa"):
        m = crt.Model(scheme, nlayers=60)
        args = {k: m._p[k] for k in m.scheme["args"]}
        t_gpu = best(lambda: m.scheme["solver"](**args), 10)
        q = dict(big, clump=1.0)
        args_big = {k: q[k] for k in m.scheme["args"]}
        t_gpu_big = best(lambda: m.scheme["solver"](**args_big), 5)
        if scheme == "4s":
            t_cpu = t_cpu_big = float("nan")  # the oracle's solve_bvp takes ~1 s / ~20 s; see BASELINE.md
        else:
            t_cpu = best(lambda: oracle.run(scheme, m._p), 5)
            t_cpu_big = best(lambda: oracle.run(scheme, q), 3)
        rows.append(dict(scheme=scheme, gpu_ms_60x107=t_gpu, cpu_port_ms_60x107=t_cpu, gpu_ms_60x2100=t_gpu_big,
                         cpu_port_ms_60x2100=t_cpu_big))
        if scheme in ("bl", "n79", "zq", "zq_pa"):  # host prologue with the vectorised Gauss-Legendre tau_d (opt-in extension)
            from crt1d_b200.solvers import common

            common.use_gl_for_quad(True)
            rows[-1]["gpu_ms_60x107_gl_prologue"] = best(lambda: m.scheme["solver"](**args), 10)
            rows[-1]["gpu_ms_60x2100_gl_prologue"] = best(lambda: m.scheme["solver"](**args_big), 5)
            common.use_gl_for_quad(False)
        print(rows[-1])
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "plugin_latency.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
