"""The CPU arm of bench.py: the reference's own solver functions timed on the host cores.

`kind = "reference"`: the UNMODIFIED zmoon/crt1d `solve_<id>` functions from `baseline/_ref` (staged by
`baseline/stage_reference.py` during `build()`), one call per scenario exactly as `Model.run` makes it
(ref model.py:305-310), with the reference's own `leaf_angle.G_ellipsoidal_approx` as `G_fn` / `K_b_fn`.
`import crt1d` itself needs xarray / matplotlib (absent here), so the solver sub-package is imported through a stub
parent package that skips `crt1d/__init__.py` (recipe A) -- the solver code runs unmodified.
`kind = "port"` (stated fallback when `baseline/_ref` is missing): the band-vectorised numpy oracle, ~30x faster per
core than the shipped code.

NOT product code: imported by bench.py's `--impl reference` and `cpu_baseline` legs only.  One worker process per
host core (spawn), whole scenarios per task, BLAS threads pinned to 1.
"""
import math
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_PKG = os.path.join(HERE, "_ref", "crt1d")

# seconds per 2100-band, 60-level scenario on one core (sizes the samples only)
REF_COST = {"2s": 0.113, "bf": 0.111, "g77": 0.108, "bl": 0.30, "zq": 1.82, "n79": 1.25, "zq_pa": 8.25, "4s": 20.6}
PORT_COST = {"2s": 0.012, "bf": 0.012, "g77": 0.012, "bl": 0.15, "zq": 0.45, "n79": 0.25, "zq_pa": 1.0, "4s": 25.0}


def reference_staged():
    return os.path.isfile(os.path.join(REF_PKG, "solvers", "_solve_2s.py")) and os.path.isfile(os.path.join(REF_PKG, "variables.yml"))


def import_staged_reference():
    """(solvers, leaf_angle) modules of the staged reference."""
    if "crt1d" not in sys.modules or not getattr(sys.modules["crt1d"], "_refarm_stub", False):
        pkg = types.ModuleType("crt1d")
        pkg.__path__ = [REF_PKG]
        pkg._refarm_stub = True
        sys.modules["crt1d"] = pkg
    import crt1d.leaf_angle as leaf_angle
    import crt1d.solvers as solvers

    return solvers, leaf_angle


_W = {}


def _init_worker(kind, n_z):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from crt1d_b200 import sweep  # host-side numpy generator of the synthetic sweep (inputs only)

    _W["spec"] = sweep.synthetic_sweep_spec(seed=0, n_z=n_z)
    _W["kind"] = kind
    if kind == "reference":
        S, LA = import_staged_reference()
        x = LA.mla_to_x_approx(57)
        G_fn = lambda psi: LA.G_ellipsoidal_approx(psi, x)  # noqa: E731
        _W["schemes"] = S.AVAILABLE_SCHEMES
        _W["fn"] = {"G_fn": G_fn, "K_b_fn": lambda psi: G_fn(psi) / np.cos(psi)}
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import crt_oracle

        _W["oracle"] = crt_oracle


def _solve_one(scheme, s):
    q = _W["spec"].scenario_params(int(s))
    if _W["kind"] == "reference":
        q.update(_W["fn"])
        sd = _W["schemes"][scheme]
        return sd["solver"](**{k: q[k] for k in sd["args"]})
    return _W["oracle"].run(scheme, q)


def _work(task):
    scheme, idx = task
    import warnings

    t0 = time.perf_counter()
    units = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for s in idx:
            units += _solve_one(scheme, s)["F"].size
    return units, time.perf_counter() - t0


class RefArm:
    """A pool of one process per core that stays up across steps (spawn + imports are not part of a step)."""

    def __init__(self, scheme, n_z=60, cores=None, kind="auto"):
        import multiprocessing as mp

        self.scheme, self.n_z = scheme, n_z
        self.cores = cores or os.cpu_count() or 1
        self.kind = ("reference" if reference_staged() else "port") if kind == "auto" else kind
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_init_worker, initargs=(self.kind, n_z))
        self.pool.map(_work, [(scheme, [])] * self.cores)  # workers up, modules imported

    def auto_sample(self, wall_s):
        per = (REF_COST if self.kind == "reference" else PORT_COST).get(self.scheme, 0.1) * self.n_z / 60.0
        return int(max(self.cores, min(4096, wall_s * self.cores / per)))

    def step(self, n_sample):
        """One bounded sample: `n_sample` scenarios strided over the 10^6-scenario sweep, split over the cores.
        Returns (units per second = units / slowest worker's busy time, units, busy seconds)."""
        idx = np.linspace(0, 999_999, n_sample).astype(np.int64)
        parts = [idx[i::self.cores] for i in range(self.cores)]
        res = self.pool.map(_work, [(self.scheme, p) for p in parts if len(p)], chunksize=1)
        units, busy = sum(r[0] for r in res), max(r[1] for r in res)
        return units / busy, units, busy

    def describe(self, n_sample, busy):
        what = ("the reference's own solve_%s (unmodified zmoon/crt1d from baseline/_ref, one call per scenario)" % self.scheme
                if self.kind == "reference" else "numpy oracle port (baseline/_ref not staged)")
        return (f"{n_sample} scenarios strided over the 10^6-scenario sweep, {self.scheme}, 2100 bands x {self.n_z} levels, "
                f"{what}, {self.cores} processes; busy {busy:.1f} s")

    def close(self):
        self.pool.close()
        self.pool.join()


def cfg1_default_case(repeats=5):
    """BASELINE.json configs[0]: the default case (n_z = 60, n_wl = 107), scheme 2s, ONE core, best of `repeats`
    (BASELINE.md section 4).  In-process; reference code when staged, else the port."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from crt1d_b200 import cases

    p = dict(cases.load_default_case(60))
    kind = "reference" if reference_staged() else "port"
    if kind == "reference":
        S, LA = import_staged_reference()
        x = LA.mla_to_x_approx(57)
        G_fn = lambda psi: LA.G_ellipsoidal_approx(psi, x)  # noqa: E731
        p["G_fn"], p["K_b_fn"] = G_fn, (lambda psi: G_fn(psi) / np.cos(psi))
        sd = S.AVAILABLE_SCHEMES["2s"]
        call = lambda: sd["solver"](**{k: p[k] for k in sd["args"]})  # noqa: E731
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import crt_oracle

        G_fn = p["G_fn"]
        p["K_b_fn"] = lambda psi: G_fn(psi) / np.cos(psi)
        call = lambda: crt_oracle.run("2s", p)  # noqa: E731
    best = math.inf
    sol = None
    for _ in range(repeats + 1):  # first run warms imports / caches
        t0 = time.perf_counter()
        sol = call()
        best = min(best, time.perf_counter() - t0)
    units = sol["F"].size
    return {"workload": "crt1d default case, scheme 2s, n_z = 60, n_wl = 107 (BASELINE.json configs[0])", "kind": kind,
            "cores": 1, "best_of": repeats, "ms": best * 1e3, "value": units / best, "unit": "layer*band solves/s",
            "checksum_F": float(sol["F"].sum())}
