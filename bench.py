#!/usr/bin/env python
"""bench.py -- headline benchmark of the crt1d solver hot path on B200.

Metric (BASELINE.json): layer.band solves per second -- one unit = all required fields at one
interface level of one wavelength band of one scenario.  Workload at N = 1: BASELINE.json configs[2],
the batched 2s sweep of 10^6 scenarios (100 SZA x 100 LAI x 100 leaf/soil/sky spectra) x 2100 one-nm
bands x 60 levels, synthetic; one STEP = one full pass over the sweep
(1.26e11 units, 4.03 TB of fp64 profiles written to HBM in chunks + per-scenario absorbed PAR/NIR).
N > 1 (torchrun, one rank per GPU): every rank runs its own 10^6-scenario sweep (seed = rank) --
weak scaling, no data-path collective; the per-scenario diagnostics are all-gathered over NCCL
inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scheme 2s|4s|...]

Prints ONE JSON line on rank 0 (see the contract in the task description).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "layer_band_solves_per_s"
UNIT = "layer*band solves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scheme", default="2s")
    ap.add_argument("--scenarios", type=int, default=1_000_000, help="scenarios per GPU (cross product is truncated)")
    ap.add_argument("--chunk", type=int, default=4144, help="scenarios per kernel launch")
    ap.add_argument("--nz", type=int, default=60, help="canopy levels (1000 = the deep-canopy case, BASELINE.json configs[4])")
    ap.add_argument("--cpu-sample", type=int, default=0, help="scenarios in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every rank runs its own --scenarios sweep; strong: ONE sweep of --scenarios "
                         "block-partitioned over the ranks (BASELINE.json configs[3])")
    ap.add_argument("--profile-dtype", default="f64", choices=["f64", "f32"],
                    help="storage type of the profiles in HBM (arithmetic is always float64); f32 = the optional "
                         "reduced-precision path of BASELINE.json, NOT the headline configuration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host-sample", type=int, default=256,
                    help="scenarios in the host-profile e2e leg (crt1d_solve_host, profiles copied back to host); 0 = skip")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_spec(seed, n_scen, n_z=60):
    from crt1d_b200 import sweep

    spec = sweep.synthetic_sweep_spec(seed=seed, n_z=n_z)
    return spec if n_scen >= spec.n_scen else spec.slice(0, n_scen)


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (numpy restatement of the reference solver) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    scheme, seed, idx, n_z = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import crt_oracle as oracle

    spec = make_spec(seed, 10**9, n_z)
    t0 = time.perf_counter()
    units = 0
    for s in idx:
        sol = oracle.run(scheme, spec.scenario_params(int(s)))
        units += sol["F"].size
    return units, time.perf_counter() - t0


def cpu_baseline(scheme, n_sample, cores=None, n_z=60):
    """Time the oracle port on a strided sample of the sweep, one process per host core."""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    if n_sample <= 0:  # ~10-30 s of CPU work: per-scenario cost of the vectorised port, measured roughly
        per = {"2s": 0.012, "bf": 0.012, "g77": 0.012, "bl": 0.15, "zq": 0.45, "n79": 0.25, "4s": 25.0}.get(scheme, 0.05)
        n_sample = int(max(cores, min(4096, 15.0 * cores / (per * n_z / 60.0))))
    idx = np.linspace(0, 999_999, n_sample).astype(np.int64)
    parts = [idx[i::cores] for i in range(cores)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(scheme, 0, p, n_z) for p in parts if len(p)])
    wall = time.perf_counter() - t0
    units = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return {
        "value": units / busy, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{n_sample} scenarios strided over the 10^6-scenario sweep, {scheme}, 2100 bands x {n_z} levels, "
                  f"numpy oracle port, {cores} processes; busy {busy:.1f} s (wall incl. spawn {wall:.1f} s)",
    }


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path.  The reference is pure Python
    and /root/reference does not exist on the GPU box, so this times the oracle PORT (a band-vectorised
    numpy restatement, faster than the reference's per-band Python loop) with all host cores."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_baseline(args.scheme, max(8, (os.cpu_count() or 1)), n_z=args.nz)
    cb = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cb = cpu_baseline(args.scheme, args.cpu_sample, n_z=args.nz)
        vals.append(cb["value"])
    ms = (time.perf_counter() - t0) * 1e3 / max(1, args.steps)
    v = float(np.mean(vals))
    cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "scheme": args.scheme, "n_z": args.nz, "n_wl": 2100,
                   "note": "each step = bounded strided sample of the sweep on the host cores"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(args):
    return (f"batched {args.scheme} sweep: {args.scenarios} scenarios (SZA x LAI x PROSPECT-style spectra) x 2100 "
            f"1-nm bands x {args.nz} levels per GPU (BASELINE.json configs[{2 if args.nz == 60 else 4}])")


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from crt1d_b200 import distributed as cdist
    from crt1d_b200 import sweep

    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.scaling == "strong":  # one sweep, contiguous scenario blocks per rank (distributed.shard_batch)
        full = make_spec(0, args.scenarios, args.nz)
        spec, _ = cdist.shard_batch(full, world, rank)
        n_total_strong = full.n_scen
    else:  # weak scaling: every rank owns a full sweep (seed = rank)
        spec = make_spec(rank, args.scenarios, args.nz)
        n_total_strong = None
    pdt = torch.float32 if args.profile_dtype == "f32" else None
    runner = sweep.SweepRunner(spec, args.scheme, chunk=args.chunk, device=dev, profile_dtype=pdt)
    runner.upload()
    S = spec.n_scen
    n_bw = runner.band_w.shape[0]
    S_total = n_total_strong if args.scaling == "strong" else world * S
    gathered = None

    def one_step(events=None):
        nonlocal gathered
        n = runner.step(events)
        if world > 1:  # diagnostics of every rank's scenarios on every rank (2 doubles per scenario)
            gathered = cdist.all_gather_rows(runner.absorbed, S_total, world) if args.scaling == "strong" else _gather_equal()
        return n

    _gbuf = torch.empty((world * S, n_bw), dtype=torch.float64, device=dev) if world > 1 and args.scaling == "weak" else None

    def _gather_equal():
        dist.all_gather_into_tensor(_gbuf, runner.absorbed)
        return _gbuf

    for _ in range(args.warmup):
        one_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    events = []
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    launches = 0
    for _ in range(args.steps):
        launches += one_step(events)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = cdist.max_over_ranks(e0.elapsed_time(e1), device=dev)
    ms_per_step = ms_total / args.steps
    units_per_step = S_total * spec.n_z * spec.n_wl  # all ranks' scenarios
    value = units_per_step / (ms_per_step * 1e-3)

    # dominant kernel: average launch duration from the per-launch CUDA event pairs of the timed region
    full = [a.elapsed_time(b) for (a, b), (v, _) in zip(events, runner._calls * args.steps) if v.batch.n_scen == runner.chunk]
    k_ms = float(np.mean(full)) if full else ms_per_step / runner.n_chunks
    bpu = runner.algorithmic_bytes_per_unit()
    units_per_launch = runner.chunk * spec.n_z * spec.n_wl
    achieved = units_per_launch * bpu / (k_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.profile_dtype == "f64":
        try:
            traffic = json.load(open(tpath)).get(f"{args.scheme}_chunk{runner.chunk}")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "kernel": ("crt::solve_2s_rows_kernel<VEC=2, LV=10, 512 threads, REC> (one CTA per scenario, row-major work items)"
                   if args.scheme == "2s" and runner.chunk >= 148 and not os.environ.get("CRT1D_B200_2S_KERNEL")
                   else f"crt::solve_kernel<{args.scheme}, VEC=2> (band-tile kernel)"),
        "kernel_ms": k_ms, "bytes_per_unit": bpu,
        "units_per_launch": units_per_launch, "peak_source": peak_src,
        "kernel_share_of_step": k_ms * runner.n_chunks / ms_per_step,
    }

    # ---------------- end to end through the public API: host tables in, host diagnostics out
    e2e = None
    if not args.no_e2e:
        pinned_out = torch.empty((S, n_bw), dtype=torch.float64).pin_memory()
        r2 = sweep.SweepRunner(spec, args.scheme, chunk=args.chunk, device=dev, profile_dtype=pdt)
        r2.ring = runner.ring  # reuse the HBM profile ring (allocation is not part of a step)
        r2.pin_host()          # inputs staged once in pinned host memory; every step copies them H2D
        r2.band_w_d, r2.absorbed = runner.band_w_d, runner.absorbed

        def e2e_step():
            r2.upload()  # H2D of psi, index arrays and the spectra/LAI libraries + device prologue kernels
            r2.step()
            pinned_out.copy_(r2.absorbed, non_blocking=True)  # D2H of the per-scenario diagnostics
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = cdist.max_over_ranks((time.perf_counter() - t0) / n_e2e, device=dev)
        e2e = {
            "value": units_per_step / dt, "unit": UNIT, "h2d_bytes_per_step": int(r2.db.h2d_bytes) * world,
            "d2h_bytes_per_step": int(pinned_out.numel() * 8) * world,
            "api": "crt1d_b200.sweep.SweepRunner.upload()+step(): host scenario tables -> HBM profiles (chunk ring) "
                   "-> per-scenario absorbed PAR/NIR back in pinned host memory",
            "ms_per_step": dt * 1e3,
        }

    # ---------------- the reference-facing call with full profiles returned to the HOST (what one reference
    # solver call returns, batched): crt1d_solve_host = H2D + kernel + D2H of every profile inside one C call.
    # 4 TB per sweep cannot cross PCIe, so this leg runs a bounded sample of the same sweep and is PCIe-bound.
    e2e_host = None
    if not args.no_e2e and world == 1 and args.host_sample > 0 and args.profile_dtype == "f64":  # N = 1 only, like cpu_baseline
        from crt1d_b200 import engine
        from crt1d_b200.solvers._plugin import solve_batch_host

        n_host = min(args.host_sample, S)
        lo = max(0, min(S // 2, S - n_host))
        sub = spec.slice(lo, lo + n_host)
        pro = engine.host_prologue(sub, args.scheme)
        n_fields = 4 + len(engine.EXTRA_NAMES.get(args.scheme, ()))
        pool = torch.empty((n_fields, n_host, spec.n_z, spec.n_wl), dtype=torch.float64).pin_memory()
        pool_np, taken = pool.numpy(), [0]

        def pinned_alloc(shape):  # page-locked output arrays, reused by every call
            a = pool_np[taken[0] % n_fields].reshape(-1)[:int(np.prod(shape))].reshape(shape)
            taken[0] += 1
            return a

        legs = {}
        for name, alloc in (("pageable", np.empty), ("pinned", pinned_alloc)):
            solve_batch_host(sub, args.scheme, pro, device=local_rank, alloc=alloc)  # warm-up (workspace allocation)
            t0 = time.perf_counter()
            res = solve_batch_host(sub, args.scheme, pro, device=local_rank, alloc=alloc)
            dt_h = time.perf_counter() - t0
            d2h = int(sum(v.nbytes for v in res.values()))
            legs[name] = {"value": n_host * spec.n_z * spec.n_wl / dt_h, "ms": dt_h * 1e3, "d2h_gb_per_s": d2h / dt_h / 1e9}
        e2e_host = {
            "value": legs["pinned"]["value"], "unit": UNIT, "scenarios": n_host, "d2h_bytes": d2h,
            "pinned_output_arrays": legs["pinned"], "fresh_pageable_output_arrays": legs["pageable"],
            "api": "crt1d_solve_host (C ABI, host pointers in and out): every profile of every scenario returned to "
                   "host memory; bound by the D2H copy (PCIe with page-locked outputs; page faults of fresh numpy "
                   "arrays, as the reference-style plugin call allocates them, otherwise)",
        }

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args.scheme, args.cpu_sample, n_z=args.nz)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args), "scheme": args.scheme, "scenarios_per_gpu": S, "n_z": spec.n_z,
                "n_wl": spec.n_wl, "chunk": runner.chunk, "launches_per_step": runner.n_chunks,
                "profile_bytes_per_step_per_gpu": int(S * spec.n_z * spec.n_wl * (bpu - 5 * 8.0 / spec.n_z - 8.0 / spec.n_wl)),
                "profile_storage": args.profile_dtype,
                "l2": "no flush needed: each launch writes a %.1f GB profile chunk (>> 126 MB L2), 2-buffer ring" % (
                    runner.chunk * spec.n_z * spec.n_wl * bpu / 1e9),
                "parallelism": f"scenario-sharded x{world}, no data-path collective; NCCL all-gather of absorbed[S,2]",
            },
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "e2e_host_profiles": e2e_host, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
