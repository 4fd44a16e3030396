"""Multi-process path on CPU: world_size-2 gloo run of the sharding + gather logic used at N > 1."""
import os
import socket

import numpy as np
import pytest

from crt1d_b200 import distributed as cdist


def test_shard_bounds_partition():
    for n in (0, 1, 7, 100, 1_000_000):
        for w in (1, 2, 3, 4, 8):
            bounds = [cdist.shard_bounds(n, w, r) for r in range(w)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cdist.shard_bounds(10, 2, 2)
    for n in (0, 1, 147, 148, 149, 1000, 10**6):  # the deal: a partition too, sizes within one unit of each other
        for w in (1, 2, 3, 8):
            idx = [cdist.dealt_indices(n, w, r) for r in range(w)]
            assert np.array_equal(np.sort(np.concatenate(idx)), np.arange(n))
            assert max(len(i) for i in idx) - min(len(i) for i in idx) <= cdist.DEAL_UNIT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crt1d_b200 import sweep

        spec = sweep.synthetic_sweep_spec(seed=0, n_sza=3, n_lai=3, n_spec=3).slice(0, n_total)
        shard, (lo, hi) = cdist.shard_batch(spec, world, rank)
        assert shard.n_scen == hi - lo and np.array_equal(shard.psi, spec.psi[lo:hi])
        # stand-in for the per-scenario diagnostics a rank's kernels would produce: a function of the
        # global scenario index, so the gathered table can be checked against the unsharded answer
        local = torch.tensor([[float(s), float(spec.lai_idx[s])] for s in range(lo, hi)], dtype=torch.float64).reshape(-1, 2)
        full = cdist.all_gather_rows(local, n_total, world)
        want = np.array([[float(s), float(spec.lai_idx[s])] for s in range(n_total)]).reshape(-1, 2)
        ok = np.array_equal(full.numpy(), want)
        # the round-robin deal (strong scaling with content-dependent scenario cost): same check, global order restored
        for unit in (1, 4, 148):
            dealt, idx = cdist.deal_batch(spec, world, rank, unit)
            ok = ok and dealt.n_scen == len(idx) and np.array_equal(dealt.psi, spec.psi[idx])
            loc = torch.tensor([[float(s), float(spec.lai_idx[s])] for s in idx], dtype=torch.float64).reshape(-1, 2)
            ok = ok and np.array_equal(cdist.all_gather_dealt(loc, n_total, world, unit).numpy(), want)
            g = cdist.DealtGather(n_total, 2, world, "cpu", unit=unit)
            for _ in range(2):  # reusable
                ok = ok and np.array_equal(g(loc, torch.full((n_total, 2), -1.0, dtype=torch.float64)).numpy(), want)
        tot = cdist.all_reduce_sum(local.sum(0).clone())
        ok = ok and np.allclose(tot.numpy(), want.sum(0))
        ok = ok and cdist.max_over_ranks(float(rank + 1)) == float(world)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [27, 20, 1])
def test_world_size_2_gloo_shard_and_gather(n_total):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(2))
    assert got == [(0, True), (1, True)]
