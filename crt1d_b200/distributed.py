"""Multi-GPU execution: one process per GPU, scenarios sharded, diagnostics gathered.

The path shards naturally: every (scenario, band) column is independent, so
ranks own disjoint contiguous blocks of the scenario axis and NO collective sits on the data path.
`torch.distributed` is used for exactly two things: the barrier / max-over-ranks around timed regions
and an all-gather of the per-scenario diagnostics (2 doubles per scenario) at the end -- NCCL over
NVLink on GPUs, gloo on CPU for the host-logic tests.  Full profiles never move between GPUs.
"""
import numpy as np


def shard_bounds(n_scen, world_size, rank):
    """Contiguous block [lo, hi) of the scenario axis owned by `rank`; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_scen), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch, world_size, rank):
    """This rank's slice of a ScenarioBatch (libraries are shared/replicated, indices are sliced)."""
    lo, hi = shard_bounds(batch.n_scen, world_size, rank)
    return batch.slice(lo, hi), (lo, hi)


def all_gather_rows(local, n_total, world_size, group=None):
    """Gather per-scenario rows (tensor `(n_local, k)`, block-partitioned as `shard_bounds`) from every
    rank into one `(n_total, k)` tensor on every rank.  Pads to equal block size, as all_gather needs."""
    import torch
    import torch.distributed as dist

    k = local.shape[1]
    sizes = [shard_bounds(n_total, world_size, r) for r in range(world_size)]
    width = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((width, k), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    gathered = torch.empty((world_size * width, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = [gathered[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def all_reduce_sum(t, group=None):
    """Ensemble sums of diagnostics across ranks (in place)."""
    import torch.distributed as dist

    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value, device=None):
    """max of a Python float over all ranks (timing rule: a multi-GPU time is the slowest rank's)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def expected_gather(rows_by_rank):
    """Reference result of `all_gather_rows` for tests: plain concatenation in rank order."""
    return np.concatenate(rows_by_rank, axis=0)
