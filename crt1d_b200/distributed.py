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


# Scenarios of a sweep do not all cost the same (4s: the rare-column fix-up and the exact-exponential path depend on the
# leaf optics and the zenith angle), and a cross-product sweep orders them by parameter, so contiguous blocks give the
# ranks different work: 4s strong scaling 0.96 at N = 4 with per-rank kernel times of 189..204 ms per step.  Dealing
# small units (one wave of 148 scenarios) round-robin gives every rank the same mix.
DEAL_UNIT = 148


def dealt_indices(n_scen, world_size, rank, unit=DEAL_UNIT):
    """Global scenario indices owned by `rank` when units of `unit` consecutive scenarios are dealt round-robin
    (unit u -> rank u % world_size); ascending.  Rank sizes differ by at most one unit."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    if unit < 1:
        raise ValueError("unit must be >= 1")
    n_units = -(-int(n_scen) // unit)
    mine = np.arange(rank, n_units, world_size, dtype=np.int64)
    idx = (mine[:, None] * unit + np.arange(unit, dtype=np.int64)[None, :]).ravel()
    return idx[idx < n_scen]


def deal_batch(batch, world_size, rank, unit=DEAL_UNIT):
    """This rank's scenarios of a ScenarioBatch under the round-robin deal, and their global indices."""
    idx = dealt_indices(batch.n_scen, world_size, rank, unit)
    return batch.take(idx), idx


def all_gather_dealt(local, n_total, world_size, unit=DEAL_UNIT, group=None, out=None):
    """Gather per-scenario rows (tensor `(n_local, k)`, owned as `dealt_indices`) from every rank into one
    `(n_total, k)` tensor in GLOBAL scenario order on every rank (pads to equal size, as all_gather needs, then
    scatters each rank's rows to their global positions)."""
    import torch
    import torch.distributed as dist

    k = local.shape[1]
    idx = [dealt_indices(n_total, world_size, r, unit) for r in range(world_size)]
    width = max(len(i) for i in idx)
    padded = torch.zeros((width, k), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    gathered = torch.empty((world_size * width, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    if out is None:
        out = torch.empty((n_total, k), dtype=local.dtype, device=local.device)
    for r, i in enumerate(idx):
        out[torch.as_tensor(i, device=local.device)] = gathered[r * width: r * width + len(i)]
    return out


class DealtGather:
    """`all_gather_dealt` with everything that does not change between calls made once: the padded send buffer, the
    receive buffer and the global positions of every rank's rows as ONE device index tensor, so a call is a copy, the
    NCCL all-gather and one `index_copy_` -- no host work, no allocation (it sits inside a timed region)."""

    def __init__(self, n_total, k, world_size, device, dtype=None, unit=DEAL_UNIT, group=None):
        import torch

        dtype = dtype or torch.float64
        idx = [dealt_indices(n_total, world_size, r, unit) for r in range(world_size)]
        self.width = max(len(i) for i in idx)
        self.group = group
        self.padded = torch.zeros((self.width, k), dtype=dtype, device=device)
        self.recv = torch.empty((world_size * self.width, k), dtype=dtype, device=device)
        # row r * width + j of `recv` is global scenario idx[r][j]; padding rows are dropped by `src`
        self.src = torch.as_tensor(np.concatenate([r * self.width + np.arange(len(i)) for r, i in enumerate(idx)]), device=device)
        self.dst = torch.as_tensor(np.concatenate(idx), device=device)

    def __call__(self, local, out):
        import torch.distributed as dist

        self.padded[: local.shape[0]].copy_(local)
        dist.all_gather_into_tensor(self.recv, self.padded, group=self.group)
        out.index_copy_(0, self.dst, self.recv.index_select(0, self.src))
        return out


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
