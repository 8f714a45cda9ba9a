"""Data-parallel seam of the update path: replicate the six networks,
shard the minibatch across ranks, exchange

  1. the two per-critic sums of std (inputs of the mean_std EMA that every sample's
     TD bound and gradient weight depend on, reference dsac_v2.py:233-241) — 2 floats, SUM;
  2. the flat gradient buffer [q1 | q2 | policy | log_alpha] — one all-reduce, SUM
     (each rank already scales its loss terms by 1/global_batch);
  3. the logged accumulators (16 sums, 2 minima).

Collectives go through `torch.distributed` (NCCL over NVLink on GPUs; gloo in the CPU
tests).  `engine` is anything with grad_phase1 / grad_phase2 / state / grads — the CUDA
`Engine`, or a CPU stand-in in tests/test_dp_gloo.py.
"""
from __future__ import annotations

from . import _lib


def world(dist_module=None):
    """(dist, world_size) if a multi-rank process group is live, else (None, 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_world_size()
    return None, 1


def shard_rows(n_rows: int, rank: int, world_size: int):
    """Contiguous [lo, hi) slice of a global minibatch owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def data_parallel_gradients(engine, data, noise, dist, local_rows: int, global_rows: int) -> int:
    """Forward, exchange std sums, losses + backward with means over `global_rows`, exchange gradients."""
    engine.grad_phase1(data, noise)
    dist.all_reduce(engine.state[_lib.STATE_STDSUM:_lib.STATE_STDSUM + 2])
    engine.grad_phase2(global_rows)
    dist.all_reduce(engine.grads)
    dist.all_reduce(engine.state[_lib.STATE_ACC:_lib.STATE_ACC + 16])
    dist.all_reduce(engine.state[_lib.STATE_ACC + 16:_lib.STATE_ACC + 18], op=dist.ReduceOp.MIN)
    return global_rows


def global_rows(dist, local_rows: int, device) -> int:
    """Sum of the ranks' shard sizes (ranks may hold ragged shards)."""
    import torch
    t = torch.tensor([local_rows], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return int(t.item())
