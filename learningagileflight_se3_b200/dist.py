"""Multi-GPU plumbing: problems are independent, so the batch is sharded by problem index (one process per GPU,
torch.distributed) and only the small per-problem results (reward / 8-vector gradient / status) are gathered to
rank 0 -- the reference's own parallelism is one OS process per sample with an 8-double shared result buffer
(deep_learning.py:58,66-72).  No data-path collective exists in this path; NCCL carries the gather only.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous split of n problems over `world` ranks: rank r owns [lo, hi); sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_to_rank0(local: torch.Tensor, sizes, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather variable-length shards (first dimension) to `dst` in rank order; returns None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def sharded_call(fn: Callable[..., Dict[str, torch.Tensor]], inputs: Dict[str, np.ndarray], device=None,
                 keys=None, n_total: Optional[int] = None) -> Optional[Dict[str, torch.Tensor]]:
    """Run `fn(**shard)` on this rank's contiguous shard of the problems and gather the outputs named in `keys` to rank 0 in problem
    order.  With one process this is fn(**inputs).

    inputs: arrays with a leading problem dimension -- either the FULL batch, identical on every rank (n_total=None: this rank's
    slice [lo, hi) = shard_range(n, rank, world) is taken here), or, with n_total given, ALREADY this rank's shard of an n_total-problem
    batch (each rank generated only its own part, e.g. from (seed, index); its length must equal the shard_range size)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    first = next(iter(inputs.values()))
    n = int(n_total) if n_total is not None else first.shape[0]
    lo, hi = shard_range(n, rank, world)
    if n_total is not None and first.shape[0] != hi - lo:
        raise ValueError(f"pre-sharded input has {first.shape[0]} problems, rank {rank} of {world} owns {hi - lo} of {n}")
    shard = {}
    for k, v in inputs.items():
        part = v if n_total is not None else v[lo:hi]
        t = part if isinstance(part, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(part))
        shard[k] = t.to(device) if device is not None else t
    out = fn(**shard)
    sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    res = {}
    for k in (keys or out.keys()):
        res[k] = gather_to_rank0(out[k], sizes)
    return res if rank == 0 else None
