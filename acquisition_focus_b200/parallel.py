"""Multi-GPU plumbing: batch x view sharding over the GPUs of one box (one process per GPU).

The path shards along the naturally independent volume (batch) axis: every (volume, view) slice is
independent in forward; backward couples only the views of one volume (through dVolume) and all samples
(through the shared view parameters).  So volumes are partitioned across ranks with all views of a
volume co-located - the volume stays L2/HBM local and dVolume needs no cross-GPU reduction - and the
ONLY collective is one small ``all_reduce(sum)`` of the per-view parameter gradients ``[V, P]`` over
NCCL / NVLink per optimiser step.  The reference has no distributed code at all
(its only batch scaling is gradient accumulation, ``running/run_dl.py:444-467``).

One parity caveat is handled here: the bilinear pad value is the global minimum of the tensor passed in
one call (``utils/nifti_utils.py:200``); sharding a batch changes it for non-one-hot volumes, so
:func:`allreduce_pad` reduces ``(min, multiplicity)`` across ranks before the forward pass.

Works with any initialised ``torch.distributed`` backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of ``n_items`` volumes for ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_view_grads(d_params: torch.Tensor, group=None) -> torch.Tensor:
    """``d_params[B_local, V, P]`` (per-sample gradients of the view parameters produced by the backward
    kernel) -> ``[V, P]`` summed over the local batch and all ranks: the single collective of the path."""
    g = d_params.sum(dim=0).contiguous()
    if is_distributed():
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return g


def allreduce_pad(min_count: torch.Tensor, group=None) -> torch.Tensor:
    """Combine per-shard ``[min, multiplicity]`` into the global one (keeps the reference's whole-batch
    ``volume.min()`` semantics under sharding).  Two tiny collectives (min, then sum of the multiplicities
    of the ranks that hold the global minimum)."""
    if not is_distributed():
        return min_count
    m = min_count[:1].clone()
    dist.all_reduce(m, op=dist.ReduceOp.MIN, group=group)
    cnt = torch.where(min_count[:1] == m, min_count[1:2], torch.zeros_like(min_count[1:2]))
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return torch.cat([m, cnt])
