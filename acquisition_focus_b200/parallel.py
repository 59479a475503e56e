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
    """Combine per-shard ``[min, multiplicity]`` (or a stack ``[k, 2]`` of them) into the global one (keeps the reference's
    whole-batch ``volume.min()`` semantics under sharding).  ONE tiny collective: every rank gathers all pairs and takes the
    minimum and the summed multiplicity of the ranks that hold it (no host synchronisation)."""
    if not is_distributed():
        return min_count
    world = dist.get_world_size(group)
    flat = min_count.reshape(-1).contiguous()
    buf = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(buf, flat, group=group)
    pairs = buf.view(world, -1, 2)
    m = pairs[..., 0].min(dim=0).values
    cnt = torch.where(pairs[..., 0] == m, pairs[..., 1], torch.zeros_like(pairs[..., 1])).sum(dim=0)
    return torch.stack([m, cnt], dim=-1).view(min_count.shape)


def exchange_pads(local, group=None):
    """Per-shard pads (``functional.volume_min`` results, MinBackward records attached) -> whole-batch pads: ONE all-gather
    for all of them, the records carried over, and a hook that sums ``d(out)/d(pad)`` over the ranks in the backward pass
    before it is spread over the global minima.  Identity when not distributed.  Pass it as ``pad_exchange`` of
    :func:`functional.acquire_views` (which keeps the label slicing overlapped under the min pass)."""
    if not is_distributed():
        return list(local)
    glob = allreduce_pad(torch.stack(list(local)), group)
    out = []
    for i, loc in enumerate(local):
        g = glob[i].contiguous()
        for attr in ("_afb_mask", "_afb_mask_sig"):
            if hasattr(loc, attr):
                setattr(g, attr, getattr(loc, attr))
        g._afb_dpad_reduce = lambda d_pad, _grp=group: dist.all_reduce(d_pad, op=dist.ReduceOp.SUM, group=_grp)
        out.append(g)
    return out


def global_pads(volumes, with_mask, group=None):
    """``[min, multiplicity]`` of each tensor of ``volumes`` over the WHOLE sharded batch, ready to be passed as ``soft_pad`` /
    ``image_pad`` of :func:`functional.acquire_views`: one local min pass per tensor (``with_mask[i]``: also leave the 1-bit
    MinBackward record) + :func:`exchange_pads`.  With these pads a sharded run returns exactly what the unsharded call
    returns (``tests/test_gpu_multirank.py``)."""
    from . import functional as AF
    return exchange_pads([AF.volume_min(v, with_mask=m) for v, m in zip(volumes, with_mask)], group)
