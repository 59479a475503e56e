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


_PEER = None          # PeerCollectives instance when the NVLink peer-memory kernels are enabled (enable_peer_collectives)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of ``n_items`` volumes for ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_view_grads(d_params: torch.Tensor, group=None) -> torch.Tensor:
    """``d_params[B_local, V, P]`` (per-sample gradients of the view parameters produced by the backward
    kernel) -> ``[V, P]`` summed over the local batch and all ranks: the single collective of the path.  With the peer-memory
    kernels enabled the local sum and the all-reduce are ONE kernel (``afb_peer_collective``), else ``sum`` + NCCL."""
    if is_distributed() and _PEER is not None and group is None and d_params.dtype == torch.float32 and \
            d_params[0].numel() <= _PEER.n_max:
        return _PEER.all_reduce_sum(d_params.contiguous().view(d_params.shape[0], -1), PeerCollectives.CH_GRADS,
                                    pre_sum=d_params.shape[0]).view(d_params.shape[1:])
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
    if _PEER is not None and group is None and flat.is_cuda and flat.dtype == torch.float32:
        buf = _PEER.all_gather(flat, PeerCollectives.CH_PADS)
    else:
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
    rows = getattr(local[0], "_afb_rows", None)          # the local pads as rows of ONE [k,2] buffer (functional.acquire_views)
    if _PEER is not None and group is None and rows is not None and rows.shape[0] == len(local):
        glob = _PEER.merge_pads(rows)                    # one kernel: publish, wait, min / multiplicity merge
    else:
        glob = allreduce_pad(torch.stack(list(local)), group)
    out = []
    for i, loc in enumerate(local):
        g = glob[i].contiguous()
        for attr in ("_afb_mask", "_afb_mask_sig"):
            if hasattr(loc, attr):
                setattr(g, attr, getattr(loc, attr))
        g._afb_dpad_reduce = lambda d_pad, _grp=group: _reduce_dpad(d_pad, _grp)
        out.append(g)
    return out


def _reduce_dpad(d_pad: torch.Tensor, group=None) -> None:
    """in-place sum over the ranks of the 1-float d(out)/d(pad)."""
    if _PEER is not None and group is None and d_pad.is_cuda and d_pad.dtype == torch.float32:
        _PEER.all_reduce_sum(d_pad.view(1, -1), PeerCollectives.CH_DPAD, out=d_pad.view(-1))
    else:
        dist.all_reduce(d_pad, op=dist.ReduceOp.SUM, group=group)


class PeerCollectives:
    """The path's three tiny collectives as one single-CTA kernel each over NVLink peer memory (``csrc/afb_peer.cu``): every
    rank's buffer lives in symmetric memory (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into all
    peers over NVLink / NVSwitch - plumbing), the exchange itself (publish, bounded wait, reduce in rank order) is ours.
    Stream-ordered and CUDA-graph capturable (the epoch counters live in device memory)."""

    CH_PADS, CH_DPAD, CH_GRADS = 0, 1, 2
    N_CHANNELS = 4

    def __init__(self, device, n_max: int = 1024):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self._L, self._C = L, C
        self.device = torch.device(device)
        self.rank, self.world, self.n_max = dist.get_rank(), dist.get_world_size(), int(n_max)
        n_floats = int(L.lib().afb_peer_buffer_floats(self.N_CHANNELS, self.n_max, self.world))
        self.buf = symm.empty(n_floats, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        try:
            self.hdl = symm.rendezvous(self.buf, group=dist.group.WORLD)
        except Exception:
            symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
            self.hdl = symm.rendezvous(self.buf, group=dist.group.WORLD)
        self.ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        self.epoch = torch.zeros(self.N_CHANNELS, dtype=torch.int32, device=self.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier()                      # every rank's zero-initialised signal words are in place before the first publish
        torch.cuda.synchronize(self.device)

    def _call(self, op, channel, n, pre_sum, src, out):
        L, C = self._L, self._C
        with torch.cuda.device(self.device):
            L.check(L.lib().afb_peer_collective(C.c_void_p(self.ptrs_dev), self.rank, self.world, op, channel, self.N_CHANNELS, n,
                                                self.n_max, pre_sum, L.ptr(src), L.ptr(out), L.ptr(self.epoch), L.ptr(self.err),
                                                L.stream_ptr(self.device)), "afb_peer_collective")

    def all_reduce_sum(self, rows: torch.Tensor, channel: int, pre_sum: int = 1, out=None) -> torch.Tensor:
        """``rows[pre_sum, n]`` fp32 -> ``[n]``: sum over the rows and over all ranks."""
        n = rows.shape[1]
        out = torch.empty(n, dtype=torch.float32, device=self.device) if out is None else out
        self._call(0, channel, n, int(pre_sum), rows, out)
        return out

    def merge_pads(self, rows: torch.Tensor) -> torch.Tensor:
        """``rows[k, 2]`` local (min, multiplicity) pairs -> the pairs of the whole sharded batch."""
        out = torch.empty_like(rows)
        self._call(2, self.CH_PADS, rows.numel(), 1, rows, out)
        return out

    def all_gather(self, flat: torch.Tensor, channel: int) -> torch.Tensor:
        """``flat[n]`` fp32 -> ``[world * n]`` (rank-major)."""
        out = torch.empty(self.world * flat.numel(), dtype=torch.float32, device=self.device)
        self._call(1, channel, flat.numel(), 1, flat, out)
        return out

    def check(self) -> None:
        """Raise if a peer ever failed to arrive (bounded spin in the kernel); synchronises."""
        if int(self.err.item()) != 0:
            raise RuntimeError(f"afb_peer_collective: peer {int(self.err.item()) - 1} did not arrive within the spin bound")


def enable_peer_collectives(device, n_max: int = 1024):
    """Route exchange_pads / reduce_view_grads through the NVLink peer-memory kernels (call once per process after
    ``init_process_group``, on every rank).  Returns the :class:`PeerCollectives` (or raises if symmetric memory is unavailable)."""
    global _PEER
    _PEER = PeerCollectives(device, n_max)
    return _PEER


def disable_peer_collectives() -> None:
    global _PEER
    _PEER = None


def global_pads(volumes, with_mask, group=None):
    """``[min, multiplicity]`` of each tensor of ``volumes`` over the WHOLE sharded batch, ready to be passed as ``soft_pad`` /
    ``image_pad`` of :func:`functional.acquire_views`: one local min pass per tensor (``with_mask[i]``: also leave the 1-bit
    MinBackward record) + :func:`exchange_pads`.  With these pads a sharded run returns exactly what the unsharded call
    returns (``tests/test_gpu_multirank.py``)."""
    from . import functional as AF
    return exchange_pads([AF.volume_min(v, with_mask=m) for v, m in zip(volumes, with_mask)], group)
