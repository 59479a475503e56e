"""Host batch -> device one-hot volumes, pipelined over PCIe.

The reference moves a batch to the GPU and then materialises ``one_hot(label)`` and its ``.float()`` with torch ops
(``running/run_dl.py:261-264``).  From HOST buffers that sequence is PCIe-bound (a 128^3 int64 label map is 16 MiB, its
two one-hot volumes are 192 MiB), so here the batch is uploaded in groups of volumes on a copy stream while the
compute stream expands the groups that have already arrived (``afb_onehot_expand``: one launch writes the int64
one-hot, the fp32 one-hot and the fp32 volume's min record).  When the last group lands only its own expansion is
left, and the ``volume.min()`` pass of the bilinear path (``utils/nifti_utils.py:200``) never has to read the soft
volume at all: ``[min, multiplicity]`` comes from the record.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from .. import functional as AF


class DeviceBatch(NamedTuple):
    label_map: torch.Tensor            # [B,D,H,W] integer (the uploaded index map)
    label: Optional[torch.Tensor]      # [B,C,D,H,W] int64 one-hot, channels-last strides (as run_dl.py:261-262)
    soft_label: torch.Tensor           # [B,C,D,H,W] fp32 one-hot, channels-last strides (as run_dl.py:263-264)
    image: Optional[torch.Tensor]      # [B,1,D,H,W]
    soft_pad: torch.Tensor             # [min, multiplicity] of soft_label (+ its MinBackward record): pass as soft_pad
    image_pad: Optional[torch.Tensor]  # [min, multiplicity] of image: pass as image_pad


def upload_one_hot(host_label: torch.Tensor, host_image: Optional[torch.Tensor], num_classes: int, device,
                   group_volumes: int = 8, want_label: bool = True) -> DeviceBatch:
    """``host_label`` [B,D,H,W] integer and ``host_image`` [B,1,D,H,W] float on the host (pinned memory makes the copies
    asynchronous) -> :class:`DeviceBatch`.  Work is enqueued on the current stream and an internal copy stream; the
    returned tensors are safe to use on the current stream."""
    assert host_label.dim() == 4 and not host_label.dtype.is_floating_point
    device = torch.device(device)
    B = host_label.shape[0]
    vox = host_label[0].numel()
    C = int(num_classes)
    per_vol = vox * C
    if per_vol % 512 != 0:
        group_volumes = B                 # ranges must end on record-chunk boundaries: fall back to one range
    group_volumes = max(1, min(int(group_volumes), B))
    with torch.cuda.device(device):
        main = torch.cuda.current_stream(device)
        copy = torch.cuda.Stream(device)
        lab = torch.empty(host_label.shape, dtype=host_label.dtype, device=device)
        img = torch.empty(host_image.shape, dtype=host_image.dtype, device=device) if host_image is not None else None
        soft = torch.empty(host_label.shape + (C,), dtype=torch.float32, device=device)
        onehot = torch.empty(host_label.shape + (C,), dtype=torch.int64, device=device) if want_label else None
        record = AF.min_record_alloc(soft.numel(), device)
        copy.wait_stream(main)            # the buffers above were allocated on `main`
        for b0 in range(0, B, group_volumes):
            b1 = min(B, b0 + group_volumes)
            with torch.cuda.stream(copy):
                lab[b0:b1].copy_(host_label[b0:b1], non_blocking=True)
                if img is not None:
                    img[b0:b1].copy_(host_image[b0:b1], non_blocking=True)
                arrived = copy.record_event()
            main.wait_event(arrived)
            AF.onehot_expand(lab[b0:b1], C, out_label=onehot[b0:b1] if want_label else None, out_soft=soft[b0:b1],
                             record=record, total_elements=soft.numel(), elem_offset=b0 * per_vol)
        soft_pad = AF.min_count_from_record(record, soft.numel())
        image_pad = AF.volume_min(img) if img is not None else None
    perm = (0, 4, 1, 2, 3)
    return DeviceBatch(lab, onehot.permute(*perm) if want_label else None, soft.permute(*perm), img, soft_pad, image_pad)


class HostInputPipeline:
    """Double-buffered :func:`upload_one_hot`: the NEXT batch crosses PCIe (and is expanded) while the CURRENT one is sliced.

    ``upload_one_hot`` allocates per call, so its copy stream has to wait for the compute stream before the first copy (the
    caching allocator may hand out memory the compute stream still uses) - the upload of step k+1 cannot start before the
    kernels of step k have finished.  Here ``depth`` sets of device buffers live for the whole pipeline and a private copy
    stream and a private expansion stream are ordered by events only:

        pipe = HostInputPipeline(num_classes, device, depth=2)
        pipe.submit(host_label, host_image)                   # batch 0
        for k in range(steps):
            pipe.submit(next_label, next_image)               # batch k+1: H2D + one-hot expansion, off the compute stream
            db = pipe.get()                                   # batch k: ready on the current stream
            ... acquire_views(db.soft_label, db.label, db.image, ..., soft_pad=db.soft_pad, image_pad=db.image_pad) ...
            pipe.release(db)                                  # after the step's last kernel has been enqueued

    A slot is reused ``depth`` submits later; ``release`` records when the compute stream is done with it."""

    def __init__(self, num_classes: int, device, depth: int = 2, group_volumes: int = 8, want_label: bool = True):
        self.C, self.device, self.depth = int(num_classes), torch.device(device), int(depth)
        self.group_volumes, self.want_label = int(group_volumes), want_label
        with torch.cuda.device(self.device):
            self.copy = torch.cuda.Stream(self.device)
            self.expand = torch.cuda.Stream(self.device)
        self.slots = [None] * self.depth
        self.free_events = [None] * self.depth        # compute stream done with the slot
        self.ready = []                                # FIFO of (slot index, DeviceBatch, ready event)
        self.n_submitted = 0

    def _slot(self, i, host_label, host_image):
        s = self.slots[i]
        shape = tuple(host_label.shape)
        if s is None or s["shape"] != shape or s["ldtype"] != host_label.dtype:
            dev, C = self.device, self.C
            with torch.cuda.device(dev):
                s = {"shape": shape, "ldtype": host_label.dtype,
                     "lab": torch.empty(shape, dtype=host_label.dtype, device=dev),
                     "img": torch.empty(host_image.shape, dtype=host_image.dtype, device=dev) if host_image is not None else None,
                     "soft": torch.empty(shape + (C,), dtype=torch.float32, device=dev),
                     "onehot": torch.empty(shape + (C,), dtype=torch.int64, device=dev) if self.want_label else None}
                s["record"] = AF.min_record_alloc(s["soft"].numel(), dev)
                torch.cuda.current_stream(dev).synchronize()          # one-time: the fresh buffers are safe on every stream
            self.slots[i] = s
        return s

    def submit(self, host_label: torch.Tensor, host_image: Optional[torch.Tensor]) -> None:
        i = self.n_submitted % self.depth
        self.n_submitted += 1
        s = self._slot(i, host_label, host_image)
        B = host_label.shape[0]
        per_vol = host_label[0].numel() * self.C
        gv = B if per_vol % 512 else max(1, min(self.group_volumes, B))
        if self.free_events[i] is not None:                 # the step that used this slot last has finished with it
            self.copy.wait_event(self.free_events[i])
            self.expand.wait_event(self.free_events[i])
        with torch.cuda.device(self.device):
            for b0 in range(0, B, gv):
                b1 = min(B, b0 + gv)
                with torch.cuda.stream(self.copy):
                    s["lab"][b0:b1].copy_(host_label[b0:b1], non_blocking=True)
                    if s["img"] is not None:
                        s["img"][b0:b1].copy_(host_image[b0:b1], non_blocking=True)
                    arrived = self.copy.record_event()
                self.expand.wait_event(arrived)
                with torch.cuda.stream(self.expand):
                    AF.onehot_expand(s["lab"][b0:b1], self.C, out_label=s["onehot"][b0:b1] if self.want_label else None,
                                     out_soft=s["soft"][b0:b1], record=s["record"], total_elements=s["soft"].numel(),
                                     elem_offset=b0 * per_vol)
            with torch.cuda.stream(self.expand):
                soft_pad = AF.min_count_from_record(s["record"], s["soft"].numel())
                image_pad = AF.volume_min(s["img"]) if s["img"] is not None else None
                done = self.expand.record_event()
        perm = (0, 4, 1, 2, 3)
        db = DeviceBatch(s["lab"], s["onehot"].permute(*perm) if self.want_label else None, s["soft"].permute(*perm), s["img"],
                         soft_pad, image_pad)
        self.ready.append((i, db, done))

    def get(self) -> DeviceBatch:
        i, db, done = self.ready.pop(0)
        torch.cuda.current_stream(self.device).wait_event(done)
        self._last = i
        return db

    def release(self, db: DeviceBatch = None) -> None:
        self.free_events[self._last] = torch.cuda.current_stream(self.device).record_event()
