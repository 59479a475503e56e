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
