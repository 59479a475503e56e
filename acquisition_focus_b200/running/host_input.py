"""Host batch -> device one-hot volumes, pipelined over PCIe.

The reference moves a batch to the GPU and then materialises ``one_hot(label)`` and its ``.float()`` with torch ops
(``running/run_dl.py:261-264``).  From HOST buffers that sequence is PCIe-bound (a 128^3 int64 label map is 16 MiB, its
two one-hot volumes are 192 MiB), so here the batch is uploaded in groups of volumes on a copy stream while the
compute stream expands the groups that have already arrived (``afb_onehot_expand``: one launch writes the int64
one-hot, the fp32 one-hot and the fp32 volume's min record).  When the last group lands only its own expansion is
left, and the ``volume.min()`` pass of the bilinear path (``utils/nifti_utils.py:200``) never has to read the soft
volume at all: ``[min, multiplicity]`` comes from the record.

Integer label maps wider than one byte (the reference's datasets hand ``torch.long``) are packed to uint8 on the host cores
first (``afb_host_narrow_labels``, several threads at memory speed, values outside [0, 255] raise): 1 instead of 8 bytes per
voxel cross PCIe.  In :class:`HostInputPipeline` that pass and the enqueueing of a batch's copies run on a background thread,
group by group, while the main thread runs the previous step.
"""
from __future__ import annotations

import ctypes as C
import os
import queue
import threading
import time
from typing import NamedTuple, Optional

import torch

from .. import _lib as L
from .. import functional as AF


def default_narrow_threads() -> int:
    """Host threads for the label packing pass: the CPUs this process may run on, shared between the ranks of one box."""
    env = os.environ.get("AFB_NARROW_THREADS")
    if env:
        return max(1, int(env))
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return max(1, min(16, n // max(1, world)))


def narrow_labels_host(host_label: torch.Tensor, out: torch.Tensor, n_threads: Optional[int] = None) -> None:
    """``out`` (uint8, host, same number of elements) <- ``host_label`` (int64 | int32 | int16, host, contiguous).
    Raises ``ValueError`` when a label lies outside [0, 255].  Releases the GIL while it runs."""
    assert not host_label.is_cuda and not out.is_cuda and out.dtype == torch.uint8
    assert host_label.is_contiguous() and out.is_contiguous() and out.numel() == host_label.numel()
    bad = C.c_int(0)
    L.check(L.lib().afb_host_narrow_labels(host_label.data_ptr(), L.DTYPES[host_label.dtype], host_label.numel(), out.data_ptr(),
                                           int(n_threads or default_narrow_threads()), C.byref(bad)), "afb_host_narrow_labels")
    if bad.value:
        raise ValueError("label map holds values outside [0, 255]: cannot be packed to uint8 for the upload")


_NARROWABLE = (torch.int64, torch.int32, torch.int16)


def balanced_pack_fraction(image_bytes: float, label_bytes: float, label_elem_size: int, pack_rate: float, link_rate: float) -> float:
    """Share f of a batch's label volumes to pack on the host so that the cores and the link finish together:
    cores  f * L / pack_rate        (pack_rate in bytes of the wide labels per second)
    link   (I + (1 - f) * L + f * L / e) / link_rate        (images, unpacked maps, packed maps; e = bytes per wide label)
    =>  f = (I + L) / (L * (link_rate / pack_rate + 1 - 1/e)), clamped to [0, 1]."""
    if label_bytes <= 0 or pack_rate <= 0 or link_rate <= 0:
        return 0.0
    f = (image_bytes + label_bytes) / (label_bytes * (link_rate / pack_rate + 1.0 - 1.0 / label_elem_size))
    return min(1.0, max(0.0, f))


class DeviceBatch(NamedTuple):
    label_map: torch.Tensor            # [B,D,H,W] integer (the uploaded index map; uint8 when it was packed on the host)
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

    MIN_PACK_THREADS = 8

    def __init__(self, num_classes: int, device, depth: int = 2, group_volumes: int = 8, want_label: bool = True,
                 narrow_labels="auto", narrow_threads: Optional[int] = None, pack_volumes: int = 64):
        """``narrow_labels``: False (upload the label maps as they are), True (pack all of them to uint8 on the host first),
        "split" (the host cores pack a share of each batch WHILE the link carries the images and the rest of the maps as int64;
        the share follows the measured packing and link rates) or "auto" (default: "split" when there are >= 2 host threads).
        Measured, 64 volumes of 128^3 per step on one GPU (profiles/r2_e2e_modes.jsonl): 16 host threads - int64 upload 29.4 ms,
        split 13.4 ms (all 64 packed); 3 host threads - int64 upload 29.4, all packed 23.9, split 20.6 ms (44 of 64 packed)."""
        self.C, self.device, self.depth = int(num_classes), torch.device(device), int(depth)
        self.group_volumes, self.want_label = int(group_volumes), want_label
        self.narrow_threads = int(narrow_threads or default_narrow_threads())
        if narrow_labels == "auto":
            narrow_labels = "split" if self.narrow_threads >= 2 else False
        self.split = narrow_labels == "split"
        self.narrow = bool(narrow_labels) and self.C <= 256
        # split uploads: share of a batch's volumes that is packed on the host (the rest crosses PCIe as int64 meanwhile);
        # adapted from the measured packing and link rates so that both finish together
        self.pack_fraction = 1.0 if self.narrow_threads >= self.MIN_PACK_THREADS else 0.5
        self._r_pack = self._r_link = None             # bytes/s: int64 bytes packed per second, bytes copied per second
        self.pack_volumes = int(os.environ.get("AFB_PACK_VOLUMES", pack_volumes))
        with torch.cuda.device(self.device):
            self.copy = torch.cuda.Stream(self.device)
            self.expand = torch.cuda.Stream(self.device)
        self.slots = [None] * self.depth
        self.free_events = [None] * self.depth        # compute stream done with the slot
        self.ready = []                                # FIFO of (slot index, 'enqueued' flag of the worker | None, result holder)
        self.n_submitted = 0
        self.h2d_bytes_last = 0                        # bytes the last submitted batch moves over PCIe
        self.pack_seconds_last = self.enqueue_seconds_last = 0.0      # host time of the last batch: packing / whole enqueue
        self.packed_volumes_last = None                # split uploads: volumes of the last batch that were packed
        self._jobs = None                              # FIFO of the ONE worker thread (started with the first packed batch):
        self._worker = None                            # batches are packed one after the other, never two at a time

    def _work(self):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            args, finished = job
            (self._enqueue_split if self.split else self._enqueue)(*args)
            finished.set()

    def close(self) -> None:
        if self._worker is not None:
            self._jobs.put(None)
            self._worker.join()
            self._worker = None

    def __del__(self):
        try:
            self.close()
        except Exception:          # noqa: BLE001 - interpreter shutdown
            pass

    def _slot(self, i, host_label, host_image, narrow):
        s = self.slots[i]
        shape = tuple(host_label.shape)
        ldtype = torch.uint8 if narrow else host_label.dtype
        if s is None or s["shape"] != shape or s["ldtype"] != ldtype:
            dev, C_ = self.device, self.C
            with torch.cuda.device(dev):
                s = {"shape": shape, "ldtype": ldtype,
                     "lab": torch.empty(shape, dtype=ldtype, device=dev),
                     "img": torch.empty(host_image.shape, dtype=host_image.dtype, device=dev) if host_image is not None else None,
                     "soft": torch.empty(shape + (C_,), dtype=torch.float32, device=dev),
                     "onehot": torch.empty(shape + (C_,), dtype=torch.int64, device=dev) if self.want_label else None,
                     # pinned staging buffer of the packed labels; `copied` = its last H2D copy has finished
                     "lab8_host": torch.empty(shape, dtype=torch.uint8).pin_memory() if narrow else None, "copied": None,
                     # split uploads: the volumes that are NOT packed arrive here as they are on the host
                     "lab_raw": torch.empty(shape, dtype=host_label.dtype, device=dev) if (narrow and self.split) else None,
                     "rate": None}
                s["record"] = AF.min_record_alloc(s["soft"].numel(), dev)
                torch.cuda.current_stream(dev).synchronize()          # one-time: the fresh buffers are safe on every stream
            self.slots[i] = s
        return s

    def _enqueue(self, i, s, host_label, host_image, narrow, holder):
        """Everything one batch needs, in order; runs on the caller's thread (plain labels) or on a worker thread (packing)."""
        try:
            t_pack, t_begin = 0.0, time.perf_counter()
            B = host_label.shape[0]
            per_vol = host_label[0].numel() * self.C
            gv = B if per_vol % 512 else max(1, min(self.group_volumes, B))
            with torch.cuda.device(self.device):
                if self.free_events[i] is not None:             # the step that used this slot last has finished with it
                    self.copy.wait_event(self.free_events[i])
                    self.expand.wait_event(self.free_events[i])
                if narrow and s["copied"] is not None:
                    s["copied"].synchronize()                   # the staging buffer's previous contents have left the host
                packed_to = 0
                for b0 in range(0, B, gv):
                    b1 = min(B, b0 + gv)
                    if s["img"] is not None:                    # the image does not wait for the packing pass
                        with torch.cuda.stream(self.copy):
                            s["img"][b0:b1].copy_(host_image[b0:b1], non_blocking=True)
                    if narrow and b1 > packed_to:
                        # packed in larger pieces than the upload groups: one call over many volumes runs at memory speed, and the
                        # uploads of this batch overlap the packing of the NEXT one anyway (measured on a 16-core host, 64-volume
                        # batches, uploads running: 8 / 16 / 32 / 64 volumes per call 16.6 / 13.1-14.1 / 12.1-13.9 / 13.2 ms; 8.9 ms alone)
                        p1 = min(B, packed_to + max(gv, self.pack_volumes))
                        t0 = time.perf_counter()
                        narrow_labels_host(host_label[packed_to:p1], s["lab8_host"][packed_to:p1], self.narrow_threads)
                        t_pack += time.perf_counter() - t0
                        packed_to = p1
                    with torch.cuda.stream(self.copy):
                        s["lab"][b0:b1].copy_(s["lab8_host"][b0:b1] if narrow else host_label[b0:b1], non_blocking=True)
                        arrived = self.copy.record_event()
                    self.expand.wait_event(arrived)
                    with torch.cuda.stream(self.expand):
                        AF.onehot_expand(s["lab"][b0:b1], self.C, out_label=s["onehot"][b0:b1] if self.want_label else None,
                                         out_soft=s["soft"][b0:b1], record=s["record"], total_elements=s["soft"].numel(),
                                         elem_offset=b0 * per_vol)
                s["copied"] = arrived
                with torch.cuda.stream(self.expand):
                    soft_pad = AF.min_count_from_record(s["record"], s["soft"].numel())
                    image_pad = AF.volume_min(s["img"]) if s["img"] is not None else None
                    done = self.expand.record_event()
            perm = (0, 4, 1, 2, 3)
            holder["db"] = DeviceBatch(s["lab"], s["onehot"].permute(*perm) if self.want_label else None, s["soft"].permute(*perm),
                                       s["img"], soft_pad, image_pad)
            holder["done"] = done
            self.pack_seconds_last, self.enqueue_seconds_last = t_pack, time.perf_counter() - t_begin
        except BaseException as e:          # noqa: BLE001 - re-raised on the caller's thread by get()
            holder["error"] = e

    def _enqueue_split(self, i, s, host_label, host_image, narrow, holder):
        """Split upload of one batch (worker thread): the link carries the images and the int64 label maps of the volumes that are
        not packed WHILE the host cores pack the others; the packed maps follow.  The share is set so that both finish together:
        with L, I bytes of int64 labels and images, r_pack the packing rate (int64 bytes/s) and r_link the copy rate,
        f = (I + L) / (L * (r_link / r_pack + 7/8)), clamped to [0, 1] and rounded to whole volumes; both rates are measured on
        every batch (host clock around the packing calls, CUDA events around the first phase of copies)."""
        try:
            t_begin = time.perf_counter()
            B = host_label.shape[0]
            vox = host_label[0].numel()
            per_vol = vox * self.C
            gv = B if per_vol % 512 else max(1, min(self.group_volumes, B))
            esz = host_label.element_size()
            L_bytes = B * vox * esz
            I_bytes = host_image.numel() * host_image.element_size() if host_image is not None else 0
            with torch.cuda.device(self.device):
                if self.free_events[i] is not None:
                    self.copy.wait_event(self.free_events[i])
                    self.expand.wait_event(self.free_events[i])
                if s["copied"] is not None:
                    s["copied"].synchronize()
                    if s["rate"] is not None:                   # link rate seen by this slot's previous batch
                        e0, e1, nbytes = s["rate"]
                        ms = e0.elapsed_time(e1)
                        if ms > 0 and nbytes > 0:
                            r = nbytes / (ms * 1e-3)
                            self._r_link = r if self._r_link is None else 0.5 * (self._r_link + r)
                if self._r_pack and self._r_link and L_bytes > 0:
                    f = balanced_pack_fraction(I_bytes, L_bytes, esz, self._r_pack, self._r_link)
                    self.pack_fraction = min(1.0, max(0.0, 0.5 * (self.pack_fraction + f)))
                n_pack = int(round(self.pack_fraction * B))
                n_pack = (n_pack // gv) * gv if per_vol % 512 else n_pack      # ragged volumes: whole groups only
                n_pack = min(B, max(0, n_pack))
                lab8, raw = s["lab"], s["lab_raw"]
                # ---- phase 1 (link): images + the int64 maps of volumes [n_pack, B) ----
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.copy)
                bytes1 = 0
                arrived = {}
                for b0 in range(0, B, gv):
                    b1 = min(B, b0 + gv)
                    with torch.cuda.stream(self.copy):
                        if s["img"] is not None:
                            s["img"][b0:b1].copy_(host_image[b0:b1], non_blocking=True)
                            bytes1 += (b1 - b0) * host_image[0].numel() * host_image.element_size()
                        r0 = max(b0, n_pack)
                        if r0 < b1:
                            raw[r0:b1].copy_(host_label[r0:b1], non_blocking=True)
                            bytes1 += (b1 - r0) * vox * esz
                            arrived[("raw", b0)] = self.copy.record_event()
                e1.record(self.copy)
                s["rate"] = (e0, e1, bytes1)
                def expand(kind, lo, hi, src, b0):
                    self.expand.wait_event(arrived[(kind, b0)])
                    with torch.cuda.stream(self.expand):
                        AF.onehot_expand(src[lo:hi], self.C, out_label=s["onehot"][lo:hi] if self.want_label else None,
                                         out_soft=s["soft"][lo:hi], record=s["record"], total_elements=s["soft"].numel(),
                                         elem_offset=lo * per_vol)
                        if kind == "raw":
                            lab8[lo:hi].copy_(src[lo:hi])              # label_map is uint8 for the whole batch

                for b0 in range(0, B, gv):                              # enqueued now: they run while the cores pack
                    b1 = min(B, b0 + gv)
                    if max(b0, n_pack) < b1:
                        expand("raw", max(b0, n_pack), b1, raw, b0)
                # ---- meanwhile (host cores): pack volumes [0, n_pack) ----
                t_pack = 0.0
                if n_pack > 0:
                    t0 = time.perf_counter()
                    narrow_labels_host(host_label[:n_pack], s["lab8_host"][:n_pack], self.narrow_threads)
                    t_pack = time.perf_counter() - t0
                    r = n_pack * vox * esz / max(t_pack, 1e-9)
                    self._r_pack = r if self._r_pack is None else 0.5 * (self._r_pack + r)
                # ---- phase 2 (link): the packed maps, expanded group by group as they arrive ----
                for b0 in range(0, B, gv):
                    p1 = min(b0 + gv, n_pack)
                    if b0 < p1:
                        with torch.cuda.stream(self.copy):
                            lab8[b0:p1].copy_(s["lab8_host"][b0:p1], non_blocking=True)
                            arrived[("packed", b0)] = self.copy.record_event()
                        expand("packed", b0, p1, lab8, b0)
                with torch.cuda.stream(self.copy):
                    last = self.copy.record_event()
                s["copied"] = last
                with torch.cuda.stream(self.expand):
                    soft_pad = AF.min_count_from_record(s["record"], s["soft"].numel())
                    image_pad = AF.volume_min(s["img"]) if s["img"] is not None else None
                    done = self.expand.record_event()
            perm = (0, 4, 1, 2, 3)
            holder["db"] = DeviceBatch(lab8, s["onehot"].permute(*perm) if self.want_label else None, s["soft"].permute(*perm),
                                       s["img"], soft_pad, image_pad)
            holder["done"] = done
            holder["h2d_bytes"] = I_bytes + n_pack * vox + (B - n_pack) * vox * esz
            self.h2d_bytes_last = holder["h2d_bytes"]
            self.packed_volumes_last = n_pack
            self.pack_seconds_last, self.enqueue_seconds_last = t_pack, time.perf_counter() - t_begin
        except BaseException as e:          # noqa: BLE001 - re-raised on the caller's thread by get()
            holder["error"] = e

    def submit(self, host_label: torch.Tensor, host_image: Optional[torch.Tensor]) -> None:
        """Start the upload of one batch.  With label packing the host pass and the enqueueing run on a worker thread:
        ``host_label`` / ``host_image`` must stay alive and unchanged until :meth:`get` has returned this batch."""
        i = self.n_submitted % self.depth
        self.n_submitted += 1
        narrow = self.narrow and host_label.dtype in _NARROWABLE and host_label.is_contiguous()
        s = self._slot(i, host_label, host_image, narrow)
        self.h2d_bytes_last = host_label.numel() * (1 if narrow else host_label.element_size()) + (
            host_image.numel() * host_image.element_size() if host_image is not None else 0)
        holder = {}
        if narrow:
            if self._worker is None:
                self._jobs = queue.Queue()
                self._worker = threading.Thread(target=self._work, daemon=True)
                self._worker.start()
            finished = threading.Event()
            self._jobs.put(((i, s, host_label, host_image, True, holder), finished))
        else:
            finished = None
            if self._worker is not None:          # keep the order of the streams' work: drain the packed batches first
                for _, f, _ in self.ready:
                    if f is not None:
                        f.wait()
            self._enqueue(i, s, host_label, host_image, False, holder)
        self.ready.append((i, finished, holder))

    def get(self) -> DeviceBatch:
        i, finished, holder = self.ready.pop(0)
        if finished is not None:
            finished.wait()
        if "error" in holder:
            raise holder["error"]
        torch.cuda.current_stream(self.device).wait_event(holder["done"])
        self._last = i
        return holder["db"]

    def release(self, db: DeviceBatch = None) -> None:
        self.free_events[self._last] = torch.cuda.current_stream(self.device).record_event()
