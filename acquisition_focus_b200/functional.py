"""Autograd-aware Python entry points over the C ABI (libafb200.so).

Three levels, all backed by the same two CUDA kernels (``afb_slice_fwd`` / ``afb_slice_bwd``):

* :func:`affine_grid_sample` - ``F.affine_grid`` + ``F.grid_sample`` (5-D, zeros padding,
  ``align_corners=False``) in one launch, never materialising the grid.
* :func:`slice_with_pre_affine` - the body of the reference's ``nifti_grid_sample``
  (``utils/nifti_utils.py:112-207``): fp64 affine bookkeeping fused into the sampler prologue,
  min-shift semantics, returns ``(out, grid_affine, nii_affine)``.
* :func:`acquire_views` - the tail of ``AffineTransformModule.forward``
  (``models/learnable_transform.py:259-306``) for all B x V slices at once, from raw view
  parameters (R6 | offset logits | zoom logit), gradients w.r.t. the parameters computed by the
  analytic chain in the backward kernel's epilogue.

plus :func:`r6_to_matrix` (``utils/transform_utils.py:27-58``) and :func:`embed_slices`
(``models/hybrid_unet.py:71-94``).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _is_dense(t: torch.Tensor) -> bool:
    """True if the tensor covers a compact block of memory in some dim permutation."""
    dims = sorted(range(t.dim()), key=lambda d: (t.stride(d), t.size(d)), reverse=True)
    expect = 1
    for d in reversed(dims):
        if t.size(d) == 1:
            continue
        if t.stride(d) != expect:
            return False
        expect *= t.size(d)
    return True


def _layout_sig(t: torch.Tensor):
    return (t.data_ptr(), t.numel(), tuple(t.stride()), t.dtype)


def _zeros_like_strided(t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    out = torch.empty_strided(t.shape, t.stride(), dtype=dtype, device=t.device)
    return out.zero_()


def volume_min(volume: torch.Tensor, with_mask: bool = False, out: Optional[torch.Tensor] = None,
               workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``volume.min()`` of nifti_utils.py:200 as a device tensor ``[min, multiplicity]`` (fp32).

    ``with_mask`` (fp32 / bf16 / fp16 volumes): the same pass also leaves a 1-bit-per-voxel record of where the minimum sits
    (``afb_volume_min_mask``), attached to the result as ``._afb_mask``; the backward then rebuilds MinBackward
    without re-reading the volume (half the HBM traffic of the dVolume fill).  The record refers to the memory
    layout of ``volume`` at call time."""
    L.require_cuda(volume, "volume")
    if not _is_dense(volume):
        volume = volume.contiguous()
    lib = L.lib()
    dev = volume.device
    with torch.cuda.device(dev):
        # (`out`: a row of a caller's [k,2] buffer; `workspace`: pre-allocated by a caller that launches on a side stream -
        # allocating on a side stream makes the caching allocator pay a cudaMalloc per step)
        ws = torch.empty(int(lib.afb_volume_min_workspace_bytes()), dtype=torch.uint8, device=dev) if workspace is None else workspace
        out = torch.empty(2, dtype=torch.float32, device=dev) if out is None else out
        if with_mask and volume.dtype in (torch.float32, torch.bfloat16, torch.float16):
            mask = torch.empty(int(lib.afb_min_mask_bytes(volume.numel())), dtype=torch.uint8, device=dev)
            if volume.dtype == torch.float32:
                L.check(lib.afb_volume_min_mask(L.ptr(volume), volume.numel(), L.ptr(out), L.ptr(mask), L.ptr(ws),
                                                L.stream_ptr(dev)), "afb_volume_min_mask")
            else:       # same record, 8-element lane vectors: only afb_min_grad_fill_mask_half understands its bit order
                L.check(lib.afb_volume_min_mask_half(L.ptr(volume), L.DTYPES[volume.dtype], volume.numel(), L.ptr(out), L.ptr(mask),
                                                     L.ptr(ws), L.stream_ptr(dev)), "afb_volume_min_mask_half")
            out._afb_mask = mask
            out._afb_mask_sig = _layout_sig(volume)     # the record's bit order follows THIS memory layout and dtype
        else:
            L.check(lib.afb_volume_min(L.ptr(volume), L.DTYPES[volume.dtype], volume.numel(), L.ptr(out), L.ptr(ws),
                                       L.stream_ptr(dev)), "afb_volume_min")
    return out


def min_record_alloc(n_elements: int, device) -> torch.Tensor:
    """Uninitialised chunk record (``afb_min_mask_bytes``) for an fp32 tensor of ``n_elements``."""
    return torch.empty(int(L.lib().afb_min_mask_bytes(int(n_elements))), dtype=torch.uint8, device=device)


def onehot_expand(label_map: torch.Tensor, num_classes: int, out_label: Optional[torch.Tensor] = None,
                  out_soft: Optional[torch.Tensor] = None, record: Optional[torch.Tensor] = None,
                  total_elements: Optional[int] = None, elem_offset: int = 0) -> None:
    """``one_hot(label_map, C)`` (int64, into ``out_label``) and its ``.float()`` (into ``out_soft``) in one launch
    (``running/run_dl.py:261-264``); both outputs are contiguous ``[..., C]`` buffers (view them with
    ``.permute(0,4,1,2,3)`` to get the reference's ``[B,C,D,H,W]`` tensors).  With ``record`` the chunk record of the
    soft volume is written at ``elem_offset`` of a tensor of ``total_elements`` fp32 values (see
    :func:`min_count_from_record`), which replaces the separate ``volume.min()`` pass."""
    L.require_cuda(label_map, "label_map")
    assert not label_map.dtype.is_floating_point and label_map.is_contiguous()
    n_vox = label_map.numel()
    for t, dt in ((out_label, torch.int64), (out_soft, torch.float32)):
        if t is not None:
            assert t.is_cuda and t.dtype == dt and t.is_contiguous() and t.numel() == n_vox * num_classes
    tot = int(total_elements) if total_elements is not None else n_vox * num_classes
    dev = label_map.device
    with torch.cuda.device(dev):
        L.check(L.lib().afb_onehot_expand(L.ptr(label_map), L.DTYPES[label_map.dtype], n_vox, int(num_classes), L.ptr(out_label),
                                          L.ptr(out_soft), L.ptr(record), tot, int(elem_offset), L.stream_ptr(dev)),
                "afb_onehot_expand")


def min_count_from_record(record: torch.Tensor, n_elements: int) -> torch.Tensor:
    """``[min, multiplicity]`` of an fp32 tensor from its chunk record alone; the record rides along as ``._afb_mask`` so
    that the result can be passed as ``soft_pad`` / ``pad`` and serve MinBackward (same contract as
    ``volume_min(..., with_mask=True)``)."""
    lib = L.lib()
    dev = record.device
    with torch.cuda.device(dev):
        ws = torch.empty(int(lib.afb_volume_min_workspace_bytes()), dtype=torch.uint8, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        L.check(lib.afb_min_count_from_mask(L.ptr(record), int(n_elements), L.ptr(out), L.ptr(ws), L.stream_ptr(dev)),
                "afb_min_count_from_mask")
    out._afb_mask = record
    return out


@dataclass
class ViewSpec:
    """Python mirror of ``afb_views`` (include/afb200.h)."""
    kind: int
    V: int = 1
    theta: Optional[torch.Tensor] = None
    pre: Optional[torch.Tensor] = None
    params: Optional[torch.Tensor] = None
    gpre: Optional[torch.Tensor] = None
    init: Optional[torch.Tensor] = None
    R: int = 0
    spat: int = 1
    offset_clip: float = 0.0
    zoom_clip: float = 0.0
    nii_affine: Optional[torch.Tensor] = None
    fov_mm: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    state: Optional[torch.Tensor] = None          # output of afb_view_prologue, shared by all samplers of one acquisition

    def struct(self) -> L.AfbViews:
        s = L.AfbViews()
        s.kind, s.V = self.kind, self.V
        s.theta = None if self.theta is None else self.theta.data_ptr()
        s.pre = None if self.pre is None else self.pre.data_ptr()
        s.pre_is_f64 = int(self.pre is not None and self.pre.dtype == torch.float64)
        s.params = None if self.params is None else self.params.data_ptr()
        s.gpre = None if self.gpre is None else self.gpre.data_ptr()
        s.init = None if self.init is None else self.init.data_ptr()
        s.R, s.spat, s.offset_clip, s.zoom_clip = int(self.R), int(self.spat), float(self.offset_clip), float(self.zoom_clip)
        s.nii_affine = None if self.nii_affine is None else self.nii_affine.data_ptr()
        s.fov_mm = (C.c_double * 3)(*[float(v) for v in self.fov_mm])
        s.state = None if self.state is None else self.state.data_ptr()
        return s

    def diff_input(self) -> torch.Tensor:
        return {L.AFFINE_GRID: self.theta, L.AFFINE_PRE: self.pre, L.AFFINE_PARAMS: self.params}[self.kind]

    def replace(self, **kw) -> "ViewSpec":
        d = dict(self.__dict__)
        d.update(kw)
        return ViewSpec(**d)

    def with_diff_input(self, t: torch.Tensor) -> "ViewSpec":
        return self.replace(**{{L.AFFINE_GRID: "theta", L.AFFINE_PRE: "pre", L.AFFINE_PARAMS: "params"}[self.kind]: t})


def _prep(t: Optional[torch.Tensor], dtype, device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.detach().to(device=device, dtype=dtype).contiguous()


def prepare_views(spec: ViewSpec, B: int, in_size, out_size, device, launch_stream=None):
    """Run the view prologue ONCE for all S = B*V slices (``afb_view_prologue``): returns the spec with its
    device-side state attached plus ``(grid_affine[B,V,4,4] fp32, nii_affine[B,V,4,4] fp64 | None, theta | None)``.
    ``launch_stream``: enqueue the (latency-bound, one warp per slice) kernel there instead of on the current stream - the
    buffers are still allocated on the current stream and the CALLER joins the streams before anything reads them."""
    lib = L.lib()
    S = B * spec.V
    D, H, W = (int(v) for v in in_size)
    Do, Ho, Wo = (int(v) for v in out_size)
    with torch.cuda.device(device):
        state = torch.empty(S * int(lib.afb_view_state_bytes()), dtype=torch.uint8, device=device)
        ga = torch.empty((B, spec.V, 4, 4), dtype=torch.float32, device=device)
        nii = torch.empty((B, spec.V, 4, 4), dtype=torch.float64, device=device) if spec.kind != L.AFFINE_GRID else None
        th = torch.empty((B, spec.V, 4, 4), dtype=torch.float32, device=device) if spec.kind == L.AFFINE_PARAMS else None
        vs = spec.replace(state=None).struct()
        st = L.stream_ptr(device) if launch_stream is None else launch_stream.cuda_stream
        L.check(lib.afb_view_prologue(C.byref(vs), B, D, H, W, Do, Ho, Wo, L.ptr(state), L.ptr(ga), L.ptr(nii), L.ptr(th), st),
                "afb_view_prologue")
    return spec.replace(state=state), ga, nii, th


def _slice_forward_raw(volume, spec: ViewSpec, out_size, mode, pad_mode, pad_value, pad_dev, out=None):
    """One sampler launch over all slices; ``spec.state`` must be attached (see prepare_views)."""
    lib = L.lib()
    B, Cc = volume.shape[:2]
    Do, Ho, Wo = (int(v) for v in out_size)
    dev = volume.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((B, spec.V, Cc, Do, Ho, Wo), dtype=volume.dtype, device=dev)
        vd, vs = L.volume_desc(volume), spec.struct()
        L.check(lib.afb_slice_fwd(C.byref(vd), C.byref(vs), Do, Ho, Wo, mode, pad_mode, float(pad_value), L.ptr(pad_dev),
                                  L.ptr(out), L.stream_ptr(dev)), "afb_slice_fwd")
    return out


class _SliceFn(torch.autograd.Function):
    """out, grid_affine, nii_affine, theta = f(volume, view_input).  ``prepared`` (optional) carries the output of
    prepare_views so that several samplings of one acquisition share a single prologue launch."""

    @staticmethod
    def forward(ctx, volume, view_input, spec: ViewSpec, out_size, mode, pad_mode, pad_value, pad_dev, prepared, fused=None):
        L.require_cuda(volume, "volume")
        ctx.set_materialize_grads(False)
        if prepared is None:
            spec = spec.with_diff_input(view_input.detach().contiguous())
            prepared = prepare_views(spec, volume.shape[0], volume.shape[2:], out_size, volume.device)
        spec, ga, nii, th = prepared
        y_label = y_image = None
        if fused is not None:          # (label volume | None, image volume | None, image pad triple): ONE launch for all three
            out, y_label, y_image = _slice_forward3_raw(volume.detach(), fused[0], fused[1], spec, out_size,
                                                        (pad_mode, pad_value, pad_dev), fused[2])
        else:
            out = _slice_forward_raw(volume.detach(), spec, out_size, mode, pad_mode, pad_value, pad_dev)
        ctx.spec, ctx.out_size, ctx.mode = spec, tuple(int(v) for v in out_size), mode
        ctx.pad_mode, ctx.pad_value = pad_mode, pad_value
        ctx.save_for_backward(volume.detach(), pad_dev if pad_dev is not None else torch.empty(0))
        ctx.pad_mask = getattr(pad_dev, "_afb_mask", None) if pad_dev is not None else None
        sig = getattr(pad_dev, "_afb_mask_sig", None) if pad_dev is not None else None
        if ctx.pad_mask is not None and sig is not None and sig != _layout_sig(volume):
            ctx.pad_mask = None      # record built for another tensor / layout / dtype: MinBackward re-reads the volume instead
        if ctx.pad_mask is not None and volume.dtype == torch.float32 and ctx.pad_mask.numel() != int(L.lib().afb_min_mask_bytes(volume.numel())):
            ctx.pad_mask = None
        # sharded batches (parallel.global_pad): d(out)/d(pad) must be summed over the ranks before it is spread over the minima
        ctx.dpad_reduce = getattr(pad_dev, "_afb_dpad_reduce", None) if pad_dev is not None else None
        ctx.in_dtype = view_input.dtype
        if fused is None:
            ga = ga.clone()      # several Function calls may share one prologue: each owns its differentiable output
        nii = nii if nii is not None else torch.empty(0, device=volume.device)
        th = th if th is not None else torch.empty(0, device=volume.device)
        y_label = y_label if y_label is not None else torch.empty(0, device=volume.device)
        y_image = y_image if y_image is not None else torch.empty(0, device=volume.device)
        ctx.mark_non_differentiable(nii, th, y_label, y_image)
        if mode == L.NEAREST or not volume.dtype.is_floating_point:
            ctx.mark_non_differentiable(out)
        return out, ga, nii, th, y_label, y_image

    @staticmethod
    def backward(ctx, g_out, g_ga, _g_nii, _g_th, _g_yl=None, _g_yi=None):
        volume, pad_dev = ctx.saved_tensors
        pad_dev = pad_dev if pad_dev.numel() else None
        spec: ViewSpec = ctx.spec
        lib = L.lib()
        dev = volume.device
        B = volume.shape[0]
        S = B * spec.V
        none = (None,) * 8
        sample_grad = g_out is not None and ctx.mode == L.BILINEAR and volume.dtype.is_floating_point
        need_vol = ctx.needs_input_grad[0] and sample_grad
        need_aff = ctx.needs_input_grad[1]
        if not need_vol and not need_aff:
            return (None, None) + none
        if not sample_grad and g_ga is None:
            d_aff = torch.zeros(spec.diff_input().shape, dtype=ctx.in_dtype, device=dev) if need_aff else None
            return (None, d_aff) + none
        Do, Ho, Wo = ctx.out_size
        with torch.cuda.device(dev):
            st = L.stream_ptr(dev)
            vd, vs = L.volume_desc(volume), spec.struct()
            go = g_out.contiguous().float() if sample_grad else None
            gga = g_ga.contiguous().float() if g_ga is not None else None
            # every entry of d_aff is written by the chain kernel; the fp64 accumulators and d_pad share ONE zeroed buffer
            d_aff = torch.empty(spec.diff_input().shape, dtype=torch.float32, device=dev) if need_aff else None
            ws_bytes = int(lib.afb_slice_bwd_workspace_bytes(S))
            zbuf = torch.zeros(ws_bytes + 16, dtype=torch.uint8, device=dev)
            ws, d_pad0 = zbuf[:ws_bytes], zbuf[ws_bytes:ws_bytes + 4].view(torch.float32)

            def theta_half(stream_ptr):      # re-gather + dTheta reduction + analytic chain: does not touch dVolume
                L.check(lib.afb_slice_bwd(C.byref(vd), C.byref(vs), Do, Ho, Wo, ctx.pad_mode, float(ctx.pad_value), L.ptr(pad_dev),
                                          L.ptr(go), L.ptr(gga), None, L.ptr(d_aff), None, None, L.ptr(ws), stream_ptr),
                        "afb_slice_bwd")

            d_vol = None
            if need_vol:
                # Optional split backward (AFB_BWD_SPLIT=1): the dTheta half is independent of the MinBackward fill, so it can run
                # on a side stream UNDER the fill, the scatter half (REDs only, no gather) following the fill.  Measured on the
                # B200 at 64 volumes x 6 views it LOSES to the single kernel (2.823 vs 2.680 ms/step: the two halves repeat the
                # coordinate work and the overlap under the fill is small), so it is off by default; the scatter-only kernel is
                # what runs when only dVolume is wanted.
                split = need_aff and os.environ.get("AFB_BWD_SPLIT", "0") == "1"
                if split:
                    main, side = torch.cuda.current_stream(dev), _side_stream(dev)
                    side.wait_event(main.record_event())             # go, d_aff, ws are ready at this point
                    with torch.cuda.stream(side):
                        theta_half(L.stream_ptr(dev))
                if ctx.pad_mode == L.PAD_DEVICE:
                    # MinBackward fused with the zero fill: d_pad first (geometry + grad_out only), then
                    # d_vol = (vol == min) ? d_pad / count : 0, then the scatter adds on top
                    d_pad = d_pad0
                    L.check(lib.afb_slice_pad_grad(C.byref(vd), C.byref(vs), Do, Ho, Wo, L.ptr(go), L.ptr(d_pad), st),
                            "afb_slice_pad_grad")
                    if ctx.dpad_reduce is not None:
                        ctx.dpad_reduce(d_pad)
                    d_vol = torch.empty_strided(volume.shape, volume.stride(), dtype=torch.float32, device=dev)
                    if ctx.pad_mask is not None:        # 1-bit record left by the forward's min pass: no volume re-read
                        fill = lib.afb_min_grad_fill_mask if volume.dtype == torch.float32 else lib.afb_min_grad_fill_mask_half
                        L.check(fill(L.ptr(ctx.pad_mask), volume.numel(), L.ptr(pad_dev), L.ptr(d_pad), L.ptr(d_vol), st),
                                "afb_min_grad_fill_mask")
                    else:
                        L.check(lib.afb_min_grad_fill(L.ptr(volume), L.DTYPES[volume.dtype], volume.numel(), L.ptr(pad_dev),
                                                      L.ptr(d_pad), L.ptr(d_vol), st), "afb_min_grad_fill")
                else:
                    d_vol = _zeros_like_strided(volume)
                if split or not need_aff:
                    L.check(lib.afb_slice_scatter(C.byref(vd), C.byref(vs), Do, Ho, Wo, L.ptr(go), L.ptr(d_vol), st), "afb_slice_scatter")
                    if split:
                        main.wait_stream(side)
                else:                                    # one kernel does both halves
                    L.check(lib.afb_slice_bwd(C.byref(vd), C.byref(vs), Do, Ho, Wo, ctx.pad_mode, float(ctx.pad_value), L.ptr(pad_dev),
                                              L.ptr(go), L.ptr(gga), L.ptr(d_vol), L.ptr(d_aff), None, None, L.ptr(ws), st),
                            "afb_slice_bwd")
            else:
                theta_half(st)
        if d_aff is not None and d_aff.dtype != ctx.in_dtype:
            d_aff = d_aff.to(ctx.in_dtype)
        if d_vol is not None and d_vol.dtype != volume.dtype:
            # same dense block, same strides: one flat cast kernel instead of torch's strided element-wise copy
            d_half = torch.empty_strided(volume.shape, volume.stride(), dtype=volume.dtype, device=dev)
            with torch.cuda.device(dev):
                L.check(lib.afb_cast_from_f32(L.ptr(d_vol), L.ptr(d_half), L.DTYPES[volume.dtype], volume.numel(), L.stream_ptr(dev)),
                        "afb_cast_from_f32")
            d_vol = d_half
        return (d_vol, d_aff) + none


def _pad_args(volume, mode, pad):
    """pad: 'zero' | 'global_min' | float | device tensor [min, count]."""
    pad_mode, pad_value, pad_dev = L.PAD_ZERO, 0.0, None
    if mode == L.BILINEAR:
        if isinstance(pad, torch.Tensor):
            pad_mode, pad_dev = L.PAD_DEVICE, pad
        elif pad == "global_min":
            want_dvol = volume.requires_grad and torch.is_grad_enabled()
            pad_mode, pad_dev = L.PAD_DEVICE, volume_min(volume, with_mask=want_dvol)
        elif pad == "zero":
            pass
        else:
            pad_mode, pad_value = L.PAD_VALUE, float(pad)
    return pad_mode, pad_value, pad_dev


def _maybe_channels_last(volume, spec, out_size):
    """Planar (NCDHW-contiguous) float volumes with several channels take the generic samplers (a serial channel loop, scalar
    gathers and scalar atomics); the channels-last kernels move 4-8 channels per instruction.  When a call samples about as
    many locations as the volume has voxels (3-D -> 3-D resamples, many views) one transposing copy pays for itself:
    measured on the B200 (bench variant `f1_resample_3d`): 128^3 -> 128^3, C = 8, B = 2: planar 1.9x slower than
    transpose + channels-last.  Autograd carries the gradient back through the copy."""
    C_ = volume.shape[1]
    if volume.dim() != 5 or not volume.dtype.is_floating_point or C_ < 4 or C_ % 4 or volume.stride(1) == 1:
        return volume
    n_out = int(out_size[0]) * int(out_size[1]) * int(out_size[2]) * spec.V
    if 4 * n_out < volume.shape[2] * volume.shape[3] * volume.shape[4] or os.environ.get("AFB_NO_TRANSPOSE", "0") == "1":
        return volume
    return volume.contiguous(memory_format=torch.channels_last_3d)


def _run_slice(volume, view_input, spec, out_size, mode, pad, prepared=None):
    volume = _maybe_channels_last(volume, spec, out_size)
    if not _is_dense(volume):
        volume = volume.contiguous()
    pad_mode, pad_value, pad_dev = _pad_args(volume, mode, pad)
    return _SliceFn.apply(volume, view_input, spec, out_size, mode, pad_mode, pad_value, pad_dev, prepared)[:4]


def _fwd3_ok(soft, label, image) -> bool:
    """Layouts afb_slice_fwd3 takes: fp32 channels-last soft label (C % 4 == 0), channels-last integer label with 16-byte channel
    vectors, fp32 image; same B, D, H, W."""
    if soft.dtype != torch.float32 or soft.stride(1) != 1 or soft.shape[1] % 4 != 0 or soft.data_ptr() % 16:
        return False
    if any(st % 4 for st in (soft.stride(0), soft.stride(2), soft.stride(3), soft.stride(4))):
        return False
    for t in (label, image):
        if t is not None and (t.shape[0] != soft.shape[0] or tuple(t.shape[2:]) != tuple(soft.shape[2:])):
            return False
    if label is not None:
        if label.dtype not in (torch.int64, torch.int32, torch.int16, torch.uint8) or label.stride(1) != 1:
            return False
        n = 16 // label.element_size()
        if label.shape[1] % n or label.data_ptr() % 16 or any(st % n for st in (label.stride(0), label.stride(2), label.stride(3), label.stride(4))):
            return False
    if image is not None and image.dtype != torch.float32:
        return False
    return True


def _slice_forward3_raw(soft, label, image, spec: ViewSpec, out_size, pad_s, pad_i):
    """One launch for the soft-label (bilinear), label (nearest) and image (bilinear) slicings of one acquisition."""
    lib = L.lib()
    B, Cc = soft.shape[:2]
    Do, Ho, Wo = (int(v) for v in out_size)
    dev = soft.device
    with torch.cuda.device(dev):
        y_soft = torch.empty((B, spec.V, Cc, Do, Ho, Wo), dtype=torch.float32, device=dev)
        y_label = torch.empty((B, spec.V, label.shape[1], Do, Ho, Wo), dtype=label.dtype, device=dev) if label is not None else None
        y_image = torch.empty((B, spec.V, image.shape[1], Do, Ho, Wo), dtype=torch.float32, device=dev) if image is not None else None
        vs = spec.struct()
        ds = L.volume_desc(soft)
        dl = L.volume_desc(label) if label is not None else None
        di = L.volume_desc(image) if image is not None else None
        L.check(lib.afb_slice_fwd3(C.byref(ds), C.byref(dl) if dl is not None else None, C.byref(di) if di is not None else None,
                                   C.byref(vs), Do, Ho, Wo, pad_s[0], float(pad_s[1]), L.ptr(pad_s[2]), pad_i[0], float(pad_i[1]),
                                   L.ptr(pad_i[2]), L.ptr(y_soft), L.ptr(y_label), L.ptr(y_image), L.stream_ptr(dev)), "afb_slice_fwd3")
    return y_soft, y_label, y_image


# ------------------------------------------------------------------------------------------------
# public ops
# ------------------------------------------------------------------------------------------------
def affine_grid_sample(volume: torch.Tensor, theta: torch.Tensor, size: Sequence[int], mode: str = "bilinear",
                       pad="zero") -> torch.Tensor:
    """``F.grid_sample(volume, F.affine_grid(theta, [N,C,*size], False), mode, 'zeros', False)``.

    volume ``[N,C,D,H,W]``, theta ``[N,3,4]`` fp32 -> ``[N,C,*size]``; differentiable w.r.t. both."""
    L.require_cuda(volume, "volume")
    m = {"bilinear": L.BILINEAR, "nearest": L.NEAREST}[mode]
    spec = ViewSpec(kind=L.AFFINE_GRID, V=1)
    out, _, _, _ = _run_slice(volume, theta.to(volume.device, torch.float32), spec, size, m, pad)
    return out[:, 0]


def slice_with_pre_affine(volume, nii_affine, pre_affine, fov_mm, fov_vox, is_label=False, pad="global_min"):
    """Body of the reference's ``nifti_grid_sample`` (affine bookkeeping fused into the sampler).

    volume ``[B,C,D,H,W]``, nii_affine ``[B,4,4]`` fp64, pre_affine ``[B,4,4]`` fp32/fp64,
    fov_mm (D,H,W) floats (<=0: keep the input FOV), fov_vox (Do,Ho,Wo).
    Returns ``(out[B,C,Do,Ho,Wo], grid_affine[B,4,4] fp32, nii_affine_out[B,4,4] fp64)``."""
    L.require_cuda(volume, "volume")
    dev = volume.device
    spec = ViewSpec(kind=L.AFFINE_PRE, V=1, nii_affine=_prep(nii_affine, torch.float64, dev),
                    fov_mm=tuple(float(v) for v in fov_mm))
    pre = pre_affine.to(dev)
    if pre.dtype not in (torch.float32, torch.float64):
        pre = pre.float()
    out, ga, nii, _ = _run_slice(volume, pre, spec, fov_vox, L.NEAREST if is_label else L.BILINEAR, pad)
    return out[:, 0], ga[:, 0], nii[:, 0]


_SIDE_STREAMS = {}
# one fused launch vs three launches (label / image overlapped under the min pass on a side stream), measured on the B200 at
# 64 volumes x 6 views (profiles/r2_ab_fwd3.json): 0.451 vs 0.582 ms stand-alone, 2.726 vs 2.757 ms per whole step
_FWD3_DEFAULT = "1"


def _side_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(key)
    return _SIDE_STREAMS[key]


def acquire_views(x_soft_label, x_label, x_image, nifti_affine, gpre, params, init, *, offset_clip, zoom_clip,
                  spat, slice_fov_mm, slice_fov_vox, soft_pad="global_min", image_pad="global_min", overlap_streams=None,
                  pad_exchange=None, fused_forward=None):
    """Fused tail of ``AffineTransformModule.forward`` for all views at once.

    x_soft_label ``[B,C,D,H,W]`` float (grad flows), x_label ``[B,C,D,H,W]`` int (nearest, no grad) or None,
    x_image ``[B,1,D,H,W]`` float (no grad) or None, gpre ``[B,V,4,4]`` fp32, params ``[B,V,6+3R+1]`` fp32
    (R6 | offset logits | zoom logit: the MLP-head output), init ``[V,10]`` fp32.
    Returns ``(y_soft[B,V,C,Do,Ho,Wo], y_label, y_image, grid_affine[B,V,4,4], nii_affine[B,V,4,4], theta[B,V,4,4])``.
    Views are concatenated batch-major, i.e. ``y_soft.flatten(1,2).squeeze(-1)`` is the ``[B, V*C, H, W]`` encoder
    input the reference builds with ``torch.cat(slices, dim=1)`` (running/run_dl.py:325).
    ``pad_exchange`` (sharded batches, ``parallel.exchange_pads``): maps the local ``[min, multiplicity]`` pads of
    (soft label, image) to those of the whole batch, so that every rank pads with the reference's whole-batch
    ``volume.min()`` (utils/nifti_utils.py:200)."""
    L.require_cuda(x_soft_label, "x_soft_label")
    dev = x_soft_label.device
    if overlap_streams is None:
        overlap_streams = os.environ.get("AFB_OVERLAP", "1") != "0"
    B, V = gpre.shape[0], gpre.shape[1]
    NP = params.shape[-1]
    R = (NP - 7) // 3
    spec = ViewSpec(kind=L.AFFINE_PARAMS, V=V, gpre=_prep(gpre, torch.float32, dev).view(B * V, 4, 4),
                    init=_prep(init, torch.float32, dev), R=R, spat=int(spat), offset_clip=float(offset_clip),
                    zoom_clip=float(zoom_clip), nii_affine=_prep(nifti_affine, torch.float64, dev),
                    fov_mm=tuple(float(v) for v in slice_fov_mm))
    p = params.to(dev, torch.float32).reshape(B * V, NP)
    spec = spec.replace(params=p.detach().contiguous())
    has_l = x_label is not None and x_label.numel() > 0
    has_i = x_image is not None and x_image.numel() > 0
    if fused_forward is None:
        fused_forward = os.environ.get("AFB_FWD3", _FWD3_DEFAULT) == "1"
    pads_inside = soft_pad == "global_min" and (not has_i or image_pad == "global_min")
    # one prologue for everything below; when this call also runs the min passes, the prologue (one warp per slice, latency
    # bound) and the image's min pass go to the side stream UNDER the soft volume's min pass
    hide = bool(overlap_streams and fused_forward and pads_inside and (has_l or has_i))
    if hide:
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        side.wait_event(main.record_event())
    prepared = prepare_views(spec, B, x_soft_label.shape[2:], slice_fov_vox, dev, launch_stream=side if hide else None)
    side_joined = not hide

    def no_grad_slicings():
        yl = yi = None
        with torch.no_grad():
            if x_label is not None and x_label.numel() > 0:
                yl = _run_slice(x_label, p.detach(), spec, slice_fov_vox, L.NEAREST, "zero", prepared)[0]
            if x_image is not None and x_image.numel() > 0:
                yi = _run_slice(x_image, p.detach(), spec, slice_fov_vox, L.BILINEAR, image_pad, prepared)[0]
        return yl, yi

    if fused_forward and (has_l or has_i) and (pad_exchange is None or (soft_pad == "global_min" and (not has_i or image_pad == "global_min"))):
        # ONE launch for the three slicings (coordinates / corners / weights once per output location): the min passes first
        # (pads), then afb_slice_fwd3.  Falls through to the per-volume launches when a layout does not qualify.
        xs = x_soft_label if _is_dense(x_soft_label) else x_soft_label.contiguous()
        xl = (x_label if _is_dense(x_label) else x_label.contiguous()) if has_l else None
        xi = (x_image.detach() if _is_dense(x_image) else x_image.detach().contiguous()) if has_i else None
        if _fwd3_ok(xs, xl, xi):
            if soft_pad == "global_min" and (not has_i or image_pad == "global_min"):
                # the min passes: the soft volume's (HBM bound, the longest kernel of the forward) on the caller's stream, the
                # image's UNDER it on the side stream; both write rows of one [k,2] buffer so that a sharded batch turns them
                # into whole-batch pads with ONE small exchange kernel
                want_dvol = xs.requires_grad and torch.is_grad_enabled()
                rows = torch.empty((2 if has_i else 1, 2), dtype=torch.float32, device=dev)
                if hide:
                    pi = None
                    if has_i:
                        ws_i = torch.empty(int(L.lib().afb_volume_min_workspace_bytes()), dtype=torch.uint8, device=dev)   # on `main`
                        with torch.cuda.stream(side):
                            pi = volume_min(xi, out=rows[1], workspace=ws_i)
                    ps = volume_min(xs.detach(), with_mask=want_dvol, out=rows[0])
                    main.wait_stream(side)
                    side_joined = True
                else:
                    ps = volume_min(xs.detach(), with_mask=want_dvol, out=rows[0])
                    pi = volume_min(xi, out=rows[1]) if has_i else None
                local = [ps] + ([pi] if has_i else [])
                local[0]._afb_rows = rows
                pads = pad_exchange(local) if pad_exchange is not None else local
                soft_pad, image_pad = pads[0], (pads[1] if has_i else image_pad)
            if not side_joined:
                main.wait_stream(side)
                side_joined = True
            pm_s, pv_s, pd_s = _pad_args(xs, L.BILINEAR, soft_pad)
            pad_i = _pad_args(xi, L.BILINEAR, image_pad) if has_i else (L.PAD_ZERO, 0.0, None)
            y_soft, ga, nii, theta, y_label, y_image = _SliceFn.apply(xs, p, spec, slice_fov_vox, L.BILINEAR, pm_s, pv_s, pd_s, prepared,
                                                                      (xl, xi, pad_i))
            return y_soft, (y_label if has_l else None), (y_image if has_i else None), ga, nii, theta
    if not side_joined:                        # (a layout did not qualify for the fused launch)
        main.wait_stream(side)
        side_joined = True
    if pad_exchange is not None and soft_pad == "global_min" and (not has_i or image_pad == "global_min"):
        # sharded batch: label slicing (needs no pad) on the side stream UNDER the local min passes; then ONE small exchange
        # turns the local pads into whole-batch pads; soft and image slicings follow on the caller's stream
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        Do, Ho, Wo = (int(v) for v in slice_fov_vox)
        xs = x_soft_label if _is_dense(x_soft_label) else x_soft_label.contiguous()
        xl = (x_label if _is_dense(x_label) else x_label.contiguous()) if has_l else None
        xi = (x_image.detach() if _is_dense(x_image) else x_image.detach().contiguous()) if has_i else None
        y_label = torch.empty((B, V, xl.shape[1], Do, Ho, Wo), dtype=xl.dtype, device=dev) if has_l else None
        forked = main.record_event()
        if has_l:
            side.wait_event(forked)
            with torch.cuda.stream(side):
                _slice_forward_raw(xl, prepared[0], slice_fov_vox, L.NEAREST, L.PAD_ZERO, 0.0, None, out=y_label)
        want_dvol = xs.requires_grad and torch.is_grad_enabled()
        local = [volume_min(xs.detach(), with_mask=want_dvol)] + ([volume_min(xi)] if has_i else [])
        pads = pad_exchange(local)
        y_soft, ga, nii, theta = _run_slice(xs, p, spec, slice_fov_vox, L.BILINEAR, pads[0], prepared)
        y_image = None
        if has_i:
            y_image = _slice_forward_raw(xi, prepared[0], slice_fov_vox, L.BILINEAR, L.PAD_DEVICE, 0.0, pads[1])
        main.wait_stream(side)
    elif overlap_streams and not isinstance(soft_pad, torch.Tensor):
        # the label / image slicings (gather kernels, latency bound) run on a side stream UNDER the soft volume's min pass
        # (HBM bound), then the soft slicing follows on the caller's stream.  Their outputs are allocated on the caller's
        # stream and the side stream is joined before returning, so the caching allocator needs no cross-stream bookkeeping.
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        Do, Ho, Wo = (int(v) for v in slice_fov_vox)
        xl = (x_label if _is_dense(x_label) else x_label.contiguous()) if has_l else None
        xi = (x_image.detach() if _is_dense(x_image) else x_image.detach().contiguous()) if has_i else None
        y_label = torch.empty((B, V, xl.shape[1], Do, Ho, Wo), dtype=xl.dtype, device=dev) if has_l else None
        y_image = torch.empty((B, V, xi.shape[1], Do, Ho, Wo), dtype=xi.dtype, device=dev) if has_i else None
        ws_i = torch.empty(int(L.lib().afb_volume_min_workspace_bytes()) + 16, dtype=torch.uint8, device=dev) if has_i else None
        def side_work():
            side.wait_event(forked)
            with torch.cuda.stream(side):
                if has_l:
                    _slice_forward_raw(xl, prepared[0], slice_fov_vox, L.NEAREST, L.PAD_ZERO, 0.0, None, out=y_label)
                if has_i:
                    pm, pv, pd = L.PAD_ZERO, 0.0, None
                    if isinstance(image_pad, torch.Tensor):
                        pm, pd = L.PAD_DEVICE, image_pad
                    elif image_pad == "global_min":
                        pd = ws_i[:8].view(torch.float32)
                        L.check(L.lib().afb_volume_min(L.ptr(xi), L.DTYPES[xi.dtype], xi.numel(), L.ptr(pd), L.ptr(ws_i[16:]),
                                                       L.stream_ptr(dev)), "afb_volume_min")
                        pm = L.PAD_DEVICE
                    elif image_pad != "zero":
                        pm, pv = L.PAD_VALUE, float(image_pad)
                    _slice_forward_raw(xi, prepared[0], slice_fov_vox, L.BILINEAR, pm, pv, pd, out=y_image)

        forked = main.record_event()                # everything the side stream needs exists at this point
        side_work()                                 # (enqueueing the min pass first instead makes no difference: measured)
        y_soft, ga, nii, theta = _run_slice(x_soft_label, p, spec, slice_fov_vox, L.BILINEAR, soft_pad, prepared)
        main.wait_stream(side)
    else:
        y_soft, ga, nii, theta = _run_slice(x_soft_label, p, spec, slice_fov_vox, L.BILINEAR, soft_pad, prepared)
        y_label, y_image = no_grad_slicings()
    return y_soft, y_label, y_image, ga, nii, theta


class _OneHotSliceFn(torch.autograd.Function):
    """y_soft, y_label, grid_affine = f(params) for an integer label map (no volume gradient exists)."""

    @staticmethod
    def forward(ctx, labels, params, spec: ViewSpec, num_classes, out_size, label_out, prepared):
        ctx.set_materialize_grads(False)
        spec, ga, nii, th = prepared
        lib = L.lib()
        dev = labels.device
        B = labels.shape[0]
        Do, Ho, Wo = (int(v) for v in out_size)
        with torch.cuda.device(dev):
            y_soft = torch.empty((B, spec.V, num_classes, Do, Ho, Wo), dtype=torch.float32, device=dev)
            if label_out == 1:
                y_label = torch.empty((B, spec.V, num_classes, Do, Ho, Wo), dtype=torch.int64, device=dev)
            elif label_out == 2:
                y_label = torch.empty((B, spec.V, Do, Ho, Wo), dtype=torch.uint8, device=dev)
            else:
                y_label = None
            vd, vs = L.volume_desc(labels), spec.struct()
            L.check(lib.afb_slice_onehot_fwd(C.byref(vd), int(num_classes), C.byref(vs), Do, Ho, Wo, L.ptr(y_soft), L.ptr(y_label),
                                             int(label_out), L.stream_ptr(dev)), "afb_slice_onehot_fwd")
        ctx.spec, ctx.out_size, ctx.num_classes = spec, (Do, Ho, Wo), int(num_classes)
        ctx.save_for_backward(labels)
        ctx.in_dtype = params.dtype
        ga = ga.clone()
        if y_label is None:
            y_label = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(y_label)
        return y_soft, y_label, ga

    @staticmethod
    def backward(ctx, g_soft, _g_label, g_ga):
        (labels,) = ctx.saved_tensors
        spec: ViewSpec = ctx.spec
        if not ctx.needs_input_grad[1] or (g_soft is None and g_ga is None):
            return (None,) * 7
        lib = L.lib()
        dev = labels.device
        S = labels.shape[0] * spec.V
        Do, Ho, Wo = ctx.out_size
        with torch.cuda.device(dev):
            d_aff = torch.zeros(spec.params.shape, dtype=torch.float32, device=dev)
            ws = torch.zeros(int(lib.afb_slice_bwd_workspace_bytes(S)), dtype=torch.uint8, device=dev)
            go = g_soft.contiguous().float() if g_soft is not None else None
            gga = g_ga.contiguous().float() if g_ga is not None else None
            vd, vs = L.volume_desc(labels), spec.struct()
            L.check(lib.afb_slice_onehot_bwd(C.byref(vd), ctx.num_classes, C.byref(vs), Do, Ho, Wo, L.ptr(go), L.ptr(gga),
                                             L.ptr(d_aff), None, L.ptr(ws), L.stream_ptr(dev)), "afb_slice_onehot_bwd")
        return (None, d_aff.to(ctx.in_dtype), None, None, None, None, None)


def acquire_views_from_labels(label_map, x_image, nifti_affine, gpre, params, init, *, num_classes, offset_clip, zoom_clip,
                              spat, slice_fov_mm, slice_fov_vox, label_out="onehot", image_pad="global_min"):
    """Same acquisition as :func:`acquire_views`, but from the INTEGER label map ``[B,D,H,W]`` (uint8/int16/int32/int64)
    instead of its materialised fp32 / int64 one-hot volumes (``running/run_dl.py:261-264``): ``y_soft`` is bitwise what
    :func:`acquire_views` returns for ``one_hot(label_map).float()``, ``y_label`` its nearest one-hot (``label_out=
    'onehot'``, int64 like the reference) or the compact uint8 index slice (``'index'``) or nothing (``None``).
    Gradients flow to ``params`` only - the reference's training case, where the volume never requires grad.
    Returns ``(y_soft, y_label, y_image, grid_affine, nii_affine, theta)``."""
    L.require_cuda(label_map, "label_map")
    assert label_map.dim() == 4 and not label_map.dtype.is_floating_point
    dev = label_map.device
    B, V = gpre.shape[0], gpre.shape[1]
    NP = params.shape[-1]
    R = (NP - 7) // 3
    lab5 = label_map[:, None]
    if not _is_dense(lab5):
        lab5 = lab5.contiguous()
    p = params.to(dev, torch.float32).reshape(B * V, NP)
    spec = ViewSpec(kind=L.AFFINE_PARAMS, V=V, gpre=_prep(gpre, torch.float32, dev).view(B * V, 4, 4),
                    init=_prep(init, torch.float32, dev), R=R, spat=int(spat), offset_clip=float(offset_clip),
                    zoom_clip=float(zoom_clip), nii_affine=_prep(nifti_affine, torch.float64, dev),
                    fov_mm=tuple(float(v) for v in slice_fov_mm), params=p.detach().contiguous())
    prepared = prepare_views(spec, B, lab5.shape[2:], slice_fov_vox, dev)
    lo = {"onehot": 1, "index": 2, None: 0}[label_out]
    y_soft, y_label, ga = _OneHotSliceFn.apply(lab5, p, spec, int(num_classes), slice_fov_vox, lo, prepared)
    y_image = None
    if x_image is not None and x_image.numel() > 0:
        with torch.no_grad():
            y_image = _run_slice(x_image, p.detach(), spec, slice_fov_vox, L.BILINEAR, image_pad, prepared)[0]
    return y_soft, (y_label if lo else None), y_image, ga, prepared[2], prepared[3]


def onehot_resample_with_pre_affine(label_map, nii_affine, pre_affine, fov_mm, fov_vox, num_classes):
    """``nifti_grid_sample(one_hot(label_map).float(), ..., is_label=False)`` without materialising the one-hot input
    (e.g. the pre-oriented prescan volume fed to the LocalizationNet, ``models/learnable_transform.py:248-255``).
    label_map ``[B,D,H,W]`` integer -> ``[B,num_classes,Do,Ho,Wo]`` fp32; no gradient (the reference calls it under no_grad)."""
    L.require_cuda(label_map, "label_map")
    dev = label_map.device
    B = label_map.shape[0]
    lab5 = label_map[:, None]
    if not _is_dense(lab5):
        lab5 = lab5.contiguous()
    pre = pre_affine.detach().to(dev)
    if pre.dtype not in (torch.float32, torch.float64):
        pre = pre.float()
    spec = ViewSpec(kind=L.AFFINE_PRE, V=1, nii_affine=_prep(nii_affine, torch.float64, dev), fov_mm=tuple(float(v) for v in fov_mm),
                    pre=pre.contiguous())
    prepared = prepare_views(spec, B, lab5.shape[2:], fov_vox, dev)
    Do, Ho, Wo = (int(v) for v in fov_vox)
    with torch.cuda.device(dev):
        out = torch.empty((B, 1, int(num_classes), Do, Ho, Wo), dtype=torch.float32, device=dev)
        vd, vs = L.volume_desc(lab5), prepared[0].struct()
        L.check(L.lib().afb_slice_onehot_fwd(C.byref(vd), int(num_classes), C.byref(vs), Do, Ho, Wo, L.ptr(out), None, 0,
                                             L.stream_ptr(dev)), "afb_slice_onehot_fwd")
    return out[:, 0], prepared[1][:, 0], prepared[2][:, 0]


class _R6Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ortho):
        L.require_cuda(ortho, "ortho")
        o = ortho.detach().float().contiguous()
        N = o.shape[0]
        mat = torch.empty((N, 4, 4), dtype=torch.float32, device=o.device)
        with torch.cuda.device(o.device):
            L.check(L.lib().afb_r6_fwd(L.ptr(o), N, L.ptr(mat), L.stream_ptr(o.device)), "afb_r6_fwd")
        ctx.save_for_backward(o)
        ctx.in_dtype = ortho.dtype
        return mat.to(ortho.dtype)

    @staticmethod
    def backward(ctx, g):
        (o,) = ctx.saved_tensors
        d = torch.empty_like(o)
        with torch.cuda.device(o.device):
            L.check(L.lib().afb_r6_bwd(L.ptr(o), L.ptr(g.float().contiguous()), o.shape[0], L.ptr(d),
                                       L.stream_ptr(o.device)), "afb_r6_bwd")
        return d.to(ctx.in_dtype)


def r6_to_matrix(ortho: torch.Tensor) -> torch.Tensor:
    """``compute_rotation_matrix_from_ortho6d`` (utils/transform_utils.py:27-58): [N,6] -> [N,4,4]."""
    return _R6Fn.apply(ortho)


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _int_array(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


class _EmbedMultiFn(torch.autograd.Function):
    """(out_0 .. out_{n-1}) = f(affines, x_0 .. x_{n-1}): every stage of one U-Net pass in ONE forward and ONE backward launch."""

    @staticmethod
    def forward(ctx, affines, V, *xs):
        ctx.set_materialize_grads(False)          # a stage whose output takes no part in the loss is skipped in the backward
        n = len(xs)
        dev = xs[0].device
        for x in xs:
            L.require_cuda(x, "x")
        xd = [x.detach().float().contiguous() for x in xs]
        ad = affines.detach().to(dev, torch.float32).contiguous()
        B = xd[0].shape[0]
        cs = [x.shape[1] // V for x in xd]
        Ss = [x.shape[2] for x in xd]
        outs = [torch.empty((B, x.shape[1], S_, S_, S_), dtype=torch.float32, device=dev) for x, S_ in zip(xd, Ss)]
        lib = L.lib()
        with torch.cuda.device(dev):
            ws = torch.empty(int(lib.afb_embed_workspace_bytes(B * V)), dtype=torch.uint8, device=dev)
            L.check(lib.afb_embed_multi_fwd(n, _ptr_array(xd), _int_array(cs), _int_array(Ss), _ptr_array(outs), L.ptr(ad), B, V,
                                            L.ptr(ws), L.stream_ptr(dev)), "afb_embed_multi_fwd")
        ctx.save_for_backward(ad, *xd)
        ctx.V, ctx.cs, ctx.Ss = V, cs, Ss
        ctx.x_dtypes, ctx.a_dtype = [x.dtype for x in xs], affines.dtype
        return tuple(o.to(x.dtype) for o, x in zip(outs, xs))

    @staticmethod
    def backward(ctx, *gs):
        ad, *xd = ctx.saved_tensors
        V, n = ctx.V, len(xd)
        need_a = ctx.needs_input_grad[0]
        need_x = [ctx.needs_input_grad[2 + i] for i in range(n)]
        if not (need_a or any(need_x)) or all(g is None for g in gs):
            return (None,) * (2 + n)
        dev = ad.device
        B = xd[0].shape[0]
        lib = L.lib()
        with torch.cuda.device(dev):
            go = [None if g is None else g.float().contiguous() for g in gs]
            # every element of dx is written (gather form, no atomics) - but only for stages that received a gradient
            dxs = [torch.empty_like(x) if (nx and g is not None) else None for x, nx, g in zip(xd, need_x, go)]
            da = torch.zeros_like(ad) if need_a else None
            ws = torch.zeros(int(lib.afb_embed_workspace_bytes(B * V)), dtype=torch.uint8, device=dev)
            L.check(lib.afb_embed_multi_bwd(n, _ptr_array(go), _ptr_array(xd), _int_array(ctx.cs), _int_array(ctx.Ss),
                                            _ptr_array(dxs) if any(d is not None for d in dxs) else None, L.ptr(ad), B, V, L.ptr(da),
                                            L.ptr(ws), L.stream_ptr(dev)), "afb_embed_multi_bwd")
        d_xs = [None if d is None else d.to(dt) for d, dt in zip(dxs, ctx.x_dtypes)]
        return (da.to(ctx.a_dtype) if da is not None else None, None, *d_xs)


def embed_slices_multi(xs, affines: torch.Tensor, n_views: int):
    """``[SkipConnector.forward(x, affines) for x in xs]`` (models/hybrid_unet.py:40-43: the six encoder skips share the view
    affines) in ONE launch forward and ONE backward.  ``xs[i]`` ``[B, V*c_i, S_i, S_i]`` -> ``[B, V*c_i, S_i, S_i, S_i]``."""
    return list(_EmbedMultiFn.apply(affines, n_views, *xs))


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, affines, V):
        L.require_cuda(x, "x")
        xd = x.detach().float().contiguous()
        ad = affines.detach().to(x.device, torch.float32).contiguous()
        B, CV, S, _ = xd.shape
        c = CV // V
        out = torch.empty((B, CV, S, S, S), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = torch.empty(int(L.lib().afb_embed_workspace_bytes(B * V)), dtype=torch.uint8, device=x.device)
            L.check(L.lib().afb_embed_fwd(L.ptr(xd), L.ptr(ad), B, V, c, S, L.ptr(out), L.ptr(ws), L.stream_ptr(x.device)),
                    "afb_embed_fwd")
        ctx.save_for_backward(xd, ad)
        ctx.V = V
        ctx.x_dtype, ctx.a_dtype = x.dtype, affines.dtype
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        xd, ad = ctx.saved_tensors
        V = ctx.V
        B, CV, S, _ = xd.shape
        c = CV // V
        lib = L.lib()
        need_x, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_a):
            return None, None, None
        with torch.cuda.device(xd.device):
            dx = torch.empty_like(xd) if need_x else None       # every element is written (gather, no atomics)
            da = torch.zeros_like(ad) if need_a else None
            ws = torch.zeros(int(lib.afb_embed_workspace_bytes(B * V)), dtype=torch.uint8, device=xd.device)
            L.check(lib.afb_embed_bwd(L.ptr(g.float().contiguous()), L.ptr(xd), L.ptr(ad), B, V, c, S, L.ptr(dx), L.ptr(da),
                                      L.ptr(ws), L.stream_ptr(xd.device)), "afb_embed_bwd")
        return (dx.to(ctx.x_dtype) if dx is not None else None), (da.to(ctx.a_dtype) if da is not None else None), None


def embed_slices(x: torch.Tensor, affines: torch.Tensor, n_views: int) -> torch.Tensor:
    """``SkipConnector.forward`` (models/hybrid_unet.py:71-94).

    x ``[B, V*c, S, S]``, affines ``[V, B, 4, 4]`` (stacked ``b_grid_affines``) -> ``[B, V*c, S, S, S]``.
    One stage: zero kernel + slab kernel (stand-alone on a par with the batched kernel: 0.41 vs 0.42 ms at S = 128, c = 16,
    B x V = 12); all stages of a U-Net pass at once: :func:`embed_slices_multi` (one launch each way, 0.64 vs 0.93 ms forward,
    1.06 vs 1.96 ms forward + backward).  The backward is the batched gather kernel in both cases.
    ``AFB_EMBED_SINGLE_PASS=1`` routes this call through the batched kernels too (A/B)."""
    if os.environ.get("AFB_EMBED_SINGLE_PASS", "0") == "1":
        return _EmbedMultiFn.apply(affines, n_views, x)[0]
    return _EmbedFn.apply(x, affines, n_views)


class _Rot3Fn(torch.autograd.Function):
    """angle-axis / normal-vector -> homogeneous rotation (utils/transform_utils.py:62-178): one kernel each way."""

    @staticmethod
    def forward(ctx, params, kind):
        L.require_cuda(params, "params")
        p = params.detach().float().contiguous()
        N = p.shape[0]
        mat = torch.empty((N, 4, 4), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            L.check(L.lib().afb_rot3_fwd(int(kind), L.ptr(p), N, L.ptr(mat), L.stream_ptr(p.device)), "afb_rot3_fwd")
        ctx.save_for_backward(p)
        ctx.kind, ctx.in_dtype = int(kind), params.dtype
        return mat.to(params.dtype)

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        d = torch.empty_like(p)
        with torch.cuda.device(p.device):
            L.check(L.lib().afb_rot3_bwd(ctx.kind, L.ptr(p), L.ptr(g.float().contiguous()), p.shape[0], L.ptr(d),
                                         L.stream_ptr(p.device)), "afb_rot3_bwd")
        return d.to(ctx.in_dtype), None


def angle_axis_to_matrix(angle_axis: torch.Tensor) -> torch.Tensor:
    """``angle_axis_to_rotation_matrix`` (utils/transform_utils.py:106-178): ``[N,3] -> [N,4,4]``, differentiable."""
    return _Rot3Fn.apply(angle_axis, L.ROT_ANGLE_AXIS)


def normal_to_matrix(normals: torch.Tensor) -> torch.Tensor:
    """``normal_to_rotation_matrix`` (utils/transform_utils.py:62-103): ``[N,3]`` (columns nz, ny, nx) ``-> [N,4,4]``."""
    return _Rot3Fn.apply(normals, L.ROT_NORMAL)


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        L.require_cuda(x, "x")
        xd = x.detach().float().contiguous()
        h, w = xd.shape[-3], xd.shape[-2]
        n = xd.numel() // (h * w)
        out = torch.empty(xd.shape[:-3] + (H, W, 1), dtype=torch.float32, device=xd.device)
        with torch.cuda.device(xd.device):
            L.check(L.lib().afb_upsample2d_fwd(L.ptr(xd), n, h, w, int(H), int(W), L.ptr(out), L.stream_ptr(xd.device)), "afb_upsample2d_fwd")
        ctx.shape, ctx.HW, ctx.in_dtype = tuple(xd.shape), (int(H), int(W)), x.dtype
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        h, w = ctx.shape[-3], ctx.shape[-2]
        gd = g.float().contiguous()
        dx = torch.empty(ctx.shape, dtype=torch.float32, device=gd.device)
        with torch.cuda.device(gd.device):
            L.check(L.lib().afb_upsample2d_bwd(L.ptr(gd), dx.numel() // (h * w), h, w, ctx.HW[0], ctx.HW[1], L.ptr(dx),
                                               L.stream_ptr(gd.device)), "afb_upsample2d_bwd")
        return dx.to(ctx.in_dtype), None, None


def upsample_slices(x: torch.Tensor, size) -> torch.Tensor:
    """``F.interpolate(x, size=[H, W, 1], mode='trilinear', align_corners=False)`` for one-voxel-thin slices ``[..., h, w, 1]``
    (the up-sampling of low-resolution slices to the hires in-plane size, running/run_dl.py:193-197); differentiable."""
    assert x.shape[-1] == 1 and int(size[-1]) == 1, "slices are one voxel thin"
    return _UpsampleFn.apply(x, int(size[0]), int(size[1]))


def compose_pre_affine(base_affine: torch.Tensor, view_affine: torch.Tensor, aug_affine: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``base_affine.inverse() @ view_affine (@ aug_affine)`` -> fp32 ``[B,4,4]`` (running/run_dl.py:227-234 and the augmentation
    product of :208-223), in fp64 inside one kernel like the reference's fp64 torch chain.  Not differentiable (the reference
    builds these affines from dataset constants and host-side random draws)."""
    L.require_cuda(base_affine, "base_affine")
    dev = base_affine.device
    base = base_affine.detach().to(torch.float64).contiguous()
    view = view_affine.detach().to(dev)
    if view.dtype not in (torch.float32, torch.float64):
        view = view.float()
    view = view.contiguous()
    aug = None if aug_affine is None else aug_affine.detach().to(dev, torch.float32).contiguous()
    B = base.shape[0]
    out = torch.empty((B, 4, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().afb_compose_pre_affine(L.ptr(base), L.ptr(view), int(view.dtype == torch.float64), L.ptr(aug), B, L.ptr(out),
                                               None, L.stream_ptr(dev)), "afb_compose_pre_affine")
    return out
