"""Mirror of the reference's per-batch caller of the hot path, ``running/run_dl.py:208-329``
(``apply_affine_augmentation``, ``get_input_affine_for_atm``, ``get_reconstruction_model_input``), on top of the
fused kernels: what the reference does with 14 ``nifti_grid_sample`` calls and ~10^4 ATen launches per batch
(B=2, V=3) becomes 2 hires resamples + ONE acquisition over all B x V slices straight from the integer label map.

Same signature and return tuple as the reference: ``(b_input[B, V*C, H, W], b_target[B, C, D, H, W] long,
grid_affines: list[V] of [B,4,4])``; ``config`` is any object with the reference's attribute names
(``config_dict.json``).  ``label_slice_type == 'from-gt'`` with R6 view modules takes the fused all-views route;
``'from-segmented'`` (the caller supplies ``segment_fn``, e.g. the reference's nnU-Net wrapper), the other rotation
parameterisations and ``rotate_slice_to_min_principle`` go view by view through :func:`get_transformed`
(``run_dl.py:146-204``), like the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import functional as AF
from ..utils.nifti_utils import get_zooms, nifti_grid_sample
from ..utils.transform_utils import get_random_affine


def draw_augmentation_affines(B, zoom_strength=0.1, offset_strength=0.1, rotation_strength=0.1, generator=None):
    """The ``[B,4,4]`` random affines of run_dl.py:208-223 (host RNG, the reference's draw sequence)."""
    return torch.stack([get_random_affine(rotation_strength, zoom_strength, offset_strength, generator=generator) for _ in range(B)])


def apply_affine_augmentation(affine_list, zoom_strength=0.1, offset_strength=0.1, rotation_strength=0.1, generator=None):
    """run_dl.py:208-223: right-multiply every affine of the list by ONE random affine per batch element.  Host RNG with the
    reference's draw sequence: with ``generator=None`` (the global torch RNG the reference uses) a seeded run reproduces the
    reference's augmentation affines bit for bit (golden ``model_input_s32.npz``)."""
    B = affine_list[0].shape[0]
    b_affine = torch.stack([get_random_affine(rotation_strength, zoom_strength, offset_strength, generator=generator)
                            for _ in range(B)])
    return [a @ b_affine.to(a) for a in affine_list]


def get_input_affine_for_atm(atm, base_affine, b_view_affines, aug_affine=None):
    """run_dl.py:227-234: ``Gpre = base_affine^-1 @ view_affine[view_id]`` (or the module's random affine for 'RND'); one kernel
    (fp64 inside, like the reference's fp64 chain) instead of a cuSOLVER LU + matmul; ``aug_affine`` folds the augmentation
    product of :208-223 into the same launch.  CPU tensors (host-side tests) take the torch expression."""
    B = base_affine.shape[0]
    if atm.view_id == "RND":
        out = atm.random_grid_affine.repeat(B, 1, 1).to(base_affine)
        return out if aug_affine is None else out @ aug_affine.to(out)
    view = torch.as_tensor(b_view_affines[atm.view_id]).view(B, 4, 4)
    if base_affine.is_cuda:
        return AF.compose_pre_affine(base_affine, view.to(base_affine.device), aug_affine)
    out = base_affine.inverse() @ view.to(base_affine)
    return out if aug_affine is None else out @ aug_affine.to(out)


def get_transformed(config, phase, label, soft_label, nifti_affine, grid_affine_pre_mlp, atm, image=None, segment_fn=None):
    """run_dl.py:146-204, one view: ``atm`` slices the soft label / label / image volumes; for ``label_slice_type ==
    'from-segmented'`` outside training the label slice is replaced by ``segment_fn(image_slice[B,C,1,D,H], zooms)`` (no
    gradient any more, as in the reference); low-resolution slices are up-sampled to the hires FOV.
    Returns ``(image_slc, soft_label_slc, grid_affine)``."""
    img_is_invalid = image is None or image.dim() == 0
    B, num_classes, D, H, W = label.shape
    if img_is_invalid:
        image = torch.zeros(B, 1, D, H, W, device=label.device)
    soft_label_slc, label_slc, image_slc, grid_affine, atm_nii_affine = atm(
        soft_label.view(B, num_classes, D, H, W), label.view(B, num_classes, D, H, W), image.view(B, 1, D, H, W),
        nifti_affine, grid_affine_pre_mlp)
    if getattr(config, "label_slice_type", "from-gt") == "from-segmented" and phase != "train":
        assert not img_is_invalid and segment_fn is not None
        with torch.no_grad():
            pred_slc = segment_fn(image_slc.permute(0, 1, 4, 2, 3), get_zooms(atm_nii_affine)).long()      # B C D H 1 -> B C 1 D H
            if pred_slc.shape[-1] != 1:
                pred_slc = pred_slc[:, 0][..., None]                                                     # B 1 D H -> B D H 1
            soft_label_slc = label_slc = F.one_hot(pred_slc, num_classes).permute(0, 4, 1, 2, 3).to(soft_label_slc)
    if list(config.slice_fov_vox) != list(config.hires_fov_vox):
        tgt = list(config.hires_fov_vox[:2]) + [1]
        image_slc = AF.upsample_slices(image_slc, tgt)             # F.interpolate(..., 'trilinear') of :196-197 as one kernel
        soft_label_slc = AF.upsample_slices(soft_label_slc, tgt)
    if img_is_invalid:
        image_slc = torch.empty([])
    return image_slc, soft_label_slc, grid_affine


def _fused_route_ok(config, modules, atm_container=None) -> bool:
    """The fused all-views acquisition needs this package's container (a reference ``ATModulesContainer`` whose modules call
    the patched ``nifti_grid_sample`` still works, view by view), 'from-gt' label slices, R6 modules, no re-alignment."""
    if atm_container is not None and not hasattr(atm_container, "acquire_from_labels"):
        return False
    if not getattr(config, "use_affine_theta", True) or not all(getattr(m, "use_affine_theta", True) for m in modules):
        return False          # 'ref' stage (running/stages.py:76-82): init affines only, the MLP heads are not applied
    return getattr(config, "label_slice_type", "from-gt") == "from-gt" and \
        all(getattr(m, "optim_method", None) == "R6-vector" and not getattr(m, "rotate_slice_to_min_principle", False)
            for m in modules)


_ONEHOT_MAX_CLASSES = 16      # afb_onehot.cu MAXC: channels the one-hot-from-index kernels hold in registers


def get_reconstruction_model_input(batch, phase, config, num_classes, atm_container, segment_fn=None, generator=None):
    """run_dl.py:238-329.  ``batch``: {'label' [B,D,H,W] int, 'image' [B,D,H,W] float, 'additional_data': {'nifti_affine'
    [B,4,4], 'gt_view_affines' | 'prescan_view_affines': {view name: [B,4,4], 'centroids': [B,4,4]}}}, CUDA tensors."""
    b_label, b_image = batch["label"], batch["image"]
    key = "gt_view_affines" if config.clinical_view_affine_type == "from-gt" else "prescan_view_affines"
    b_view_affines = batch["additional_data"][key]
    nifti_affine = batch["additional_data"]["nifti_affine"]
    base_affine = torch.as_tensor(b_view_affines["centroids"]).to(nifti_affine)

    with torch.no_grad():                                                              # :251-259
        hires_mm, hires_vox = torch.as_tensor(config.hires_fov_mm), torch.as_tensor(config.hires_fov_vox)
        b_label, _, nifti_affine = nifti_grid_sample(b_label.unsqueeze(1), nifti_affine, target_fov_mm=hires_mm,
                                                     target_fov_vox=hires_vox, is_label=True, pre_grid_sample_affine=base_affine)
        b_image, _, _ = nifti_grid_sample(b_image.unsqueeze(1), nifti_affine, target_fov_mm=hires_mm, target_fov_vox=hires_vox,
                                          is_label=False, pre_grid_sample_affine=base_affine)
        b_label = b_label.squeeze(1)                                                   # [B,D,H,W] integer, stays an index map
    B, D, H, W = b_label.shape

    for atm in atm_container:
        atm.use_affine_theta = config.use_affine_theta
    active = list(atm_container.get_active_view_modules())                            # :269-271
    b_aug = None
    if config.do_augment_input_orientation and phase in config.aug_phases:            # :273-278: ONE draw per batch element
        s = config.sample_augment_strength
        b_aug = draw_augmentation_affines(b_label.shape[0], rotation_strength=0.1 * s, zoom_strength=0.2 * s, offset_strength=0.0,
                                          generator=generator)
    input_grid_affines = [get_input_affine_for_atm(m, base_affine, b_view_affines, b_aug).to(torch.float32) for m in active]

    # the one-hot-from-index kernels take 2..16 classes (pad value 0 = min of a one-hot needs C >= 2) and <= 65535 slices
    fused = _fused_route_ok(config, active, atm_container) and 2 <= num_classes <= _ONEHOT_MAX_CLASSES and B * len(active) <= 65535
    if not fused:
        return _per_view_route(config, phase, num_classes, active, input_grid_affines, b_label, b_image, nifti_affine, segment_fn,
                               generator)

    # per-view MLP heads; gradient context per view as in :283-289
    mlp_outs = []
    for i, (m, ga_in) in enumerate(zip(active, input_grid_affines)):
        with_grad = config.view_optimization_mode == "opt-all" or \
            (config.view_optimization_mode == "opt-current-fix-previous" and i == len(active) - 1)
        with torch.enable_grad() if with_grad else torch.no_grad():
            mlp_outs.append(m.mlp_head_from_labels(b_label, num_classes, nifti_affine, ga_in))
    y_soft, _, y_image, grid_affines, _ = atm_container.acquire_from_labels(
        b_label, num_classes, b_image, nifti_affine, input_grid_affines, mlp_outs=mlp_outs, modules=active, label_out=None)

    if list(config.slice_fov_vox) != list(config.hires_fov_vox):                     # :193-197 up-sample low-res slices
        tgt = list(config.hires_fov_vox[:2]) + [1]
        y_soft = AF.upsample_slices(y_soft, tgt)                   # all B x V x C planes in one launch

    output_grid_affines = [grid_affines[:, v] for v in range(len(active))]
    if config.do_augment_recon_orientation and phase in config.aug_phases:           # :303-309
        s = config.sample_augment_strength
        output_grid_affines = [apply_affine_augmentation([g], rotation_strength=0.1 * s, zoom_strength=0.2 * s,
                                                         offset_strength=0.0, generator=generator)[0] for g in output_grid_affines]
    n_views, n_active = len(config.base_views), len(active)
    slices = [y_soft[:, v] for v in range(n_active)] + [y_soft[:, n_active - 1]] * (n_views - n_active)      # :321-323
    output_grid_affines = output_grid_affines + [output_grid_affines[-1]] * (n_views - n_active)
    b_input = torch.cat(slices, dim=1).squeeze(-1) if n_views != n_active else y_soft.flatten(1, 2).squeeze(-1)   # :325
    assert b_input.dim() == 4
    b_target = F.one_hot(b_label.long(), num_classes).permute(0, 4, 1, 2, 3)          # :261-262 (a view, as in the reference)
    return b_input, b_target, output_grid_affines


def _per_view_route(config, phase, num_classes, active, input_grid_affines, b_label, b_image, nifti_affine, segment_fn, generator):
    """The reference's loop over views (run_dl.py:261-329) on the materialised one-hot volumes: used whenever a view cannot
    take the fused all-views acquisition ('from-segmented', non-R6 parameterisations, in-plane re-alignment)."""
    label = F.one_hot(b_label.long(), num_classes).permute(0, 4, 1, 2, 3)
    soft = label.float()
    B, C, D, H, W = label.shape
    slices, grid_affines = [], []
    for i, (atm, input_ga) in enumerate(zip(active, input_grid_affines)):
        with_grad = config.view_optimization_mode == "opt-all" or \
            (config.view_optimization_mode == "opt-current-fix-previous" and i == len(active) - 1)
        with torch.enable_grad() if with_grad else torch.no_grad():
            _, label_slc, output_ga = get_transformed(config, phase, label, soft, nifti_affine, input_ga, atm,
                                                      image=b_image.view(B, 1, D, H, W), segment_fn=segment_fn)
            if config.do_augment_recon_orientation and phase in config.aug_phases:
                s = config.sample_augment_strength
                output_ga = apply_affine_augmentation([output_ga], rotation_strength=0.1 * s, zoom_strength=0.2 * s,
                                                      offset_strength=0.0, generator=generator)[0].to(nifti_affine)
            slices.append(label_slc)
            grid_affines.append(output_ga)
    n_views, n_active = len(config.base_views), len(active)
    slices = slices + [slices[-1]] * (n_views - n_active)
    grid_affines = grid_affines + [grid_affines[-1]] * (n_views - n_active)
    b_input = torch.cat(slices, dim=1).squeeze(-1)
    assert b_input.dim() == 4
    return b_input, label, grid_affines
