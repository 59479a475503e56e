"""Drop-in for the view-parameter tail of the reference's ``models/learnable_transform.py``.

``AffineTransformModule`` keeps the reference's constructor arguments, parameter names
(``init_theta_ap``, ``init_theta_t_offsets``, ``init_theta_zp``), ``forward`` signature and return
tuple (reference ``:64-333``).  What changes is where the work happens: the R6 -> rotation,
soft-argmax offset, tanh zoom, ``T@R@Z``, ``Gpre @ theta`` composition (``:144-230, :262-289``), the
fp64 NIfTI bookkeeping and the three slicings (``:287-306``) run inside the fused CUDA sampler, and
their gradients in its backward epilogue.  The LocalizationNet (dense 3-D convolutions, cuDNN) is out
of scope of this path and is passed in (or a small default is built).

``ATModulesContainer.acquire`` additionally slices ALL views of a batch in one launch
(B x V slices), which is what ``running/run_dl.py:283-312`` loops over in Python.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from .. import functional as AF
from ..utils.nifti_utils import nifti_grid_sample
from ..utils.transform_utils import (angle_axis_to_rotation_matrix, compute_rotation_matrix_from_ortho6d,
                                     get_random_affine, normal_to_rotation_matrix)


class DefaultLocalizationNet(nn.Module):
    """Small 3-D conv encoder + linear head producing ``[B, 6 + 3R + 1]`` (placeholder for the reference's
    LocalizationNet, ``models/learnable_transform.py:38-60``; any module with that output works)."""

    def __init__(self, input_channels: int, output_size: int, size_3d: Sequence[int]):
        super().__init__()
        self.conv_net = nn.Sequential(
            nn.Conv3d(input_channels, 16, 5, stride=2, padding=2), nn.InstanceNorm3d(16), nn.LeakyReLU(),
            nn.Conv3d(16, 32, 3, stride=2, padding=1), nn.InstanceNorm3d(32), nn.LeakyReLU(),
            nn.Conv3d(32, 32, 3, stride=2, padding=1), nn.InstanceNorm3d(32), nn.LeakyReLU(),
            nn.AdaptiveAvgPool3d(2))
        self.fc = nn.Linear(32 * 8, output_size)
        nn.init.zeros_(self.fc.weight)
        nn.init.zeros_(self.fc.bias)

    def forward(self, x):
        return self.fc(self.conv_net(x).flatten(1))


class AffineTransformModule(nn.Module):
    def __init__(self, input_channels, volume_fov_mm, volume_fov_vox, slice_fov_mm, slice_fov_vox,
                 optim_method="angle-axis", use_affine_theta=True, offset_clip_value=1.0, zoom_clip_value=2.0,
                 view_id=None, align_corners=False, rotate_slice_to_min_principle=False, localization_net=None):
        super().__init__()
        assert volume_fov_mm[0] == volume_fov_mm[1] == volume_fov_mm[2]
        assert volume_fov_vox[0] == volume_fov_vox[1] == volume_fov_vox[2]
        assert optim_method in ["angle-axis", "normal-vector", "R6-vector"], \
            f"optim_method must be 'angle-axis', 'normal-vector' or 'R6-vector', not {optim_method}"
        if align_corners:
            raise NotImplementedError("align_corners=True is never set by the reference's configs and is not part of "
                                      "the accelerated path")
        self.optim_method = optim_method
        if optim_method == "R6-vector":                       # fused into the CUDA view prologue
            self.ap_space, self.optim_function = 6, compute_rotation_matrix_from_ortho6d
            self.init_theta_ap = nn.Parameter(torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0]]), requires_grad=False)
        else:                                                 # 3-parameter closed forms -> the same CUDA sampler
            self.ap_space = 3
            self.optim_function = angle_axis_to_rotation_matrix if optim_method == "angle-axis" else normal_to_rotation_matrix
            self.init_theta_ap = nn.Parameter(torch.zeros(3), requires_grad=False)
        self.volume_fov_mm = torch.as_tensor(volume_fov_mm)
        self.volume_fov_vox = torch.as_tensor(volume_fov_vox)
        self.slice_fov_vox = torch.as_tensor(slice_fov_vox)
        self.slice_fov_mm = torch.as_tensor(slice_fov_mm)
        self.use_affine_theta = use_affine_theta
        self.align_corners = align_corners
        self.rotate_slice_to_min_principle = rotate_slice_to_min_principle
        self.offset_clip_value = float(offset_clip_value)
        self.zoom_clip_value = float(zoom_clip_value)
        self.spat = int(self.volume_fov_vox[0])
        # vox_range / arra: width of the soft-argmax support in voxels (reference :112-116)
        self.vox_range = int(round(self.get_vox_offsets_from_gs_offsets(self.offset_clip_value)
                                   - self.get_vox_offsets_from_gs_offsets(-self.offset_clip_value)))
        self.arra = torch.arange(0, self.vox_range) + (self.spat - self.vox_range) // 2
        out_size = self.ap_space + 3 * self.vox_range + 1
        self.localization_net = localization_net if localization_net is not None else \
            DefaultLocalizationNet(input_channels, out_size, [int(v) for v in self.volume_fov_vox])
        self.view_id = view_id
        self.init_theta_t_offsets = nn.Parameter(torch.zeros([3]), requires_grad=False)
        self.init_theta_zp = nn.Parameter(torch.ones([1, 1]), requires_grad=False)
        self.last_theta = None
        self.last_grid_affine = None
        self.last_transformed_nifti_affine = None
        self.random_grid_affine = get_random_affine(rotation_strength=4.0, zoom_strength=0.0)[None]   # global RNG (:136)

    # -- same small API as the reference ---------------------------------------------------------
    def set_init_theta_ap(self, init_theta_ap):
        self.init_theta_ap.data = init_theta_ap.data

    def set_init_theta_t_offsets(self, init_theta_t_offsets):
        self.init_theta_t_offsets.data = init_theta_t_offsets.data

    def set_init_theta_zp(self, init_theta_zp):
        self.init_theta_zp.data = init_theta_zp.data

    def get_vox_offsets_from_gs_offsets(self, gs_offsets):
        return ((gs_offsets + 1.0) * self.spat - 1.0) / 2.0

    def init_vector(self) -> torch.Tensor:
        """``[10]``: init_theta_ap | init_theta_t_offsets | init_theta_zp, the layout afb_views.init wants."""
        assert self.optim_method == "R6-vector", "the fused view prologue takes R6 parameters"
        return torch.cat([self.init_theta_ap.reshape(6).float(), self.init_theta_t_offsets.reshape(3).float(),
                          self.init_theta_zp.reshape(1).float()])

    def get_init_affines(self):
        """Init rotation / translation / zoom 4x4s (reference :144-161); host-side convenience."""
        dev = self.init_theta_t_offsets.device
        if dev.type != "cuda":
            raise RuntimeError("get_init_affines needs the module on a CUDA device (no CPU fallback)")
        theta_a = self.optim_function(self.init_theta_ap.view(1, self.ap_space))
        theta_t = torch.eye(4, device=dev)[None].clone()
        theta_t[0, :3, 3] = self.init_theta_t_offsets
        z = self.init_theta_zp.view(1)
        theta_z = torch.diag_embed(torch.cat([z, z, z, torch.ones(1, device=dev)]))[None]
        return theta_a.float(), theta_t.float(), theta_z.float()

    def get_gs_offsets_from_theta_tp(self, theta_tp):
        """Soft-argmax offsets (reference :163-176, ``align_corners=False``): ``[B,3,R] -> [B,3]``."""
        pos = (torch.softmax(theta_tp, dim=2) * self.arra.to(theta_tp).view(1, 1, self.vox_range)).sum(-1)
        return (2.0 * pos + 1.0) / self.spat - 1.0

    def theta_from_mlp_out(self, mlp_out):
        """``T @ R @ Z`` from the MLP-head output with torch ops (reference :193-230, :262-272); the R6 method has this
        fused into the CUDA prologue instead (``AF.acquire_views``), the other two parameterisations come through here."""
        B, A, R = mlp_out.shape[0], self.ap_space, self.vox_range
        dev = mlp_out.device
        ap = mlp_out[:, :A] + self.init_theta_ap.view(1, A).to(dev)
        zp = mlp_out[:, -1:] + self.init_theta_zp.view(1, 1).to(dev)
        if self.optim_method == "normal-vector":
            ap = ap / ap.norm(dim=1).view(-1, 1)
        offs = self.get_gs_offsets_from_theta_tp(mlp_out[:, A:-1].view(B, 3, R))
        if self.offset_clip_value == 0.0:
            offs = 0.0 * offs
        theta_t = torch.eye(4, device=dev)[None].repeat(B, 1, 1)
        theta_t[:, :3, 3] = offs
        z = self.zoom_clip_value * -(zp.tanh()) + 1.0
        theta_z = torch.diag_embed(torch.cat([z, z, z, torch.ones(B, 1, device=dev)], dim=-1))
        a0, t0, z0 = self.get_init_affines()
        return (t0 @ theta_t) @ (a0 @ self.optim_function(ap)) @ (z0 @ theta_z)

    def mlp_head(self, x_soft_label, nifti_affine, grid_affine_pre_mlp):
        """LocalizationNet input (pre-oriented prescan volume, reference :248-255) -> ``[B, 6+3R+1]``."""
        with torch.no_grad():
            x_pre, _, _ = nifti_grid_sample(x_soft_label, nifti_affine, target_fov_mm=self.volume_fov_mm,
                                            target_fov_vox=self.volume_fov_vox, is_label=False,
                                            pre_grid_sample_affine=grid_affine_pre_mlp)
        return self.localization_net(x_pre)

    def mlp_head_from_labels(self, label_map, num_classes, nifti_affine, grid_affine_pre_mlp):
        """Same as :meth:`mlp_head` but from the integer label map ``[B,D,H,W]`` (no one-hot volume is materialised)."""
        with torch.no_grad():
            x_pre, _, _ = AF.onehot_resample_with_pre_affine(label_map, nifti_affine, grid_affine_pre_mlp,
                                                             self.volume_fov_mm.tolist(), self.volume_fov_vox.tolist(), num_classes)
        return self.localization_net(x_pre)

    def forward(self, x_soft_label, x_label, x_image, nifti_affine, grid_affine_pre_mlp, theta_override=None):
        soft_none = x_soft_label is None or x_soft_label.numel() == 0
        assert not soft_none
        B = x_soft_label.shape[0]
        dev = x_soft_label.device
        gpre = grid_affine_pre_mlp.to(dev, torch.float32)
        if theta_override is not None or not self.use_affine_theta or self.optim_method != "R6-vector":
            if theta_override is not None:
                theta = theta_override.detach().clone().to(dev, torch.float32)       # non-differentiable (:260)
            elif self.use_affine_theta:
                theta = self.theta_from_mlp_out(self.mlp_head(x_soft_label, nifti_affine, gpre))     # differentiable
            else:
                a, t, z = self.get_init_affines()
                theta = (t @ a @ z).repeat(B, 1, 1)
            self.last_theta = theta
            pre = gpre @ theta
            kw = dict(target_fov_mm=self.slice_fov_mm, target_fov_vox=self.slice_fov_vox, pre_grid_sample_affine=pre)
            y_soft, grid_affine, nii = nifti_grid_sample(x_soft_label, nifti_affine, is_label=False, **kw)
            y_label = y_image = None
            with torch.no_grad():
                if x_label is not None and x_label.numel() > 0:
                    y_label = nifti_grid_sample(x_label, nifti_affine, is_label=True, **kw)[0]
                if x_image is not None and x_image.numel() > 0:
                    y_image = nifti_grid_sample(x_image, nifti_affine, is_label=False, **kw)[0]
        else:
            mlp_out = self.mlp_head(x_soft_label, nifti_affine, gpre)
            y_soft, y_label, y_image, grid_affine, nii, theta = AF.acquire_views(
                x_soft_label, x_label, x_image, nifti_affine, gpre[:, None], mlp_out[:, None], self.init_vector()[None],
                offset_clip=self.offset_clip_value, zoom_clip=self.zoom_clip_value, spat=self.spat,
                slice_fov_mm=self.slice_fov_mm.tolist(), slice_fov_vox=self.slice_fov_vox.tolist())
            y_soft, grid_affine, nii, theta = y_soft[:, 0], grid_affine[:, 0], nii[:, 0], theta[:, 0]
            y_label = None if y_label is None else y_label[:, 0]
            y_image = None if y_image is None else y_image[:, 0]
            self.last_theta = theta
        if self.rotate_slice_to_min_principle:                                           # reference :315-328
            # quirk kept on purpose: the reference threads ONE NIfTI affine through the three calls, so the alignment is
            # applied to it once per non-empty input (soft, label, image), not once
            y_soft, align_affine, nii = rotate_slice_to_min_principle(y_soft, nii, is_label=False)
            with torch.no_grad():
                if y_label is not None:
                    y_label, _, nii = rotate_slice_to_min_principle(y_label, nii, is_label=True, align_affine_override=align_affine)
                if y_image is not None:
                    y_image, _, nii = rotate_slice_to_min_principle(y_image, nii, is_label=False, align_affine_override=align_affine)
            grid_affine = grid_affine @ align_affine
        self.last_grid_affine = grid_affine
        self.last_transformed_nifti_affine = nii
        return y_soft, y_label, y_image, grid_affine, nii


def min_principle_align_affines(x_input: torch.Tensor) -> torch.Tensor:
    """``[B,C,H,W,1]`` slices -> ``[B,4,4]`` torch-grid affines that turn each slice so that the axis of least inertia of its
    foreground (``argmax`` over channels != 0) lies along the first in-plane axis (reference :346-355 with
    ``utils/torch_sparse_tensor_utils.py:34-56,79-85`` and ``functional/clinical_cardiac_views.py:66-100``).

    The second moments of the foreground are batched device reductions (fp64); the 3x3 eigenproblem and the frame are a few
    dozen flops per sample and run on the host with ``torch.linalg.eig`` exactly as the reference does (one small D2H per
    call), which also pins the sign convention of the eigenvectors to the reference's LAPACK path."""
    assert x_input.dim() == 5 and x_input.shape[-1] == 1
    B, _, H, W, _ = x_input.shape
    dev = x_input.device
    fg = (x_input.argmax(1)[..., 0] != 0).to(torch.float64)                              # [B,H,W]
    hh = torch.arange(H, device=dev, dtype=torch.float64).view(1, H, 1)
    ww = torch.arange(W, device=dev, dtype=torch.float64).view(1, 1, W)
    n = fg.sum((1, 2))
    ch, cw = (fg * hh).sum((1, 2)) / n, (fg * ww).sum((1, 2)) / n
    dh, dw = hh - ch.view(B, 1, 1), ww - cw.view(B, 1, 1)
    shh, sww, shw = (fg * dh * dh).sum((1, 2)), (fg * dw * dw).sum((1, 2)), (fg * dh * dw).sum((1, 2))
    zero = torch.zeros_like(shh)
    inertia = torch.stack([sww, -shw, zero, -shw, shh, zero, zero, zero, shh + sww], dim=1).view(B, 3, 3)
    inertia, center = inertia.float().cpu(), torch.stack([ch, cw, torch.full_like(ch, 0.5)], dim=1).float().cpu()
    shape = torch.tensor([H, W, 1], dtype=torch.float32)
    out = torch.zeros(B, 4, 4)
    for b in range(B):
        eig = torch.linalg.eig(inertia[b])
        main = eig.eigenvectors.real.T[eig.eigenvalues.real.argsort()][0]
        main = main / torch.linalg.norm(main)
        two = torch.linalg.cross(main, torch.tensor([0.0, 0.0, 1.0]))
        two = two / torch.linalg.norm(two)
        normal = torch.linalg.cross(main, two)
        normal = normal / torch.linalg.norm(normal)
        two = torch.linalg.cross(normal, main)
        frame = torch.stack([two, main, normal], dim=0)                                  # pixel-space rows
        out[b, :3, :3] = frame.flip(0, 1).T                                              # pixel -> torch-grid axis convention
        out[b, :3, 3] = (2.0 * center[b] / shape - 1.0).flip(0)
        out[b, 3, 3] = 1.0
    return out.to(dev)


def rotate_slice_to_min_principle(x_input, nii_affine, is_label=False, align_affine_override=None):
    """Reference :337-366: re-align a one-voxel-thin slice volume ``[B,C,H,W,1]`` in-plane; same signature and return tuple
    ``(y_output, b_align_affines, transformed_nii_affine)``.  The resample runs in the CUDA sampler (differentiable w.r.t.
    ``x_input``); the alignment affine itself carries no gradient, as in the reference."""
    assert x_input.shape[-1] == 1
    if align_affine_override is None:
        with torch.no_grad():
            b_align_affines = min_principle_align_affines(x_input)
    else:
        b_align_affines = align_affine_override
    return nifti_grid_sample(x_input, nii_affine, pre_grid_sample_affine=b_align_affines, is_label=is_label)


class ATModulesContainer(nn.ModuleList):
    """One AffineTransformModule per base view (reference :370-414) + a fused all-views acquisition."""

    def __init__(self, config, num_classes, localization_net_factory=None):
        super().__init__()
        for view_id in config.base_views:
            self.add_new_atm(view_id, config, num_classes, localization_net_factory)
        self.is_optimized = nn.Parameter(torch.tensor(len(config.base_views) * [False]), requires_grad=False)

    def add_new_atm(self, view_id, config, num_classes, localization_net_factory=None):
        net = localization_net_factory() if localization_net_factory is not None else None
        self.append(AffineTransformModule(
            num_classes, torch.tensor(config.prescan_fov_mm), torch.tensor(config.prescan_fov_vox),
            torch.tensor(config.slice_fov_mm), torch.tensor(config.slice_fov_vox),
            offset_clip_value=config.offset_clip_value, zoom_clip_value=config.zoom_clip_value,
            optim_method=config.affine_theta_optim_method, view_id=view_id,
            rotate_slice_to_min_principle=config.rotate_slice_to_min_principle, localization_net=net))

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        sd.update({"is_optimized": self.get_active_views()})
        return sd

    def get_all_atms_requires_grad(self):
        return torch.tensor([next(atm.localization_net.parameters()).requires_grad for atm in self])

    def get_active_views(self):
        return self.is_optimized.cpu() | self.get_all_atms_requires_grad()

    def get_active_view_modules(self):
        return [m for m, a in zip(self, self.get_active_views()) if a]

    def get_next_non_optimized_view_module(self):
        """reference :407-411 (used by the training loop, running/run_dl.py:117,125,416)."""
        next_idx = self.__get_next_non_optimized_idx__()
        if next_idx is not None:
            return self[next_idx]
        return None

    def __get_next_non_optimized_idx__(self):
        if False in self.is_optimized:
            return (self.is_optimized == False).nonzero()[0]          # noqa: E712  (tensor comparison, as in the reference)
        return None

    def acquire(self, x_soft_label, x_label, x_image, nifti_affine, view_pre_affines, mlp_outs: Optional[list] = None):
        """All views in ONE launch per volume kind.  ``view_pre_affines``: list (len V) of ``[B,4,4]``;
        ``mlp_outs``: optional list of ``[B, 6+3R+1]`` (else each module's LocalizationNet is run).
        Returns ``(y_soft[B,V,C,H,W,1], y_label, y_image, grid_affines[B,V,4,4], nii[B,V,4,4])``."""
        atms = list(self)
        assert all(m.optim_method == "R6-vector" for m in atms), "the fused all-views acquisition takes R6 parameters"
        assert all(m.use_affine_theta for m in atms), "use_affine_theta=False (init affines only) goes view by view"
        dev = x_soft_label.device
        gpre = torch.stack([g.to(dev, torch.float32) for g in view_pre_affines], dim=1)
        if mlp_outs is None:
            mlp_outs = [m.mlp_head(x_soft_label, nifti_affine, g) for m, g in zip(atms, view_pre_affines)]
        params = torch.stack(mlp_outs, dim=1)
        init = torch.stack([m.init_vector() for m in atms]).to(dev)
        a0 = atms[0]
        y_soft, y_label, y_image, ga, nii, theta = AF.acquire_views(
            x_soft_label, x_label, x_image, nifti_affine, gpre, params, init, offset_clip=a0.offset_clip_value,
            zoom_clip=a0.zoom_clip_value, spat=a0.spat, slice_fov_mm=a0.slice_fov_mm.tolist(),
            slice_fov_vox=a0.slice_fov_vox.tolist())
        for v, m in enumerate(atms):
            m.last_theta, m.last_grid_affine, m.last_transformed_nifti_affine = theta[:, v], ga[:, v], nii[:, v]
        return y_soft, y_label, y_image, ga, nii

    def acquire_from_labels(self, label_map, num_classes, x_image, nifti_affine, view_pre_affines, mlp_outs: Optional[list] = None,
                            modules: Optional[list] = None, label_out="onehot"):
        """:meth:`acquire` from the INTEGER label map ``[B,D,H,W]``: neither ``one_hot(label).float()`` nor the int64
        one-hot volume is materialised (``running/run_dl.py:261-264``); gradients flow to the view parameters only, which is
        all the reference's training step consumes.  ``modules``: the active view modules (default: all)."""
        atms = list(self) if modules is None else list(modules)
        assert all(m.optim_method == "R6-vector" for m in atms), "the fused all-views acquisition takes R6 parameters"
        assert all(m.use_affine_theta for m in atms), "use_affine_theta=False (init affines only) goes view by view"
        dev = label_map.device
        gpre = torch.stack([g.to(dev, torch.float32) for g in view_pre_affines], dim=1)
        if mlp_outs is None:
            mlp_outs = [m.mlp_head_from_labels(label_map, num_classes, nifti_affine, g) for m, g in zip(atms, view_pre_affines)]
        params = torch.stack(mlp_outs, dim=1)
        init = torch.stack([m.init_vector() for m in atms]).to(dev)
        a0 = atms[0]
        y_soft, y_label, y_image, ga, nii, theta = AF.acquire_views_from_labels(
            label_map, x_image, nifti_affine, gpre, params, init, num_classes=num_classes, offset_clip=a0.offset_clip_value,
            zoom_clip=a0.zoom_clip_value, spat=a0.spat, slice_fov_mm=a0.slice_fov_mm.tolist(),
            slice_fov_vox=a0.slice_fov_vox.tolist(), label_out=label_out)
        for v, m in enumerate(atms):
            m.last_theta, m.last_grid_affine, m.last_transformed_nifti_affine = theta[:, v], ga[:, v], nii[:, v]
        return y_soft, y_label, y_image, ga, nii
