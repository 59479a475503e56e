"""Drop-in for the hot-path part of the reference's ``acquisition_focus/utils/nifti_utils.py``.

Same names, argument meaning and error behaviour as the reference for

* ``nifti_grid_sample``                    (reference ``utils/nifti_utils.py:112-207``)
* ``get_zooms``                            (``:254-256``)
* ``rescale_rot_components_with_diag``     (``:27-32``)

but the work (fp64 affine bookkeeping ``:36-71``, ``affine_grid`` ``:182``, min-shift ``:200-203`` and
``grid_sample`` ``:87-94`` incl. their autograd) runs in the CUDA kernels of libafb200.so.  CUDA tensors only:
there is no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import functional as AF


def get_zooms(nii_affine: torch.Tensor) -> torch.Tensor:
    """Column norms of the 3x3 block: voxel spacings of a NIfTI affine."""
    assert nii_affine.dim() == 3
    return torch.linalg.vector_norm(nii_affine[:, :3, :3], dim=1)


def rescale_rot_components_with_diag(affine: torch.Tensor, scaler: torch.Tensor) -> torch.Tensor:
    """``affine @ diag(scaler, 1)``: scale the three rotation columns."""
    ones = torch.ones_like(scaler[:, :1])
    return affine * torch.cat([scaler, ones], dim=1)[:, None, :]


def nifti_grid_sample(volume: torch.Tensor, volume_nii_affine: torch.Tensor,
                      ras_transform_affine: torch.Tensor = None,
                      target_fov_mm: torch.Tensor = None, target_fov_vox: torch.Tensor = None,
                      is_label: bool = False, pre_grid_sample_affine: torch.Tensor = None, dtype=torch.float32):
    """Resample ``volume[B,C,D,H,W]`` into a target field of view, keeping track of its NIfTI affine.

    Returns ``(transformed[B,C,*target_fov_vox], grid_affine[B,4,4], transformed_nii_affine[B,4,4])`` exactly like
    the reference; ``grid_affine`` is differentiable w.r.t. ``pre_grid_sample_affine`` and ``transformed`` w.r.t.
    both the volume and the affine.  Out-of-field samples evaluate to ``volume.min()`` for ``is_label=False``
    (the reference's min-shift) and to 0 for ``is_label=True``.
    """
    assert volume.dim() == 5
    assert isinstance(volume, torch.Tensor) and isinstance(volume_nii_affine, torch.Tensor)
    if pre_grid_sample_affine is not None:
        assert isinstance(pre_grid_sample_affine, torch.Tensor)
    B, C, D, H, W = volume.shape
    if target_fov_vox is None:
        fov_vox = (D, H, W)
    else:
        fov_vox = tuple(int(v) for v in torch.as_tensor(target_fov_vox).tolist())
    if target_fov_mm is None:
        fov_mm = (0.0, 0.0, 0.0)          # library convention: keep the input field of view
    else:
        fov_mm = tuple(float(v) for v in torch.as_tensor(target_fov_mm).tolist())
    if pre_grid_sample_affine is not None:
        assert pre_grid_sample_affine.dim() == 3 and B == pre_grid_sample_affine.shape[0]
    if ras_transform_affine is not None:
        raise Warning("Providing a RAS space transform matrix is experimental and might produce wrong results.")
    assert volume_nii_affine.dim() == 3 and B == volume_nii_affine.shape[0]
    if pre_grid_sample_affine is None:
        pre_grid_sample_affine = torch.eye(4, dtype=torch.float64, device=volume.device)[None].repeat(B, 1, 1)
    elif pre_grid_sample_affine.shape[0] != B:
        pre_grid_sample_affine = pre_grid_sample_affine.expand(B, 4, 4)

    out, grid_affine, nii = AF.slice_with_pre_affine(
        volume, volume_nii_affine, pre_grid_sample_affine, fov_mm, fov_vox, is_label=is_label,
        pad="global_min")
    if "int" in str(volume.dtype):
        grid_affine = grid_affine.to(dtype)
    elif volume.dtype != torch.float32:
        grid_affine = grid_affine.to(volume.dtype)
    return out, grid_affine, nii
