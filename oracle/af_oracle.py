"""ORACLE (test infrastructure only; the product never imports this).

CPU restatement ("port") of the reference's differentiable view-acquisition
path, written against torch-CPU so that it can travel to the GPU box (where
``/root/reference`` does not exist).  It is only ever used as

* the checker in ``tests/`` and ``__graft_entry__.smoke()``, and
* the timed ``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``.

Each function cites the reference lines it follows (paths relative to
``/root/reference/acquisition_focus``).  The sampling arithmetic itself is
delegated -- exactly as the reference does -- to ATen's ``affine_grid`` /
``grid_sample`` (third-party, pinned torch 2.0.0, ``pyproject.toml:11``; here
torch 2.11 CPU); ``oracle/aten_np.py`` restates those two primitives bitwise in
numpy.  The op *sequence* of the reference is kept on purpose (fp64 affine
bookkeeping, whole-volume min-shift, re-entrant checkpoint around
``grid_sample``) so that timing this port on host cores is representative of
the reference's own CPU path.

PARITY PINNING: the reference holds no tests, fixtures or golden vectors for
this path (``/root/reference/tests/__init__.py`` is empty).  This port is
therefore pinned against outputs of the reference itself, run in the build
container: ``oracle/make_golden.py`` -> ``tests/golden/*.npz`` (checked by
``tests/test_oracle_golden.py`` everywhere) and live, function by function,
in ``tests/test_oracle_vs_reference.py`` wherever ``/root/reference`` exists.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

warnings.filterwarnings("ignore", message=".*use_reentrant.*")
warnings.filterwarnings("ignore", message=".*None of the inputs have requires_grad.*")


# ----------------------------------------------------------------------------
# a1  R6 -> rotation                                   utils/transform_utils.py:27-58
# ----------------------------------------------------------------------------
def r6_to_matrix(ortho: torch.Tensor) -> torch.Tensor:
    """Gram-Schmidt on two 3-vectors; columns of R are (x, y, z); homogeneous 4x4."""
    a, b = ortho[:, 0:3], ortho[:, 3:6]
    x = a / a.norm(dim=1, keepdim=True)
    z = torch.linalg.cross(x, b, dim=1)
    z = z / z.norm(dim=1, keepdim=True)
    y = torch.linalg.cross(z, x, dim=1)
    rot = torch.stack([x, y, z], dim=2)                       # [B,3,3], columns x,y,z
    out = torch.zeros(ortho.shape[0], 4, 4, dtype=ortho.dtype, device=ortho.device)
    out[:, :3, :3] = rot
    out[:, 3, 3] = 1.0
    return out


# ----------------------------------------------------------------------------
# f4  the two non-default rotation parameterisations    utils/transform_utils.py:62-178
# ----------------------------------------------------------------------------
def normal_to_matrix(normals: torch.Tensor) -> torch.Tensor:
    """``normal_to_rotation_matrix`` (:62-103): the input columns are read as (nz, ny, nx); rows of R are the in-plane
    direction (ny, -nx, 0)/d, its complement (nx nz, ny nz, -d^2)/d and the normal (nx, ny, nz), d = sqrt(nx^2+ny^2).
    Degenerate (d = 0, e.g. the module's all-zero init) divides by zero exactly as the reference does."""
    nz, ny, nx = normals[:, 0], normals[:, 1], normals[:, 2]
    d = torch.sqrt(nx ** 2 + ny ** 2)
    zero, one = torch.zeros_like(nx), torch.ones_like(nx)
    rows = [ny / d, -nx / d, zero, zero,
            nx * nz / d, ny * nz / d, -d, zero,
            nx, ny, nz, zero,
            zero, zero, zero, one]
    return torch.stack(rows, dim=1).view(-1, 4, 4)


def angle_axis_to_matrix(angle_axis: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """``angle_axis_to_rotation_matrix`` (:106-178, after ceres/rotation.h): Rodrigues' formula with the reference's
    double-eps normalisation ``w = r / (sqrt(|r|^2 + eps) + eps)`` where |r|^2 > eps, first-order ``I + [r]x`` otherwise;
    both branches are evaluated and blended with 0/1 masks (so both contribute to autograd, as in the reference)."""
    r = angle_axis
    theta2 = (r.to(torch.float32)[:, None, :] @ r.to(torch.float32)[:, :, None]).squeeze(1)        # [N,1]  (:156-159)
    theta = torch.sqrt(theta2 + eps)                                                                # :127
    w = r / (theta + eps)                                                                           # :128
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c, s = torch.cos(theta), torch.sin(theta)
    full = torch.cat([c + wx * wx * (1.0 - c), wx * wy * (1.0 - c) - wz * s, wy * s + wx * wz * (1.0 - c),
                      wz * s + wx * wy * (1.0 - c), c + wy * wy * (1.0 - c), -wx * s + wy * wz * (1.0 - c),
                      -wy * s + wx * wz * (1.0 - c), wx * s + wy * wz * (1.0 - c), c + wz * wz * (1.0 - c)], dim=1).view(-1, 3, 3)
    rx, ry, rz = r[:, 0:1], r[:, 1:2], r[:, 2:3]
    k1 = torch.ones_like(rx)
    taylor = torch.cat([k1, -rz, ry, rz, k1, -rx, -ry, rx, k1], dim=1).view(-1, 3, 3)              # :146-151
    big = (theta2 > eps).view(-1, 1, 1)
    out = torch.eye(4).to(r.device).type_as(r).view(1, 4, 4).repeat(r.shape[0], 1, 1)
    out[..., :3, :3] = big.type_as(theta2) * full + (~big).type_as(theta2) * taylor                 # :166-177
    return out


AP_SPACE = {"R6-vector": 6, "angle-axis": 3, "normal-vector": 3}


def rotation_of(optim_method: str, ap: torch.Tensor) -> torch.Tensor:
    return {"R6-vector": r6_to_matrix, "angle-axis": angle_axis_to_matrix, "normal-vector": normal_to_matrix}[optim_method](ap)


# ----------------------------------------------------------------------------
# a7  affine bookkeeping                               utils/nifti_utils.py:7-83,98-108,254-256
# ----------------------------------------------------------------------------
def column_norms(m: torch.Tensor) -> torch.Tensor:
    """``get_zooms`` (nifti_utils.py:254-256): L2 norm of each of the 3 columns."""
    return m[:, :3, :3].pow(2).sum(dim=1).sqrt()


def _swap02(m: torch.Tensor) -> torch.Tensor:
    """``switch_0_2_mat_dim`` (nifti_utils.py:19-23): conjugation J.M.J with J
    exchanging axes 0 and 2 (D<->W), i.e. reverse the first three rows and cols."""
    idx = torch.tensor([2, 1, 0, 3], device=m.device)
    return m.index_select(1, idx).index_select(2, idx)


def _scale_columns(m: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """``rescale_rot_components_with_diag`` (nifti_utils.py:27-32): m @ diag(s,1)."""
    d = torch.zeros_like(m)
    d[:, 0, 0], d[:, 1, 1], d[:, 2, 2] = s[:, 0], s[:, 1], s[:, 2]
    d[:, 3, 3] = 1.0
    return m @ d


def grid_and_nii_affine(nii_affine, fov_vox_in, fov_mm_out, fov_vox_out, pre_affine):
    """nifti_utils.py:36-71 with the no-op RAS matrix of :98-108 folded in.

    All arguments float64; ``fov_*`` in (D,H,W) order.  Returns the torch-space
    4x4 grid affine G' and the NIfTI affine of the resampled array.
    """
    B = nii_affine.shape[0]
    centre = torch.eye(4, dtype=nii_affine.dtype, device=nii_affine.device).repeat(B, 1, 1)
    centre[:, :3, 3] += fov_vox_in / 2.0                                    # :103-105
    ras = nii_affine @ centre                                               # :107
    m = nii_affine.inverse() @ ras                                          # :40
    zooms_in = column_norms(nii_affine)                                     # :44
    fov_mm_in = zooms_in * fov_vox_in                                       # :43
    m[:, :3, 3] = m[:, :3, 3] * 2.0 / fov_vox_in - 1.0                      # :48, :81-83
    m = _swap02(m)                                                          # :49
    m = m @ pre_affine                                                      # :52
    scale = ((1.0 / column_norms(m)) * (fov_mm_out / fov_mm_in)).flip(1)    # :57 (flip applies to the product)
    g = _scale_columns(m, scale)                                            # :55-58
    n = _swap02(g.clone())                                                  # :61-62
    n = _scale_columns(n, fov_mm_in / (fov_vox_out * zooms_in))             # :63-65
    n[:, :3, 3] = (n[:, :3, 3] + 1.0) / 2.0 * fov_vox_in                    # :66, :75-77
    half = nii_affine[:, :3, :3] @ n[:, :3, :3] @ (-(fov_vox_out - 1.0) / 2.0).to(nii_affine)   # :67
    n = nii_affine @ n                                                      # :69
    n[:, :3, 3] += half                                                     # :70
    return g, n


# ----------------------------------------------------------------------------
# a8/a9/a10  the sampler                               utils/nifti_utils.py:87-94,112-207
# ----------------------------------------------------------------------------
def _sample(volume, grid, mode):
    """``do_sample`` (:87-94): checkpointed (re-entrant) when anything needs grad."""
    if volume.requires_grad or grid.requires_grad:
        return checkpoint(F.grid_sample, volume, grid, mode, "zeros", False, use_reentrant=True)
    return F.grid_sample(volume, grid, mode=mode, padding_mode="zeros", align_corners=False)


def nifti_grid_sample(volume, volume_nii_affine, ras_transform_affine=None, target_fov_mm=None,
                      target_fov_vox=None, is_label=False, pre_grid_sample_affine=None,
                      dtype=torch.float32):
    """Same signature, argument meaning and error behaviour as nifti_utils.py:112-207."""
    assert volume.dim() == 5                                                               # :128-129
    assert isinstance(volume, torch.Tensor) and isinstance(volume_nii_affine, torch.Tensor)  # :131
    if pre_grid_sample_affine is not None:
        assert isinstance(pre_grid_sample_affine, torch.Tensor)
    dev, in_dtype = volume.device, volume.dtype
    B, C, D, H, W = volume.shape
    fov_vox_in = torch.tensor([D, H, W], device=dev).double()                               # :138
    if target_fov_mm is None:
        target_fov_mm = column_norms(volume_nii_affine) * fov_vox_in                        # :140-141
    if target_fov_vox is None:
        target_fov_vox = torch.as_tensor(volume.shape[-3:])                                 # :142-143
    out_shape = torch.Size([B, C] + [int(v) for v in target_fov_vox.tolist()])             # :145
    if pre_grid_sample_affine is not None:
        assert pre_grid_sample_affine.dim() == 3 and B == pre_grid_sample_affine.shape[0]  # :147-149
    target_fov_mm = target_fov_mm.to(dev).double()
    target_fov_vox = target_fov_vox.to(dev).double()
    volume_nii_affine = volume_nii_affine.to(dev).double()
    if ras_transform_affine is not None:                                                    # :157-158
        raise Warning("Providing a RAS space transform matrix is experimental and might produce wrong results.")
    assert volume_nii_affine.dim() == 3 and B == volume_nii_affine.shape[0]                 # :162-164
    if pre_grid_sample_affine is None:
        pre_grid_sample_affine = torch.eye(4)[None]                                         # :166-167
    pre_grid_sample_affine = pre_grid_sample_affine.to(dev).double()

    grid_affine, nii_out = grid_and_nii_affine(
        volume_nii_affine, fov_vox_in, target_fov_mm, target_fov_vox, pre_grid_sample_affine)

    if "int" in str(in_dtype):                                                              # :175-179
        volume = volume.to(dtype=dtype)
        grid_affine = grid_affine.to(dtype=dtype)
    else:
        grid_affine = grid_affine.to(volume)

    grid = F.affine_grid(grid_affine[:, :3, :].view(B, 3, 4), out_shape, align_corners=False)  # :182-184
    if is_label:                                                                            # :186-192
        sampled = _sample(volume, grid, "nearest")
    else:                                                                                   # :194-203
        lowest = volume.min()
        sampled = _sample(volume - lowest, grid, "bilinear") + lowest
    sampled = sampled.view(out_shape).to(dtype=in_dtype)                                    # :205
    return sampled, grid_affine, nii_out


# ----------------------------------------------------------------------------
# a2-a5  view-parameter tail of AffineTransformModule   models/learnable_transform.py
# ----------------------------------------------------------------------------
def offset_vox_range(offset_clip_value: float, spat: int) -> int:
    """learnable_transform.py:112-115 (align_corners False branch of :178-186)."""
    def vox(g):
        return ((g + 1.0) * spat - 1.0) / 2.0
    return int(round(vox(offset_clip_value) - vox(-offset_clip_value)))


def offset_positions(spat: int, vox_range: int) -> torch.Tensor:
    """``arra`` learnable_transform.py:116."""
    return torch.arange(0, vox_range) + (spat - vox_range) // 2


def _translation(offs: torch.Tensor) -> torch.Tensor:
    t = torch.eye(4, dtype=offs.dtype, device=offs.device).repeat(offs.shape[0], 1, 1)
    t[:, :3, 3] = offs
    return t


def _zoom(z: torch.Tensor) -> torch.Tensor:
    ones = torch.ones_like(z)
    return torch.diag_embed(torch.cat([z, z, z, ones], dim=-1))


def view_theta(mlp_out, init_ap, init_t_offsets, init_zp, offset_clip_value, zoom_clip_value, spat, optim_method="R6-vector"):
    """learnable_transform.py:144-161 (init affines), :188-230 (batch affines),
    :262-272 (composition).  ``mlp_out[B, A+3R+1]`` is what LocalizationNet returns (A = 6 for the default R6-vector
    parameterisation, 3 for angle-axis / normal-vector).  ``use_affine_theta=True``, ``align_corners=False``."""
    B = mlp_out.shape[0]
    R = offset_vox_range(offset_clip_value, spat)
    A = AP_SPACE[optim_method]
    assert mlp_out.shape[1] == A + 3 * R + 1
    # init affines (:144-161)
    a0 = rotation_of(optim_method, init_ap.view(1, A)).to(torch.float32).repeat(B, 1, 1)
    t0 = _translation(init_t_offsets.view(1, 3).to(torch.float32)).repeat(B, 1, 1)
    z0 = _zoom(init_zp.view(1, 1).to(torch.float32)).repeat(B, 1, 1)
    # batch affines (:193-230); inits are added to the raw parameters as well (:198-199)
    ap = mlp_out[:, :A] + init_ap.view(1, A)
    tp = mlp_out[:, A:-1].view(B, 3, R)
    zp = mlp_out[:, -1:] + init_zp.view(1, 1)
    if optim_method == "normal-vector":
        ap = ap / ap.norm(dim=1).view(-1, 1)                                                # :204-205
    a_b = rotation_of(optim_method, ap)
    pos = (F.softmax(tp, dim=2) * offset_positions(spat, R).to(tp).view(1, 1, R)).sum(-1)   # :165-168
    offs = (2.0 * pos + 1.0) / spat - 1.0                                                   # :174
    if offset_clip_value == 0.0:
        offs = 0.0 * offs                                                                   # :211-212
    t_b = _translation(offs)
    z_b = _zoom(zoom_clip_value * -(zp.tanh()) + 1.0)                                       # :220
    theta_a, theta_t, theta_z = a0 @ a_b, t0 @ t_b, z0 @ z_b                                # :268-270
    return theta_t @ theta_a @ theta_z                                                      # :272


def atm_tail_forward(x_soft_label, x_label, x_image, nifti_affine, grid_affine_pre_mlp, theta,
                     slice_fov_mm, slice_fov_vox):
    """learnable_transform.py:284-306,333: three slicings with ``Gpre @ theta``."""
    pre = grid_affine_pre_mlp.to(theta) @ theta                                             # :284,:289
    y_soft, grid_affine, nii = nifti_grid_sample(
        x_soft_label, nifti_affine, target_fov_mm=slice_fov_mm, target_fov_vox=slice_fov_vox,
        is_label=False, pre_grid_sample_affine=pre)
    y_label = y_image = None
    with torch.no_grad():
        if x_label is not None and x_label.numel() > 0:
            y_label, _, _ = nifti_grid_sample(
                x_label, nifti_affine, target_fov_mm=slice_fov_mm, target_fov_vox=slice_fov_vox,
                is_label=True, pre_grid_sample_affine=pre)
        if x_image is not None and x_image.numel() > 0:
            y_image, _, _ = nifti_grid_sample(
                x_image, nifti_affine, target_fov_mm=slice_fov_mm, target_fov_vox=slice_fov_vox,
                is_label=False, pre_grid_sample_affine=pre)
    return y_soft, y_label, y_image, grid_affine, nii


# ----------------------------------------------------------------------------
# a13  in-plane re-alignment of a slice                 models/learnable_transform.py:337-366
#      (+ utils/torch_sparse_tensor_utils.py:34-56,79-85, functional/clinical_cardiac_views.py:66-100)
# ----------------------------------------------------------------------------
def min_principle_align_affine(label_slice: torch.Tensor) -> torch.Tensor:
    """``label_slice[H,W,1]`` integer (argmax of one sample's slice) -> 4x4 torch-grid affine that turns the slice so that
    the axis of least inertia of its foreground lies along the first in-plane axis.  Host arithmetic like the reference:
    inertia tensor of the foreground index cloud (torch_sparse_tensor_utils.py:34-56), eigenvectors by ``torch.linalg.eig``
    sorted by eigenvalue (:79-85), orthonormal frame (two, main, normal) through the centroid with its depth set to 0.5
    (learnable_transform.py:346-353, clinical_cardiac_views.py:75-100), pixel -> grid convention (:66-71)."""
    idx = label_slice.to_sparse()._indices()                                  # [3, n] foreground positions
    center = idx.float().mean(1)
    d = idx - center.view(3, 1)
    r2 = torch.linalg.vector_norm(d, 2, dim=0) ** 2
    inertia = torch.zeros(3, 3)
    for i in range(3):
        for j in range(3):
            inertia[i, j] = (r2 * float(i == j) - d[i] * d[j]).sum()
    center[-1] = 0.5
    eig = torch.linalg.eig(inertia)
    main = eig.eigenvectors.real.T[eig.eigenvalues.real.argsort()][0].clone()  # axis of least inertia
    two = torch.linalg.cross(main, torch.tensor([0.0, 0.0, 1.0]))
    main = main / torch.linalg.norm(main, 2)
    two = two / torch.linalg.norm(two, 2)
    normal = torch.linalg.cross(main, two)
    normal = normal / torch.linalg.norm(normal, 2)
    two = torch.linalg.cross(normal, main)
    pix = torch.eye(4)
    pix[:3, :3] = torch.stack([two, main, normal], dim=0)
    pix[:3, -1] = center
    out = pix.clone()
    out[:3, :3] = out[:3, :3].flip(0, 1).T
    out[:3, -1] = (2.0 * out[:3, -1] / torch.as_tensor(label_slice.shape) - 1.0).flip(0)
    return out


def rotate_slice_to_min_principle(x_input, nii_affine, is_label=False, align_affine_override=None):
    """learnable_transform.py:337-366: per-sample alignment affine from ``argmax`` over the channels (unless overridden),
    then a same-FOV resample of the one-voxel-thin slice volume with it."""
    assert x_input.shape[-1] == 1
    if align_affine_override is None:
        aff = torch.zeros(x_input.shape[0], 4, 4).to(x_input.device)
        with torch.no_grad():
            for b, lbl in enumerate(x_input):
                aff[b] = min_principle_align_affine(lbl.argmax(0))
    else:
        aff = align_affine_override
    return nifti_grid_sample(x_input, nii_affine, pre_grid_sample_affine=aff, is_label=is_label)


# ----------------------------------------------------------------------------
# a6  clinical composition + augmentation              running/run_dl.py:208-234
# ----------------------------------------------------------------------------
def input_affine_for_view(base_affine, view_affine):
    """run_dl.py:227-234: ``Gpre = base_affine^-1 @ view_affine``."""
    return base_affine.inverse() @ view_affine.to(base_affine)


def random_aug_affine(gen: torch.Generator, rotation_strength=0.2, zoom_strength=0.2, offset_strength=0.0):
    """utils/transform_utils.py:6-23 with an explicit generator (host RNG)."""
    rz = torch.rand(1, generator=gen) * zoom_strength - zoom_strength / 2 + 1.0
    n = torch.cat([rotation_strength * torch.randn(2, generator=gen), torch.ones(1)])
    n = n / n.norm(2)
    one = torch.cat([torch.ones(1), rotation_strength * torch.randn(2, generator=gen)])
    two = torch.linalg.cross(n, one)
    two = two / two.norm(2)
    one = torch.linalg.cross(two, n)
    r = torch.eye(4)
    r[:3, :3] = torch.stack([one, two, n])
    z = torch.diag(torch.cat([rz, rz, rz, torch.ones(1)]))
    t = torch.eye(4)
    t[:3, 3] = offset_strength * torch.randn(3, generator=gen)
    return z @ r @ t


def get_random_affine(rotation_strength=0.2, zoom_strength=0.2, offset_strength=0.0):
    """utils/transform_utils.py:6-23 on the GLOBAL torch RNG with the reference's exact draw sequence (``rand(1)``,
    ``randn(2)``, ``randn(2)``, ``randn(3)`` - the last one also when ``offset_strength`` is 0): seeded identically, this
    reproduces the reference's augmentation affines bit for bit."""
    rz = torch.rand(1) * zoom_strength - zoom_strength / 2 + 1.0
    n = torch.tensor((rotation_strength * torch.randn(2)).tolist() + [1.0])
    n = n / n.norm(2)
    one = torch.tensor([1.0] + (rotation_strength * torch.randn(2)).tolist())
    two = torch.linalg.cross(n, one)
    two = two / two.norm(2)
    one = torch.linalg.cross(two, n)
    r = torch.eye(4)
    r[:3, :3] = torch.stack([one, two, n])
    z = torch.diag(torch.tensor([rz.item(), rz.item(), rz.item(), 1.0]))
    t = torch.eye(4)
    t[:3, 3] = offset_strength * torch.randn(3)
    return z @ r @ t


def apply_affine_augmentation(affine_list, zoom_strength=0.1, offset_strength=0.1, rotation_strength=0.1):
    """run_dl.py:208-223: one random affine per batch element, right-multiplied onto every affine of the list."""
    B = affine_list[0].shape[0]
    b_affine = torch.stack([get_random_affine(rotation_strength=rotation_strength, zoom_strength=zoom_strength,
                                              offset_strength=offset_strength) for _ in range(B)])
    return [a @ b_affine.to(a) for a in affine_list]


# ----------------------------------------------------------------------------
# a12  slice -> 3-D embedding                           models/hybrid_unet.py:71-94
# ----------------------------------------------------------------------------
def skip_connector(x, b_grid_affines, n_views):
    """Place each view's 2-D map on the W = S//2 plane of a zero S^3 volume and
    resample it with the inverse of the column-normalised slicing affine."""
    B, C, S, _ = x.shape
    c = C // n_views
    shape = torch.Size([B, c, S, S, S])
    mid = torch.zeros(B, C, S, S, S).to(x)                                                  # :75
    mid[..., S // 2] = x                                                                    # :76
    outs = []
    for vx, ga in zip(torch.chunk(mid, n_views, dim=1), b_grid_affines):                    # :77,:80
        ga = _scale_columns(ga, 1.0 / column_norms(ga))                                     # :83
        grid = F.affine_grid(ga.to(torch.float32).inverse()[:, :3, :].view(B, 3, 4), shape,
                             align_corners=False)                                           # :85-87
        outs.append(checkpoint(F.grid_sample, vx, grid.to(vx), "bilinear", "zeros", False,
                               use_reentrant=True))                                         # :88-90
    return torch.cat(outs, dim=1)                                                           # :93


# ----------------------------------------------------------------------------
# sparse closed form of the embedding -- used to cross-check the
# dense restatement above and as a fast checker at large S.
# ----------------------------------------------------------------------------
def skip_connector_sparse(x, b_grid_affines, n_views):
    B, C, S, _ = x.shape
    c = C // n_views
    dev = x.device
    lin = (torch.linspace(-1, 1, S, device=dev) * (S - 1) / S).double()
    zd, yh, xw = torch.meshgrid(lin, lin, lin, indexing="ij")
    base = torch.stack([xw, yh, zd, torch.ones_like(xw)], dim=-1).reshape(-1, 4)           # [S^3,4]
    outs = []
    for v, ga in enumerate(b_grid_affines):
        ga = ga.double()
        a = _scale_columns(ga, 1.0 / column_norms(ga))
        ainv = a.inverse()[:, :3, :]                                                        # [B,3,4]
        p = torch.einsum("nk,brk->bnr", base, ainv)                                         # [B,S^3,3]
        ix = ((p[..., 0] + 1) * S - 1) / 2
        iy = ((p[..., 1] + 1) * S - 1) / 2
        iz = ((p[..., 2] + 1) * S - 1) / 2
        wx = (1 - (ix - S // 2).abs()).clamp(min=0)
        y0 = iy.floor(); z0 = iz.floor()
        fy = iy - y0; fz = iz - z0
        xv = x[:, v * c:(v + 1) * c].double()                                               # [B,c,S(row=D),S(col=H)]
        acc = torch.zeros(B, c, S ** 3, dtype=torch.float64, device=dev)
        for dz, wz in ((0, 1 - fz), (1, fz)):
            for dy, wy in ((0, 1 - fy), (1, fy)):
                r = (z0 + dz).long(); q = (y0 + dy).long()
                ok = (r >= 0) & (r < S) & (q >= 0) & (q < S)
                idx = (r.clamp(0, S - 1) * S + q.clamp(0, S - 1))                           # [B,S^3]
                val = torch.gather(xv.reshape(B, c, S * S), 2, idx[:, None].expand(B, c, -1))
                acc += val * (wz * wy * ok)[:, None]
        outs.append((acc * wx[:, None]).view(B, c, S, S, S))
    return torch.cat(outs, dim=1).to(x.dtype)


# ----------------------------------------------------------------------------
# a11  per-batch caller                                   running/run_dl.py:238-329
# ----------------------------------------------------------------------------
def reconstruction_model_input(b_label, b_image, nifti_affine, base_affine, view_affines, mlp_outs, init, hires_fov_mm,
                               hires_fov_vox, slice_fov_mm, slice_fov_vox, num_classes, offset_clip, zoom_clip, spat,
                               aug_affine=None, augment_input=False, augment_recon=False, sample_augment_strength=1.0):
    """Restatement of ``get_reconstruction_model_input`` for ``label_slice_type='from-gt'``, ``opt-all``, all views active:
    hires resample of label (nearest) and image (bilinear) with ``base_affine`` (:251-259; the image call receives the
    UPDATED nifti affine, as in the reference), one-hot (:261-264), ``Gpre = base^-1 @ view`` (:227-234), the input
    augmentation (:273-278: ONE draw per batch element shared by all views, global RNG) or a given ``aug_affine``, per-view
    ATM tail (:283-312) with the up-sampling of low-resolution slices (:193-197) and the per-view reconstruction
    augmentation of the returned affine (:303-309), ``cat`` (:325).  ``mlp_outs[v]`` stands in for the LocalizationNet
    output of view v.  Pinned against the reference's own function: golden ``model_input_s32*.npz``."""
    with torch.no_grad():
        lab, _, nii = nifti_grid_sample(b_label.unsqueeze(1), nifti_affine, target_fov_mm=hires_fov_mm, target_fov_vox=hires_fov_vox,
                                        is_label=True, pre_grid_sample_affine=base_affine)
        img, _, _ = nifti_grid_sample(b_image.unsqueeze(1), nii, target_fov_mm=hires_fov_mm, target_fov_vox=hires_fov_vox,
                                      is_label=False, pre_grid_sample_affine=base_affine)
        lab = lab.squeeze(1)
    label = F.one_hot(lab, num_classes).permute(0, 4, 1, 2, 3)
    soft = label.float()
    gpres = [input_affine_for_view(base_affine, va).to(nii) for va in view_affines]
    if aug_affine is not None:
        gpres = [g @ aug_affine.to(g) for g in gpres]
    if augment_input:
        gpres = apply_affine_augmentation(gpres, rotation_strength=0.1 * sample_augment_strength,
                                          zoom_strength=0.2 * sample_augment_strength, offset_strength=0.0)
    slices, affines = [], []
    for v, gpre in enumerate(gpres):
        theta = view_theta(mlp_outs[v], init[v:v + 1, :6], init[v, 6:9], init[v:v + 1, 9:], offset_clip, zoom_clip, spat)
        ys, yl, yi, ga, _ = atm_tail_forward(soft, label, img, nii, gpre.float(), theta, slice_fov_mm, slice_fov_vox)
        if [int(v_) for v_ in slice_fov_vox.tolist()] != [int(v_) for v_ in hires_fov_vox.tolist()]:       # :193-197
            tgt = [int(v_) for v_ in hires_fov_vox.tolist()[:2]] + [1]
            ys = F.interpolate(ys, size=tgt, mode="trilinear", align_corners=False)
        if augment_recon:                                                                                # :303-309
            ga = apply_affine_augmentation([ga], rotation_strength=0.1 * sample_augment_strength,
                                           zoom_strength=0.2 * sample_augment_strength, offset_strength=0.0)[0].to(nii)
        slices.append(ys)
        affines.append(ga)
    return torch.cat(slices, dim=1).squeeze(-1), label, affines
