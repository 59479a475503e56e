"""Drop-ins for the rotation parameterisations of the reference's ``utils/transform_utils.py``.

``compute_rotation_matrix_from_ortho6d`` (``:27-58``, the default ``R6-vector`` method) is a CUDA kernel pair and is also
fused into the view prologue.  ``angle_axis_to_rotation_matrix`` (``:106-178``) and ``normal_to_rotation_matrix``
(``:62-103``), the two non-default ``optim_method``s, are three-parameter closed forms evaluated with a
handful of batched torch ops on the device; their output feeds the same CUDA sampler through the differentiable
``pre_grid_sample_affine`` input.
"""
from __future__ import annotations

import torch

from .. import functional as AF


def get_random_affine(rotation_strength: float = 0.2, zoom_strength: float = 0.2, offset_strength: float = 0.0,
                      generator=None) -> torch.Tensor:
    """Random ``zoom @ rotation @ translation`` 4x4 of the input / reconstruction augmentation (reference ``:6-23``).

    Host RNG, as in the reference.  The draws are issued in the reference's order and shapes - ``rand(1)`` (zoom),
    ``randn(2)`` (plane normal), ``randn(2)`` (in-plane axis), ``randn(3)`` (offset, drawn even when its strength is 0) - so
    with ``generator=None`` (the global torch RNG, which is what the reference uses) a run seeded with ``torch.manual_seed``
    reproduces the reference's augmentation stream bit for bit (``tests/test_oracle_vs_reference.py``)."""
    g = generator
    zoom = torch.rand(1, generator=g) * zoom_strength - zoom_strength / 2 + 1.0
    normal = torch.cat([rotation_strength * torch.randn(2, generator=g), torch.ones(1)])
    normal = normal / normal.norm(2)
    u = torch.cat([torch.ones(1), rotation_strength * torch.randn(2, generator=g)])
    v = torch.linalg.cross(normal, u)
    v = v / v.norm(2)
    u = torch.linalg.cross(v, normal)
    rot = torch.eye(4)
    rot[:3, :3] = torch.stack([u, v, normal])
    scale = torch.diag(torch.cat([zoom, zoom, zoom, torch.ones(1)]))
    shift = torch.eye(4)
    shift[:3, 3] = offset_strength * torch.randn(3, generator=g)
    return scale @ rot @ shift


def compute_rotation_matrix_from_ortho6d(ortho: torch.Tensor) -> torch.Tensor:
    """R6 -> homogeneous 4x4 rotation (Gram-Schmidt, columns x,y,z); ``[B,6] -> [B,4,4]``, differentiable."""
    return AF.r6_to_matrix(ortho)


def _homogeneous(rot3: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(rot3.shape[0], 4, 4, dtype=rot3.dtype, device=rot3.device)
    out[:, :3, :3] = rot3
    out[:, 3, 3] = 1.0
    return out


def _skew(v: torch.Tensor) -> torch.Tensor:
    """``[N,3] -> [N,3,3]`` cross-product matrix."""
    x, y, z = v.unbind(dim=1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=1).view(-1, 3, 3)


def angle_axis_to_rotation_matrix(angle_axis: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Axis-angle vector ``[N,3] -> [N,4,4]``.  ``R = c I + (1-c) w w^T + s [w]x`` with the reference's regularised axis
    ``w = r / (sqrt(|r|^2 + eps) + eps)`` where ``|r|^2 > eps``, and the first-order ``I + [r]x`` elsewhere."""
    r = angle_axis
    theta2 = (r.float() * r.float()).sum(dim=1, keepdim=True)
    theta = torch.sqrt(theta2 + eps)
    w = r / (theta + eps)
    c, s = torch.cos(theta)[..., None], torch.sin(theta)[..., None]
    eye = torch.eye(3, dtype=r.dtype, device=r.device)[None]
    full = c * eye + (1.0 - c) * (w[:, :, None] * w[:, None, :]) + s * _skew(w)
    first_order = eye + _skew(r)
    big = (theta2 > eps).view(-1, 1, 1).to(full.dtype)
    # blended with 0/1 weights (not torch.where) so that autograd sees both branches like the reference's mask arithmetic
    return _homogeneous(big * full + (1.0 - big) * first_order)


def normal_to_rotation_matrix(normals: torch.Tensor) -> torch.Tensor:
    """Plane normal ``[N,3]`` (columns nz, ny, nx) ``-> [N,4,4]``: rows = in-plane axis, its complement, the normal."""
    nz, ny, nx = normals.unbind(dim=1)
    d = torch.sqrt(nx * nx + ny * ny)
    o = torch.zeros_like(nx)
    rot = torch.stack([ny / d, -nx / d, o, nx * nz / d, ny * nz / d, -d, nx, ny, nz], dim=1).view(-1, 3, 3)
    return _homogeneous(rot)
