"""Drop-ins for the rotation parameterisations of the reference's ``utils/transform_utils.py``.

``compute_rotation_matrix_from_ortho6d`` (``:27-58``, the default ``R6-vector`` method) is a CUDA kernel pair and is also
fused into the view prologue.  ``angle_axis_to_rotation_matrix`` (``:106-178``) and ``normal_to_rotation_matrix``
(``:62-103``), the two non-default ``optim_method``s, are one kernel each way too (``afb_rot3_fwd/bwd``); their
output feeds the same CUDA sampler through the differentiable ``pre_grid_sample_affine`` input.  ``get_random_affine``
(``:6-23``) stays host code like the reference's (host RNG), with the reference's draw sequence.
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


def angle_axis_to_rotation_matrix(angle_axis: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Axis-angle vector ``[N,3] -> [N,4,4]`` (reference ``:106-178``): ``R = c I + (1-c) w w^T + s [w]x`` with the reference's
    regularised axis ``w = r / (sqrt(|r|^2 + eps) + eps)`` where ``|r|^2 > eps`` and the first-order ``I + [r]x`` elsewhere.
    One CUDA kernel forward, one backward (``afb_rot3_fwd/bwd``)."""
    assert eps == 1e-6, "the kernel carries the reference's eps"
    return AF.angle_axis_to_matrix(angle_axis)


def normal_to_rotation_matrix(normals: torch.Tensor) -> torch.Tensor:
    """Plane normal ``[N,3]`` (columns nz, ny, nx) ``-> [N,4,4]`` (reference ``:62-103``): rows = in-plane axis, its complement,
    the normal.  One CUDA kernel forward, one backward."""
    return AF.normal_to_matrix(normals)
