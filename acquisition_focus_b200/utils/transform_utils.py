"""Drop-in for ``compute_rotation_matrix_from_ortho6d`` (reference ``utils/transform_utils.py:27-58``)."""
from __future__ import annotations

import torch

from .. import functional as AF


def compute_rotation_matrix_from_ortho6d(ortho: torch.Tensor) -> torch.Tensor:
    """R6 -> homogeneous 4x4 rotation (Gram-Schmidt, columns x,y,z); ``[B,6] -> [B,4,4]``, differentiable."""
    return AF.r6_to_matrix(ortho)
