"""Drop-in for ``SkipConnector`` of the reference's ``models/hybrid_unet.py:65-94``."""
from __future__ import annotations

import torch

from .. import functional as AF


class SkipConnector(torch.nn.Module):
    """Re-embed each view's 2-D feature map into the 3-D reconstruction FOV (one fused CUDA kernel per call;
    neither the zero-padded ``x_mid`` volume nor the sampling grid is materialised)."""

    def __init__(self, n_views):
        self.dtype = torch.float32
        self.n_views = n_views
        super().__init__()

    def forward(self, x, b_grid_affines):
        B, C, SPAT, _ = x.shape
        assert C % self.n_views == 0 and len(b_grid_affines) == self.n_views
        affines = torch.stack([ga.to(x.device, self.dtype) for ga in b_grid_affines], dim=0)   # [V,B,4,4]
        return AF.embed_slices(x, affines, self.n_views)

    def embed_all(self, skips, b_grid_affines):
        """``[self(s, b_grid_affines) for s in skips]`` - what ``HybridUnet.forward`` does with the encoder's skip list
        (reference ``:40-43``) - as ONE launch forward and ONE backward over all stages."""
        x0 = skips[0]
        affines = torch.stack([ga.to(x0.device, self.dtype) for ga in b_grid_affines], dim=0)
        return AF.embed_slices_multi(list(skips), affines, self.n_views)
