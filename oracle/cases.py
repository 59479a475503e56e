"""ORACLE / test infrastructure: deterministic input builders for parity cases.

Shared by ``oracle/make_golden.py`` (which feeds them to the *unmodified
reference* in the build container and stores the outputs under
``tests/golden/``) and by ``tests/`` (which rebuild the same inputs anywhere,
including the GPU box, and feed them to the CUDA path and to the oracle port).
Only small things (parameters, 4x4 affines) are stored in fixtures; volumes are
regenerated from closed forms / seeded numpy PCG64 streams.
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_synthetic():
    # product-side input generator (numpy only); loaded by path so that the oracle
    # never imports the CUDA-backed package __init__.
    spec = importlib.util.spec_from_file_location(
        "_afb_synthetic", os.path.join(_ROOT, "acquisition_focus_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synthetic = _load_synthetic()

# the case builders live with the product's synthetic-input module (bench variants use them too); re-exported here
pattern, randn, randint = synthetic.pattern, synthetic.randn, synthetic.randint
phantom_batch, one_hot_volumes = synthetic.phantom_batch, synthetic.one_hot_volumes
atm_case, embed_case = synthetic.atm_case, synthetic.embed_case


def rotated_nii_affine(B: int, seed: int, spacing=(1.5, 1.2, 2.0)) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    out = np.zeros((B, 4, 4))
    for b in range(B):
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        out[b, :3, :3] = q @ np.diag(spacing)
        out[b, :3, 3] = rng.standard_normal(3) * 10
        out[b, 3, 3] = 1.0
    return torch.from_numpy(out)


def random_pre_affine(B: int, seed: int, strength: float = 0.2) -> torch.Tensor:
    """Non-orthogonal 4x4 with unequal column norms (exercises the flip quirk of
    utils/nifti_utils.py:57)."""
    rng = np.random.default_rng(seed)
    P = np.tile(np.eye(4), (B, 1, 1)) + strength * rng.standard_normal((B, 4, 4))
    P[:, 3, :] = [0, 0, 0, 1]
    return torch.from_numpy(P.astype(np.float32))


def model_input_setup(S=32, slice_vox=None, aug=True):
    """Inputs of the a11 goldens (shared with tests/): config (the reference's config_dict.json names), batch, MLP-head
    outputs.  ``slice_vox`` < S exercises the low-resolution-slice up-sampling of run_dl.py:193-197."""
    B, V, C = 2, 3, 8
    names = ["p2CH", "p4CH", "4CH"]
    sv = S if slice_vox is None else slice_vox
    case = atm_case(sv, B, V, seed=91)            # params sized for the SLICE module: R = round(0.2 * prescan S)
    case_vol = atm_case(S, B, V, seed=91)
    views = synthetic.phantom_view_affines()
    gen = torch.Generator().manual_seed(5)
    base = torch.stack([synthetic.random_aug_affine(gen, 0.2, 0.1, 0.02) for _ in range(B)])
    cfg = dict(clinical_view_affine_type="from-gt", label_slice_type="from-gt", hires_fov_mm=[192.0] * 3, hires_fov_vox=[S] * 3,
               prescan_fov_mm=[192.0] * 3, prescan_fov_vox=[S] * 3, slice_fov_mm=[192.0, 192.0, 192.0 / sv], slice_fov_vox=[sv, sv, 1],
               use_affine_theta=True, do_augment_input_orientation=aug, do_augment_recon_orientation=aug, aug_phases=["train"],
               sample_augment_strength=1.0, view_optimization_mode="opt-all", base_views=names, offset_clip_value=0.2,
               zoom_clip_value=0.0, affine_theta_optim_method="R6-vector", rotate_slice_to_min_principle=False)
    batch = {"label": case_vol["lab"], "image": case_vol["image"][:, 0],
             "additional_data": {"nifti_affine": case_vol["nii"],
                                 "gt_view_affines": {**{n: views[n][None].repeat(B, 1, 1) for n in names}, "centroids": base}}}
    params = [p.clone() for p in case_vol["params"]]     # [B, 6 + 3R + 1] with R = round(0.2 * S) (prescan size)
    return cfg, batch, params, (B, V, C, names)
