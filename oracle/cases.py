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


def pattern(shape, k: float = 1.0) -> torch.Tensor:
    """Deterministic pseudo-random upstream gradient in [-1,1] (closed form)."""
    n = int(np.prod(shape))
    v = np.cos(np.arange(n, dtype=np.float64) * 0.6180339887498949 * k + 0.3)
    return torch.from_numpy(v.astype(np.float32).reshape(shape))


def randn(shape, seed: int) -> torch.Tensor:
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32))


def randint(lo, hi, shape, seed: int) -> torch.Tensor:
    return torch.from_numpy(np.random.default_rng(seed).integers(lo, hi, size=shape, dtype=np.int64))


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


def phantom_batch(S: int, B: int, num_classes: int = 8):
    """B label maps [B,S,S,S] int64: the phantom and axis-permuted/flipped variants."""
    base = synthetic.heart_phantom(S)
    variants = [base, np.ascontiguousarray(base.transpose(1, 0, 2)[::-1]),
                np.ascontiguousarray(base[:, ::-1, :]), np.ascontiguousarray(base.transpose(2, 1, 0))]
    lab = np.stack([variants[b % len(variants)] for b in range(B)])
    return torch.from_numpy(lab)


def one_hot_volumes(lab: torch.Tensor, num_classes: int = 8):
    """As running/run_dl.py:261-264: ``rearrange(one_hot(label), 'B D H W OH -> B OH D H W')``
    (a channels-last *view*) and its ``.float()``."""
    oh = torch.nn.functional.one_hot(lab, num_classes).permute(0, 4, 1, 2, 3)
    return oh, oh.float()


def atm_case(S: int, B: int, V: int, seed: int, zoom_clip: float = 0.0, offset_clip: float = 0.2,
             num_classes: int = 8):
    """cfg-2 style inputs: per view a p2CH pre-affine with per-sample augmentation,
    MLP-head outputs (R6 | offset logits | zoom logit)."""
    gen = torch.Generator().manual_seed(seed)
    views = synthetic.phantom_view_affines()
    R = int(round(offset_clip * S))
    lab = phantom_batch(S, B, num_classes)
    label_oh, soft = one_hot_volumes(lab, num_classes)
    img = torch.stack([torch.from_numpy(synthetic.phantom_image(lab[b].numpy(), seed=seed + b)) for b in range(B)])[:, None]
    nii = synthetic.default_nifti_affine(B, 192.0 / S)
    gpre, params = [], []
    for v in range(V):
        aug = torch.stack([synthetic.random_aug_affine(gen, 0.1, 0.2, 0.0) for _ in range(B)])
        gpre.append(views["p2CH"][None].repeat(B, 1, 1) @ aug)
        r6 = torch.tensor([1.0, 0, 0, 0, 1.0, 0]) + 0.3 * torch.randn(B, 6, generator=gen)
        params.append(torch.cat([r6, torch.randn(B, 3 * R, generator=gen), torch.randn(B, 1, generator=gen)], dim=1))
    return dict(S=S, B=B, V=V, R=R, lab=lab, label=label_oh, soft=soft, image=img, nii=nii,
                gpre=gpre, params=params, zoom_clip=zoom_clip, offset_clip=offset_clip,
                slice_fov_mm=torch.tensor([192.0, 192.0, 192.0 / S]), slice_fov_vox=torch.tensor([S, S, 1]),
                volume_fov_mm=torch.tensor([192.0, 192.0, 192.0]), volume_fov_vox=torch.tensor([S, S, S]))


def embed_case(S: int, c: int, V: int, B: int, seed: int):
    gen = torch.Generator().manual_seed(seed)
    x = randn((B, V * c, S, S), seed)
    views = synthetic.phantom_view_affines()
    names = ["p2CH", "p4CH", "SA-1", "4CH", "2CH", "axial"]
    gas = []
    for v in range(V):
        aug = torch.stack([synthetic.random_aug_affine(gen, 0.3, 0.3, 0.05) for _ in range(B)])
        gas.append(views[names[v % len(names)]][None].repeat(B, 1, 1) @ aug)
    return dict(S=S, c=c, V=V, B=B, x=x, affines=gas)


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
