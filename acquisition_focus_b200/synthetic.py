"""Synthetic inputs of the reference's shapes (there is no dataset in this image).

* :func:`heart_phantom` - a nested-ellipsoid cardiac label map (BG + 7
  structures = 8 classes like MMWHS) with an oblique LV long axis, so that the
  clinical view affines the reference derives from it
  (``functional/clinical_cardiac_views.py:223-364``) are genuinely oblique.
* :func:`phantom_image` - piecewise-constant intensities + seeded noise.
* :func:`phantom_view_affines` - the torch-grid view affines (p2CH, p4CH, 2CH,
  4CH, SA-k, axial, ...) that ``get_clinical_cardiac_view_affines`` returns for
  the 128^3 phantom.  They are *inputs* (constants) of the hot path; they were
  computed once with the unmodified reference (``oracle/make_golden.py``) and
  are stored in ``data/phantom_view_affines.json``.
* :func:`random_aug_affine` - host-side RNG augmentation affine of the same
  family as ``utils/transform_utils.py:6-23`` (rotation about a perturbed
  normal, isotropic zoom, optional offset).

Everything here is numpy/torch-CPU host code producing *inputs*; no sampling
arithmetic lives here.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

CLASS_DICT = {"MYO": 1, "LA": 2, "LV": 3, "RA": 4, "RV": 5, "AO": 6, "PA": 7}
NUM_CLASSES = 8

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _ellipsoid(zz, yy, xx, centre, axes_rot, radii):
    d = np.stack([zz - centre[0], yy - centre[1], xx - centre[2]], axis=-1) @ axes_rot
    return ((d / np.asarray(radii)) ** 2).sum(-1) <= 1.0


def heart_phantom(size: int = 128) -> np.ndarray:
    """[S,S,S] int64 label map, indices (D,H,W); deterministic closed form."""
    S = size
    lin = (np.arange(S, dtype=np.float64) + 0.5) / S * 2.0 - 1.0
    zz, yy, xx = np.meshgrid(lin, lin, lin, indexing="ij")
    # oblique long axis
    a = np.array([0.55, 0.45, 0.70]); a /= np.linalg.norm(a)
    b = np.cross(a, [0.0, 0.0, 1.0]); b /= np.linalg.norm(b)
    c = np.cross(a, b)
    rot = np.stack([a, b, c], axis=1)
    lab = np.zeros((S, S, S), dtype=np.int64)
    c_lv = np.array([0.05, -0.05, 0.10])
    lab[_ellipsoid(zz, yy, xx, c_lv, rot, (0.46, 0.30, 0.30))] = CLASS_DICT["MYO"]
    lab[_ellipsoid(zz, yy, xx, c_lv, rot, (0.38, 0.21, 0.21))] = CLASS_DICT["LV"]
    c_rv = c_lv + 0.40 * b + 0.02 * c
    m_rv = _ellipsoid(zz, yy, xx, c_rv, rot, (0.40, 0.17, 0.27)) & (lab == 0)
    lab[m_rv] = CLASS_DICT["RV"]
    c_la = c_lv - 0.58 * a + 0.03 * c
    m_la = _ellipsoid(zz, yy, xx, c_la, rot, (0.17, 0.22, 0.22)) & (lab == 0)
    lab[m_la] = CLASS_DICT["LA"]
    c_ra = c_rv - 0.55 * a
    m_ra = _ellipsoid(zz, yy, xx, c_ra, rot, (0.17, 0.18, 0.22)) & (lab == 0)
    lab[m_ra] = CLASS_DICT["RA"]
    c_ao = c_la - 0.25 * a - 0.20 * c
    m_ao = _ellipsoid(zz, yy, xx, c_ao, rot, (0.22, 0.07, 0.07)) & (lab == 0)
    lab[m_ao] = CLASS_DICT["AO"]
    c_pa = c_ra - 0.22 * a + 0.22 * c
    m_pa = _ellipsoid(zz, yy, xx, c_pa, rot, (0.20, 0.07, 0.07)) & (lab == 0)
    lab[m_pa] = CLASS_DICT["PA"]
    return lab


def phantom_image(label: np.ndarray, seed: int = 0, noise: float = 0.15) -> np.ndarray:
    """float32 intensity image for a label map: class means + seeded Gaussian noise."""
    means = np.array([0.05, 0.55, 0.85, 0.95, 0.80, 0.90, 0.75, 0.70], dtype=np.float32)
    rng = np.random.default_rng(seed)
    img = means[label] + noise * rng.standard_normal(label.shape, dtype=np.float32)
    return img.astype(np.float32)


def phantom_view_affines(size: int = 128) -> dict:
    """View name -> 4x4 float32 torch-grid affine (normalised coordinates, hence
    valid for any cubic resampling of the phantom)."""
    with open(os.path.join(_DATA, "phantom_view_affines.json")) as fh:
        raw = json.load(fh)
    return {k: torch.tensor(v, dtype=torch.float32) for k, v in raw["views"].items()}


def random_aug_affine(gen: torch.Generator, rotation_strength: float = 0.2, zoom_strength: float = 0.2,
                      offset_strength: float = 0.0) -> torch.Tensor:
    """Random zoom @ rotation @ translation 4x4 (input augmentation, host RNG)."""
    zoom = float(torch.rand(1, generator=gen)) * zoom_strength - zoom_strength / 2 + 1.0
    normal = torch.cat([rotation_strength * torch.randn(2, generator=gen), torch.ones(1)])
    normal = normal / normal.norm()
    u = torch.cat([torch.ones(1), rotation_strength * torch.randn(2, generator=gen)])
    v = torch.linalg.cross(normal, u)
    v = v / v.norm()
    u = torch.linalg.cross(v, normal)
    out = torch.eye(4)
    out[:3, :3] = torch.stack([u, v, normal]) * zoom
    if offset_strength != 0.0:
        shift = torch.eye(4)
        shift[:3, 3] = offset_strength * torch.randn(3, generator=gen)
        out = out @ shift
    return out


def default_nifti_affine(batch: int, spacing_mm: float = 1.5) -> torch.Tensor:
    a = torch.eye(4, dtype=torch.float64) * spacing_mm
    a[3, 3] = 1.0
    return a[None].repeat(batch, 1, 1)

# ------------------------------------------------------------------------------------------------
# deterministic case builders (inputs of tests, smoke and the bench variants): closed forms / seeded PCG64 streams
# ------------------------------------------------------------------------------------------------
def pattern(shape, k: float = 1.0) -> torch.Tensor:
    """Deterministic pseudo-random upstream gradient in [-1,1] (closed form)."""
    n = int(np.prod(shape))
    v = np.cos(np.arange(n, dtype=np.float64) * 0.6180339887498949 * k + 0.3)
    return torch.from_numpy(v.astype(np.float32).reshape(shape))


def randn(shape, seed: int) -> torch.Tensor:
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32))


def randint(lo, hi, shape, seed: int) -> torch.Tensor:
    return torch.from_numpy(np.random.default_rng(seed).integers(lo, hi, size=shape, dtype=np.int64))


def phantom_batch(S: int, B: int, num_classes: int = 8):
    """B label maps [B,S,S,S] int64: the phantom and axis-permuted/flipped variants."""
    base = heart_phantom(S)
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
    views = phantom_view_affines()
    R = int(round(offset_clip * S))
    lab = phantom_batch(S, B, num_classes)
    label_oh, soft = one_hot_volumes(lab, num_classes)
    img = torch.stack([torch.from_numpy(phantom_image(lab[b].numpy(), seed=seed + b)) for b in range(B)])[:, None]
    nii = default_nifti_affine(B, 192.0 / S)
    gpre, params = [], []
    for v in range(V):
        aug = torch.stack([random_aug_affine(gen, 0.1, 0.2, 0.0) for _ in range(B)])
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
    views = phantom_view_affines()
    names = ["p2CH", "p4CH", "SA-1", "4CH", "2CH", "axial"]
    gas = []
    for v in range(V):
        aug = torch.stack([random_aug_affine(gen, 0.3, 0.3, 0.05) for _ in range(B)])
        gas.append(views[names[v % len(names)]][None].repeat(B, 1, 1) @ aug)
    return dict(S=S, c=c, V=V, B=B, x=x, affines=gas)
