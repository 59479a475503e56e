"""Live check of the oracle port against the unmodified reference (build container only:
skipped wherever /root/reference is not mounted, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import af_oracle as O
from oracle import cases
from oracle.ref_import import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not mounted")


def test_nifti_grid_sample_random_affines():
    R = load_reference()
    B = 2
    vol = cases.randn((B, 2, 18, 22, 26), 101)
    nii = cases.rotated_nii_affine(B, 102, spacing=(0.9, 1.7, 1.3))
    P = cases.random_pre_affine(B, 103, 0.3)
    kw = dict(target_fov_mm=torch.tensor([25.0, 21.0, 2.0]), target_fov_vox=torch.tensor([14, 10, 1]))
    for is_label, v in ((False, vol), (True, cases.randint(0, 9, vol.shape, 104))):
        a = R.nifti_grid_sample(v, nii, is_label=is_label, pre_grid_sample_affine=P, **kw)
        b = O.nifti_grid_sample(v, nii, is_label=is_label, pre_grid_sample_affine=P, **kw)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and torch.equal(x, y)


def test_error_behaviour_matches():
    R = load_reference()
    vol = torch.zeros(1, 1, 4, 4, 4)
    nii = torch.eye(4)[None]
    for fn in (R.nifti_grid_sample, O.nifti_grid_sample):
        with pytest.raises(Warning):
            fn(vol, nii, ras_transform_affine=torch.eye(4)[None])
        with pytest.raises(AssertionError):
            fn(vol[0], nii)
        with pytest.raises(AssertionError):
            fn(vol, nii, pre_grid_sample_affine=torch.eye(4)[None].repeat(2, 1, 1))


def test_random_aug_affine_family():
    """same distribution family as utils/transform_utils.py:6-23: zoom*rotation, det = zoom^3."""
    gen = torch.Generator().manual_seed(0)
    a = cases.synthetic.random_aug_affine(gen, 0.3, 0.2, 0.0)
    r = a[:3, :3]
    n = r.norm(dim=0)
    assert torch.allclose(n, n[0].expand(3), atol=1e-6)
    assert torch.allclose((r / n) @ (r / n).T, torch.eye(3), atol=1e-5)


def test_skip_connector_matches_reference():
    R = load_reference()
    case = cases.embed_case(8, 2, 2, 2, seed=7)
    a = R.SkipConnector(2)(case["x"], case["affines"])
    b = O.skip_connector(case["x"], case["affines"], 2)
    assert torch.equal(a, b)


def test_install_patches_and_restores_the_reference_bindings():
    """install() rebinds the hot-path names inside the imported reference modules, uninstall() puts the originals back."""
    R = load_reference()
    import acquisition_focus_b200.install as inst
    import acquisition_focus_b200 as afb
    lt, hu = R.learnable_transform, R.hybrid_unet
    orig = (lt.nifti_grid_sample, lt.compute_rotation_matrix_from_ortho6d, hu.SkipConnector)
    done = inst.install()
    try:
        assert lt.nifti_grid_sample is afb.nifti_grid_sample and hu.SkipConnector is afb.SkipConnector
        assert lt.compute_rotation_matrix_from_ortho6d is afb.compute_rotation_matrix_from_ortho6d
        assert len(done) >= 3
        with pytest.raises(Exception):                         # patched functions take CUDA tensors only: no CPU fallback
            lt.nifti_grid_sample(torch.zeros(1, 1, 4, 4, 4), torch.eye(4)[None].double())
    finally:
        inst.uninstall()
    assert (lt.nifti_grid_sample, lt.compute_rotation_matrix_from_ortho6d, hu.SkipConnector) == orig
