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


def test_random_affine_stream_bitwise():
    """a6: utils/transform_utils.py:6-23 draws from the global torch RNG; the oracle restatement and the PRODUCT's
    get_random_affine issue the same draws in the same order, so a seeded run reproduces the reference bit for bit."""
    R = load_reference()
    from acquisition_focus_b200.utils.transform_utils import get_random_affine as product
    for seed, (r, z, o) in enumerate([(0.1, 0.2, 0.0), (0.3, 0.2, 0.1), (4.0, 0.0, 0.0)]):
        outs = []
        for fn in (R.get_random_affine, O.get_random_affine, product):
            torch.manual_seed(seed)
            outs.append(torch.stack([fn(rotation_strength=r, zoom_strength=z, offset_strength=o) for _ in range(3)]))
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_reconstruction_model_input_live():
    """a11: the restatement against the reference's own get_reconstruction_model_input, executed unmodified (imports that
    are absent here stubbed by oracle.ref_import.load_run_dl), augmentation on, seeded global RNG: bitwise."""
    from oracle.ref_import import load_run_dl
    R, RD = load_reference(), load_run_dl()
    cfg, batch, params, (B, V, C, names) = cases.model_input_setup(S=32)
    config = RD.DotDict(cfg)
    container = R.learnable_transform.ATModulesContainer(config, C)

    class Stub(torch.nn.Module):
        def __init__(self, p):
            super().__init__()
            self.p = torch.nn.Parameter(p)

        def forward(self, x):
            return self.p
    for v, atm in enumerate(container):
        atm.localization_net = Stub(params[v].clone())
    torch.manual_seed(7)
    a_in, a_t, a_aff = RD.get_reconstruction_model_input(batch, "train", config, C, container, None)
    init = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(V, 1)
    ad = batch["additional_data"]
    torch.manual_seed(7)
    b_in, b_t, b_aff = O.reconstruction_model_input(
        batch["label"], batch["image"], ad["nifti_affine"], ad["gt_view_affines"]["centroids"].to(ad["nifti_affine"]),
        [ad["gt_view_affines"][n] for n in names], params, init, torch.tensor(cfg["hires_fov_mm"]), torch.tensor(cfg["hires_fov_vox"]),
        torch.tensor(cfg["slice_fov_mm"]), torch.tensor(cfg["slice_fov_vox"]), C, 0.2, 0.0, 32, augment_input=True, augment_recon=True)
    assert torch.equal(a_in, b_in) and torch.equal(a_t, b_t)
    for x, y in zip(a_aff, b_aff):
        assert torch.equal(x, y)


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
