"""Pin the oracle port (oracle/af_oracle.py) against golden vectors minted from the unmodified
reference (oracle/make_golden.py).  Forward results must be bitwise equal on the same torch
build; gradients within 1e-5 of the gradient scale."""
import os

import numpy as np
import pytest
import torch

from oracle import af_oracle as O
from oracle import cases


def _close(a, b, rel=1e-5):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() <= rel * scale


def test_r6(golden_dir):
    g = np.load(os.path.join(golden_dir, "r6.npz"))
    o = torch.from_numpy(g["ortho"]).requires_grad_(True)
    m = O.r6_to_matrix(o)
    assert np.array_equal(m.detach().numpy(), g["mat"])
    (m * cases.pattern(m.shape, 1.0)).sum().backward()
    assert _close(o.grad.numpy(), g["d_ortho"])


@pytest.mark.parametrize("tag,kw", [
    ("slice", dict(target_fov_mm=torch.tensor([30.0, 20.0, 1.5]), target_fov_vox=torch.tensor([16, 12, 1]))),
    ("vol3d", dict(target_fov_mm=torch.tensor([28.0, 30.0, 33.0]), target_fov_vox=torch.tensor([9, 10, 11]))),
    ("same", dict())])
def test_slice_small(golden_dir, tag, kw):
    g = np.load(os.path.join(golden_dir, "slice_small.npz"))
    B, C, D, H, W = 2, 3, 20, 24, 28
    vol = cases.randn((B, C, D, H, W), 21).requires_grad_(True)
    lab = cases.randint(0, 6, (B, C, D, H, W), 22)
    nii = torch.from_numpy(g["nii"]); P = torch.from_numpy(g["P"]).requires_grad_(True)
    y, ga, na = O.nifti_grid_sample(vol, nii, is_label=False, pre_grid_sample_affine=P, **kw)
    assert np.array_equal(y.detach().numpy(), g[f"{tag}_y"])
    assert np.array_equal(ga.detach().numpy(), g[f"{tag}_ga"])
    assert np.allclose(na.detach().numpy(), g[f"{tag}_nii"], rtol=1e-12, atol=1e-12)
    ((y * cases.pattern(y.shape, 1.0)).sum() + (ga * cases.pattern(ga.shape, 2.0)).sum()).backward()
    assert _close(vol.grad.numpy(), g[f"{tag}_dvol"])
    assert _close(P.grad.numpy(), g[f"{tag}_dP"])
    yl, _, _ = O.nifti_grid_sample(lab, nii, is_label=True, pre_grid_sample_affine=P.detach(), **kw)
    assert yl.dtype == torch.int64 and np.array_equal(yl.numpy(), g[f"{tag}_ylabel"])


def test_slice_cfg1_128(golden_dir):
    g = np.load(os.path.join(golden_dir, "slice_cfg1_128.npz"))
    syn = cases.synthetic
    vol = torch.from_numpy(syn.phantom_image(syn.heart_phantom(128), seed=5))[None, None].requires_grad_(True)
    r6 = torch.from_numpy(g["r6"]).requires_grad_(True)
    P = torch.from_numpy(g["gpre"]) @ O.r6_to_matrix(r6)
    y, ga, na = O.nifti_grid_sample(vol, syn.default_nifti_affine(1), target_fov_mm=torch.tensor([192.0, 192.0, 1.5]),
                                    target_fov_vox=torch.tensor([128, 128, 1]), pre_grid_sample_affine=P)
    assert np.array_equal(y.detach().numpy(), g["y"]) and np.array_equal(ga.detach().numpy(), g["ga"])
    (y * cases.pattern(y.shape, 1.0)).sum().backward()
    assert _close(r6.grad.numpy(), g["d_r6"])
    dv = vol.grad[0, 0]
    assert _close(dv.sum(0).numpy(), g["dvol_sum_d"]) and _close(dv.sum(2).numpy(), g["dvol_sum_w"])


@pytest.mark.parametrize("tag,zc", [("atm_s32", 0.0), ("atm_s32_zoom", 0.3)])
def test_atm_s32(golden_dir, tag, zc):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    case = cases.atm_case(32, 2, 3, seed=41, zoom_clip=zc)
    for v in range(3):
        assert np.array_equal(case["gpre"][v].numpy(), g["gpre"][v])
        params = case["params"][v].clone().requires_grad_(True)
        soft = case["soft"].clone().requires_grad_(True)
        theta = O.view_theta(params, torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0]]), torch.zeros(3), torch.ones(1, 1),
                             case["offset_clip"], zc, 32)
        assert np.array_equal(theta.detach().numpy(), g[f"theta{v}"])
        ys, yl, yi, ga, na = O.atm_tail_forward(soft, case["label"], case["image"], case["nii"], case["gpre"][v], theta,
                                                case["slice_fov_mm"], case["slice_fov_vox"])
        assert np.array_equal(ys.detach().numpy(), g[f"ys{v}"])
        assert np.array_equal(yl.numpy().astype(np.uint8), g[f"yl{v}"])
        assert np.array_equal(yi.numpy(), g[f"yi{v}"])
        assert np.array_equal(ga.detach().numpy(), g[f"ga{v}"])
        ((ys * cases.pattern(ys.shape, 1.0 + v)).sum() + (ga * cases.pattern(ga.shape, 2.0 + v)).sum()).backward()
        assert _close(params.grad.numpy(), g[f"dparams{v}"])
        assert _close(soft.grad.sum(-1).numpy(), g[f"dsoft_sum_w{v}"])


@pytest.mark.parametrize("tag,shape", [("embed_s16", (16, 3, 2, 2)), ("embed_s8", (8, 4, 3, 2)), ("embed_s32", (32, 4, 6, 1))])
def test_embed(golden_dir, tag, shape):
    S, c, V, B = shape
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    case = cases.embed_case(S, c, V, B, seed=51 + S)
    x = case["x"].clone().requires_grad_(True)
    gas = [a.clone().requires_grad_(True) for a in case["affines"]]
    assert np.array_equal(np.stack([a.detach().numpy() for a in gas]), g["affines"])
    out = O.skip_connector(x, gas, V)
    assert np.array_equal(out.detach().numpy(), g["out"])
    (out * cases.pattern(out.shape, 1.0)).sum().backward()
    assert _close(x.grad.numpy(), g["dx"])
    assert _close(np.stack([a.grad.numpy() for a in gas]), g["d_affines"], rel=1e-4)
    sp = O.skip_connector_sparse(x.detach(), [a.detach() for a in gas], V)
    assert _close(sp.numpy(), g["out"], rel=1e-5)


# ---- the two non-default rotation parameterisations --------------------------------------------------
@pytest.mark.parametrize("tag,fn", [("aa", "angle_axis"), ("nv", "normal")])
def test_rotation_params(golden_dir, tag, fn):
    """oracle restatement (bitwise forward) against the reference's transform_utils.py:62-178; the product's kernels
    (afb_rot3_fwd/bwd) are checked against the same golden on the GPU (tests/test_gpu_aux.py)."""
    g = np.load(os.path.join(golden_dir, "rotation_params.npz"))
    oracle_fn = {"angle_axis": O.angle_axis_to_matrix, "normal": O.normal_to_matrix}[fn]
    for f, exact in ((oracle_fn, True),):
        x = torch.from_numpy(g[f"{tag}_in"]).requires_grad_(True)
        m = f(x)
        if exact:
            assert np.array_equal(m.detach().numpy(), g[f"{tag}_mat"])
        else:
            assert _close(m.detach().numpy(), g[f"{tag}_mat"], rel=1e-6)
        (m * cases.pattern(m.shape, 1.0)).sum().backward()
        assert _close(x.grad.numpy(), g[f"{tag}_grad"])


@pytest.mark.parametrize("method,init_ap", [("angle-axis", None), ("normal-vector", [0.2, -0.1, 1.0])])
def test_atm_other_parameterisations(golden_dir, method, init_ap):
    g = np.load(os.path.join(golden_dir, "atm_s32_" + method.replace("-", "_") + ".npz"))
    case = cases.atm_case(32, 2, 2, seed=71)
    ia = torch.tensor(init_ap) if init_ap is not None else torch.zeros(3)
    for v in range(2):
        params = torch.from_numpy(g["params"][v]).requires_grad_(True)
        theta = O.view_theta(params, ia, torch.zeros(3), torch.ones(1, 1), case["offset_clip"], 0.0, 32, optim_method=method)
        assert np.array_equal(theta.detach().numpy(), g[f"theta{v}"])
        ys, yl, yi, ga, na = O.atm_tail_forward(case["soft"], case["label"], case["image"], case["nii"], case["gpre"][v], theta,
                                                case["slice_fov_mm"], case["slice_fov_vox"])
        assert np.array_equal(ys.detach().numpy(), g[f"ys{v}"]) and np.array_equal(ga.detach().numpy(), g[f"ga{v}"])
        assert np.array_equal(yl.numpy().astype(np.uint8), g[f"yl{v}"]) and np.array_equal(yi.numpy(), g[f"yi{v}"])
        ((ys * cases.pattern(ys.shape, 1.0 + v)).sum() + (ga * cases.pattern(ga.shape, 2.0 + v)).sum()).backward()
        assert _close(params.grad.numpy(), g[f"dparams{v}"])


# ---- in-plane re-alignment of the slices ---------------------------------------------------------------
def test_atm_rotate_slice_to_min_principle(golden_dir):
    """oracle restatement of learnable_transform.py:315-328,337-366 against the reference's outputs, and the product's
    batched alignment-affine helper (device-agnostic torch + host eig) against the oracle's per-sample loop."""
    from acquisition_focus_b200.models.learnable_transform import min_principle_align_affines
    g = np.load(os.path.join(golden_dir, "atm_s32_rotate.npz"))
    case = cases.atm_case(32, 2, 2, seed=81)
    for v in range(2):
        params = case["params"][v].clone().requires_grad_(True)
        soft = case["soft"].clone().requires_grad_(True)
        theta = O.view_theta(params, torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0]]), torch.zeros(3), torch.ones(1, 1), case["offset_clip"], 0.0, 32)
        ys, yl, yi, ga, na = O.atm_tail_forward(soft, case["label"], case["image"], case["nii"], case["gpre"][v], theta,
                                                case["slice_fov_mm"], case["slice_fov_vox"])
        helper = min_principle_align_affines(ys.detach())
        loop = torch.stack([O.min_principle_align_affine(sl.argmax(0)) for sl in ys.detach()])
        assert _close(helper.numpy(), loop.numpy(), rel=1e-6)
        ys, align, na2 = O.rotate_slice_to_min_principle(ys, na, is_label=False)      # `align` is the resample's grid affine (:362)
        with torch.no_grad():       # the reference threads one NIfTI affine through all three calls (:317-326)
            yl, _, na2 = O.rotate_slice_to_min_principle(yl, na2, is_label=True, align_affine_override=align)
            yi, _, na2 = O.rotate_slice_to_min_principle(yi, na2, is_label=False, align_affine_override=align)
        ga = ga @ align
        assert np.array_equal(ys.detach().numpy(), g[f"ys{v}"]) and np.array_equal(ga.detach().numpy(), g[f"ga{v}"])
        assert np.array_equal(yl.numpy().astype(np.uint8), g[f"yl{v}"]) and np.array_equal(yi.numpy(), g[f"yi{v}"])
        assert np.allclose(na2.detach().numpy(), g[f"na{v}"], rtol=1e-12, atol=1e-12)
        ((ys * cases.pattern(ys.shape, 1.0 + v)).sum() + (ga * cases.pattern(ga.shape, 2.0 + v)).sum()).backward()
        assert _close(params.grad.numpy(), g[f"dparams{v}"])
        assert _close(soft.grad.sum(-1).numpy(), g[f"dsoft_sum_w{v}"])


MODEL_INPUT_CASES = [("model_input_s32", dict(S=32)), ("model_input_s32_lowres", dict(S=32, slice_vox=16)),
                     ("model_input_s32_noaug", dict(S=32, aug=False))]


@pytest.mark.parametrize("tag,kw", MODEL_INPUT_CASES)
def test_reconstruction_model_input(golden_dir, tag, kw):
    """a11 / a6: the restatement of running/run_dl.py:208-329 against outputs of the reference's OWN
    get_reconstruction_model_input (executed unmodified by oracle/make_golden.py::gold_model_input), including the
    augmentation affines drawn from the global torch RNG (seed 123) and the up-sampling of low-resolution slices."""
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    cfg, batch, params, (B, V, C, names) = cases.model_input_setup(**kw)
    mlp = [p.clone().requires_grad_(True) for p in params]
    init = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(V, 1)
    ad = batch["additional_data"]
    torch.manual_seed(123)
    b_in, b_t, affs = O.reconstruction_model_input(
        batch["label"], batch["image"], ad["nifti_affine"], ad["gt_view_affines"]["centroids"].to(ad["nifti_affine"]),
        [ad["gt_view_affines"][n] for n in names], mlp, init, torch.tensor(cfg["hires_fov_mm"]), torch.tensor(cfg["hires_fov_vox"]),
        torch.tensor(cfg["slice_fov_mm"]), torch.tensor(cfg["slice_fov_vox"]), C, cfg["offset_clip_value"], cfg["zoom_clip_value"],
        cfg["prescan_fov_vox"][0], augment_input=cfg["do_augment_input_orientation"], augment_recon=cfg["do_augment_recon_orientation"])
    loss = (b_in * cases.pattern(b_in.shape, 1.0)).sum()
    for v, a in enumerate(affs):
        loss = loss + (a * cases.pattern(a.shape, 2.0 + v)).sum()
    loss.backward()
    assert np.array_equal(b_in.detach().numpy(), g["b_input"])
    assert np.array_equal(b_t.argmax(1).numpy().astype(np.uint8), g["b_target_argmax"])
    assert np.array_equal(torch.stack(affs).detach().numpy(), g["grid_affines"])
    assert _close(torch.stack([m.grad for m in mlp]).numpy(), g["dparams"])
