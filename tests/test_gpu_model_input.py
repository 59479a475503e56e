"""Mirror of the reference's per-batch caller (running/run_dl.py:238-329) on the fused kernels vs its oracle restatement."""
import types

import pytest
import torch

from oracle import af_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu


class _Stub(torch.nn.Module):
    def __init__(self, p):
        super().__init__()
        self.p = torch.nn.Parameter(p)

    def forward(self, x):
        return self.p


S, B, V, C = 32, 2, 3, 8
NAMES = ["p2CH", "p4CH", "4CH"]


def _setup(label_slice_type="from-gt"):
    import acquisition_focus_b200 as afb
    case = cases.atm_case(S, B, V, seed=81)
    views = cases.synthetic.phantom_view_affines()
    gen = torch.Generator().manual_seed(3)
    base = torch.stack([cases.synthetic.random_aug_affine(gen, 0.2, 0.1, 0.02) for _ in range(B)])
    cfg = types.SimpleNamespace(
        clinical_view_affine_type="from-gt", label_slice_type=label_slice_type, hires_fov_mm=[192.0] * 3, hires_fov_vox=[S] * 3,
        prescan_fov_mm=[192.0] * 3, prescan_fov_vox=[S] * 3, slice_fov_mm=[192.0, 192.0, 192.0 / S], slice_fov_vox=[S, S, 1],
        use_affine_theta=True, do_augment_input_orientation=False, do_augment_recon_orientation=False, aug_phases=["train"],
        sample_augment_strength=1.0, view_optimization_mode="opt-all", base_views=NAMES, offset_clip_value=0.2, zoom_clip_value=0.0,
        affine_theta_optim_method="R6-vector", rotate_slice_to_min_principle=False)
    nets = iter([_Stub(case["params"][v].clone()) for v in range(V)])
    container = afb.ATModulesContainer(cfg, C, localization_net_factory=lambda: next(nets)).cuda()
    assert all(container.get_active_views())
    batch = {"label": case["lab"].cuda(), "image": case["image"][:, 0].cuda(),
             "additional_data": {"nifti_affine": case["nii"].cuda(),
                                 "gt_view_affines": {**{n: views[n][None].repeat(B, 1, 1).cuda() for n in NAMES}, "centroids": base.cuda()}}}
    return case, views, base, cfg, container, batch


def test_per_view_route_equals_fused_route(monkeypatch):
    """The view-by-view route through get_transformed (run_dl.py:146-204, used for 'from-segmented', non-R6 parameterisations
    and in-plane re-alignment) gives what the fused all-views acquisition gives."""
    from acquisition_focus_b200.running import model_input as MI
    _, _, _, cfg, container, batch = _setup()
    fused = MI.get_reconstruction_model_input(batch, "train", cfg, C, container)
    go = cases.pattern(fused[0].shape, 1.0).cuda()
    (fused[0] * go).sum().backward()
    g_fused = [container[v].localization_net.p.grad.clone() for v in range(V)]
    for v in range(V):
        container[v].localization_net.p.grad = None
    monkeypatch.setattr(MI, "_fused_route_ok", lambda *a, **k: False)
    per_view = MI.get_reconstruction_model_input(batch, "train", cfg, C, container)
    (per_view[0] * go).sum().backward()
    assert torch.equal(per_view[0], fused[0]) and torch.equal(per_view[1], fused[1])
    for a, b in zip(per_view[2], fused[2]):
        assert torch.equal(a, b)
    for v in range(V):
        g = container[v].localization_net.p.grad
        assert (g - g_fused[v]).abs().max().item() <= 1e-5 * g_fused[v].abs().max().item()


def test_from_segmented_uses_the_callers_segmenter():
    """label_slice_type='from-segmented' outside training (run_dl.py:172-190): the label slices are one-hot(segment_fn(image
    slice, zooms)) and carry no gradient; segment_fn is the caller's (the reference passes its nnU-Net wrapper)."""
    from acquisition_focus_b200.running import model_input as MI
    _, _, _, cfg_gt, container, batch = _setup()
    seen = []

    def segment_fn(image_slc, zooms):                      # [B,1,1,D,H] image slice, [B,3] voxel sizes -> [B,1,D,H] labels
        assert image_slc.shape == (B, 1, 1, S, S) and zooms.shape == (B, 3)
        seen.append(image_slc)
        return (image_slc[:, :, 0] * 7.0).clamp(0, C - 1).round()

    _, _, _, cfg_seg, container2, _ = _setup("from-segmented")
    b_input, b_target, grid_affines = MI.get_reconstruction_model_input(batch, "val", cfg_seg, C, container2, segment_fn=segment_fn)
    assert len(seen) == V and b_input.shape == (B, V * C, S, S) and not b_input.requires_grad
    for v in range(V):
        want = torch.nn.functional.one_hot((seen[v][:, 0, 0] * 7.0).clamp(0, C - 1).round().long(), C).permute(0, 3, 1, 2).float()
        assert torch.equal(b_input[:, v * C:(v + 1) * C], want)
    # in training the segmenter is not consulted (reference: `phase != 'train'`)
    b_train, _, _ = MI.get_reconstruction_model_input(batch, "train", cfg_seg, C, container2, segment_fn=segment_fn)
    assert len(seen) == V and b_train.requires_grad
    fused, _, _ = MI.get_reconstruction_model_input(batch, "train", cfg_gt, C, container)
    assert torch.equal(b_train, fused)


def test_reconstruction_model_input_matches_oracle():
    from acquisition_focus_b200.running.model_input import get_reconstruction_model_input
    case, views, base, cfg, container, batch = _setup()
    names = NAMES
    b_input, b_target, grid_affines = get_reconstruction_model_input(batch, "train", cfg, C, container)
    assert b_input.shape == (B, V * C, S, S) and b_target.shape == (B, C, S, S, S) and len(grid_affines) == V
    go = cases.pattern(b_input.shape, 1.0)
    (b_input * go.cuda()).sum().backward()

    mlp = [case["params"][v].clone().requires_grad_(True) for v in range(V)]
    init = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(V, 1)
    r_input, r_target, r_affines = O.reconstruction_model_input(
        case["lab"], case["image"][:, 0], case["nii"], base.double(), [views[n][None].repeat(B, 1, 1) for n in names], mlp, init,
        torch.tensor(cfg.hires_fov_mm), torch.tensor(cfg.hires_fov_vox), torch.tensor(cfg.slice_fov_mm), torch.tensor(cfg.slice_fov_vox),
        C, 0.2, 0.0, S)
    (r_input * go).sum().backward()
    assert torch.equal(b_target.cpu(), r_target)                                   # hires nearest resample: bit-exact
    for a, b in zip(grid_affines, r_affines):
        assert (a.detach().cpu() - b.detach()).abs().max().item() <= 2e-6 * b.abs().max().item()
    assert (b_input.detach().cpu() - r_input.detach()).abs().max().item() <= 2e-5
    for v in range(V):
        g, r = container[v].localization_net.p.grad.cpu(), mlp[v].grad
        assert (g - r).abs().max().item() <= 1e-4 * r.abs().max().item()


@pytest.mark.parametrize("tag,kw", [("model_input_s32", dict(S=32)), ("model_input_s32_lowres", dict(S=32, slice_vox=16)),
                                    ("model_input_s32_noaug", dict(S=32, aug=False))])
def test_reconstruction_model_input_golden(golden_dir, tag, kw):
    """a11 / a6 against the reference ITSELF: goldens minted by executing running/run_dl.py:238-329 unmodified
    (oracle/make_golden.py::gold_model_input), with input AND reconstruction augmentation drawn from the global torch RNG
    (seed 123: the product issues the reference's draw sequence, so the augmentation affines are bitwise the same) and, in
    the `lowres` case, 16x16 slices up-sampled to the 32^3 hires FOV (run_dl.py:193-197)."""
    import os
    import numpy as np
    import acquisition_focus_b200 as afb
    from acquisition_focus_b200.running.model_input import get_reconstruction_model_input
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    cfgd, batch, params, (B_, V_, C_, names) = cases.model_input_setup(**kw)
    cfg = types.SimpleNamespace(**cfgd)
    nets = iter([_Stub(params[v].clone()) for v in range(V_)])
    container = afb.ATModulesContainer(cfg, C_, localization_net_factory=lambda: next(nets)).cuda()
    ad = batch["additional_data"]
    cb = {"label": batch["label"].cuda(), "image": batch["image"].cuda(),
          "additional_data": {"nifti_affine": ad["nifti_affine"].cuda(), "gt_view_affines": {k: v.cuda() for k, v in ad["gt_view_affines"].items()}}}
    torch.manual_seed(123)
    b_input, b_target, grid_affines = get_reconstruction_model_input(cb, "train", cfg, C_, container)
    loss = (b_input * cases.pattern(b_input.shape, 1.0).cuda()).sum()
    for v, a in enumerate(grid_affines):
        loss = loss + (a * cases.pattern(a.shape, 2.0 + v).cuda()).sum()
    loss.backward()
    assert np.array_equal(b_target.argmax(1).cpu().numpy().astype(np.uint8), g["b_target_argmax"])      # hires nearest: bit-exact
    ga = torch.stack([a.detach().cpu() for a in grid_affines]).numpy()
    e_ga = np.abs(ga - g["grid_affines"]).max() / np.abs(g["grid_affines"]).max()
    e_in = np.abs(b_input.detach().cpu().numpy() - g["b_input"]).max()
    dp = torch.stack([container[v].localization_net.p.grad.cpu() for v in range(V_)]).numpy()
    e_dp = np.abs(dp - g["dparams"]).max() / np.abs(g["dparams"]).max()
    print(f"{tag}: grid_affine {e_ga:.2e}  b_input {e_in:.2e}  dparams {e_dp:.2e}")
    assert e_ga <= 2e-6 and e_in <= 2e-5 and e_dp <= 1e-4, (e_ga, e_in, e_dp)
