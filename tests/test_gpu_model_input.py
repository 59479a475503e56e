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


def test_reconstruction_model_input_matches_oracle():
    import acquisition_focus_b200 as afb
    from acquisition_focus_b200.running.model_input import get_reconstruction_model_input
    S, B, V, C = 32, 2, 3, 8
    case = cases.atm_case(S, B, V, seed=81)
    views = cases.synthetic.phantom_view_affines()
    names = ["p2CH", "p4CH", "4CH"]
    gen = torch.Generator().manual_seed(3)
    base = torch.stack([cases.synthetic.random_aug_affine(gen, 0.2, 0.1, 0.02) for _ in range(B)])
    cfg = types.SimpleNamespace(
        clinical_view_affine_type="from-gt", label_slice_type="from-gt", hires_fov_mm=[192.0] * 3, hires_fov_vox=[S] * 3,
        prescan_fov_mm=[192.0] * 3, prescan_fov_vox=[S] * 3, slice_fov_mm=[192.0, 192.0, 192.0 / S], slice_fov_vox=[S, S, 1],
        use_affine_theta=True, do_augment_input_orientation=False, do_augment_recon_orientation=False, aug_phases=["train"],
        sample_augment_strength=1.0, view_optimization_mode="opt-all", base_views=names, offset_clip_value=0.2, zoom_clip_value=0.0,
        affine_theta_optim_method="R6-vector", rotate_slice_to_min_principle=False)
    nets = iter([_Stub(case["params"][v].clone()) for v in range(V)])
    container = afb.ATModulesContainer(cfg, C, localization_net_factory=lambda: next(nets)).cuda()
    assert all(container.get_active_views())
    batch = {"label": case["lab"].cuda(), "image": case["image"][:, 0].cuda(),
             "additional_data": {"nifti_affine": case["nii"].cuda(),
                                 "gt_view_affines": {**{n: views[n][None].repeat(B, 1, 1).cuda() for n in names}, "centroids": base.cuda()}}}
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
