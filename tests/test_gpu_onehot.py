"""One-hot label slicing straight from the integer label map: must be bitwise what the dense path
(and therefore the reference) produces on the materialised one-hot volumes, for every integer storage type."""
import pytest
import torch

from oracle import af_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu
INIT = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]])


@pytest.fixture(scope="module")
def afb():
    import acquisition_focus_b200 as m
    return m


def _kw(case):
    return dict(offset_clip=case["offset_clip"], zoom_clip=case["zoom_clip"], spat=case["S"],
                slice_fov_mm=case["slice_fov_mm"].tolist(), slice_fov_vox=case["slice_fov_vox"].tolist())


@pytest.mark.parametrize("S,dt", [(32, torch.uint8), (32, torch.int64), (48, torch.int16), (128, torch.uint8), (128, torch.int32)])
def test_from_labels_bitwise_vs_dense_path(afb, S, dt):
    case = cases.atm_case(S, 2, 3, seed=60 + S)
    V = case["V"]
    gpre = torch.stack(case["gpre"], 1).cuda()
    init = INIT.repeat(V, 1).cuda()
    p1 = torch.stack(case["params"], 1).cuda().requires_grad_(True)
    p2 = p1.detach().clone().requires_grad_(True)
    dense = afb.acquire_views(case["soft"].cuda(), case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), gpre, p1, init, **_kw(case))
    lab = case["lab"].to(dt).cuda()
    fast = afb.acquire_views_from_labels(lab, case["image"].cuda(), case["nii"].cuda(), gpre, p2, init, num_classes=8, **_kw(case))
    assert torch.equal(fast[3], dense[3])                        # same grid affine
    assert torch.equal(fast[0], dense[0])                        # y_soft bitwise
    assert fast[1].dtype == torch.int64 and torch.equal(fast[1], dense[1])
    assert torch.equal(fast[2], dense[2])
    go = cases.pattern(dense[0].shape, 1.0).cuda()
    gg = cases.pattern(dense[3].shape, 2.0).cuda()
    ((dense[0] * go).sum() + (dense[3] * gg).sum()).backward()
    ((fast[0] * go).sum() + (fast[3] * gg).sum()).backward()
    scale = p1.grad.abs().max().item()
    assert (p1.grad - p2.grad).abs().max().item() <= 1e-5 * scale
    # compact index output == argmax of the one-hot output (0 out of field)
    idx = afb.acquire_views_from_labels(lab, None, case["nii"].cuda(), gpre, p2.detach(), init, num_classes=8, label_out="index", **_kw(case))[1]
    assert idx.dtype == torch.uint8 and torch.equal(idx.long(), dense[1].argmax(2))


def test_cfg5_256_dense_and_label_paths_vs_oracle(afb):
    """configs[4] size: 256^3 volume, 256^2 slices, C=8, fp32 and bf16 storage; checked against the oracle port on
    the same inputs (forward bitwise given the same grid affine; dTheta 1e-4; bf16 after rounding, 2^-8)."""
    import torch.nn.functional as F
    S, V, C = 256, 2, 8
    gen = torch.Generator().manual_seed(5)
    lab = torch.randint(0, C, (1, S // 8, S // 8, S // 8), generator=gen)
    lab = lab.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3)          # blocky 256^3 label map
    label, soft = cases.one_hot_volumes(lab, C)
    syn = cases.synthetic
    gpre = torch.stack([syn.phantom_view_affines()["p2CH"] @ syn.random_aug_affine(gen, 0.3, 0.2, 0.0) for _ in range(V)])[None]
    R = 51
    params = torch.cat([torch.tensor([1.0, 0, 0, 0, 1.0, 0]) + 0.3 * torch.randn(1, V, 6, generator=gen),
                        torch.randn(1, V, 3 * R, generator=gen), torch.randn(1, V, 1, generator=gen)], -1)
    nii = syn.default_nifti_affine(1, 0.75)
    kw = dict(offset_clip=0.2, zoom_clip=0.0, spat=S, slice_fov_mm=[192.0, 192.0, 0.75], slice_fov_vox=[S, S, 1])
    init = INIT.repeat(V, 1)
    p_fast = params.cuda().requires_grad_(True)
    ys, yl, _, ga, _, _ = afb.acquire_views_from_labels(lab.to(torch.uint8).cuda(), None, nii.cuda(), gpre.cuda(), p_fast, init.cuda(),
                                                        num_classes=C, **kw)
    go = cases.pattern(ys.shape, 1.0)
    (ys * go.cuda()).sum().backward()
    p_dense = params.cuda().requires_grad_(True)
    yd = afb.acquire_views(soft.cuda(), None, None, nii.cuda(), gpre.cuda(), p_dense, init.cuda(), **kw)[0]
    (yd * go.cuda()).sum().backward()
    assert torch.equal(ys, yd)
    assert (p_fast.grad - p_dense.grad).abs().max().item() <= 1e-5 * p_dense.grad.abs().max().item()
    yb = afb.acquire_views(soft.to(torch.bfloat16).cuda(), None, None, nii.cuda(), gpre.cuda(), params.cuda(), init.cuda(), **kw)[0]
    assert yb.dtype == torch.bfloat16 and (yb.float() - yd.detach().to(torch.bfloat16).float()).abs().max().item() <= 2.0 ** -7
    for v in range(V):          # oracle: the kernel's own grid affine through ATen (bitwise), dTheta through the full chain
        grid = F.affine_grid(ga[:, v, :3, :].cpu(), [1, C, S, S, 1], align_corners=False)
        assert torch.equal(ys[:, v].detach().cpu(), F.grid_sample(soft, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
        assert torch.equal(yl[:, v].cpu(), F.grid_sample(label.float(), grid, mode="nearest", padding_mode="zeros", align_corners=False).long())
        p = params[:, v].clone().requires_grad_(True)
        theta = O.view_theta(p, INIT[:, :6], INIT[0, 6:9], INIT[:, 9:], 0.2, 0.0, S)
        rs = O.atm_tail_forward(soft, None, None, nii, gpre[:, v], theta, torch.tensor(kw["slice_fov_mm"]), torch.tensor(kw["slice_fov_vox"]))[0]
        (rs * go[:, v]).sum().backward()
        assert (p_fast.grad[:, v].cpu() - p.grad).abs().max().item() <= 1e-4 * p.grad.abs().max().item()


def test_from_labels_vs_oracle(afb):
    """Directly against the oracle port of the reference (one view at a time, as the reference loops)."""
    case = cases.atm_case(32, 2, 3, seed=71, zoom_clip=0.25)
    gpre = torch.stack(case["gpre"], 1).cuda()
    params = torch.stack(case["params"], 1).cuda().requires_grad_(True)
    ys, yl, yi, ga, nii, th = afb.acquire_views_from_labels(case["lab"].cuda(), case["image"].cuda(), case["nii"].cuda(), gpre, params,
                                                            INIT.repeat(3, 1).cuda(), num_classes=8, **_kw(case))
    loss = 0
    for v in range(3):
        loss = loss + (ys[:, v] * cases.pattern(ys[:, v].shape, 1.0 + v).cuda()).sum()
    loss.backward()
    for v in range(3):
        p = case["params"][v].clone().requires_grad_(True)
        theta = O.view_theta(p, INIT[:, :6], INIT[0, 6:9], INIT[:, 9:], 0.2, 0.25, 32)
        rs, rl, ri, rga, rn = O.atm_tail_forward(case["soft"], case["label"], case["image"], case["nii"], case["gpre"][v], theta,
                                                 case["slice_fov_mm"], case["slice_fov_vox"])
        (rs * cases.pattern(rs.shape, 1.0 + v)).sum().backward()
        assert (ga[:, v].cpu() - rga).abs().max().item() <= 2e-6 * rga.abs().max().item()
        assert (ys[:, v].detach().cpu() - rs.detach()).abs().max().item() <= 2e-5
        assert (params.grad[:, v].cpu() - p.grad).abs().max().item() <= 1e-4 * p.grad.abs().max().item()
        assert (yl[:, v].cpu() != rl).float().mean().item() < 5e-3
