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
