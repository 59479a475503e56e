"""Parity of the fused view-parameter path (raw R6 / offset logits / zoom -> slices, all B x V at once)
and of the slice -> 3-D embedding against golden vectors from the unmodified reference and the
oracle port."""
import os

import numpy as np
import pytest
import torch

from oracle import af_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu

FWD_REL = 1e-5
GRAD_REL = 1e-4
INIT = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]])


@pytest.fixture(scope="module")
def afb():
    import acquisition_focus_b200 as m
    return m


def close(a, b, rel, what=""):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= rel * scale, f"{what}: max abs err {err:.3e} > {rel:g} * scale {scale:.3e}"


def _acquire(afb, case, soft_requires_grad=True):
    B, V = case["B"], case["V"]
    soft = case["soft"].cuda().requires_grad_(soft_requires_grad)
    params = torch.stack(case["params"], dim=1).cuda().requires_grad_(True)         # [B,V,NP]
    gpre = torch.stack(case["gpre"], dim=1).cuda()
    out = afb.acquire_views(soft, case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), gpre, params,
                            INIT.repeat(V, 1).cuda(), offset_clip=case["offset_clip"], zoom_clip=case["zoom_clip"],
                            spat=case["S"], slice_fov_mm=case["slice_fov_mm"].tolist(),
                            slice_fov_vox=case["slice_fov_vox"].tolist())
    return soft, params, out


@pytest.mark.parametrize("tag,zc", [("atm_s32", 0.0), ("atm_s32_zoom", 0.3)])
def test_acquire_views_golden_s32(afb, golden_dir, tag, zc):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    case = cases.atm_case(32, 2, 3, seed=41, zoom_clip=zc)
    soft, params, (ys, yl, yi, ga, nii, theta) = _acquire(afb, case)
    V = case["V"]
    loss = 0
    for v in range(V):
        close(theta[:, v], g[f"theta{v}"], 1e-6, "theta")
        close(ga[:, v], g[f"ga{v}"], 2e-6, "grid_affine")
        assert np.allclose(nii[:, v].cpu().numpy(), g[f"na{v}"], rtol=1e-5, atol=1e-4)
        close(ys[:, v], g[f"ys{v}"], 2e-5, "y_soft")        # G' agrees to ~1e-6 => values to ~1e-5 of scale
        close(yi[:, v], g[f"yi{v}"], 1e-4, "y_image")
        mism = (yl[:, v].cpu().to(torch.uint8) != torch.from_numpy(g[f"yl{v}"])).float().mean().item()
        assert mism < 5e-3, f"label mismatch fraction {mism}"
        loss = loss + (ys[:, v] * cases.pattern(ys[:, v].shape, 1.0 + v).cuda()).sum() \
                    + (ga[:, v] * cases.pattern(ga[:, v].shape, 2.0 + v).cuda()).sum()
    loss.backward()
    dsoft = 0
    for v in range(V):
        close(params.grad[:, v], g[f"dparams{v}"], GRAD_REL, "dparams")
    # dVolume accumulates over the three views (autograd sums); golden stores per-view projections
    want_w = sum(torch.from_numpy(g[f"dsoft_sum_w{v}"]) for v in range(V))
    want_d = sum(torch.from_numpy(g[f"dsoft_sum_d{v}"]) for v in range(V))
    close(soft.grad.sum(-1), want_w, GRAD_REL, "dsoft_sum_w")
    close(soft.grad.sum(2), want_d, GRAD_REL, "dsoft_sum_d")


def test_acquire_views_self_consistent_labels(afb):
    """Bit-exact nearest / argmax given the SAME grid affine: feed the kernel's own G' to the oracle sampler."""
    case = cases.atm_case(32, 2, 3, seed=45)
    soft, params, (ys, yl, yi, ga, nii, theta) = _acquire(afb, case, soft_requires_grad=False)
    import torch.nn.functional as F
    for v in range(case["V"]):
        th = ga[:, v, :3, :].cpu()
        grid = F.affine_grid(th, [case["B"], 8, 32, 32, 1], align_corners=False)
        ref_l = F.grid_sample(case["label"].float(), grid, mode="nearest", padding_mode="zeros", align_corners=False).long()
        assert torch.equal(yl[:, v].cpu(), ref_l)
        ref_s = F.grid_sample(case["soft"], grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        assert torch.equal(ys[:, v].detach().cpu(), ref_s)
        assert torch.equal(ys[:, v].argmax(1).cpu(), ref_s.argmax(1))


def test_acquire_views_golden_s128(afb, golden_dir):
    """configs[1] shapes: B=2 x V=3 p2CH views, 8-class one-hot label + image, 128^3 -> 128^2."""
    g = np.load(os.path.join(golden_dir, "atm_s128.npz"))
    case = cases.atm_case(128, 2, 3, seed=43)
    soft, params, (ys, yl, yi, ga, nii, theta) = _acquire(afb, case, soft_requires_grad=False)
    loss = 0
    for v in range(3):
        close(ga[:, v], g[f"ga{v}"], 2e-6, "grid_affine")
        assert np.allclose(nii[:, v].cpu().numpy(), g[f"na{v}"], rtol=1e-5, atol=1e-4)
        close(ys[:, v, 3], g[f"ys_ch3_{v}"], 1e-4, "y_soft ch3")
        close(yi[:, v], g[f"yi{v}"], 2e-4, "y_image")
        for got, want in ((yl[:, v].cpu().to(torch.uint8), g[f"yl{v}"]), (ys[:, v].argmax(1).cpu().to(torch.uint8), g[f"ys_argmax{v}"])):
            mism = (got != torch.from_numpy(want)).float().mean().item()
            assert mism < 2e-3, f"label mismatch fraction {mism}"
        loss = loss + (ys[:, v] * cases.pattern(ys[:, v].shape, 1.0 + v).cuda()).sum() \
                    + (ga[:, v] * cases.pattern(ga[:, v].shape, 2.0 + v).cuda()).sum()
    loss.backward()
    for v in range(3):
        close(params.grad[:, v], g[f"dparams{v}"], GRAD_REL, "dparams")
    # encoder-input layout of running/run_dl.py:325 comes for free
    assert ys.flatten(1, 2).squeeze(-1).shape == (2, 24, 128, 128)


def test_module_forward_matches_oracle(afb):
    """AffineTransformModule drop-in (LocalizationNet stubbed by a fixed MLP-head output)."""
    case = cases.atm_case(32, 2, 1, seed=47, zoom_clip=0.2)

    class Stub(torch.nn.Module):
        def __init__(self, p):
            super().__init__(); self.p = torch.nn.Parameter(p)
        def forward(self, x):
            return self.p
    atm = afb.AffineTransformModule(8, case["volume_fov_mm"], case["volume_fov_vox"], case["slice_fov_mm"], case["slice_fov_vox"],
                                    optim_method="R6-vector", offset_clip_value=0.2, zoom_clip_value=0.2, view_id="p2CH",
                                    localization_net=Stub(case["params"][0].clone())).cuda()
    assert atm.vox_range == case["R"]
    ys, yl, yi, ga, nii = atm(case["soft"].cuda(), case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), case["gpre"][0].cuda())
    p = case["params"][0].clone().requires_grad_(True)
    theta = O.view_theta(p, INIT[:, :6], INIT[0, 6:9], INIT[:, 9:], 0.2, 0.2, 32)
    rs, rl, ri, rga, rn = O.atm_tail_forward(case["soft"], case["label"], case["image"], case["nii"], case["gpre"][0], theta,
                                             case["slice_fov_mm"], case["slice_fov_vox"])
    close(ga, rga, 2e-6); close(ys, rs, 2e-5); close(yi, ri, 1e-4)
    assert ys.shape == rs.shape and yl.shape == rl.shape and yl.dtype == rl.dtype
    go = cases.pattern(ys.shape, 1.0)
    (ys * go.cuda()).sum().backward(); (rs * go).sum().backward()
    close(atm.localization_net.p.grad, p.grad, GRAD_REL)
    # theta_override path (non-differentiable theta, reference :259-260)
    ys2, _, _, ga2, _ = atm(case["soft"].cuda(), None, None, case["nii"].cuda(), case["gpre"][0].cuda(), theta_override=theta.detach().cuda())
    close(ga2, rga, 2e-6); close(ys2, rs, 2e-5)


@pytest.mark.parametrize("method,init_ap", [("angle-axis", None), ("normal-vector", [0.2, -0.1, 1.0])])
def test_module_other_parameterisations_golden(afb, golden_dir, method, init_ap):
    """optim_method 'angle-axis' / 'normal-vector' through the drop-in module (closed-form rotation with
    torch ops on the device -> CUDA sampler via the differentiable pre-affine) against the reference's own outputs."""
    g = np.load(os.path.join(golden_dir, "atm_s32_" + method.replace("-", "_") + ".npz"))
    case = cases.atm_case(32, 2, 2, seed=71)

    class Stub(torch.nn.Module):
        def __init__(self, p):
            super().__init__(); self.p = torch.nn.Parameter(p)
        def forward(self, x):
            return self.p
    for v in range(2):
        atm = afb.AffineTransformModule(8, case["volume_fov_mm"], case["volume_fov_vox"], case["slice_fov_mm"], case["slice_fov_vox"],
                                        optim_method=method, offset_clip_value=case["offset_clip"], zoom_clip_value=0.0, view_id="p2CH",
                                        localization_net=Stub(torch.from_numpy(g["params"][v]).clone())).cuda()
        assert atm.ap_space == 3 and atm.vox_range == case["R"]
        if init_ap is not None:
            atm.set_init_theta_ap(torch.tensor(init_ap).cuda())
        ys, yl, yi, ga, nii = atm(case["soft"].cuda(), case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), case["gpre"][v].cuda())
        close(atm.last_theta, g[f"theta{v}"], 2e-6, "theta")
        close(ga, g[f"ga{v}"], 2e-6, "grid affine")
        close(nii, g[f"na{v}"], 1e-6, "nii affine")
        close(ys, g[f"ys{v}"], 2e-5, "soft slice")
        close(yi, g[f"yi{v}"], 1e-4, "image slice")
        mism = (yl.cpu().numpy().astype(np.uint8) != g[f"yl{v}"]).mean()
        assert mism < 2e-3, mism                      # only pixels within an ulp of a rounding tie may flip (see DESIGN)
        ((ys * cases.pattern(ys.shape, 1.0 + v).cuda()).sum() + (ga * cases.pattern(ga.shape, 2.0 + v).cuda()).sum()).backward()
        close(atm.localization_net.p.grad, g[f"dparams{v}"], GRAD_REL, "dparams")


def test_module_rotate_slice_to_min_principle_golden(afb, golden_dir):
    """AffineTransformModule(rotate_slice_to_min_principle=True) against the reference's own outputs
    (aligned soft / label / image slices, composed grid affine, NIfTI affine, parameter and volume gradients through BOTH
    resamplings)."""
    g = np.load(os.path.join(golden_dir, "atm_s32_rotate.npz"))
    case = cases.atm_case(32, 2, 2, seed=81)

    class Stub(torch.nn.Module):
        def __init__(self, p):
            super().__init__(); self.p = torch.nn.Parameter(p)
        def forward(self, x):
            return self.p
    for v in range(2):
        atm = afb.AffineTransformModule(8, case["volume_fov_mm"], case["volume_fov_vox"], case["slice_fov_mm"], case["slice_fov_vox"],
                                        optim_method="R6-vector", offset_clip_value=case["offset_clip"], zoom_clip_value=0.0,
                                        view_id="p2CH", rotate_slice_to_min_principle=True,
                                        localization_net=Stub(case["params"][v].clone())).cuda()
        soft = case["soft"].cuda().requires_grad_(True)
        ys, yl, yi, ga, nii = atm(soft, case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), case["gpre"][v].cuda())
        close(ga, g[f"ga{v}"], 5e-6, "grid affine")
        close(nii, g[f"na{v}"], 1e-5, "nii affine")
        close(ys, g[f"ys{v}"], 1e-4, "soft slice")            # two chained resamplings, alignment affine from fp32 moments
        close(yi, g[f"yi{v}"], 1e-3, "image slice")
        assert (yl.cpu().numpy().astype(np.uint8) != g[f"yl{v}"]).mean() < 5e-3
        ((ys * cases.pattern(ys.shape, 1.0 + v).cuda()).sum() + (ga * cases.pattern(ga.shape, 2.0 + v).cuda()).sum()).backward()
        close(atm.localization_net.p.grad, g[f"dparams{v}"], 5e-4, "dparams")
        close(soft.grad.sum(-1), g[f"dsoft_sum_w{v}"], 5e-4, "dsoft")


@pytest.mark.parametrize("tag,shape", [("embed_s16", (16, 3, 2, 2)), ("embed_s8", (8, 4, 3, 2)), ("embed_s32", (32, 4, 6, 1))])
def test_embed_golden(afb, golden_dir, tag, shape):
    S, c, V, B = shape
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    case = cases.embed_case(S, c, V, B, seed=51 + S)
    x = case["x"].cuda().requires_grad_(True)
    gas = [a.cuda().requires_grad_(True) for a in case["affines"]]
    sc = afb.SkipConnector(V)
    out = sc(x, gas)
    assert out.shape == (B, V * c, S, S, S)
    close(out, g["out"], 2e-5, "embed out")
    (out * cases.pattern(out.shape, 1.0).cuda()).sum().backward()
    close(x.grad, g["dx"], GRAD_REL, "dx")
    close(torch.stack([a.grad for a in gas]), g["d_affines"], 2e-4, "d_affines")


@pytest.mark.parametrize("S,c", [(64, 8), (4, 16), (12, 2), (6, 3)])
def test_embed_vs_oracle_sizes(afb, S, c):
    V, B = 3, 2
    case = cases.embed_case(S, c, V, B, seed=100 + S)
    ref = O.skip_connector(case["x"], case["affines"], V)
    out = afb.SkipConnector(V)(case["x"].cuda(), [a.cuda() for a in case["affines"]])
    close(out, ref, 2e-5)
    assert ((out != 0).float().mean() - (ref != 0).float().mean()).abs().item() < 1e-3


def test_cuda_graph_replay_matches_eager(afb):
    """The whole fwd+bwd step is capturable (stream-ordered, allocation-free kernels) and replays identically."""
    from acquisition_focus_b200.graphs import GraphedStep
    case = cases.atm_case(32, 2, 3, seed=49)
    soft = case["soft"].cuda().requires_grad_(True)
    params = torch.stack(case["params"], dim=1).cuda().requires_grad_(True)
    gpre = torch.stack(case["gpre"], dim=1).cuda()
    label, image, nii = case["label"].cuda(), case["image"].cuda(), case["nii"].cuda()
    init = INIT.repeat(3, 1).cuda()
    go = cases.pattern((2, 3, 8, 32, 32, 1), 1.0).cuda()

    def step():
        soft.grad = None; params.grad = None
        ys, yl, yi, ga, nii_o, th = afb.acquire_views(soft, label, image, nii, gpre, params, init, offset_clip=0.2, zoom_clip=0.0,
                                                      spat=32, slice_fov_mm=case["slice_fov_mm"].tolist(),
                                                      slice_fov_vox=case["slice_fov_vox"].tolist())
        ys.backward(go)
        return ys, yl, yi, ga, soft.grad, params.grad
    eager = [t.detach().clone() for t in step()]
    g = GraphedStep(step)
    for _ in range(2):
        outs = g()
    torch.cuda.synchronize()
    for a, b in zip(outs, eager):
        if a.dtype.is_floating_point:
            close(a, b, 1e-6)
        else:
            assert torch.equal(a, b)
    # new parameter values through the static buffer
    with torch.no_grad():
        params.add_(0.05)
    outs = [t.detach().clone() for t in g()]
    eager = [t.detach().clone() for t in step()]
    close(outs[0], eager[0], 1e-6); close(outs[5], eager[5], 1e-5)


def test_embed_all_stages_one_launch_equals_per_stage(afb, monkeypatch):
    """HybridUnet.forward embeds every encoder skip with the same affines (hybrid_unet.py:40-43): the batched call (one launch
    forward, one backward, d_affines summed over the stages inside the kernel) against stage-by-stage calls, and the
    single-pass forward (zero-then-patch per CTA) against the zero-kernel + slab-kernel pair of the one-stage call (bitwise)."""
    V, B = 3, 2
    stages = [(4, 32), (8, 16), (16, 8), (16, 4), (3, 12)]            # (c, S), incl. a non-multiple-of-4 size
    case0 = cases.embed_case(32, 4, V, B, seed=91)
    gas = [a.cuda().requires_grad_(True) for a in case0["affines"]]
    xs = [cases.randn((B, V * c, S, S), 400 + S).cuda().requires_grad_(True) for c, S in stages]
    sc = afb.SkipConnector(V)
    outs = sc.embed_all(xs, gas)
    gos = [cases.pattern(o.shape, 1.0 + i).cuda() for i, o in enumerate(outs)]
    torch.autograd.backward(outs, gos)
    dx_multi = [x.grad.clone() for x in xs]
    da_multi = torch.stack([a.grad.clone() for a in gas])
    for x in xs:
        x.grad = None
    for a in gas:
        a.grad = None
    da_sum = 0
    for i, x in enumerate(xs):
        o = sc(x, gas)
        assert torch.equal(o, outs[i])
        monkeypatch.setenv("AFB_EMBED_SINGLE_PASS", "1")                  # one stage through the batched kernels
        assert torch.equal(sc(x.detach(), [a.detach() for a in gas]), outs[i])
        monkeypatch.delenv("AFB_EMBED_SINGLE_PASS")
        o.backward(gos[i])
        assert torch.equal(x.grad, dx_multi[i])
    da_sum = torch.stack([a.grad for a in gas])
    close(da_multi, da_sum, 1e-6, "d_affines summed over stages")
    # a stage whose output takes no part in the loss is skipped, its dx is None / zero
    xs2 = [x.detach().clone().requires_grad_(True) for x in xs[:2]]
    o2 = sc.embed_all(xs2, [a.detach() for a in gas])
    o2[1].backward(gos[1])
    assert xs2[0].grad is None and torch.equal(xs2[1].grad, dx_multi[1])


@pytest.mark.parametrize("S", [32, 128])
def test_fused_three_way_forward_bitwise(afb, S):
    """afb_slice_fwd3 (soft + label + image slicings of one acquisition in ONE launch) returns bitwise what the three separate
    launches return, and the backward through it is the same backward."""
    case = cases.atm_case(S, 2, 3, seed=57)
    V = case["V"]
    gpre = torch.stack(case["gpre"], dim=1).cuda()
    res = []
    for fused in (False, True):
        soft = case["soft"].cuda().requires_grad_(True)
        params = torch.stack(case["params"], dim=1).cuda().requires_grad_(True)
        out = afb.acquire_views(soft, case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), gpre, params,
                                INIT.repeat(V, 1).cuda(), offset_clip=0.2, zoom_clip=0.0, spat=S,
                                slice_fov_mm=case["slice_fov_mm"].tolist(), slice_fov_vox=case["slice_fov_vox"].tolist(),
                                fused_forward=fused)
        (out[0] * cases.pattern(out[0].shape, 1.0).cuda()).sum().backward()
        res.append((out, soft.grad, params.grad))
    (a, da_s, da_p), (b, db_s, db_p) = res
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y)
    close(db_p, da_p, 1e-6, "dparams")
    close(db_s, da_s, 1e-6, "dsoft")
    # label-only / image-only combinations
    soft = case["soft"].cuda()
    params = torch.stack(case["params"], dim=1).cuda()
    kw = dict(offset_clip=0.2, zoom_clip=0.0, spat=S, slice_fov_mm=case["slice_fov_mm"].tolist(), slice_fov_vox=case["slice_fov_vox"].tolist())
    o1 = afb.acquire_views(soft, case["label"].cuda(), None, case["nii"].cuda(), gpre, params, INIT.repeat(V, 1).cuda(), fused_forward=True, **kw)
    o2 = afb.acquire_views(soft, None, case["image"].cuda(), case["nii"].cuda(), gpre, params, INIT.repeat(V, 1).cuda(), fused_forward=True, **kw)
    assert torch.equal(o1[0], a[0]) and torch.equal(o1[1], a[1]) and o1[2] is None
    assert torch.equal(o2[0], a[0]) and torch.equal(o2[2], a[2]) and o2[1] is None


@pytest.mark.parametrize("C,label_dtype", [(16, torch.uint8), (4, torch.int32), (8, torch.int16), (8, torch.int64)])
def test_fused_three_way_forward_other_layouts(afb, C, label_dtype):
    """afb_slice_fwd3 with other channel counts / label storage types (16-byte channel vectors: 16 x u8, 8 x i16, 4 x i32,
    2 x i64) and a 2-channel image, against the per-volume launches (bitwise)."""
    S, B, V = 32, 2, 2
    gen = torch.Generator().manual_seed(5)
    lab = torch.randint(0, C, (B, S, S, S), generator=gen)
    onehot = torch.nn.functional.one_hot(lab, C).permute(0, 4, 1, 2, 3)
    soft = (onehot.float() + 0.25 * torch.rand(onehot.shape, generator=gen).permute(0, 1, 2, 3, 4)).cuda()      # channels-last view
    label = onehot.to(label_dtype).cuda()
    image = torch.randn(B, 2, S, S, S, generator=gen).cuda()
    case = cases.atm_case(S, B, V, seed=59)
    gpre = torch.stack(case["gpre"], dim=1).cuda()
    params = torch.stack(case["params"], dim=1).cuda()
    kw = dict(offset_clip=0.2, zoom_clip=0.0, spat=S, slice_fov_mm=case["slice_fov_mm"].tolist(), slice_fov_vox=case["slice_fov_vox"].tolist())
    a = afb.acquire_views(soft, label, image, case["nii"].cuda(), gpre, params, INIT.repeat(V, 1).cuda(), fused_forward=False, **kw)
    b = afb.acquire_views(soft, label, image, case["nii"].cuda(), gpre, params, INIT.repeat(V, 1).cuda(), fused_forward=True, **kw)
    assert b[1].dtype == label_dtype and b[2].shape == (B, V, 2, S, S, 1)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and torch.equal(x, y)
