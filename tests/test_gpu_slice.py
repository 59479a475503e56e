"""Parity of the CUDA slice / volume extraction path (through the C ABI) against the oracle and the
golden vectors minted from the unmodified reference.

Tolerances (north_star): forward within 1e-5 relative (of the output scale) in fp32 - in practice the
kernel reproduces torch-CPU ATen bitwise, which is asserted where it must hold; gradients within 1e-4
of the gradient scale; nearest-neighbour / integer outputs bit-exact.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import af_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu

FWD_REL = 1e-5
GRAD_REL = 1e-4


@pytest.fixture(scope="module")
def afb():
    import acquisition_focus_b200 as m
    assert torch.cuda.is_available()
    return m


def close(a, b, rel):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= rel * scale, f"max abs err {err:.3e} > {rel:g} * scale {scale:.3e}"


SHAPES = [(2, 3, 20, 24, 28, 16, 12, 1), (1, 2, 32, 32, 32, 32, 32, 1), (1, 1, 16, 16, 16, 8, 9, 10),
          (1, 1, 5, 6, 7, 1, 1, 1), (2, 1, 33, 17, 9, 7, 1, 5), (1, 4, 64, 64, 64, 64, 64, 1)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("scale", [0.3, 1.5])
def test_affine_grid_sample_bitwise_vs_aten_cpu(afb, shape, scale):
    N, C, D, H, W, Do, Ho, Wo = shape
    vol = cases.randn((N, C, D, H, W), sum(shape))
    th = torch.eye(3, 4)[None].repeat(N, 1, 1) + scale * cases.randn((N, 3, 4), sum(shape) + 1)
    grid = F.affine_grid(th, [N, C, Do, Ho, Wo], align_corners=False)
    ref = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    out = afb.affine_grid_sample(vol.cuda(), th.cuda(), (Do, Ho, Wo), "bilinear")
    assert out.shape == ref.shape
    assert torch.equal(out.cpu(), ref), f"{(out.cpu() != ref).sum().item()} of {ref.numel()} differ"
    lab = cases.randint(0, 200, (N, C, D, H, W), sum(shape) + 2)
    ref = F.grid_sample(lab.float(), grid, mode="nearest", padding_mode="zeros", align_corners=False).long()
    for dt in (torch.int64, torch.int32, torch.int16, torch.uint8):
        out = afb.affine_grid_sample(lab.to(dt).cuda(), th.cuda(), (Do, Ho, Wo), "nearest")
        assert out.dtype == dt and torch.equal(out.cpu().long(), ref)


STRUCTURED = {
    "identity": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]],                    # coordinates land exactly on voxel centres
    "flip_x": [[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]],
    "swap_xy": [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0]],
    "swap_xz_flip": [[0, 0, -1, 0], [0, 1, 0, 0], [1, 0, 0, 0]],
    "half_voxel_shift": [[1, 0, 0, 1.0 / 32], [0, 1, 0, -1.0 / 32], [0, 0, 1, 0]],   # exactly between two voxels: .5 ties
    "zoom2": [[0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 0.5, 0]],
}


@pytest.mark.parametrize("name", sorted(STRUCTURED))
@pytest.mark.parametrize("sizes", [((32, 32, 32), (32, 32, 32)), ((32, 32, 32), (8, 8, 8)), ((32, 32, 32), (16, 16, 1)),
                                   ((128, 128, 128), (32, 32, 1))])
def test_structured_affines_tie_policy_bitwise_vs_aten_cpu(afb, name, sizes):
    """Tie policy at the sampler boundary.  Axis-aligned views, exact down-sampling (128 -> 32: every coordinate is x.5) and
    half-voxel shifts put sampling coordinates exactly on integers and halves, where `floor` / `nearbyint` (round half to
    even) decide which voxel is read.  Given the same fp32 affine the kernels must take the same decisions as ATen on the
    CPU: outputs are compared BITWISE, bilinear and nearest."""
    (D, H, W), (Do, Ho, Wo) = sizes
    th = torch.tensor(STRUCTURED[name], dtype=torch.float32)[None]
    vol = cases.randn((1, 2, D, H, W), D + Do)
    grid = F.affine_grid(th, [1, 2, Do, Ho, Wo], align_corners=False)
    ref = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    out = afb.affine_grid_sample(vol.cuda(), th.cuda(), (Do, Ho, Wo), "bilinear")
    assert torch.equal(out.cpu(), ref), f"bilinear: {(out.cpu() != ref).sum().item()} of {ref.numel()} differ"
    lab = cases.randint(0, 100, (1, 2, D, H, W), D + Do + 1)
    ref = F.grid_sample(lab.float(), grid, mode="nearest", padding_mode="zeros", align_corners=False).long()
    out = afb.affine_grid_sample(lab.cuda(), th.cuda(), (Do, Ho, Wo), "nearest")
    assert torch.equal(out.cpu(), ref), f"nearest: {(out.cpu() != ref).sum().item()} of {ref.numel()} differ"


@pytest.mark.parametrize("name", ["identity", "half_voxel_shift", "zoom2", "random"])
def test_channels_last_3d_resample_bitwise_vs_aten_cpu(afb, name, monkeypatch):
    """3-D -> 3-D resample of a channels-last fp32 volume with C = 8 (the prescan resample, learnable_transform.py:252-255): takes
    the LDG.256 forward (one 32-byte gather per corner); bitwise ATen on the CPU for the same fp32 affine, and bitwise the
    LDG.128 kernel the slices use (AFB_FWD3D_VB=16)."""
    D = 32
    if name == "random":
        th = (torch.eye(3, 4) + 0.25 * cases.randn((3, 4), 901))[None].float()
    else:
        th = torch.tensor(STRUCTURED[name], dtype=torch.float32)[None]
    vol = cases.randn((2, 8, D, D, D), 900).contiguous(memory_format=torch.channels_last_3d)
    th2 = th.repeat(2, 1, 1)
    ref = F.grid_sample(vol, F.affine_grid(th2, [2, 8, D, D, D], align_corners=False), mode="bilinear", padding_mode="zeros",
                        align_corners=False)
    vc = vol.cuda()
    assert vc.stride(1) == 1
    out = afb.affine_grid_sample(vc, th2.cuda(), (D, D, D), "bilinear")
    assert torch.equal(out.cpu(), ref), f"{(out.cpu() != ref).sum().item()} of {ref.numel()} differ"
    monkeypatch.setenv("AFB_FWD3D_VB", "16")
    out16 = afb.affine_grid_sample(vc, th2.cuda(), (D, D, D), "bilinear")
    assert torch.equal(out16, out)


@pytest.mark.parametrize("shape", SHAPES[:3] + SHAPES[5:])
def test_affine_grid_sample_backward(afb, shape):
    N, C, D, H, W, Do, Ho, Wo = shape
    vol = cases.randn((N, C, D, H, W), sum(shape) + 5).requires_grad_(True)
    th = (torch.eye(3, 4)[None].repeat(N, 1, 1) + 0.3 * cases.randn((N, 3, 4), sum(shape) + 6)).requires_grad_(True)
    ref = F.grid_sample(vol, F.affine_grid(th, [N, C, Do, Ho, Wo], align_corners=False), mode="bilinear",
                        padding_mode="zeros", align_corners=False)
    go = cases.pattern(ref.shape, 1.0)
    ref.backward(go)
    v2 = vol.detach().cuda().requires_grad_(True); t2 = th.detach().cuda().requires_grad_(True)
    out = afb.affine_grid_sample(v2, t2, (Do, Ho, Wo), "bilinear")
    out.backward(go.cuda())
    close(v2.grad, vol.grad, GRAD_REL)
    close(t2.grad, th.grad, GRAD_REL)


@pytest.mark.parametrize("tag,kw", [
    ("slice", dict(target_fov_mm=torch.tensor([30.0, 20.0, 1.5]), target_fov_vox=torch.tensor([16, 12, 1]))),
    ("vol3d", dict(target_fov_mm=torch.tensor([28.0, 30.0, 33.0]), target_fov_vox=torch.tensor([9, 10, 11]))),
    ("same", dict())])
def test_nifti_grid_sample_golden(afb, golden_dir, tag, kw):
    g = np.load(os.path.join(golden_dir, "slice_small.npz"))
    B, C, D, H, W = 2, 3, 20, 24, 28
    vol = cases.randn((B, C, D, H, W), 21).cuda().requires_grad_(True)
    lab = cases.randint(0, 6, (B, C, D, H, W), 22).cuda()
    nii = torch.from_numpy(g["nii"]).cuda()
    P = torch.from_numpy(g["P"]).cuda().requires_grad_(True)
    y, ga, na = afb.nifti_grid_sample(vol, nii, is_label=False, pre_grid_sample_affine=P, **kw)
    assert y.dtype == torch.float32 and ga.dtype == torch.float32 and na.dtype == torch.float64
    close(ga, torch.from_numpy(g[f"{tag}_ga"]), 1e-6)
    close(y, torch.from_numpy(g[f"{tag}_y"]), FWD_REL)
    assert np.allclose(na.cpu().numpy(), g[f"{tag}_nii"], rtol=1e-9, atol=1e-9)
    ((y * cases.pattern(y.shape, 1.0).cuda()).sum() + (ga * cases.pattern(ga.shape, 2.0).cuda()).sum()).backward()
    close(vol.grad, torch.from_numpy(g[f"{tag}_dvol"]), GRAD_REL)
    close(P.grad, torch.from_numpy(g[f"{tag}_dP"]), GRAD_REL)
    yl, gl, _ = afb.nifti_grid_sample(lab, nii, is_label=True, pre_grid_sample_affine=P.detach(), **kw)
    assert yl.dtype == torch.int64
    if torch.equal(ga.detach().cpu(), torch.from_numpy(g[f"{tag}_ga"])):
        assert torch.equal(yl.cpu(), torch.from_numpy(g[f"{tag}_ylabel"]))
        assert torch.equal(y.detach().cpu(), torch.from_numpy(g[f"{tag}_y"]))
    else:  # grid affine differs in the last bit (fp64 rounding of the reference's inverse): near-tie pixels may flip
        mism = (yl.cpu() != torch.from_numpy(g[f"{tag}_ylabel"])).float().mean().item()
        assert mism < 2e-3, mism


def test_nifti_grid_sample_label_bit_exact_given_same_grid_affine(afb):
    """Bit-exactness is defined at the sampler boundary: same fp32 grid affine in -> same labels out."""
    B, C, S = 2, 1, 48
    lab = cases.randint(0, 8, (B, C, S, S, S), 77)
    nii = cases.synthetic.default_nifti_affine(B, 1.5)
    P = cases.random_pre_affine(B, 78, 0.25)
    kw = dict(target_fov_mm=torch.tensor([60.0, 66.0, 1.5]), target_fov_vox=torch.tensor([40, 44, 1]))
    yl, ga, _ = afb.nifti_grid_sample(lab.cuda(), nii.cuda(), is_label=True, pre_grid_sample_affine=P.cuda(), **kw)
    ref, ga_ref, _ = O.nifti_grid_sample(lab, nii, is_label=True, pre_grid_sample_affine=P, **kw)
    assert torch.equal(ga.cpu(), ga_ref)
    assert torch.equal(yl.cpu(), ref)


def test_slice_cfg1_128_golden(afb, golden_dir):
    """configs[0]: one 128^3 fp32 image volume, one p2CH view, batch 1, R6 -> slice, fwd + bwd."""
    g = np.load(os.path.join(golden_dir, "slice_cfg1_128.npz"))
    syn = cases.synthetic
    vol = torch.from_numpy(syn.phantom_image(syn.heart_phantom(128), seed=5))[None, None].cuda().requires_grad_(True)
    r6 = torch.from_numpy(g["r6"]).cuda().requires_grad_(True)
    P = torch.from_numpy(g["gpre"]).cuda() @ afb.compute_rotation_matrix_from_ortho6d(r6)
    y, ga, na = afb.nifti_grid_sample(vol, syn.default_nifti_affine(1).cuda(), target_fov_mm=torch.tensor([192.0, 192.0, 1.5]),
                                      target_fov_vox=torch.tensor([128, 128, 1]), pre_grid_sample_affine=P)
    close(ga, torch.from_numpy(g["ga"]), 1e-6)
    close(y, torch.from_numpy(g["y"]), FWD_REL)
    (y * cases.pattern(y.shape, 1.0).cuda()).sum().backward()
    close(r6.grad, torch.from_numpy(g["d_r6"]), GRAD_REL)
    dv = vol.grad[0, 0]
    close(dv.sum(0), torch.from_numpy(g["dvol_sum_d"]), GRAD_REL)
    close(dv.sum(1), torch.from_numpy(g["dvol_sum_h"]), GRAD_REL)
    close(dv.sum(2), torch.from_numpy(g["dvol_sum_w"]), GRAD_REL)


def test_r6_golden(afb, golden_dir):
    g = np.load(os.path.join(golden_dir, "r6.npz"))
    o = torch.from_numpy(g["ortho"]).cuda().requires_grad_(True)
    m = afb.compute_rotation_matrix_from_ortho6d(o)
    close(m, torch.from_numpy(g["mat"]), 1e-6)
    (m * cases.pattern(m.shape, 1.0).cuda()).sum().backward()
    close(o.grad, torch.from_numpy(g["d_ortho"]), GRAD_REL)


@pytest.mark.parametrize("dt,rel", [(torch.bfloat16, 2.0 ** -8), (torch.float16, 2.0 ** -10)])
def test_half_storage(afb, dt, rel):
    """bf16/fp16 storage, fp32 coordinates/weights/accumulation.  The reference has no usable bf16 path
    (SURVEY 8d): oracle = reference fp32 path on the rounded volume, compared after rounding."""
    B, C, S = 1, 2, 40
    vol = cases.randn((B, C, S, S, S), 91).to(dt)
    nii = cases.synthetic.default_nifti_affine(B, 1.5)
    P = cases.random_pre_affine(B, 92, 0.2)
    kw = dict(target_fov_mm=torch.tensor([50.0, 50.0, 1.5]), target_fov_vox=torch.tensor([36, 36, 1]))
    ref, _, _ = O.nifti_grid_sample(vol.float(), nii, pre_grid_sample_affine=P, **kw)
    out, ga, _ = afb.nifti_grid_sample(vol.cuda(), nii.cuda(), pre_grid_sample_affine=P.cuda(), **kw)
    assert out.dtype == dt
    close(out.float(), ref.to(dt).float(), rel)
    v = vol.cuda().requires_grad_(True)
    out, _, _ = afb.nifti_grid_sample(v, nii.cuda(), pre_grid_sample_affine=P.cuda(), **kw)
    out.float().sum().backward()
    vr = vol.float().requires_grad_(True)
    O.nifti_grid_sample(vr, nii, pre_grid_sample_affine=P, **kw)[0].sum().backward()
    close(v.grad.float(), vr.grad, 1e-2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_channels_last_kernels_equal_generic_kernels(afb, dt):
    """The 16-byte vector (channels-last) kernels and the generic strided kernels run the same per-channel arithmetic:
    forward bitwise equal, gradients equal up to atomic ordering; for every floating storage type."""
    B, S, C = 2, 28, 8
    base = cases.randn((B, S, S, S, C), 57).to(dt)
    cl = base.permute(0, 4, 1, 2, 3).cuda()                   # channels-last view  (sC == 1)
    pl = cl.contiguous()                                       # planar copy          (generic kernel)
    assert cl.stride(1) == 1 and pl.stride(1) != 1
    nii = cases.synthetic.default_nifti_affine(B, 2.0).cuda()
    P = cases.random_pre_affine(B, 58, 0.25).cuda()
    kw = dict(target_fov_mm=torch.tensor([56.0, 56.0, 2.0]), target_fov_vox=torch.tensor([28, 28, 1]))
    outs, grads = [], []
    for v in (cl, pl):
        v = v.detach().requires_grad_(True)
        p = P.clone().requires_grad_(True)
        y, ga, _ = afb.nifti_grid_sample(v, nii, pre_grid_sample_affine=p, **kw)
        (y.float() * cases.pattern(y.shape, 1.0).cuda()).sum().backward()
        outs.append(y.detach()); grads.append((v.grad.float().contiguous(), p.grad))
    assert outs[0].dtype == dt and torch.equal(outs[0], outs[1])
    close(grads[0][0], grads[1][0], 2e-2 if dt != torch.float32 else 1e-5)      # dVolume is rounded to the storage type
    close(grads[0][1], grads[1][1], 1e-5)
    ref, _, _ = O.nifti_grid_sample(pl.float().cpu(), nii.cpu(), pre_grid_sample_affine=P.cpu(), **kw)
    close(outs[0].float(), ref.to(dt).float(), {torch.float32: 1e-6, torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -10}[dt])


def test_channels_last_volume_view(afb):
    """running/run_dl.py:261-264 hands the sampler a channels-last *view* (one_hot + rearrange)."""
    B, S, C = 2, 24, 8
    lab = cases.randint(0, C, (B, S, S, S), 55)
    label_oh, soft = cases.one_hot_volumes(lab, C)
    assert not soft.is_contiguous()
    nii = cases.synthetic.default_nifti_affine(B, 2.0)
    P = cases.random_pre_affine(B, 56, 0.2)
    kw = dict(target_fov_mm=torch.tensor([48.0, 48.0, 2.0]), target_fov_vox=torch.tensor([24, 24, 1]))
    ref, ga_ref, _ = O.nifti_grid_sample(soft, nii, pre_grid_sample_affine=P, **kw)
    s_cuda = soft.cuda()
    assert s_cuda.stride() == soft.stride()
    out, ga, _ = afb.nifti_grid_sample(s_cuda, nii.cuda(), pre_grid_sample_affine=P.cuda(), **kw)
    assert torch.equal(ga.cpu(), ga_ref) and torch.equal(out.cpu(), ref)
    refl, _, _ = O.nifti_grid_sample(label_oh, nii, is_label=True, pre_grid_sample_affine=P, **kw)
    outl, _, _ = afb.nifti_grid_sample(label_oh.cuda(), nii.cuda(), is_label=True, pre_grid_sample_affine=P.cuda(), **kw)
    assert torch.equal(outl.cpu(), refl)
    # argmax over the bilinear one-hot slice equals the oracle's (bit-exact values => bit-exact argmax)
    assert torch.equal(out.argmax(1).cpu(), ref.argmax(1))


def test_min_shift_semantics_and_min_grad(afb):
    """Out-of-field samples evaluate to volume.min(); MinBackward spreads evenly over ALL minima."""
    B, C, S = 1, 2, 16
    vol = cases.randn((B, C, S, S, S), 61)
    vol[0, 0, :2] = vol.min() - 1.0        # many equal minima
    nii = cases.synthetic.default_nifti_affine(B, 1.0)
    P = torch.eye(4)[None].clone(); P[0, 0, 3] = 0.9      # shift far out of the field of view
    kw = dict(target_fov_mm=torch.tensor([16.0, 16.0, 1.0]), target_fov_vox=torch.tensor([16, 16, 1]))
    vr = vol.clone().requires_grad_(True)
    ref, _, _ = O.nifti_grid_sample(vr, nii, pre_grid_sample_affine=P, **kw)
    ref.sum().backward()
    vc = vol.cuda().requires_grad_(True)
    out, _, _ = afb.nifti_grid_sample(vc, nii.cuda(), pre_grid_sample_affine=P.cuda(), **kw)
    assert torch.equal(out.cpu(), ref.detach())
    out.sum().backward()
    close(vc.grad, vr.grad, GRAD_REL)
    mc = afb.volume_min(vol.cuda()).cpu()
    assert mc[0].item() == vol.min().item() and mc[1].item() == float((vol == vol.min()).sum())


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n", [1, 7, 4096, 4097, 3 * 4096 + 1023, 1_000_003])
def test_min_mask_record_equals_volume_reread(afb, n, dt):
    """afb_volume_min_mask[_half] + afb_min_grad_fill_mask[_half] (1-bit record, no volume re-read) == afb_volume_min +
    afb_min_grad_fill, for fp32 and for 16-bit storage."""
    import ctypes as C
    from acquisition_focus_b200 import _lib as L
    from acquisition_focus_b200 import functional as AF
    lib = L.lib()
    g = torch.Generator().manual_seed(n)
    vol = torch.randint(0, 5, (n,), generator=g).float() * 0.5 - 1.0        # many ties at the minimum, spread over chunks
    vol[torch.randint(0, n, (max(1, n // 50),), generator=g)] = -3.25
    vol = vol.to(dt)                                                        # all values are exact in bf16 / fp16
    v = vol.cuda()
    plain, masked = AF.volume_min(v), AF.volume_min(v, with_mask=True)
    assert torch.equal(plain, masked) and hasattr(masked, "_afb_mask")
    assert plain[0].item() == vol.min().item() and plain[1].item() == float((vol == vol.min()).sum())
    d_pad = torch.tensor([2.5], device="cuda")
    a = torch.full((n,), 7.0, device="cuda"); b = torch.full((n,), 7.0, device="cuda")
    st = L.stream_ptr(v.device)
    L.check(lib.afb_min_grad_fill(L.ptr(v), L.DTYPES[dt], n, L.ptr(plain), L.ptr(d_pad), L.ptr(a), st), "fill")
    fill = lib.afb_min_grad_fill_mask if dt == torch.float32 else lib.afb_min_grad_fill_mask_half
    L.check(fill(L.ptr(masked._afb_mask), n, L.ptr(masked), L.ptr(d_pad), L.ptr(b), st), "fill_mask")
    assert torch.equal(a, b)
    assert torch.equal(AF.min_count_from_record(masked._afb_mask, n), plain)
    want = (vol == vol.min()).float() * (2.5 / float((vol == vol.min()).sum()))
    assert torch.allclose(a.cpu(), want, rtol=1e-6, atol=0)


def test_errors(afb):
    from acquisition_focus_b200._lib import AfbError
    vol = torch.zeros(1, 1, 4, 4, 4)
    nii = torch.eye(4)[None].double()
    with pytest.raises(AfbError):
        afb.nifti_grid_sample(vol, nii)                                  # CPU tensor: no fallback
    with pytest.raises(Warning):
        afb.nifti_grid_sample(vol.cuda(), nii.cuda(), ras_transform_affine=torch.eye(4)[None])
    with pytest.raises(AssertionError):
        afb.nifti_grid_sample(vol[0].cuda(), nii.cuda())
    with pytest.raises(AssertionError):
        afb.nifti_grid_sample(vol.cuda(), nii.cuda(), pre_grid_sample_affine=torch.eye(4)[None].repeat(2, 1, 1).cuda())


def test_identity_resample_and_linearity_fullsize(afb):
    """Size-independent properties at the full 128^3 size."""
    vol = cases.randn((1, 2, 128, 128, 128), 71).cuda()
    nii = cases.synthetic.default_nifti_affine(1).cuda()
    out, ga, na = afb.nifti_grid_sample(vol, nii)
    close(out, vol, 1e-5)
    lab = cases.randint(0, 8, (1, 1, 128, 128, 128), 72).cuda()
    outl, _, _ = afb.nifti_grid_sample(lab, nii, is_label=True)
    assert torch.equal(outl, lab)
    # the reference's bookkeeping moves the origin by half a voxel even for a no-op resample; keep that quirk
    want = O.nifti_grid_sample(torch.zeros(1, 1, 128, 128, 128), nii.cpu())[2]
    assert np.allclose(na.cpu().numpy(), want.numpy(), atol=1e-9)
    th = (torch.eye(3, 4)[None] + 0.2 * cases.randn((1, 3, 4), 73)).cuda()
    a = afb.affine_grid_sample(vol[:, :1], th, (128, 128, 1)); b = afb.affine_grid_sample(vol[:, 1:], th, (128, 128, 1))
    c = afb.affine_grid_sample(2.0 * vol[:, :1] + vol[:, 1:], th, (128, 128, 1))
    close(c, 2.0 * a + b, 1e-5)


def test_edge_shapes_vs_oracle(afb):
    """Edge cases of the sampler boundary: a one-voxel-thin input (the W_in = 1 volumes of rotate_slice_to_min_principle,
    learnable_transform.py:337-366), a single output location, and a slice lying completely outside the volume
    (every corner masked: the output is the pad value, volume.min(), and MinBackward receives the whole gradient)."""
    B = 2
    nii = cases.synthetic.default_nifti_affine(B, 1.5)
    P = cases.random_pre_affine(B, 33, 0.15)
    # (1) thin input, 2-D -> 2-D resample, bilinear and nearest
    thin = cases.randn((B, 3, 16, 20, 1), 31)
    kw = dict(target_fov_mm=torch.tensor([20.0, 26.0, 1.5]), target_fov_vox=torch.tensor([12, 14, 1]))
    ref = O.nifti_grid_sample(thin, nii, pre_grid_sample_affine=P, **kw)
    out = afb.nifti_grid_sample(thin.cuda(), nii.cuda(), pre_grid_sample_affine=P.cuda(), **kw)
    close(out[1], ref[1], 2e-6); close(out[0], ref[0], 2e-5); close(out[2], ref[2], 1e-9)
    lab = cases.randint(0, 5, (B, 1, 16, 20, 1), 32)
    refl = O.nifti_grid_sample(lab, nii, is_label=True, pre_grid_sample_affine=P, **kw)[0]
    outl = afb.nifti_grid_sample(lab.cuda(), nii.cuda(), is_label=True, pre_grid_sample_affine=P.cuda(), **kw)[0]
    assert outl.dtype == torch.int64 and (outl.cpu() != refl).float().mean().item() < 5e-3
    # (2) one output location
    vol = cases.randn((B, 2, 9, 10, 11), 34)
    kw1 = dict(target_fov_mm=torch.tensor([1.5, 1.5, 1.5]), target_fov_vox=torch.tensor([1, 1, 1]))
    close(afb.nifti_grid_sample(vol.cuda(), nii.cuda(), pre_grid_sample_affine=P.cuda(), **kw1)[0],
          O.nifti_grid_sample(vol, nii, pre_grid_sample_affine=P, **kw1)[0], 2e-5)
    # (3) slice shifted far outside the field of view
    far = torch.eye(4)[None].repeat(B, 1, 1)
    far[:, :3, 3] = torch.tensor([5.0, -4.0, 6.0])
    kw2 = dict(target_fov_mm=torch.tensor([9.0, 9.0, 1.5]), target_fov_vox=torch.tensor([8, 8, 1]))
    v1 = vol.clone().cuda().requires_grad_(True)
    o = afb.nifti_grid_sample(v1, nii.cuda(), pre_grid_sample_affine=far.cuda(), **kw2)[0]
    assert torch.equal(o, torch.full_like(o, vol.min().item()))
    o.sum().backward()
    v2 = vol.clone().requires_grad_(True)
    O.nifti_grid_sample(v2, nii, pre_grid_sample_affine=far, **kw2)[0].sum().backward()
    close(v1.grad, v2.grad, 1e-5)
    assert v1.grad.flatten()[vol.flatten().argmin()].item() == float(o.numel())
