"""The small callers either side of the samplers as kernels: rotation parameterisations against the
reference-minted golden, slice up-sampling against ATen's upsample_trilinear3d (torch CPU), the clinical composition against
the reference's fp64 torch expression."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import af_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("tag,fn", [("aa", "angle_axis_to_rotation_matrix"), ("nv", "normal_to_rotation_matrix")])
def test_rotation_parameterisation_kernels_golden(golden_dir, tag, fn):
    """f4: afb_rot3_fwd/bwd vs utils/transform_utils.py:62-178 (golden minted from the reference; includes r = 0 and |r|^2 < eps)."""
    from acquisition_focus_b200.utils import transform_utils as T
    g = np.load(os.path.join(golden_dir, "rotation_params.npz"))
    x = torch.from_numpy(g[f"{tag}_in"]).cuda().requires_grad_(True)
    m = getattr(T, fn)(x)
    assert m.shape == (x.shape[0], 4, 4)
    e_f = _rel(m, g[f"{tag}_mat"])
    (m * cases.pattern(m.shape, 1.0).cuda()).sum().backward()
    e_g = _rel(x.grad, g[f"{tag}_grad"])
    print(f"{fn}: fwd {e_f:.2e} grad {e_g:.2e}")
    assert e_f <= 1e-6 and e_g <= 1e-5


@pytest.mark.parametrize("shape,size", [((2, 3, 8, 16, 16, 1), (32, 32)), ((4, 8, 32, 32, 1), (128, 128)), ((1, 2, 5, 7, 1), (16, 12)),
                                        ((2, 2, 16, 16, 1), (16, 16)), ((1, 1, 12, 10, 1), (30, 25))])
def test_upsample_slices_vs_aten(shape, size):
    """f3: F.interpolate(x, size=[H,W,1], mode='trilinear', align_corners=False) (running/run_dl.py:193-197) against ATen-CPU
    (1e-6 of scale: ATen's vectorised CPU kernel contracts / associates the three lerps differently in the last bit);
    backward against autograd of the same ATen op."""
    from acquisition_focus_b200 import functional as AF
    x = cases.randn(shape, 601)
    tgt = list(size) + [1]
    xr = x.clone().requires_grad_(True)
    ref = F.interpolate(xr.flatten(0, -5) if xr.dim() > 5 else xr, size=tgt, mode="trilinear", align_corners=False)
    xg = x.cuda().requires_grad_(True)
    out = AF.upsample_slices(xg, tgt)
    assert tuple(out.shape) == tuple(shape[:-3]) + tuple(tgt)
    go = cases.pattern(ref.shape, 1.0)
    (ref * go).sum().backward()
    (out * go.view(out.shape).cuda()).sum().backward()
    e_f, e_g = _rel(out.detach().cpu().view(ref.shape), ref), _rel(xg.grad.cpu(), xr.grad)
    print(f"upsample {shape}->{size}: fwd {e_f:.2e} grad {e_g:.2e}")
    assert e_f <= 1e-6 and e_g <= 1e-5


def test_compose_pre_affine_vs_reference_expression():
    """a6: Gpre = base^-1 @ view @ aug (running/run_dl.py:227-234 + :208-223) in one kernel vs the reference's fp64 torch chain."""
    from acquisition_focus_b200 import functional as AF
    from acquisition_focus_b200.utils.transform_utils import get_random_affine
    B = 5
    gen = torch.Generator().manual_seed(9)
    base = torch.stack([cases.synthetic.random_aug_affine(gen, 0.4, 0.3, 0.05) for _ in range(B)]).double()
    base[1, :3, :3] = base[1, :3, :3].flip(0)                      # needs pivoting
    view = cases.synthetic.phantom_view_affines()["p4CH"][None].repeat(B, 1, 1)
    torch.manual_seed(3)
    aug = torch.stack([get_random_affine(0.1, 0.2, 0.0) for _ in range(B)])
    want = (O.input_affine_for_view(base, view) @ aug.to(base)).float()
    got = AF.compose_pre_affine(base.cuda(), view.cuda(), aug.cuda())
    assert got.dtype == torch.float32 and _rel(got, want) <= 2e-7
    got2 = AF.compose_pre_affine(base.cuda(), view.double().cuda(), None)
    assert _rel(got2, O.input_affine_for_view(base, view).float()) <= 2e-7


def test_clinical_cardiac_view_affines_on_gpu(monkeypatch):
    """f4: get_clinical_cardiac_view_affines (functional/clinical_cardiac_views.py:223-364) with its voxel passes as CUDA kernels
    (group moments, extent bisection, nearest slices through the sampler) against the affines the REFERENCE minted for the same
    phantom (data/phantom_view_affines.json), and - other size, 5 SA slices, uint8 / int64 label maps - against the same host
    logic driven by numpy / oracle restatements of the three device passes."""
    from acquisition_focus_b200 import clinical_cardiac_views as CV
    from oracle.clinical_np import cpu_extent as _cpu_extent, cpu_moments as _cpu_moments
    syn = cases.synthetic
    lab = torch.from_numpy(syn.heart_phantom(128))
    nii = torch.diag(torch.tensor([1.5, 1.5, 1.5, 1.0]))
    got = CV.get_clinical_cardiac_view_affines(lab.cuda(), nii, syn.CLASS_DICT, num_sa_slices=3, return_unrolled=True)
    want = syn.phantom_view_affines()
    assert list(got.keys()) == list(want.keys())
    worst = max((got[k] - want[k]).abs().max().item() for k in want)
    print(f"clinical views vs reference-minted golden: max abs diff {worst:.2e}")
    assert worst <= 2e-5
    # the device passes themselves against their restatements
    lab64 = torch.from_numpy(np.ascontiguousarray(np.roll(syn.heart_phantom(64), 3, axis=1)))
    masks = [CV._group_mask((1, 3)), CV._group_mask((1, 2, 3)), CV._group_mask(syn.CLASS_DICT.values())]
    for dt in (torch.uint8, torch.int64):
        c1, m1, i1 = CV._moments(lab64.to(dt).cuda(), masks)
        c2, m2, i2 = _cpu_moments(lab64, masks)
        assert np.array_equal(c1, c2) and torch.equal(m1, m2) and torch.equal(i1, i2)          # exact integer moments
    d = torch.tensor([0.48, -0.6, 0.64])
    d = d / d.norm()
    p1, p2 = CV._extent_along_axis(lab64.to(torch.uint8).cuda(), masks[0], m1[0], d)
    q1, q2 = _cpu_extent(lab64, masks[0], m1[0], d)
    assert torch.allclose(p1, q1, atol=1e-6) and torch.allclose(p2, q2, atol=1e-6)
    nii64 = torch.diag(torch.tensor([3.0, 3.0, 3.0, 1.0]))
    g5 = CV.get_clinical_cardiac_view_affines(lab64.cuda(), nii64, syn.CLASS_DICT, num_sa_slices=5)
    monkeypatch.setattr(CV, "_moments", _cpu_moments)
    monkeypatch.setattr(CV, "_extent_along_axis", _cpu_extent)
    monkeypatch.setattr(CV, "nifti_grid_sample", O.nifti_grid_sample)
    import acquisition_focus_b200._lib as L
    monkeypatch.setattr(L, "require_cuda", lambda *a, **k: None)
    w5 = CV.get_clinical_cardiac_view_affines(lab64, nii64, syn.CLASS_DICT, num_sa_slices=5)
    for k in w5:
        a, b = (torch.stack(g5[k]), torch.stack(w5[k])) if k == "ALL_SA" else (g5[k], w5[k])
        assert (a - b).abs().max().item() <= 1e-6, k
