"""Measured parity of the RAW-PARAMETER path (MLP-head output -> slices), the inputs the reference actually takes.

north_star asks for 1e-5 (forward), 1e-4 (gradients) and bit-exact nearest / argmax labels.  At the sampler boundary (same
fp32 grid affine) the kernels are bitwise equal to ATen (tests/test_gpu_slice.py).  Upstream of it the fp32 parameter chain
(softmax, norms, 4x4 products) is evaluated by a different program than torch's, so G' can differ in the last bits; this
file MEASURES what that does, three ways, and asserts against the measurements instead of loose constants:

* ours (CUDA kernels)                    vs the oracle port on CPU (pinned bitwise to the unmodified reference);
* the reference's own op sequence on CUDA (the same oracle code with its tensors on the GPU = ATen's sm_100 kernels)
                                          vs the same code on CPU: how far the reference is from ITSELF across devices;
* every label pixel that differs is proven to sit on a rounding tie: its fp64 coordinate is within delta of k+0.5 on some
  axis (delta derived from the measured G' difference) and our value is the label of the other tie candidate.

Thresholds live in tests/golden/parity_thresholds.json (= 2x the errors measured on the B200, profiles/r2_parity_measured.json);
every run rewrites gpurun_out/parity_measured.json with the current measurements.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import af_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INIT = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]])
_THRESH_PATH = os.path.join(ROOT, "tests", "golden", "parity_thresholds.json")
_RECORD = {}


def _thresholds():
    with open(_THRESH_PATH) as fh:
        return json.load(fh)


def _record(tag, d):
    _RECORD[tag] = d
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_measured.json")
    prev = {}
    if os.path.exists(path):
        try:
            prev = json.load(open(path))
        except Exception:
            prev = {}
    prev.update(_RECORD)
    with open(path, "w") as fh:
        json.dump(prev, fh, indent=1, sort_keys=True)


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _oracle_run(case, device, with_dvol):
    """The reference's op sequence (oracle port) on `device`: per view slices, G', gradients."""
    V = case["V"]
    init = INIT.to(device)
    nii, label, image = case["nii"].to(device), case["label"].to(device), case["image"].to(device)
    fov_mm, fov_vox = case["slice_fov_mm"].to(device), case["slice_fov_vox"].to(device)
    res = []
    dsoft = None
    for v in range(V):
        p = case["params"][v].clone().to(device).requires_grad_(True)
        s = case["soft"].clone().to(device).requires_grad_(with_dvol)
        th = O.view_theta(p, init[:1, :6], init[0, 6:9], init[:1, 9:], case["offset_clip"], case["zoom_clip"], case["S"])
        ys, yl, yi, ga, nii_o = O.atm_tail_forward(s, label, image, nii, case["gpre"][v].to(device), th, fov_mm, fov_vox)
        ((ys * cases.pattern(ys.shape, 1.0 + v).to(device)).sum() + (ga * cases.pattern(ga.shape, 2.0 + v).to(device)).sum()).backward()
        if with_dvol:
            dsoft = s.grad.cpu() if dsoft is None else dsoft + s.grad.cpu()
        res.append(dict(ys=ys.detach().cpu(), yl=yl.cpu(), yi=yi.cpu(), ga=ga.detach().cpu(), dparams=p.grad.cpu(), theta=th.detach().cpu()))
    return res, dsoft


def _ours_run(afb, case, with_dvol):
    V = case["V"]
    soft = case["soft"].cuda().requires_grad_(with_dvol)
    params = torch.stack(case["params"], dim=1).cuda().requires_grad_(True)
    gpre = torch.stack(case["gpre"], dim=1).cuda()
    ys, yl, yi, ga, nii, theta = afb.acquire_views(
        soft, case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), gpre, params, INIT.repeat(V, 1).cuda(),
        offset_clip=case["offset_clip"], zoom_clip=case["zoom_clip"], spat=case["S"], slice_fov_mm=case["slice_fov_mm"].tolist(),
        slice_fov_vox=case["slice_fov_vox"].tolist())
    loss = 0
    for v in range(V):
        loss = loss + (ys[:, v] * cases.pattern(ys[:, v].shape, 1.0 + v).cuda()).sum() + (ga[:, v] * cases.pattern(ga[:, v].shape, 2.0 + v).cuda()).sum()
    loss.backward()
    res = [dict(ys=ys[:, v].detach().cpu(), yl=yl[:, v].cpu(), yi=yi[:, v].cpu(), ga=ga[:, v].detach().cpu(),
                dparams=params.grad[:, v].cpu(), theta=theta[:, v].cpu()) for v in range(V)]
    return res, (soft.grad.cpu() if with_dvol else None)


def _coords64(ga, S_out, size_in):
    """fp64 un-normalised source coordinates [B,Do,Ho,3 (x,y,z)] of a 128x128x1 slice for grid affine ga [B,4,4]."""
    B = ga.shape[0]
    Do, Ho = S_out
    z = (2.0 * torch.arange(Do, dtype=torch.float64) + 1.0) / Do - 1.0
    y = (2.0 * torch.arange(Ho, dtype=torch.float64) + 1.0) / Ho - 1.0
    zz, yy = torch.meshgrid(z, y, indexing="ij")
    base = torch.stack([torch.zeros_like(zz), yy, zz, torch.ones_like(zz)], dim=-1)          # x = 0 for Wo = 1
    g = torch.einsum("ijk,brk->bijr", base, ga.double()[:, :3, :])
    return ((g + 1.0) * size_in - 1.0) / 2.0


def _tie_proof(ref, ours, lab_map, S):
    """Every nearest-label pixel that differs from the reference sits on a rounding tie.  Returns (n_mismatch, delta,
    worst distance to a tie among the mismatching pixels, number of pixels within delta of a tie)."""
    B = ref["ga"].shape[0]
    dG = (ours["ga"].double() - ref["ga"].double()).abs()[:, :3, :]
    # |d coord| <= S/2 * (sum_j |dG[r,j]| * |base_j| + |dG[r,3]|), |base_j| <= 1; plus the two fp32 evaluations' own rounding
    # (each <= 4 ulp of the coordinate magnitude S)
    delta_b = (S / 2.0) * dG.sum(-1).max(-1).values + 8.0 * S * 2.0 ** -24                      # [B]
    c = _coords64(ref["ga"], (S, S), float(S))                                                  # reference coordinates
    frac_dist = ((c - torch.floor(c)) - 0.5).abs()                                              # distance to a k+0.5 tie per axis
    near = (frac_dist <= delta_b.view(B, 1, 1, 1)).any(-1)                                      # [B,S,S]
    rl, ol = ref["yl"][..., 0], ours["yl"][..., 0]                                              # [B,C,S,S] one-hot int64
    mism = (rl != ol).any(1)                                                                    # [B,S,S]
    n_mis = int(mism.sum())
    worst = 0.0
    for b, i, j in mism.nonzero().tolist():
        d = frac_dist[b, i, j].min().item()
        worst = max(worst, d)
        assert d <= delta_b[b].item(), f"label pixel {(b, i, j)} differs but is {d:.3e} > delta {delta_b[b].item():.3e} from a tie"
        # our value must be the label at one of the tie candidates (nearest index with each near-tie axis rounded either way)
        cands = [[]]
        for ax in range(3):
            x = c[b, i, j, ax].item()
            opts = {int(np.floor(x + 0.5))}
            if frac_dist[b, i, j, ax].item() <= delta_b[b].item():
                opts |= {int(np.floor(x)), int(np.floor(x)) + 1}
            cands = [cc + [o] for cc in cands for o in sorted(opts)]
        got = ol[b, :, i, j]
        ok = False
        for xn, yn, zn in cands:
            inb = 0 <= xn < S and 0 <= yn < S and 0 <= zn < S
            want = F.one_hot(lab_map[b, zn, yn, xn], got.numel()) if inb else torch.zeros_like(got)
            ok = ok or torch.equal(got, want.to(got.dtype))
        assert ok, f"label pixel {(b, i, j)}: value is not the label of any tie candidate"
    return n_mis, float(delta_b.max()), worst, int(near.sum())


def _argmax_proof(ref_ys, our_ys, err_abs):
    """Pixels whose channel argmax differs must be near-ties of the two largest reference values (gap <= 2 * max err)."""
    ra, oa = ref_ys.argmax(1), our_ys.argmax(1)
    mism = ra != oa
    top2 = ref_ys.topk(2, dim=1).values
    gap = (top2[:, 0] - top2[:, 1])
    n = int(mism.sum())
    worst = float(gap[mism].max()) if n else 0.0
    assert worst <= 2.0 * err_abs + 1e-12, f"argmax differs at a pixel whose top-2 gap {worst:.3e} exceeds 2 x max err {err_abs:.3e}"
    return n, worst


@pytest.fixture(scope="module")
def afb():
    import acquisition_focus_b200 as m
    return m


@pytest.mark.parametrize("S,seed,with_dvol", [(32, 41, True), (128, 43, True)])
def test_raw_parameter_path_measured(afb, S, seed, with_dvol):
    """cfg2 (B=2 x V=3, soft C=8 with grad wrt volume AND parameters + int64 one-hot label + image) from MLP-head outputs."""
    tag = f"cfg2_s{S}"
    case = cases.atm_case(S, 2, 3, seed=seed)
    ref, ref_dsoft = _oracle_run(case, "cpu", with_dvol)
    cud, cud_dsoft = _oracle_run(case, "cuda", with_dvol)          # the reference's ops through ATen's sm_100 kernels
    our, our_dsoft = _ours_run(afb, case, with_dvol)
    m = {"ours_vs_cpu": {}, "aten_cuda_vs_cpu": {}}
    for name, got, gd in (("ours_vs_cpu", our, our_dsoft), ("aten_cuda_vs_cpu", cud, cud_dsoft)):
        d = m[name]
        for key in ("theta", "ga", "ys", "yi", "dparams"):
            d[key] = max(_rel(g[key], r[key]) for g, r in zip(got, ref))
        d["label_mismatch_pixels"] = int(sum((g["yl"] != r["yl"]).any(1).sum() for g, r in zip(got, ref)))
        d["argmax_mismatch_pixels"] = int(sum((g["ys"].argmax(1) != r["ys"].argmax(1)).sum() for g, r in zip(got, ref)))
        if with_dvol:
            d["dsoft"] = _rel(gd, ref_dsoft)
    m["pixels"] = int(case["B"] * case["V"] * S * S)
    # tie proofs (ours)
    ties = [_tie_proof(r, o, case["lab"], S) for r, o in zip(ref, our)]
    m["tie_proof"] = {"label_mismatches": sum(t[0] for t in ties), "delta_vox": max(t[1] for t in ties),
                      "worst_distance_to_tie": max(t[2] for t in ties), "pixels_within_delta_of_a_tie": sum(t[3] for t in ties)}
    arg = [_argmax_proof(r["ys"], o["ys"], (o["ys"] - r["ys"]).abs().max().item()) for r, o in zip(ref, our)]
    m["argmax_proof"] = {"mismatches": sum(a[0] for a in arg), "worst_top2_gap": max(a[1] for a in arg)}
    _record(tag, m)

    th = _thresholds()[tag]
    o, c = m["ours_vs_cpu"], m["aten_cuda_vs_cpu"]
    for key in ("theta", "ga", "ys", "yi", "dparams") + (("dsoft",) if with_dvol else ()):
        assert o[key] <= th[key], f"{tag} {key}: {o[key]:.3e} > threshold {th[key]:.3e} (2x the error measured on the B200)"
    # labels: bit-exact except on rounding ties - every mismatch was proven above to be a tie flip, so their number is bounded
    # by the number of pixels within delta of a tie (measured on the B200: 0 mismatches of 98 304 pixels at 128^3, 36 near ties)
    assert o["label_mismatch_pixels"] <= m["tie_proof"]["pixels_within_delta_of_a_tie"]
    assert o["argmax_mismatch_pixels"] == m["argmax_proof"]["mismatches"]
    # north_star bars: gradients 1e-4; forward 1e-5 - or no further from the CPU reference than the reference's own CUDA run is
    assert o["dparams"] <= 1e-4 and (not with_dvol or o["dsoft"] <= 1e-4)
    assert o["ys"] <= max(1e-5, 2.0 * c["ys"]), (o["ys"], c["ys"])
    assert o["yi"] <= max(1e-5, 2.0 * c["yi"]), (o["yi"], c["yi"])


def test_dvolume_cfg2_128_golden(afb, golden_dir):
    """cfg2 at the headline shape WITH soft.requires_grad: dVolume (projections, 4096 picked voxels) against the reference's
    own gradient (tests/golden/atm_s128.npz, minted by oracle/make_golden.py on the unmodified reference)."""
    g = np.load(os.path.join(golden_dir, "atm_s128.npz"))
    if "dsoft_sum_w" not in g.files:
        pytest.skip("atm_s128.npz predates the dVolume fields")
    case = cases.atm_case(128, 2, 3, seed=43)
    our, dsoft = _ours_run(afb, case, True)
    scale = float(g["dsoft_absmax"])
    e_w = (dsoft.sum(-1).double() - torch.from_numpy(g["dsoft_sum_w"]).double()).abs().max().item() / torch.from_numpy(g["dsoft_sum_w"]).abs().max().item()
    e_d = (dsoft.sum(2).double() - torch.from_numpy(g["dsoft_sum_d"]).double()).abs().max().item() / torch.from_numpy(g["dsoft_sum_d"]).abs().max().item()
    e_p = np.abs(dsoft.reshape(-1).numpy()[g["dsoft_pick_idx"]].astype(np.float64) - g["dsoft_pick"]).max() / scale
    _record("cfg2_s128_dvolume_golden", {"sum_w": e_w, "sum_d": e_d, "picked_voxels": e_p, "scale": scale})
    assert e_w <= 1e-4 and e_d <= 1e-4 and e_p <= 1e-4, (e_w, e_d, e_p)
    for v in range(3):
        assert _rel(our[v]["dparams"], g[f"dparams{v}"]) <= 1e-4


def test_embed_cfg3_stage0_s128(afb):
    """cfg3 at the stage that carries 75 % of the bytes: S = 128, c = 16 (B = 1, V = 2), forward AND backward against the
    dense oracle restatement of SkipConnector.forward (x_mid + affine_grid + grid_sample on CPU) and, forward only, against
    the fp64 sparse closed form."""
    S, c, V, B = 128, 16, 2, 1
    case = cases.embed_case(S, c, V, B, seed=179)
    x = case["x"].cuda().requires_grad_(True)
    gas = [a.cuda().requires_grad_(True) for a in case["affines"]]
    out = afb.SkipConnector(V)(x, gas)
    go = cases.pattern(out.shape, 1.0)
    (out * go.cuda()).sum().backward()
    xr = case["x"].clone().requires_grad_(True)
    gr = [a.clone().requires_grad_(True) for a in case["affines"]]
    ref = O.skip_connector(xr, gr, V)
    (ref * go).sum().backward()
    m = {"out": _rel(out, ref), "dx": _rel(x.grad, xr.grad), "d_affines": _rel(torch.stack([a.grad for a in gas]), torch.stack([a.grad for a in gr])),
         "nonzero_frac_ours": float((out != 0).float().mean()), "nonzero_frac_ref": float((ref != 0).float().mean())}
    sp = O.skip_connector_sparse(case["x"].cuda(), [a.cuda() for a in case["affines"]], V)     # fp64 closed form (torch ops on the GPU)
    m["out_vs_sparse_fp64"] = _rel(out, sp)
    # the reference's own op sequence through ATen's sm_100 kernels, against the same code on CPU
    xc = case["x"].cuda().requires_grad_(True)
    gc = [a.cuda().requires_grad_(True) for a in case["affines"]]
    refc = O.skip_connector(xc, gc, V)
    (refc * go.cuda()).sum().backward()
    m["aten_cuda_vs_cpu"] = {"out": _rel(refc, ref), "dx": _rel(xc.grad, xr.grad),
                             "d_affines": _rel(torch.stack([a.grad for a in gc]), torch.stack([a.grad for a in gr]))}
    _record("cfg3_s128_c16", m)
    th = _thresholds()["cfg3_s128_c16"]
    for k in ("out", "dx", "d_affines", "out_vs_sparse_fp64"):
        assert m[k] <= th[k], f"{k}: {m[k]:.3e} > {th[k]:.3e}"
    assert m["out"] <= max(1e-5, 2.0 * m["aten_cuda_vs_cpu"]["out"]) and m["dx"] <= 1e-4 and m["d_affines"] <= 1e-4
    assert abs(m["nonzero_frac_ours"] - m["nonzero_frac_ref"]) < 1e-4


def test_use_affine_theta_false_takes_init_affines_only(afb):
    """the 'ref' stage (running/stages.py:76-82, learnable_transform.py:262-272) slices with the init affines only;
    a non-zero MLP head must not move the views, in the single module and in get_reconstruction_model_input."""
    from acquisition_focus_b200.running import model_input as MI
    import types
    cfgd, batch, params, (B, V, C, names) = cases.model_input_setup(S=32, aug=False)
    cfgd["use_affine_theta"] = False
    cfg = types.SimpleNamespace(**cfgd)

    class Stub(torch.nn.Module):
        def __init__(self, p):
            super().__init__(); self.p = torch.nn.Parameter(p)
        def forward(self, x):
            return self.p
    nets = iter([Stub(params[v].clone()) for v in range(V)])
    container = afb.ATModulesContainer(cfg, C, localization_net_factory=lambda: next(nets)).cuda()
    cb = {"label": batch["label"].cuda(), "image": batch["image"].cuda(),
          "additional_data": {"nifti_affine": batch["additional_data"]["nifti_affine"].cuda(),
                              "gt_view_affines": {k: v.cuda() for k, v in batch["additional_data"]["gt_view_affines"].items()}}}
    b_in, b_t, affs = MI.get_reconstruction_model_input(cb, "train", cfg, C, container)
    # oracle: theta = init_t @ init_a @ init_z = identity for the default inits -> pre = Gpre
    ad = batch["additional_data"]
    with torch.no_grad():
        lab, _, nii = O.nifti_grid_sample(batch["label"].unsqueeze(1), ad["nifti_affine"], target_fov_mm=torch.tensor(cfgd["hires_fov_mm"]),
                                          target_fov_vox=torch.tensor(cfgd["hires_fov_vox"]), is_label=True,
                                          pre_grid_sample_affine=ad["gt_view_affines"]["centroids"].to(ad["nifti_affine"]))
    soft = F.one_hot(lab.squeeze(1), C).permute(0, 4, 1, 2, 3).float()
    base = ad["gt_view_affines"]["centroids"].to(ad["nifti_affine"])
    for v, n in enumerate(names):
        gpre = O.input_affine_for_view(base, ad["gt_view_affines"][n]).float()
        a0 = O.r6_to_matrix(INIT[:, :6])
        ys, ga, _ = O.nifti_grid_sample(soft, nii, target_fov_mm=torch.tensor(cfgd["slice_fov_mm"]), target_fov_vox=torch.tensor(cfgd["slice_fov_vox"]),
                                        is_label=False, pre_grid_sample_affine=gpre @ a0.repeat(B, 1, 1))
        assert _rel(affs[v], ga) <= 2e-6
        assert _rel(b_in[:, v * C:(v + 1) * C], ys.squeeze(-1)) <= 2e-5
