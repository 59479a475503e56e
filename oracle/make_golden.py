"""ORACLE / test infrastructure: mint golden vectors from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python -m oracle.make_golden

The reference ships no tests or fixtures for the view-acquisition path, so the
known answers are produced by importing the reference itself
(``oracle/ref_import.py``) and running it on CPU on the deterministic inputs of
``oracle/cases.py``.  Outputs:

* ``acquisition_focus_b200/data/phantom_view_affines.json`` - the clinical view
  affines ``get_clinical_cardiac_view_affines`` (``functional/
  clinical_cardiac_views.py:223-364``) returns for the 128^3 phantom (inputs of
  the hot path, used by bench/smoke/tests as constants).
* ``tests/golden/*.npz`` - forward outputs and gradients of
  ``compute_rotation_matrix_from_ortho6d``, ``nifti_grid_sample``,
  ``AffineTransformModule.forward`` (LocalizationNet replaced by a stub that
  returns given MLP-head outputs) and ``SkipConnector.forward``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "acquisition_focus_b200", "data")


def _np(t):
    return t.detach().cpu().numpy()


def write_view_affines(R):
    from oracle.cases import synthetic
    lab = synthetic.heart_phantom(128)
    nii = torch.diag(torch.tensor([1.5, 1.5, 1.5, 1.0]))
    views = R.get_clinical_cardiac_view_affines(torch.as_tensor(lab), nii, synthetic.CLASS_DICT,
                                                num_sa_slices=3, return_unrolled=True)
    os.makedirs(DATA, exist_ok=True)
    with open(os.path.join(DATA, "phantom_view_affines.json"), "w") as fh:
        json.dump({"source": "get_clinical_cardiac_view_affines (reference, CPU) on synthetic.heart_phantom(128), "
                             "nifti affine diag(1.5,1.5,1.5,1), num_sa_slices=3",
                   "views": {k: [[float(x) for x in row] for row in v.to(torch.float32).tolist()]
                             for k, v in views.items()}}, fh, indent=1)
    print("view affines:", list(views.keys()))


class _Stub(torch.nn.Module):
    def __init__(self, out):
        super().__init__()
        self.out = out

    def forward(self, x):
        return self.out


def gold_r6(R):
    from oracle import cases
    o = cases.randn((7, 6), 11).requires_grad_(True)
    m = R.compute_rotation_matrix_from_ortho6d(o)
    go = cases.pattern(m.shape, 1.0)
    (m * go).sum().backward()
    np.savez_compressed(os.path.join(GOLD, "r6.npz"), ortho=_np(o), mat=_np(m), d_ortho=_np(o.grad))


def gold_slice_small(R):
    from oracle import cases
    B, C, D, H, W = 2, 3, 20, 24, 28
    vol = cases.randn((B, C, D, H, W), 21)
    lab = cases.randint(0, 6, (B, C, D, H, W), 22)
    nii = cases.rotated_nii_affine(B, 23)
    P = cases.random_pre_affine(B, 24)
    out = {}
    for tag, kw in (("slice", dict(target_fov_mm=torch.tensor([30.0, 20.0, 1.5]), target_fov_vox=torch.tensor([16, 12, 1]))),
                    ("vol3d", dict(target_fov_mm=torch.tensor([28.0, 30.0, 33.0]), target_fov_vox=torch.tensor([9, 10, 11]))),
                    ("same", dict())):
        v = vol.clone().requires_grad_(True)
        p = P.clone().requires_grad_(True)
        y, ga, na = R.nifti_grid_sample(v, nii, is_label=False, pre_grid_sample_affine=p, **kw)
        go = cases.pattern(y.shape, 1.0)
        gg = cases.pattern(ga.shape, 2.0)
        ((y * go).sum() + (ga * gg).sum()).backward()
        yl, _, _ = R.nifti_grid_sample(lab, nii, is_label=True, pre_grid_sample_affine=P, **kw)
        out.update({f"{tag}_y": _np(y), f"{tag}_ga": _np(ga), f"{tag}_nii": _np(na), f"{tag}_dvol": _np(v.grad),
                    f"{tag}_dP": _np(p.grad), f"{tag}_ylabel": _np(yl)})
    np.savez_compressed(os.path.join(GOLD, "slice_small.npz"), nii=_np(nii), P=_np(P), **out)


def gold_slice_cfg1(R):
    """configs[0]: one 128^3 fp32 image volume, one p2CH view, batch 1, R6 -> slice fwd+bwd."""
    from oracle import cases
    syn = cases.synthetic
    lab = syn.heart_phantom(128)
    vol = torch.from_numpy(syn.phantom_image(lab, seed=5))[None, None].requires_grad_(True)
    r6 = (torch.tensor([[1.0, 0, 0, 0, 1.0, 0]]) + 0.3 * cases.randn((1, 6), 31)).requires_grad_(True)
    gpre = syn.phantom_view_affines()["p2CH"][None]
    nii = syn.default_nifti_affine(1)
    P = gpre @ R.compute_rotation_matrix_from_ortho6d(r6)
    y, ga, na = R.nifti_grid_sample(vol, nii, target_fov_mm=torch.tensor([192.0, 192.0, 1.5]),
                                    target_fov_vox=torch.tensor([128, 128, 1]), is_label=False,
                                    pre_grid_sample_affine=P)
    go = cases.pattern(y.shape, 1.0)
    (y * go).sum().backward()
    dv = vol.grad[0, 0]
    np.savez_compressed(os.path.join(GOLD, "slice_cfg1_128.npz"), r6=_np(r6), gpre=_np(gpre), y=_np(y), ga=_np(ga),
                        nii=_np(na), d_r6=_np(r6.grad), dvol_sum_d=_np(dv.sum(0)), dvol_sum_h=_np(dv.sum(1)),
                        dvol_sum_w=_np(dv.sum(2)), dvol_absmax=np.float32(dv.abs().max().item()))


def gold_rotation_params(R):
    """f4: the two non-default parameterisations, forward + gradient (utils/transform_utils.py:62-178)."""
    from oracle import cases
    tu = R.transform_utils
    aa = cases.randn((9, 3), 61)
    aa[0] = 0.0
    aa[1] = 1e-4 * aa[1]                      # below the eps threshold: first-order branch
    nv = cases.randn((9, 3), 62)
    out = {}
    for tag, fn, x in (("aa", tu.angle_axis_to_rotation_matrix, aa), ("nv", tu.normal_to_rotation_matrix, nv)):
        xx = x.clone().requires_grad_(True)
        m = fn(xx)
        (m * cases.pattern(m.shape, 1.0)).sum().backward()
        out.update({f"{tag}_in": _np(x), f"{tag}_mat": _np(m), f"{tag}_grad": _np(xx.grad)})
    np.savez_compressed(os.path.join(GOLD, "rotation_params.npz"), **out)


def _run_atm(R, case, optim_method="R6-vector", init_ap=None, rotate=False):
    B, V, S = case["B"], case["V"], case["S"]
    res = []
    for v in range(V):
        atm = R.AffineTransformModule(8, case["volume_fov_mm"], case["volume_fov_vox"], case["slice_fov_mm"],
                                      case["slice_fov_vox"], optim_method=optim_method, rotate_slice_to_min_principle=rotate,
                                      offset_clip_value=case["offset_clip"], zoom_clip_value=case["zoom_clip"],
                                      view_id="p2CH")
        assert atm.vox_range == case["R"]
        if init_ap is not None:
            atm.set_init_theta_ap(init_ap)
        params = case["params"][v].clone().requires_grad_(True)
        atm.localization_net = _Stub(params)
        soft = case["soft"].clone().requires_grad_(True)
        ys, yl, yi, ga, na = atm(soft, case["label"], case["image"], case["nii"], case["gpre"][v])
        from oracle import cases as C
        go = C.pattern(ys.shape, 1.0 + v)
        gg = C.pattern(ga.shape, 2.0 + v)
        ((ys * go).sum() + (ga * gg).sum()).backward()
        res.append(dict(ys=ys, yl=yl, yi=yi, ga=ga, na=na, dparams=params.grad, dsoft=soft.grad, theta=atm.last_theta))
    return res


def gold_atm(R):
    from oracle import cases
    # 32^3, full outputs and gradients, with and without zoom clip
    for tag, zc in (("atm_s32", 0.0), ("atm_s32_zoom", 0.3)):
        case = cases.atm_case(32, 2, 3, seed=41, zoom_clip=zc)
        res = _run_atm(R, case)
        out = {"gpre": np.stack([_np(g) for g in case["gpre"]]), "params": np.stack([_np(p) for p in case["params"]])}
        for v, r in enumerate(res):
            out.update({f"ys{v}": _np(r["ys"]), f"yl{v}": _np(r["yl"]).astype(np.uint8), f"yi{v}": _np(r["yi"]),
                        f"ga{v}": _np(r["ga"]), f"na{v}": _np(r["na"]), f"dparams{v}": _np(r["dparams"]),
                        f"theta{v}": _np(r["theta"]),
                        # dense dVolume: store its projections (it is dense because the min-shift
                        # gradient is spread evenly over all voxels equal to the minimum)
                        f"dsoft_sum_w{v}": _np(r["dsoft"].sum(-1)), f"dsoft_sum_d{v}": _np(r["dsoft"].sum(2))})
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)
    # 128^3 (configs[1] shapes): label slices (bit-exact targets), argmax of soft slices, image slices
    case = cases.atm_case(128, 2, 3, seed=43)
    res = _run_atm(R, case)
    out = {"gpre": np.stack([_np(g) for g in case["gpre"]]), "params": np.stack([_np(p) for p in case["params"]])}
    for v, r in enumerate(res):
        out.update({f"yl{v}": _np(r["yl"]).astype(np.uint8), f"ys_argmax{v}": _np(r["ys"].argmax(1)).astype(np.uint8),
                    f"yi{v}": _np(r["yi"]), f"ga{v}": _np(r["ga"]), f"na{v}": _np(r["na"]),
                    f"dparams{v}": _np(r["dparams"]),
                    f"ys_ch3_{v}": _np(r["ys"][:, 3]).astype(np.float32)})
    # dVolume at the headline shape: projections of the gradient summed over the three views (what one fused acquisition
    # accumulates into soft.grad), plus its value at a fixed pseudo-random set of voxels
    dsoft = sum(r["dsoft"] for r in res)
    pick = np.random.default_rng(7).integers(0, dsoft.numel(), size=4096)
    out.update({"dsoft_sum_w": _np(dsoft.sum(-1)), "dsoft_sum_d": _np(dsoft.sum(2)), "dsoft_pick_idx": pick,
                "dsoft_pick": _np(dsoft).reshape(-1)[pick], "dsoft_absmax": np.float32(dsoft.abs().max().item())})
    np.savez_compressed(os.path.join(GOLD, "atm_s128.npz"), **out)


def gold_atm_other_params(R):
    """f4: AffineTransformModule.forward with optim_method angle-axis / normal-vector (32^3, 2 views).  The MLP-head outputs
    are the R6 cases' with the first 3 of the 6 rotation parameters kept; normal-vector gets a non-degenerate init
    (the reference's all-zero default divides by zero, learnable_transform.py:85-88 + transform_utils.py:83)."""
    from oracle import cases
    for method, init_ap in (("angle-axis", None), ("normal-vector", torch.tensor([0.2, -0.1, 1.0]))):
        case = cases.atm_case(32, 2, 2, seed=71)
        case["params"] = [torch.cat([0.4 * p[:, :3], p[:, 6:]], dim=1) for p in case["params"]]
        res = _run_atm(R, case, optim_method=method, init_ap=init_ap)
        out = {"gpre": np.stack([_np(g) for g in case["gpre"]]), "params": np.stack([_np(p) for p in case["params"]]),
               "init_ap": _np(init_ap if init_ap is not None else torch.zeros(3))}
        for v, r in enumerate(res):
            out.update({f"ys{v}": _np(r["ys"]), f"yl{v}": _np(r["yl"]).astype(np.uint8), f"yi{v}": _np(r["yi"]),
                        f"ga{v}": _np(r["ga"]), f"na{v}": _np(r["na"]), f"dparams{v}": _np(r["dparams"]),
                        f"theta{v}": _np(r["theta"])})
        np.savez_compressed(os.path.join(GOLD, "atm_s32_" + method.replace("-", "_") + ".npz"), **out)


def gold_atm_rotate(R):
    """a13: AffineTransformModule.forward with rotate_slice_to_min_principle=True (learnable_transform.py:315-328,337-366),
    32^3, 2 views: aligned slices, the composed grid affine, the updated NIfTI affine and parameter gradients."""
    from oracle import cases
    case = cases.atm_case(32, 2, 2, seed=81)
    res = _run_atm(R, case, rotate=True)
    out = {"gpre": np.stack([_np(g) for g in case["gpre"]]), "params": np.stack([_np(p) for p in case["params"]])}
    for v, r in enumerate(res):
        out.update({f"ys{v}": _np(r["ys"]), f"yl{v}": _np(r["yl"]).astype(np.uint8), f"yi{v}": _np(r["yi"]),
                    f"ga{v}": _np(r["ga"]), f"na{v}": _np(r["na"]), f"dparams{v}": _np(r["dparams"]),
                    f"dsoft_sum_w{v}": _np(r["dsoft"].sum(-1))})
    np.savez_compressed(os.path.join(GOLD, "atm_s32_rotate.npz"), **out)


def gold_embed(R):
    from oracle import cases
    for tag, (S, c, V, B) in (("embed_s16", (16, 3, 2, 2)), ("embed_s8", (8, 4, 3, 2)), ("embed_s32", (32, 4, 6, 1))):
        case = cases.embed_case(S, c, V, B, seed=51 + S)
        sc = R.SkipConnector(V)
        x = case["x"].clone().requires_grad_(True)
        gas = [g.clone().requires_grad_(True) for g in case["affines"]]
        out = sc(x, gas)
        go = cases.pattern(out.shape, 1.0)
        (out * go).sum().backward()
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), affines=np.stack([_np(g) for g in gas]),
                            out=_np(out).astype(np.float32), dx=_np(x.grad), d_affines=np.stack([_np(g.grad) for g in gas]))


def gold_model_input(R):
    """a11 + a6: the reference's own ``get_reconstruction_model_input`` (running/run_dl.py:238-329, which calls
    ``get_transformed`` :146-204, ``get_input_affine_for_atm`` :227-234 and ``apply_affine_augmentation`` :208-223 on the
    GLOBAL torch RNG) executed unmodified via ``oracle.ref_import.load_run_dl``; LocalizationNets replaced by stubs that
    return fixed MLP-head outputs.  ``torch.manual_seed(123)`` right before the call pins the augmentation stream."""
    from oracle import cases
    from oracle.ref_import import load_run_dl
    RD = load_run_dl()
    for tag, kw in (("model_input_s32", dict(S=32)), ("model_input_s32_lowres", dict(S=32, slice_vox=16)),
                    ("model_input_s32_noaug", dict(S=32, aug=False))):
        cfg, batch, params, (B, V, C, names) = cases.model_input_setup(**kw)
        config = RD.DotDict(cfg)
        container = R.learnable_transform.ATModulesContainer(config, C)
        plist = []
        for v, atm in enumerate(container):
            p = params[v].clone().requires_grad_(True)
            atm.localization_net = _Stub(p)
            # _Stub holds a plain tensor: give the module a parameter so that the view counts as active (:399-405)
            atm.localization_net.flag = torch.nn.Parameter(torch.zeros(1))
            plist.append(p)
        torch.manual_seed(123)
        b_input, b_target, grid_affines = RD.get_reconstruction_model_input(batch, "train", config, C, container, None)
        go = cases.pattern(b_input.shape, 1.0)
        loss = (b_input * go).sum()
        for v, g in enumerate(grid_affines):
            loss = loss + (g * cases.pattern(g.shape, 2.0 + v)).sum()
        loss.backward()
        out = {"b_input": _np(b_input).astype(np.float32), "b_target_argmax": _np(b_target.argmax(1)).astype(np.uint8),
               "grid_affines": np.stack([_np(g) for g in grid_affines]), "dparams": np.stack([_np(p.grad) for p in plist])}
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    R = load_reference()
    groups = {"views": write_view_affines, "r6": gold_r6, "slice_small": gold_slice_small, "slice_cfg1": gold_slice_cfg1,
              "atm": gold_atm, "embed": gold_embed, "rotation_params": gold_rotation_params, "atm_other": gold_atm_other_params,
              "atm_rotate": gold_atm_rotate, "model_input": gold_model_input}
    for name in (sys.argv[1:] or list(groups)):          # python -m oracle.make_golden [group ...]
        groups[name](R)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
