#!/usr/bin/env python
"""bench_extra.py - the other BASELINE.json configurations (parity-test cases, not the headline bench line):

  cfg1  single 128^3 fp32 image volume, one p2CH view, batch 1: R6 -> slice fwd+bwd (launch-latency bound)
  cfg2  config_dict.json default: B=2 x V=3, soft label C=8 + int64 one-hot label + image, fwd+bwd
  cfg3  slice-to-3D embedding of 6 views into the reconstruction FOV, all 6 U-Net stages, fwd+bwd
  cfg5  256^3 volume, 256^2 slices, 16 views, C=8, fp32 and bf16 storage, fwd+bwd

Each line: {"config": ..., "ms": ..., "value": ..., "unit": ..., "bytes": algorithmic bytes, "gbs": ...,
            "torch_cuda_ms": same op through ATen's sm_100 kernels on the same GPU (the Blackwell bar)}.
CUDA events, 3 warm-ups, mean of 10.  Run on the GPU box: python bench_extra.py > profiles/...json

Inputs come from the product's own synthetic module.  `oracle.af_oracle` is imported ONLY inside the `ref()` / ATen-CUDA
baseline legs (the reference's op sequence timed on the GPU as the Blackwell bar) - never on a measured product path.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def torch_slice(vol, theta, size, mode):
    grid = F.affine_grid(theta, [vol.shape[0], vol.shape[1], *size], align_corners=False)
    return F.grid_sample(vol, grid, mode=mode, padding_mode="zeros", align_corners=False)


def cfg1(afb, dev):
    vol = torch.randn(1, 1, 128, 128, 128, device=dev, requires_grad=True)
    r6 = (torch.tensor([[1.0, 0, 0, 0, 1.0, 0]]) + 0.3 * torch.randn(1, 6)).to(dev).requires_grad_(True)
    nii = torch.diag(torch.tensor([1.5, 1.5, 1.5, 1.0])).double()[None].to(dev)
    fov_mm, fov_vox = torch.tensor([192.0, 192.0, 1.5]), torch.tensor([128, 128, 1])
    go = torch.randn(1, 1, 128, 128, 1, device=dev)

    def ours():
        vol.grad = None; r6.grad = None
        y, ga, _ = afb.nifti_grid_sample(vol, nii, target_fov_mm=fov_mm, target_fov_vox=fov_vox,
                                         pre_grid_sample_affine=afb.compute_rotation_matrix_from_ortho6d(r6))
        y.backward(go)

    def ref():     # the reference's op sequence through ATen CUDA (min-shift + affine_grid + grid_sample)
        vol.grad = None; r6.grad = None
        from oracle import af_oracle as O
        y, ga, _ = O.nifti_grid_sample(vol, nii, target_fov_mm=fov_mm, target_fov_vox=fov_vox,
                                       pre_grid_sample_affine=O.r6_to_matrix(r6))
        y.backward(go)
    t, tmin = timeit(ours)
    tr, _ = timeit(ref)
    return {"config": "cfg1: 1 x 128^3 fp32, 1 view, R6 -> slice fwd+bwd (dVol + dR6)", "ms": t, "ms_min": tmin, "value": 1e3 / t,
            "unit": "slices/s", "bytes": 104 * 128 * 128 + 2 * 8 * 2 ** 20, "torch_cuda_ms": tr,
            "note": "launch-latency bound by construction (1.7 MB of gather traffic + 8 MiB dVolume fill + min pass)"}


def cfg2(afb, dev):
    from acquisition_focus_b200 import synthetic as cases          # input builders (product-side synthetic module)
    case = cases.atm_case(128, 2, 3, seed=43)
    soft = case["soft"].to(dev).requires_grad_(True)
    label, image, nii = case["label"].to(dev), case["image"].to(dev), case["nii"].to(dev)
    gpre = torch.stack(case["gpre"], 1).to(dev)
    params = torch.stack(case["params"], 1).to(dev).requires_grad_(True)
    init = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(3, 1).to(dev)
    go = torch.randn(2, 3, 8, 128, 128, 1, device=dev)

    def ours():
        soft.grad = None; params.grad = None
        ys, yl, yi, ga, nii_o, th = afb.acquire_views(soft, label, image, nii, gpre, params, init, offset_clip=0.2, zoom_clip=0.0,
                                                      spat=128, slice_fov_mm=[192.0, 192.0, 1.5], slice_fov_vox=[128, 128, 1])
        ys.backward(go)

    def ref():
        from oracle import af_oracle as O
        soft.grad = None; params.grad = None
        loss = 0
        for v in range(3):
            th = O.view_theta(params[:, v], init[:1, :6], init[0, 6:9], init[:1, 9:], 0.2, 0.0, 128)
            ys, yl, yi, ga, _ = O.atm_tail_forward(soft, label, image, nii, gpre[:, v], th, case["slice_fov_mm"], case["slice_fov_vox"])
            loss = loss + (ys * go[:, v]).sum()
        loss.backward()
    t, tmin = timeit(ours)
    tr, _ = timeit(ref, reps=3, warm=1)
    # the same step captured once into a CUDA graph (launch-latency bound otherwise)
    from acquisition_focus_b200.graphs import GraphedStep

    def graphable():
        soft.grad = None; params.grad = None
        ys, yl, yi, ga, nii_o, th = afb.acquire_views(soft, label, image, nii, gpre, params, init, offset_clip=0.2, zoom_clip=0.0,
                                                      spat=128, slice_fov_mm=[192.0, 192.0, 1.5], slice_fov_vox=[128, 128, 1])
        ys.backward(go)
        return ys, yl, yi, ga, soft.grad, params.grad
    gs = GraphedStep(graphable)
    tg, tgmin = timeit(gs)
    return {"config": "cfg2: B=2 x V=3 p2CH, soft C=8 (grad) + int64 one-hot label + image, fwd+bwd wrt volume and theta",
            "ms": t, "ms_min": tmin, "value": 6e3 / t, "unit": "slices/s", "torch_cuda_ms": tr,
            "cuda_graph_ms": tg, "cuda_graph_ms_min": tgmin, "cuda_graph_value": 6e3 / tg,
            "note": "torch_cuda_ms = oracle port of the reference's op sequence run through ATen's sm_100 CUDA kernels"}


def cfg3(afb, dev):
    from acquisition_focus_b200 import synthetic as cases          # input builders (product-side synthetic module)
    out = []
    for B in (1, 2):
        tot_f = tot_b = tot_rf = tot_rb = 0.0
        tot_bytes = 0
        for c, S in ((16, 128), (32, 64), (64, 32), (128, 16), (256, 8), (256, 4)):
            V = 6
            case = cases.embed_case(S, c, V, B, seed=300 + S)
            x = case["x"].to(dev).requires_grad_(True)
            gas = [a.to(dev).requires_grad_(True) for a in case["affines"]]
            aff = torch.stack(gas, 0)
            go = torch.randn(B, V * c, S, S, S, device=dev)
            sc = afb.SkipConnector(V)
            f, _ = timeit(lambda: afb.embed_slices(x, aff, V), reps=5, warm=2)

            def fb():
                x.grad = None
                for a in gas:
                    a.grad = None
                sc(x, gas).backward(go)
            fbt, _ = timeit(fb, reps=5, warm=2)
            from oracle import af_oracle as O
            rf, _ = timeit(lambda: O.skip_connector(x.detach(), [a.detach() for a in gas], V), reps=2, warm=1)

            def rfb():
                x.grad = None
                for a in gas:
                    a.grad = None
                O.skip_connector(x, gas, V).backward(go)
            rfbt, _ = timeit(rfb, reps=2, warm=1)
            nbytes = B * V * c * S ** 3 * 4 + B * V * c * S * S * 4
            out.append({"config": f"cfg3 stage c={c} S={S} B={B} V=6", "fwd_ms": f, "fwd_bwd_ms": fbt, "bytes_fwd": nbytes,
                        "fwd_gbs": nbytes / f / 1e6, "torch_cuda_fwd_ms": rf, "torch_cuda_fwd_bwd_ms": rfbt})
            tot_f += f; tot_b += fbt; tot_rf += rf; tot_rb += rfbt; tot_bytes += nbytes
            del x, gas, aff, go
            torch.cuda.empty_cache()
        out.append({"config": f"cfg3 all 6 stages B={B} V=6", "fwd_ms": tot_f, "fwd_bwd_ms": tot_b, "bytes_fwd": tot_bytes,
                    "fwd_gbs": tot_bytes / tot_f / 1e6, "value": 1e3 / tot_b, "unit": "embeddings/s (fwd+bwd, 6 stages)",
                    "torch_cuda_fwd_ms": tot_rf, "torch_cuda_fwd_bwd_ms": tot_rb})
    return out


def cfg3_summary(afb, dev, B=2, V=6):
    """cfg3 for the driver-run bench line: all six U-Net stages at B=2, V=6, forward and forward+backward, per stage and summed,
    with the HBM roofline fraction of each stage's forward (compulsory bytes: the [B,V*c,S^3] output written once + the feature
    maps read once) and the same op through ATen's sm_100 kernels (oracle op sequence on the GPU), stage 0 and summed."""
    from acquisition_focus_b200 import synthetic as cases          # input builders (product-side synthetic module)
    import json as _json
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(_json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    stages, tot = [], {"fwd_ms": 0.0, "fwd_bwd_ms": 0.0, "bytes_fwd": 0, "bytes_bwd": 0, "aten_fwd_ms": 0.0, "aten_fwd_bwd_ms": 0.0}
    for c, S in ((16, 128), (32, 64), (64, 32), (128, 16), (256, 8), (256, 4)):
        case = cases.embed_case(S, c, V, B, seed=300 + S)
        x = case["x"].to(dev).requires_grad_(True)
        gas = [a.to(dev).requires_grad_(True) for a in case["affines"]]
        aff = torch.stack(gas, 0)
        go = torch.randn(B, V * c, S, S, S, device=dev)
        sc = afb.SkipConnector(V)
        f, _ = timeit(lambda: afb.embed_slices(x.detach(), aff.detach(), V), reps=5, warm=2)

        def fb():
            x.grad = None
            for a in gas:
                a.grad = None
            sc(x, gas).backward(go)
        fbt, _ = timeit(fb, reps=5, warm=2)
        from oracle import af_oracle as O

        def rfb():
            x.grad = None
            for a in gas:
                a.grad = None
            O.skip_connector(x, gas, V).backward(go)
        rf, _ = timeit(lambda: O.skip_connector(x.detach(), [a.detach() for a in gas], V), reps=1, warm=1)
        rfbt, _ = timeit(rfb, reps=1, warm=1)
        nb_f = B * V * c * S ** 3 * 4 + B * V * c * S * S * 4
        nb_b = B * V * c * (int(2.5 * S * S) + 2 * S * S) * 4          # slab of grad_out read + x read + dX written
        stages.append({"c": c, "S": S, "fwd_ms": f, "fwd_bwd_ms": fbt, "bytes_fwd": nb_f, "fwd_gbs": nb_f / f / 1e6,
                       "fwd_frac_of_hbm": nb_f / f / 1e6 / hbm, "aten_cuda_fwd_ms": rf, "aten_cuda_fwd_bwd_ms": rfbt})
        tot["fwd_ms"] += f; tot["fwd_bwd_ms"] += fbt; tot["bytes_fwd"] += nb_f; tot["bytes_bwd"] += nb_b
        tot["aten_fwd_ms"] += rf; tot["aten_fwd_bwd_ms"] += rfbt
        del x, gas, aff, go
        torch.cuda.empty_cache()
    # all six stages of one U-Net pass in ONE launch each way (HybridUnet.forward embeds every encoder skip with the same affines)
    case0 = cases.embed_case(128, 16, V, B, seed=300)
    gas = [a.to(dev).requires_grad_(True) for a in case0["affines"]]
    cfgs = ((16, 128), (32, 64), (64, 32), (128, 16), (256, 8), (256, 4))
    xs = [cases.randn((B, V * c, S, S), 500 + S).to(dev).requires_grad_(True) for c, S in cfgs]
    gos = [torch.randn(B, V * c, S, S, S, device=dev) for c, S in cfgs]
    xd = [x.detach() for x in xs]
    affd = torch.stack([g.detach() for g in gas], 0)
    f1, _ = timeit(lambda: afb.embed_slices_multi(xd, affd, V), reps=5, warm=2)

    def fb_all():
        for x in xs:
            x.grad = None
        for a in gas:
            a.grad = None
        torch.autograd.backward(afb.embed_slices_multi(xs, torch.stack(gas, 0), V), gos)
    fb1, _ = timeit(fb_all, reps=5, warm=2)
    tot["one_launch_fwd_ms"], tot["one_launch_fwd_bwd_ms"] = f1, fb1
    tot["fwd_gbs"] = tot["bytes_fwd"] / min(f1, tot["fwd_ms"]) / 1e6
    tot["fwd_frac_of_hbm"] = tot["fwd_gbs"] / hbm
    tot["value"] = 1e3 / min(fb1, tot["fwd_bwd_ms"])
    tot["unit"] = "embeddings/s (fwd+bwd, 6 stages, B=2, V=6)"
    tot["note"] = "fwd_ms / fwd_bwd_ms: stage by stage (zero kernel + slab kernel per stage); one_launch_*: afb_embed_multi_fwd/bwd"
    return {"config": f"cfg3: slice-to-3D embedding, V={V} views, B={B}, six stages (c,S) = (16,128) ... (256,4)", "stages": stages,
            "all_stages": tot, "hbm_peak_gbs": hbm}


def f1_resample_3d(afb, dev, B=8, V=3):
    """the 3-D -> 3-D resample mode of the same kernels - the prescan resample that feeds the LocalizationNet
    (models/learnable_transform.py:252-255: nifti_grid_sample(soft label C=8, 128^3 -> 128^3) per view and step, under
    no_grad), the largest sampler of a training step.  Compulsory HBM bytes: the [B,8,128^3] fp32 output written once + the
    input read once.  Three routes: channels-last soft volume (what run_dl.py:261-264 hands over), planar soft volume (generic
    kernel vs one transposing copy + channels-last kernel), and straight from the uint8 label map (no one-hot input at all)."""
    import json as _json
    from acquisition_focus_b200 import synthetic as cases          # input builders (product-side synthetic module)
    from oracle import af_oracle as O
    from acquisition_focus_b200 import functional as AF
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(_json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    S, C = 128, 8
    case = cases.atm_case(S, B, V, seed=47)
    soft_cl = case["soft"].to(dev)                                   # channels-last strides (permuted one-hot)
    soft_pl = soft_cl.contiguous()                                    # planar NCDHW
    lab = case["lab"].to(dev).to(torch.uint8)
    nii = case["nii"].to(dev)
    fov_mm, fov_vox = torch.tensor([192.0] * 3), torch.tensor([S] * 3)
    out_bytes = B * C * S ** 3 * 4
    res = {"config": f"f1: prescan resample 128^3 -> 128^3, C=8 bilinear, B={B}, one call per view (V={V} calls per step; B=2 is "
                     "host-launch bound: 0.35 ms per call for ~0.1 ms of kernels)", "hbm_peak_gbs": hbm}
    # the sampler kernel alone (one launch, no min pass / prologue / wrapper), channels-last soft volume
    from acquisition_focus_b200 import _lib as L
    spec = AF.ViewSpec(kind=L.AFFINE_PRE, V=1, nii_affine=case["nii"].to(dev), fov_mm=(192.0, 192.0, 192.0), pre=case["gpre"][0].to(dev).contiguous())
    spec = AF.prepare_views(spec, B, (S, S, S), [S, S, S], dev)[0]
    pad0 = AF.volume_min(case["soft"].to(dev))
    sd = case["soft"].to(dev)
    t, _ = timeit(lambda: AF._slice_forward_raw(sd, spec, [S, S, S], L.BILINEAR, L.PAD_DEVICE, 0.0, pad0), reps=5, warm=2)
    res["kernel_only_channels_last"] = {"ms": t, "bytes": 2 * B * C * S ** 3 * 4, "gbs": 2 * B * C * S ** 3 * 4 / t / 1e6,
                                        "frac_of_hbm": 2 * B * C * S ** 3 * 4 / t / 1e6 / hbm}

    def run(vol):
        for v in range(V):
            afb.nifti_grid_sample(vol, nii, target_fov_mm=fov_mm, target_fov_vox=fov_vox, is_label=False, pre_grid_sample_affine=case["gpre"][v].to(dev))

    def run_labels():
        for v in range(V):
            AF.onehot_resample_with_pre_affine(lab, nii, case["gpre"][v].to(dev), fov_mm.tolist(), fov_vox.tolist(), C)

    def run_aten():
        for v in range(V):
            O.nifti_grid_sample(soft_cl, nii, target_fov_mm=fov_mm.to(dev), target_fov_vox=fov_vox.to(dev), is_label=False,
                                pre_grid_sample_affine=case["gpre"][v].to(dev))
    for name, fn, in_bytes in (("channels_last_soft", lambda: run(soft_cl), out_bytes), ("planar_soft_transposed_once", lambda: run(soft_pl), 3 * out_bytes),
                               ("from_uint8_labels", run_labels, B * S ** 3)):
        t, _ = timeit(fn, reps=5, warm=2)
        nb = V * (out_bytes + in_bytes)
        res[name] = {"ms_per_step": t, "bytes": nb, "gbs": nb / t / 1e6, "frac_of_hbm": nb / t / 1e6 / hbm}
    os.environ["AFB_NO_TRANSPOSE"] = "1"
    t, _ = timeit(lambda: run(soft_pl), reps=3, warm=1)
    os.environ.pop("AFB_NO_TRANSPOSE")
    res["planar_soft_generic_kernel"] = {"ms_per_step": t, "gbs": V * 2 * out_bytes / t / 1e6}
    os.environ["AFB_NO_WIDE_PATCH"] = "1"            # A/B: the slices' 8 x 4 warp patch instead of 32 x 1 rows for 3-D outputs
    t, _ = timeit(lambda: run(soft_cl), reps=5, warm=2)
    t2, _ = timeit(run_labels, reps=5, warm=2)
    os.environ.pop("AFB_NO_WIDE_PATCH")
    res["channels_last_soft_narrow_patch"] = {"ms_per_step": t}
    res["from_uint8_labels_narrow_patch"] = {"ms_per_step": t2}
    t, _ = timeit(run_aten, reps=2, warm=1)
    res["aten_cuda_reference_ops"] = {"ms_per_step": t}
    return res


def f4_clinical_views(afb, dev):
    """SURVEY 8 f4 (last item): get_clinical_cardiac_view_affines (functional/clinical_cardiac_views.py:223-364) on the 128^3
    phantom: the GPU drop-in (voxel passes as kernels, 3x3 eigenproblems on the host) next to the reference's own function on
    the host cores (sparse CPU tensors; the vendored unmodified code when oracle/_ref travelled, else not timed)."""
    import time
    from acquisition_focus_b200 import synthetic as syn
    lab = torch.from_numpy(syn.heart_phantom(128))
    nii = torch.diag(torch.tensor([1.5, 1.5, 1.5, 1.0]))
    lab_d = lab.to(dev).to(torch.uint8)
    fn = lambda: afb.get_clinical_cardiac_view_affines(lab_d, nii, syn.CLASS_DICT, num_sa_slices=3, return_unrolled=True)
    for _ in range(2):
        got = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ours = (time.perf_counter() - t0) / 5 * 1e3
    res = {"config": "f4: clinical cardiac view affines (10 views) from a 128^3 label map", "ms_wall": ours,
           "note": "wall clock incl. six small D2H reads and the host eigenproblems; 8 kernel launches"}
    try:
        import bench
        R = bench._load_vendored_reference()
        if R is not None:
            t0 = time.perf_counter()
            want = R.get_clinical_cardiac_view_affines(lab, nii, syn.CLASS_DICT, num_sa_slices=3, return_unrolled=True)
            res["reference_cpu_ms_wall"] = (time.perf_counter() - t0) * 1e3
            res["max_abs_diff_vs_reference"] = max((got[k] - want[k]).abs().max().item() for k in want)
    except Exception as e:      # noqa: BLE001
        res["reference_cpu_error"] = repr(e)
    return res


def cfg5(afb, dev):
    out = []
    S, V, C = 256, 16, 8
    lab = torch.randint(0, C, (1, S, S, S), device=dev)
    soft32 = F.one_hot(lab, C).permute(0, 4, 1, 2, 3).float()
    del lab
    nii = torch.diag(torch.tensor([0.75, 0.75, 0.75, 1.0])).double()[None].to(dev)
    gen = torch.Generator().manual_seed(5)
    from acquisition_focus_b200 import synthetic as syn
    p2 = syn.phantom_view_affines()["p2CH"]
    gpre = torch.stack([p2 @ syn.random_aug_affine(gen, 0.3, 0.2, 0.0) for _ in range(V)])[None].to(dev)
    R = 51
    params = torch.cat([torch.tensor([1.0, 0, 0, 0, 1.0, 0]) + 0.3 * torch.randn(1, V, 6, generator=gen),
                        torch.randn(1, V, 3 * R, generator=gen), torch.randn(1, V, 1, generator=gen)], -1).to(dev).requires_grad_(True)
    init = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(V, 1).to(dev)
    go = torch.randn(1, V, C, S, S, 1, device=dev)
    for dt in (torch.float32, torch.bfloat16):
        # a fresh leaf per dtype: soft32.to(float32) would alias soft32, and once that requires grad the bf16 copy made from
        # it becomes a NON-leaf whose backward also casts and accumulates into soft32.grad (0.74 ms of torch kernels)
        soft = soft32.detach().to(dt, copy=True).requires_grad_(True)

        def ours():
            soft.grad = None; params.grad = None
            ys, _, _, ga, _, _ = afb.acquire_views(soft, None, None, nii, gpre, params, init, offset_clip=0.2, zoom_clip=0.0, spat=S,
                                                   slice_fov_mm=[192.0, 192.0, 0.75], slice_fov_vox=[S, S, 1])
            ys.backward(go.to(dt))
        t, tmin = timeit(ours, reps=5, warm=2)
        e = 4 if dt == torch.float32 else 2
        out.append({"config": f"cfg5: 1 x 8 x 256^3 {str(dt).split('.')[-1]} storage, 16 views 256^2, fwd+bwd wrt volume and theta",
                    "ms": t, "ms_min": tmin, "value": V * 1e3 / t, "unit": "slices/s",
                    "bytes": V * S * S * C * (8 * e + e + 8 * e + 4 + 32) + 2 * C * S ** 3 * 4})
        del soft
        torch.cuda.empty_cache()
    return out


def main():
    import acquisition_focus_b200 as afb
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg5"]
    for name in which:
        res = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg5": cfg5}[name](afb, dev)
        for r in (res if isinstance(res, list) else [res]):
            if "bytes" in r and "ms" in r:
                r["gbs"] = r["bytes"] / r["ms"] / 1e6
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
