#!/usr/bin/env python
"""Round 2, VERDICT item 4: one fused forward launch (afb_slice_fwd3) vs the three launches, stand-alone and inside the step.
Bench workload (64 volumes x 6 views).      python profiles/ab_fwd3.py > gpurun_out/r2_ab_fwd3.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from acquisition_focus_b200 import functional as AF  # noqa: E402
from acquisition_focus_b200 import parallel as par  # noqa: E402
from acquisition_focus_b200 import _lib as L  # noqa: E402

nv, V, S = int(os.environ.get("AB_VOLUMES", "64")), 6, bench.S
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
wl = bench.Workload(AF, par, dev, nv, V, seed=1000, world=1)
nS = nv * V
spec = AF.ViewSpec(kind=L.AFFINE_PARAMS, V=V, gpre=wl.gpre.reshape(nS, 4, 4).contiguous(), init=wl.init, R=bench.R, spat=S,
                   offset_clip=bench.OFFSET_CLIP, zoom_clip=bench.ZOOM_CLIP, nii_affine=wl.nii, fov_mm=(192.0, 192.0, 1.5),
                   params=wl.params.detach().reshape(nS, bench.NP).contiguous())
spec = AF.prepare_views(spec, nv, (S, S, S), wl.fov_vox, dev)[0]
sd = wl.soft.detach()
pad_s, pad_i = AF.volume_min(sd), AF.volume_min(wl.image)
res = {"workload": f"{nv} volumes x {V} views", "ms": {}}
t_s = bench._time(lambda: AF._slice_forward_raw(sd, spec, wl.fov_vox, L.BILINEAR, L.PAD_DEVICE, 0.0, pad_s), dev)
t_l = bench._time(lambda: AF._slice_forward_raw(wl.label, spec, wl.fov_vox, L.NEAREST, L.PAD_ZERO, 0.0, None), dev)
t_i = bench._time(lambda: AF._slice_forward_raw(wl.image, spec, wl.fov_vox, L.BILINEAR, L.PAD_DEVICE, 0.0, pad_i), dev)
res["ms"].update({"slice_fwd soft": t_s, "slice_fwd label": t_l, "slice_fwd image": t_i, "three launches, summed": t_s + t_l + t_i})
res["ms"]["afb_slice_fwd3 (one launch)"] = bench._time(
    lambda: AF._slice_forward3_raw(sd, wl.label, wl.image, spec, wl.fov_vox, (L.PAD_DEVICE, 0.0, pad_s), (L.PAD_DEVICE, 0.0, pad_i)), dev)
res["ms"]["afb_slice_fwd3 soft + image only"] = bench._time(
    lambda: AF._slice_forward3_raw(sd, None, wl.image, spec, wl.fov_vox, (L.PAD_DEVICE, 0.0, pad_s), (L.PAD_DEVICE, 0.0, pad_i)), dev)


def step(fused):
    wl.soft.grad = None
    wl.params.grad = None
    ys, yl, yi, ga, _, _ = AF.acquire_views(wl.soft, wl.label, wl.image, wl.nii, wl.gpre, wl.params, wl.init, fused_forward=fused, **wl.kw)
    torch.autograd.backward([ys], [wl.go])
    return par.reduce_view_grads(wl.params.grad)


for fused in (False, True, False, True):
    res["ms"].setdefault(f"whole step, fused_forward={fused}", []).append(bench._time(lambda: step(fused), dev, reps=20, warm=5))
print(json.dumps(res, indent=1))
