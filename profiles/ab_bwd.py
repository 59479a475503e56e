#!/usr/bin/env python
"""Round 2, VERDICT item 3: the backward slicer measured, not argued.  On the bench workload (64 volumes x 6 views, C = 8 fp32
channels-last), CUDA events, mean of 5 after 2 warm-ups:

  * slice_bwd (dVolume + dTheta) at 128 registers / 2 CTAs per SM (product) vs 80 registers / 3 CTAs per SM with LDG.256 or
    LDG.128 gathers (AFB_BWD_VARIANT=1/2);
  * the scatter alone with 16 x red.global.add.v4.f32 per pixel (slice_scatter_kernel) vs the shared-memory privatisation
    probe (64 shared-memory fp32 atomics per pixel = ATOMS.CAST.SPIN loops on sm_100, collisions ignored, coalesced flush):
    an upper bound on what any exact privatised scheme could reach.

    python profiles/ab_bwd.py > gpurun_out/r2_ab_bwd.json
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from acquisition_focus_b200 import functional as AF  # noqa: E402
from acquisition_focus_b200 import _lib as L  # noqa: E402

nv, V, S, NC = int(os.environ.get("AB_VOLUMES", "64")), 6, bench.S, bench.NUM_CLASSES
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
h = bench.make_host_inputs(nv, V, seed=1000)
label, soft = bench.one_hot_volumes(h["lab"].to(dev))
del label
nii, gpre, init = h["nii"].to(dev), h["gpre"].to(dev), h["init"].to(dev)
params = h["params"].to(dev)
go = torch.cos(torch.arange(nv * V * NC * S * S, device=dev, dtype=torch.float32) * 0.618).view(nv, V, NC, S, S, 1)
lib, st = L.lib(), L.stream_ptr(dev)
nS = nv * V
spec = AF.ViewSpec(kind=L.AFFINE_PARAMS, V=V, gpre=gpre.reshape(nS, 4, 4).contiguous(), init=init, R=bench.R, spat=S,
                   offset_clip=bench.OFFSET_CLIP, zoom_clip=bench.ZOOM_CLIP, nii_affine=nii, fov_mm=(192.0, 192.0, 1.5),
                   params=params.reshape(nS, bench.NP).contiguous())
spec = AF.prepare_views(spec, nv, (S, S, S), [S, S, 1], dev)[0]
pad = AF.volume_min(soft)
d_vol = torch.zeros_like(soft)
ws = torch.zeros(int(lib.afb_slice_bwd_workspace_bytes(nS)), dtype=torch.uint8, device=dev)
d_aff = torch.zeros(nS, bench.NP, device=dev)
vd, vs = L.volume_desc(soft), spec.struct()


def bwd(with_dvol):
    L.check(lib.afb_slice_bwd(C.byref(vd), C.byref(vs), S, S, 1, L.PAD_DEVICE, 0.0, L.ptr(pad), L.ptr(go), None,
                              L.ptr(d_vol) if with_dvol else None, L.ptr(d_aff), None, None, L.ptr(ws), st), "afb_slice_bwd")


res = {"workload": f"{nv} volumes x {V} views, C=8 fp32 channels-last, 128^3 -> 128^2", "ms": {}}
for var, name in ((None, "product: LDG.256, 128 regs, 2 CTAs/SM"), ("1", "LDG.256, <=80 regs (spills), 3 CTAs/SM"),
                  ("2", "LDG.128, <=80 regs (spills), 3 CTAs/SM")):
    if var is None:
        os.environ.pop("AFB_BWD_VARIANT", None)
    else:
        os.environ["AFB_BWD_VARIANT"] = var
    res["ms"][f"slice_bwd dVolume+dTheta [{name}]"] = bench._time(lambda: bwd(True), dev)
    res["ms"][f"slice_bwd dTheta only [{name}]"] = bench._time(lambda: bwd(False), dev)
os.environ.pop("AFB_BWD_VARIANT", None)
res["ms"]["scatter only: 16 x red.global.add.v4.f32 per pixel (slice_scatter_kernel)"] = bench._time(
    lambda: L.check(lib.afb_slice_scatter(C.byref(vd), C.byref(vs), S, S, 1, L.ptr(go), L.ptr(d_vol), st), "scatter"), dev)
res["ms"]["scatter only: smem privatisation probe (64 ATOMS.CAST.SPIN per pixel, collisions ignored, coalesced flush)"] = bench._time(
    lambda: L.check(lib.afbx_slice_scatter_priv_probe(C.byref(vd), C.byref(vs), S, S, 1, L.ptr(go), L.ptr(d_vol), st), "probe"), dev)
res["red_lane_ops"] = nS * S * S * 16
print(json.dumps(res, indent=1))
