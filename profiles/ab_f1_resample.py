#!/usr/bin/env python
"""SURVEY 8 f1, the 3-D -> 3-D prescan resample (128^3 -> 128^3, C = 8 fp32 channels-last, B = 8): sampler kernel alone,
LDG.128 x 2 per corner (the slices' choice) vs one LDG.256 per corner (AFB_FWD3D_VB=32), wide 32x1 warp rows vs 8x4 patches.
    python profiles/ab_f1_resample.py > gpurun_out/r2_ab_f1_resample.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from acquisition_focus_b200 import functional as AF, _lib as L, synthetic as cases  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
S, C_, B, V = 128, 8, 8, 3
case = cases.atm_case(S, B, V, seed=47)
sd = case["soft"].to(dev)
pad0 = AF.volume_min(sd)
res = {"workload": f"128^3 -> 128^3, C=8 fp32 channels-last, B={B}", "bytes": 2 * B * C_ * S ** 3 * 4, "ms": {}}
outs = {}
for v in range(V):
    spec = AF.ViewSpec(kind=L.AFFINE_PRE, V=1, nii_affine=case["nii"].to(dev), fov_mm=(192.0, 192.0, 192.0), pre=case["gpre"][v].to(dev).contiguous())
    spec = AF.prepare_views(spec, B, (S, S, S), [S, S, S], dev)[0]
    for vb in ("16", "32"):
        for narrow in ("0", "1"):
            os.environ["AFB_FWD3D_VB"] = vb
            if narrow == "1":
                os.environ["AFB_NO_WIDE_PATCH"] = "1"
            else:
                os.environ.pop("AFB_NO_WIDE_PATCH", None)
            fn = lambda: AF._slice_forward_raw(sd, spec, [S, S, S], L.BILINEAR, L.PAD_DEVICE, 0.0, pad0)
            outs[(v, vb, narrow)] = fn().clone()
            res["ms"][f"view {v}: VB={vb}, {'8x4 patch' if narrow == '1' else '32x1 rows'}"] = bench._time(fn, dev)
res["bitwise_equal"] = all(torch.equal(outs[(v, "16", "0")], outs[(v, vb, nr)]) for v in range(V) for vb in ("16", "32") for nr in ("0", "1"))
print(json.dumps(res, indent=1))
