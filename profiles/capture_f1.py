#!/usr/bin/env python
"""Three launches of the f1 sampler (128^3 -> 128^3, C = 8 fp32 channels-last, B = 8) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:slice_fwd_cl -s 2 -c 1 -o gpurun_out/f1 python profiles/capture_f1.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from acquisition_focus_b200 import functional as AF, _lib as L, synthetic as cases  # noqa: E402

dev = torch.device("cuda", 0)
S, B = 128, 8
case = cases.atm_case(S, B, 3, seed=47)
sd = case["soft"].to(dev)
pad0 = AF.volume_min(sd)
spec = AF.ViewSpec(kind=L.AFFINE_PRE, V=1, nii_affine=case["nii"].to(dev), fov_mm=(192.0, 192.0, 192.0), pre=case["gpre"][0].to(dev).contiguous())
spec = AF.prepare_views(spec, B, (S, S, S), [S, S, S], dev)[0]
for _ in range(3):
    AF._slice_forward_raw(sd, spec, [S, S, S], L.BILINEAR, L.PAD_DEVICE, 0.0, pad0)
torch.cuda.synchronize()
