#!/usr/bin/env python
"""Launches the embedding kernels (all six stages in one launch each way, stage 0 alone) and the one-hot-from-index kernels once:
    ncu --set full --clock-control none -k 'regex:^(embed_|onehot_)' --launch-count 12 -f -o gpurun_out/r2_prof_embed_onehot python profiles/capture_embed_onehot.py"""
import sys
sys.path.insert(0, ".")
import torch
import acquisition_focus_b200 as afb
from acquisition_focus_b200 import synthetic as cases
dev = torch.device("cuda", 0)
B, V = 2, 6
case0 = cases.embed_case(128, 16, V, B, seed=300)
gas = [a.to(dev).requires_grad_(True) for a in case0["affines"]]
cfgs = ((16, 128), (32, 64), (64, 32), (128, 16), (256, 8), (256, 4))
xs = [cases.randn((B, V * c, S, S), 500 + S).to(dev).requires_grad_(True) for c, S in cfgs]
outs = afb.embed_slices_multi(xs, torch.stack(gas, 0), V)                      # one launch: all six stages
torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])              # one launch: all six stages
o0 = afb.embed_slices(xs[0], torch.stack(gas, 0), V)                            # stage 0 alone: zero kernel + slab kernel
case = cases.atm_case(128, 8, 6, seed=43)
params = torch.stack(case["params"], 1).to(dev).requires_grad_(True)
ys, yl, yi, ga, nii, th = afb.acquire_views_from_labels(case["lab"].to(dev).to(torch.uint8), case["image"].to(dev), case["nii"].to(dev),
                                                        torch.stack(case["gpre"], 1).to(dev), params,
                                                        torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(6, 1).to(dev), num_classes=8,
                                                        offset_clip=0.2, zoom_clip=0.0, spat=128, slice_fov_mm=[192.0, 192.0, 1.5],
                                                        slice_fov_vox=[128, 128, 1])
ys.sum().backward()
torch.cuda.synchronize()
