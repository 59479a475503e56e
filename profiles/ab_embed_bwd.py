#!/usr/bin/env python
"""Embedding backward kernel alone (direct C-ABI calls, no autograd glue), cfg3 at B=2, V=6, all six stages in one launch and
stage by stage.  AFB_EMBED_BWD_VARIANT picks the launch shape: 0 = 256 threads x 2 CTAs/SM (128 regs, the default),
1 = 128 x 4 (128 regs); [2 = 128 x 5 (96 regs) and 3 = 256 x 3 (80 regs) were measured once and removed from the library].
    for v in 0 1; do AFB_EMBED_BWD_VARIANT=$v python profiles/ab_embed_bwd.py; done > gpurun_out/r2_ab_embed_bwd.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from acquisition_focus_b200 import _lib as L, functional as F, synthetic as cases  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B, V = 2, 6
stages = ((16, 128), (32, 64), (64, 32), (128, 16), (256, 8), (256, 4))
aff = torch.stack([a.to(dev) for a in cases.embed_case(128, 16, V, B, seed=300)["affines"]], 0).float().contiguous()
xs = [cases.randn((B, V * c, S, S), 500 + S).to(dev) for c, S in stages]
gos = [torch.randn(B, V * c, S, S, S, device=dev) for c, S in stages]
dxs = [torch.empty_like(x) for x in xs]
da = torch.zeros_like(aff)
lib = L.lib()
ws = torch.zeros(int(lib.afb_embed_workspace_bytes(B * V)), dtype=torch.uint8, device=dev)


def run(idx):
    cs, Ss = [stages[i][0] for i in idx], [stages[i][1] for i in idx]
    args = (len(idx), F._ptr_array([gos[i] for i in idx]), F._ptr_array([xs[i] for i in idx]), F._int_array(cs), F._int_array(Ss),
            F._ptr_array([dxs[i] for i in idx]), L.ptr(aff), B, V, L.ptr(da), L.ptr(ws), L.stream_ptr(dev))
    return lambda: L.check(lib.afb_embed_multi_bwd(*args), "afb_embed_multi_bwd")


res = {"variant": int(os.environ.get("AFB_EMBED_BWD_VARIANT", "0")), "workload": f"cfg3 B={B} V={V}",
       "all_stages_ms": bench._time(run(list(range(len(stages)))), dev),
       "per_stage_ms": {f"c{c}_S{S}": bench._time(run([i]), dev) for i, (c, S) in enumerate(stages)}}
# algorithmic bytes: every d_x element written once, ~8 gradient voxels (one 32-B sector each at worst) read per (pixel, channel)
res["dx_bytes"] = sum(x.numel() * 4 for x in xs)
print(json.dumps(res))
