#!/usr/bin/env python
"""All-stages embedding forward (afb_embed_multi_fwd, one launch): output bytes per CTA (AFB_EMBED_FWD_CHUNK_KB), cfg3 B=2 V=6.
    python profiles/ab_embed_chunk.py > gpurun_out/r2_ab_embed_chunk.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import acquisition_focus_b200 as afb  # noqa: E402
from acquisition_focus_b200 import synthetic as cases  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B, V = 2, 6
stages = ((16, 128), (32, 64), (64, 32), (128, 16), (256, 8), (256, 4))
aff = torch.stack([a.to(dev) for a in cases.embed_case(128, 16, V, B, seed=300)["affines"]], 0)
xs = [cases.randn((B, V * c, S, S), 500 + S).to(dev) for c, S in stages]
nbytes = sum(B * V * c * S ** 3 * 4 + B * V * c * S * S * 4 for c, S in stages)
res = {"workload": f"cfg3 B={B} V={V}, all six stages, one launch", "bytes": nbytes, "ms": {}, "stage0_only_ms": {}}
for kb in (16, 32, 64, 128, 256, 512):
    os.environ["AFB_EMBED_FWD_CHUNK_KB"] = str(kb)
    res["ms"][str(kb)] = bench._time(lambda: afb.embed_slices_multi(xs, aff, V), dev)
    res["stage0_only_ms"][str(kb)] = bench._time(lambda: afb.embed_slices_multi(xs[:1], aff, V), dev)
os.environ.pop("AFB_EMBED_FWD_CHUNK_KB")
best = min(res["ms"], key=res["ms"].get)
res["best_kb"] = best
res["best_gbs"] = nbytes / res["ms"][best] / 1e6
res["best_frac_of_hbm"] = res["best_gbs"] / bench._hbm_peak()[0]
print(json.dumps(res, indent=1))
