#!/usr/bin/env python
"""Stage-0 embedding forward (zero kernel + slab kernel): channels per slab CTA (AFB_EMBED_SLAB_CH)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import acquisition_focus_b200 as afb
from acquisition_focus_b200 import synthetic as cases
dev = torch.device("cuda", 0)
res = {}
for (c, S) in ((16, 128), (32, 64), (64, 32)):
    B, V = 2, 6
    case = cases.embed_case(S, c, V, B, seed=300 + S)
    aff = torch.stack([a.to(dev) for a in case["affines"]], 0)
    x = case["x"].to(dev)
    for ch in ("1", "2", "4", "8", "16", "64"):
        os.environ["AFB_EMBED_SLAB_CH"] = ch
        res[f"S={S} c={c} slab channels per CTA={ch}"] = bench._time(lambda: afb.embed_slices(x, aff, V), dev, reps=10, warm=3)
os.environ.pop("AFB_EMBED_SLAB_CH")
print(json.dumps(res, indent=1))
