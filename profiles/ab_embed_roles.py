#!/usr/bin/env python
"""Stage-0 embedding forward (B=2, V=6, c=16, S=128): the two roles of the role-split kernel alone and together, the round-1
zero + slab kernel pair, and torch.zeros of the same size (the HBM write ceiling)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import acquisition_focus_b200 as afb
from acquisition_focus_b200 import synthetic as cases

dev = torch.device("cuda", 0)
B, V, c, S = 2, 6, 16, 128
case = cases.embed_case(S, c, V, B, seed=300)
aff = torch.stack([a.to(dev) for a in case["affines"]], 0)
x = case["x"].to(dev)
nb = B * V * c * S ** 3 * 4
res = {}
def t(fn):
    return bench._time(fn, dev, reps=10, warm=3)
res["torch.zeros same size"] = t(lambda: torch.zeros(B, V * c, S, S, S, device=dev))
os.environ["AFB_EMBED_LEGACY"] = "1"; res["legacy zero+slab kernels"] = t(lambda: afb.embed_slices(x, aff, V)); os.environ.pop("AFB_EMBED_LEGACY")
res["role-split, both roles"] = t(lambda: afb.embed_slices(x, aff, V))
for r in ("zero", "slab"):
    os.environ["AFB_EMBED_ROLE"] = r
    res[f"role-split, {r} role only"] = t(lambda: afb.embed_slices(x, aff, V))
os.environ.pop("AFB_EMBED_ROLE")
os.environ["AFB_EMBED_ROLE"] = "zero"; os.environ["AFB_EMBED_NOPRED"] = "1"
res["role-split, zero role only, NO slab test (stores everything)"] = t(lambda: afb.embed_slices(x, aff, V))
os.environ.pop("AFB_EMBED_ROLE"); os.environ.pop("AFB_EMBED_NOPRED")
os.environ["AFB_EMBED_NO_PDL"] = "1"; res["role-split, roles back to back (no programmatic dependent launch)"] = t(lambda: afb.embed_slices(x, aff, V)); os.environ.pop("AFB_EMBED_NO_PDL")
for z, sl in (("2", "2"), ("3", "1"), ("2", "1"), ("4", "1"), ("4", "4"), ("3", "3")):
    os.environ["AFB_EMBED_ZERO_CTAS_PER_SM"] = z; os.environ["AFB_EMBED_SLAB_CTAS_PER_SM"] = sl
    res[f"role-split, both roles, {z} zero + {sl} slab CTAs per SM"] = t(lambda: afb.embed_slices(x, aff, V))
    os.environ["AFB_EMBED_ROLE"] = "slab"
    res[f"role-split, slab role only, {sl} slab CTAs per SM"] = t(lambda: afb.embed_slices(x, aff, V))
    os.environ.pop("AFB_EMBED_ROLE")
os.environ.pop("AFB_EMBED_ZERO_CTAS_PER_SM"); os.environ.pop("AFB_EMBED_SLAB_CTAS_PER_SM")
os.environ["AFB_EMBED_NOSPLIT"] = "1"; res["zero-then-patch per CTA (v2)"] = t(lambda: afb.embed_slices(x, aff, V)); os.environ.pop("AFB_EMBED_NOSPLIT")
print(json.dumps({"bytes": nb, "ms": res, "gbs": {k: nb / v / 1e6 for k, v in res.items()}}, indent=1))
