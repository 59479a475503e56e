#!/usr/bin/env python
"""Round 2, VERDICT item 5: the embedding forward writing every sector once, and all six stages in one launch.
cfg3 at B=2, V=6.      python profiles/ab_embed.py > gpurun_out/r2_ab_embed.json"""
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
case0 = cases.embed_case(128, 16, V, B, seed=300)
gas = [a.to(dev).requires_grad_(True) for a in case0["affines"]]
aff = torch.stack(gas, 0)
xs = [cases.randn((B, V * c, S, S), 500 + S).to(dev).requires_grad_(True) for c, S in stages]
gos = [torch.randn(B, V * c, S, S, S, device=dev) for c, S in stages]
res = {"workload": f"cfg3 B={B} V={V}, stages {stages}", "stages": [], "ms": {}}
hbm = bench._hbm_peak()[0]
tot = {"legacy_fwd": 0.0, "single_pass_fwd": 0.0, "per_stage_fwd_bwd": 0.0}
for (c, S), x, go in zip(stages, xs, gos):
    nb = B * V * c * S ** 3 * 4 + B * V * c * S * S * 4
    t_old = bench._time(lambda: afb.embed_slices(x.detach(), aff.detach(), V), dev)          # zero kernel + slab kernel
    os.environ["AFB_EMBED_SINGLE_PASS"] = "1"
    t_new = bench._time(lambda: afb.embed_slices(x.detach(), aff.detach(), V), dev)          # zero-then-patch per CTA
    os.environ.pop("AFB_EMBED_SINGLE_PASS")

    def fb():
        x.grad = None
        for a in gas:
            a.grad = None
        afb.embed_slices(x, torch.stack(gas, 0), V).backward(go)
    t_fb = bench._time(fb, dev)
    res["stages"].append({"c": c, "S": S, "bytes_fwd": nb, "legacy_fwd_ms": t_old, "single_pass_fwd_ms": t_new, "legacy_gbs": nb / t_old / 1e6,
                          "single_pass_gbs": nb / t_new / 1e6, "single_pass_frac_of_hbm": nb / t_new / 1e6 / hbm, "fwd_bwd_ms": t_fb})
    tot["legacy_fwd"] += t_old; tot["single_pass_fwd"] += t_new; tot["per_stage_fwd_bwd"] += t_fb
res["ms"].update({"all stages forward, legacy (6 x 3 launches)": tot["legacy_fwd"], "all stages forward, single pass per stage (6 x 2 launches)": tot["single_pass_fwd"],
                  "all stages fwd+bwd, stage by stage": tot["per_stage_fwd_bwd"]})
xd = [x.detach() for x in xs]
res["ms"]["all stages forward, ONE launch (afb_embed_multi_fwd)"] = bench._time(lambda: afb.embed_slices_multi(xd, aff.detach(), V), dev)


def fb_all():
    for x in xs:
        x.grad = None
    for a in gas:
        a.grad = None
    outs = afb.embed_slices_multi(xs, torch.stack(gas, 0), V)
    torch.autograd.backward(outs, gos)
res["ms"]["all stages fwd+bwd, ONE launch each (afb_embed_multi_fwd/bwd)"] = bench._time(fb_all, dev)


def bwd_only(idx):
    """backward of the stages in idx alone (events around autograd.backward; the forward runs outside the timed region)"""
    ts = []
    for it in range(13):
        for x in xs:
            x.grad = None
        for a in gas:
            a.grad = None
        outs = afb.embed_slices_multi([xs[i] for i in idx], torch.stack(gas, 0), V)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.autograd.backward(outs, [gos[i] for i in idx])
        e1.record()
        torch.cuda.synchronize(dev)
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
res["ms"]["all stages backward only (one launch + autograd glue)"] = bwd_only(range(len(stages)))
res["bwd_only_ms_per_stage"] = {f"c{c}_S{S}": bwd_only([i]) for i, (c, S) in enumerate(stages)}
tot_bytes = sum(s["bytes_fwd"] for s in res["stages"])
res["all_stages_fwd_gbs_one_launch"] = tot_bytes / res["ms"]["all stages forward, ONE launch (afb_embed_multi_fwd)"] / 1e6
res["all_stages_fwd_frac_of_hbm"] = res["all_stages_fwd_gbs_one_launch"] / hbm
print(json.dumps(res, indent=1))
