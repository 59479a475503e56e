#!/usr/bin/env python
"""A/B of kernel variants selected through AFB_* environment knobs, on the bench workload (64 volumes x 6 views).
Run on the GPU box:  python profiles/ab_kernels.py "AFB_FWD_VARIANT=0,1,2" "AFB_FILL_VARIANT=0,1,2,3" ...
Each argument is one knob swept alone (the others unset); prints ms per kernel of bench.kernel_breakdown."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from acquisition_focus_b200 import functional as AF  # noqa: E402

nv, V = int(os.environ.get("AB_VOLUMES", "64")), 6
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
h = bench.make_host_inputs(nv, V, seed=1000)
label, soft = bench.one_hot_volumes(h["lab"].to(dev))
image = h["img"].to(dev)
nii, gpre, init = h["nii"].to(dev), h["gpre"].to(dev), h["init"].to(dev)
params = h["params"].to(dev).requires_grad_(True)
go = torch.cos(torch.arange(nv * V * bench.NUM_CLASSES * bench.S * bench.S, device=dev, dtype=torch.float32) * 0.618).view(
    nv, V, bench.NUM_CLASSES, bench.S, bench.S, 1)
fov_mm, fov_vox = [192.0, 192.0, 1.5], [bench.S, bench.S, 1]

z = torch.empty_like(soft)
print("torch zero_ of dVolume size: %.3f ms" % bench._time(lambda: z.zero_(), dev))
print("torch fill_(1) of dVolume size: %.3f ms" % bench._time(lambda: z.fill_(1.0), dev))
del z


def run(tag):
    out, l2 = bench.kernel_breakdown(AF, dev, soft, label, image, nii, gpre, params, init, go, fov_mm, fov_vox, nv, V)
    print(tag, json.dumps({k.split("(")[0] + ("*" if "not in step" in k else "") + ("/" + k.split("(")[1][:5] if "slice_fwd" in k else ""): round(v["ms"], 4)
                           for k, v in out.items()}), flush=True)


run("baseline")
for arg in sys.argv[1:]:
    knob, vals = arg.split("=")
    for val in vals.split(","):
        os.environ[knob] = val
        run(f"{knob}={val}")
    del os.environ[knob]
