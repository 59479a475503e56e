#!/usr/bin/env python
"""Does the soft-label min pass overlap with the label / image slicings when it runs as the low-occupancy TMA-streamed
kernel (AFB_MIN_TMA=<stages>)?  Prints stand-alone times, checks the record bit for bit, then times
[min pass || label fwd + image min + image fwd] on two streams against the same work on one stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from acquisition_focus_b200 import functional as AF, _lib as L  # noqa: E402

nv, V = int(os.environ.get("AB_VOLUMES", "64")), 6
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
h = bench.make_host_inputs(nv, V, seed=1000)
label, soft = bench.one_hot_volumes(h["lab"].to(dev))
image = h["img"].to(dev)
nii, gpre, init = h["nii"].to(dev), h["gpre"].to(dev), h["init"].to(dev)
params = h["params"].to(dev)
S, R, NP = bench.S, bench.R, bench.NP
fov_mm, fov_vox = [192.0, 192.0, 1.5], [S, S, 1]
spec = AF.ViewSpec(kind=L.AFFINE_PARAMS, V=V, gpre=gpre.reshape(nv * V, 4, 4).contiguous(), init=init, R=R, spat=S,
                   offset_clip=bench.OFFSET_CLIP, zoom_clip=bench.ZOOM_CLIP, nii_affine=nii, fov_mm=tuple(fov_mm),
                   params=params.reshape(nv * V, NP).contiguous())
spec = AF.prepare_views(spec, nv, (S, S, S), fov_vox, dev)[0]

ref = AF.volume_min(soft, with_mask=True)
ref_mask = ref._afb_mask.clone()
print("default min_mask: %.4f ms" % bench._time(lambda: AF.volume_min(soft, with_mask=True), dev), ref.tolist())
side = torch.cuda.Stream()


def others():
    AF._slice_forward_raw(label, spec, fov_vox, L.NEAREST, L.PAD_ZERO, 0.0, None)
    pad_i = AF.volume_min(image)
    AF._slice_forward_raw(image, spec, fov_vox, L.BILINEAR, L.PAD_DEVICE, 0.0, pad_i)


def sequential():
    AF.volume_min(soft, with_mask=True)
    others()


def overlapped():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    AF.volume_min(soft, with_mask=True)
    with torch.cuda.stream(side):
        others()
    main.wait_stream(side)


print("others alone: %.4f ms" % bench._time(others, dev))
print("default sequential: %.4f ms   overlapped: %.4f ms" % (bench._time(sequential, dev), bench._time(overlapped, dev)))
for cfg in sys.argv[1:]:
    stages, ctas = cfg.split("x")
    os.environ["AFB_MIN_TMA"], os.environ["AFB_MIN_TMA_CTAS"] = stages, ctas
    out = AF.volume_min(soft, with_mask=True)
    ok = torch.equal(out, ref) and torch.equal(out._afb_mask, ref_mask)
    t = bench._time(lambda: AF.volume_min(soft, with_mask=True), dev)
    print(f"TMA stages={stages} ctas/SM={ctas}: {t:.4f} ms  record identical: {ok}   sequential {bench._time(sequential, dev):.4f}"
          f"   overlapped {bench._time(overlapped, dev):.4f}", flush=True)
