#!/usr/bin/env python
"""Host-side int64 -> uint8 label narrowing (afb_host_narrow_labels) on the GPU box's host cores: GB/s of int64 read vs threads,
pinned source and destination, next to torch's own copy_ and to the H2D copy of the same int64 buffer."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from acquisition_focus_b200 import _lib as L  # noqa: E402

lib = L.lib()
nv = int(os.environ.get("AB_VOLUMES", "64"))
n = nv * 128 ** 3
src = torch.randint(0, 8, (n,), dtype=torch.int64).pin_memory()
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
bad = C.c_int(0)


def tm(f, reps=5):
    f()
    t = time.perf_counter()
    for _ in range(reps):
        f()
    return (time.perf_counter() - t) / reps * 1e3


res = {"volumes": nv, "int64_bytes": n * 8, "cpus": len(os.sched_getaffinity(0)), "torch_threads": torch.get_num_threads(), "ms": {}}
res["ms"]["torch copy_ (int64 -> uint8)"] = tm(lambda: dst.copy_(src))
for nt in (1, 2, 4, 8, 12, 16, 24, 32):
    if nt > 2 * res["cpus"]:
        break
    res["ms"][f"afb_host_narrow_labels, {nt} threads"] = tm(
        lambda: lib.afb_host_narrow_labels(src.data_ptr(), L.DTYPES[torch.int64], n, dst.data_ptr(), nt, C.byref(bad)))
if torch.cuda.is_available():
    d64 = torch.empty(n, dtype=torch.int64, device="cuda")
    d8 = torch.empty(n, dtype=torch.uint8, device="cuda")

    def h2d(dst_d, src_h):
        dst_d.copy_(src_h, non_blocking=True)
        torch.cuda.synchronize()
    res["ms"]["H2D int64 (pinned)"] = tm(lambda: h2d(d64, src))
    res["ms"]["H2D uint8 (pinned)"] = tm(lambda: h2d(d8, dst))
res["gbs_int64_read"] = {k: n * 8 / v / 1e6 for k, v in res["ms"].items() if "uint8 (pinned)" not in k}
print(json.dumps(res, indent=1))
