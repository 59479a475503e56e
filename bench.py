#!/usr/bin/env python
"""bench.py - slices/sec of the differentiable view-acquisition hot path (fwd + bwd, 128^3 -> 128^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak] [--volumes NV] [--views V]

One "step" = one pass of the hot path over one batch of synthetic input, with the reference's exact
semantics (BASELINE.json configs[1] content at configs[3] scale).  configs[3] reads "64 synthetic volumes x 6 views SHARDED
ACROSS 1/2/4/8 B200", so the default is STRONG scaling: NV = 64 volumes in total, partitioned by volume over the ranks
(`parallel.shard_range`); `--scaling weak` keeps NV volumes per GPU (reported as a second key by the default run at N > 1).

    per rank: its shard of the NV volumes x V views;  per volume an 8-class one-hot soft label [8,128^3] fp32 (bilinear, WITH
    gradient w.r.t. the volume and the view parameters), the same one-hot as int64 (nearest) and a 1-channel
    fp32 image (bilinear), all sliced to 128x128x1 from raw view parameters (R6 | 3x26 offset logits | zoom)
    composed with an augmented p2CH clinical affine:
      min pass (volume.min() of the reference's min-shift) -> fused-prologue slice forward x3 ->
      backward (dVolume scatter + dTheta reduce + analytic parameter chain) -> MinBackward pass
      -> [N>1] the whole-batch pad exchange (one all-gather of 4 floats + one all-reduce of d(out)/d(pad)) and one NCCL
         all-reduce of the [V,85] view-parameter gradients.
The per-rank step (collectives included) is captured once into a CUDA graph (`graphs.GraphedStep`) and replayed: at 8
volumes per GPU the eager step is bound by ~0.5 ms of host launch time, not by the GPU.

`value`  : slices/s with inputs resident in HBM (device timed, CUDA events, max over ranks).
`e2e`    : same metric through the public API from HOST buffers: pinned int64 index-label + image volumes are (labels packed to
           uint8 on the host cores, then) copied H2D inside the timed region, expanded to one-hot on the device (as
           running/run_dl.py:261-264 does),
           and the reduced parameter gradients + grid affines are read back D2H every step.
`--impl reference` times the oracle port of the reference's own torch-CPU path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

NUM_CLASSES = 8
S = 128
OFFSET_CLIP, ZOOM_CLIP = 0.2, 0.0
R = 26                                  # round(0.2 * 128), models/learnable_transform.py:112-115
NP = 6 + 3 * R + 1
METRIC = "slices/sec (fwd+bwd 128^3->128^2 sampling)"
UNIT = "slices/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--volumes", type=int, default=64, help="volumes in total (strong scaling) / per GPU (weak scaling)")
    ap.add_argument("--graph", default="on", choices=["on", "off"], help="replay the step from a CUDA graph")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the additional weak-scaling measurement")
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--e2e-steps", type=int, default=24, help="steps of the e2e leg (its timed region includes the pipeline fill)")
    ap.add_argument("--e2e-group", type=int, default=8, help="volumes per PCIe upload group of the e2e leg")
    ap.add_argument("--e2e-depth", type=int, default=3, help="device buffer sets of the e2e upload pipeline (batches in flight)")
    ap.add_argument("--e2e-narrow", type=int, default=-1,
                    help="1: pack the int64 host label maps to uint8 on the host cores before the upload; 0: upload int64; 2: split upload "
                         "(the cores pack a share of each batch while the link carries the rest as int64; share adapted on line); "
                         "-1 (default): pack when the rank has >= 8 host cores for it (HostInputPipeline's 'auto')")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def make_host_inputs(nv: int, views: int, seed: int, size: int = S):
    """Index-label volumes [nv,size^3] int64 + image [nv,1,size^3] fp32 on the host, view parameters."""
    from acquisition_focus_b200 import synthetic as syn
    base = syn.heart_phantom(size)
    R = int(round(OFFSET_CLIP * size))
    rng = np.random.default_rng(seed)
    variants = [base, np.ascontiguousarray(base.transpose(1, 0, 2)[::-1]), np.ascontiguousarray(base[:, ::-1, :]),
                np.ascontiguousarray(base.transpose(2, 1, 0)), np.ascontiguousarray(base[::-1, :, ::-1])]
    lab = np.stack([np.roll(variants[i % len(variants)], int(rng.integers(-6, 7)), axis=int(rng.integers(0, 3)))
                    for i in range(nv)])
    means = np.array([0.05, 0.55, 0.85, 0.95, 0.80, 0.90, 0.75, 0.70], dtype=np.float32)
    img = means[lab] + 0.15 * rng.standard_normal(lab.shape, dtype=np.float32)
    gen = torch.Generator().manual_seed(seed)
    p2ch = syn.phantom_view_affines()["p2CH"]
    gpre = torch.stack([torch.stack([p2ch @ syn.random_aug_affine(gen, 0.1, 0.2, 0.0) for _ in range(views)])
                        for _ in range(nv)])                                   # [nv,V,4,4]
    r6 = torch.tensor([1.0, 0, 0, 0, 1.0, 0]) + 0.3 * torch.randn(nv, views, 6, generator=gen)
    params = torch.cat([r6, torch.randn(nv, views, 3 * R, generator=gen), torch.randn(nv, views, 1, generator=gen)], dim=-1)
    init = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(views, 1)
    nii = syn.default_nifti_affine(nv, 192.0 / size)
    return dict(lab=torch.from_numpy(lab), img=torch.from_numpy(img)[:, None], gpre=gpre, params=params, init=init, nii=nii)


def one_hot_volumes(lab):
    """running/run_dl.py:261-264: one-hot (channels-last view) and its float copy."""
    oh = torch.nn.functional.one_hot(lab, NUM_CLASSES).permute(0, 4, 1, 2, 3)
    return oh, oh.float()


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own torch-CPU path
# ------------------------------------------------------------------------------------------------
CPU_KIND_NOTE = {"reference": "the UNMODIFIED reference code vendored to oracle/_ref by oracle/vendor_ref.py: compute_rotation_matrix_"
                              "from_ortho6d + AffineTransformModule.get_init_affines/get_batch_affines + nifti_grid_sample x3 + autograd",
                 "port": "oracle port of the reference's torch-CPU path (oracle/_ref absent)"}


def _load_vendored_reference():
    """The reference package as vendored (git-ignored) under oracle/_ref by oracle/vendor_ref.py, or None."""
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "acquisition_focus")):
        return None
    try:
        os.environ["AFB_REFERENCE_ROOT"] = ref_root
        from oracle import ref_import
        ref_import.REFERENCE_ROOT = ref_root
        return ref_import.load_reference()
    except Exception as e:      # noqa: BLE001
        print(f"[bench] oracle/_ref present but not importable ({e!r}); using the port", file=sys.stderr)
        return None


class _FixedHead(torch.nn.Module):
    """Stands in for the LocalizationNet (dense 3-D convolutions, out of scope): returns the given MLP-head output."""

    def __init__(self):
        super().__init__()
        self.out = None

    def forward(self, x):
        return self.out


def make_reference_modules(R, views):
    """One UNMODIFIED reference AffineTransformModule per view (models/learnable_transform.py:62-141), LocalizationNet stubbed."""
    fov = torch.tensor([192.0, 192.0, 192.0])
    mods = []
    for _ in range(views):
        m = R.AffineTransformModule(NUM_CLASSES, fov, torch.tensor([S, S, S]), torch.tensor([192.0, 192.0, 1.5]), torch.tensor([S, S, 1]),
                                    optim_method="R6-vector", offset_clip_value=OFFSET_CLIP, zoom_clip_value=ZOOM_CLIP, view_id="p2CH")
        m.localization_net = _FixedHead()
        mods.append(m)
    return mods


def cpu_reference_step(h, b0, b1, views, go, ref=None):
    """One step of the same workload on the host.  With `ref` = (R, modules): the reference's own functions - the body of
    AffineTransformModule.forward (learnable_transform.py:259-306) minus the prescan resample that feeds the LocalizationNet
    (neither is part of the measured workload on the GPU side)."""
    lab = h["lab"][b0:b1]
    label, soft = one_hot_volumes(lab)
    soft = soft.requires_grad_(True)
    params = h["params"][b0:b1].clone().requires_grad_(True)
    fov_mm, fov_vox = torch.tensor([192.0, 192.0, 1.5]), torch.tensor([S, S, 1])
    B = b1 - b0
    loss = 0
    if ref is not None:
        R, mods = ref
        for v in range(views):
            m = mods[v]
            m.localization_net.out = params[:, v]
            ta, tt, tz = m.get_init_affines()                                   # :262-264
            ta, tt, tz = ta.repeat(B, 1, 1), tt.repeat(B, 1, 1), tz.repeat(B, 1, 1)
            ta_b, tt_b, tz_b = m.get_batch_affines(soft.detach()[:, :, :1, :1, :1])    # :267 (only the shape is read; the stub ignores x)
            theta = (tt @ tt_b) @ (ta @ ta_b) @ (tz @ tz_b)                      # :268-272
            pre = h["gpre"][b0:b1, v].to(theta) @ theta                          # :284-289
            kw = dict(target_fov_mm=fov_mm, target_fov_vox=fov_vox, pre_grid_sample_affine=pre)
            ys, ga, _ = R.nifti_grid_sample(soft, h["nii"][b0:b1], is_label=False, **kw)          # :287
            with torch.no_grad():
                R.nifti_grid_sample(label, h["nii"][b0:b1], is_label=True, **kw)                  # :295
                R.nifti_grid_sample(h["img"][b0:b1], h["nii"][b0:b1], is_label=False, **kw)       # :302
            loss = loss + (ys * go[:B, v]).sum()
    else:
        from oracle import af_oracle as O
        for v in range(views):
            theta = O.view_theta(params[:, v], h["init"][v:v + 1, :6], h["init"][v, 6:9], h["init"][v:v + 1, 9:],
                                 OFFSET_CLIP, ZOOM_CLIP, S)
            ys, yl, yi, ga, nii = O.atm_tail_forward(soft, label, h["img"][b0:b1], h["nii"][b0:b1], h["gpre"][b0:b1, v], theta,
                                                     fov_mm, fov_vox)
            loss = loss + (ys * go[:B, v]).sum()
    loss.backward()
    return params.grad.sum(0)


def time_cpu_reference(h, views, vols_per_step, budget_s, min_steps=1, warmup=1, steps=None):
    torch.set_num_threads(os.cpu_count() or 1)
    R = _load_vendored_reference()
    ref = (R, make_reference_modules(R, views)) if R is not None else None
    kind = "reference" if ref is not None else "port"
    go = torch.from_numpy(np.cos(np.arange(vols_per_step * views * NUM_CLASSES * S * S, dtype=np.float64) * 0.618).astype(np.float32)
                          ).view(vols_per_step, views, NUM_CLASSES, S, S, 1)
    for _ in range(warmup):
        cpu_reference_step(h, 0, vols_per_step, views, go, ref)
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        cpu_reference_step(h, 0, vols_per_step, views, go, ref)
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif len(times) >= min_steps and time.perf_counter() - t_start > budget_s:
            break
    return times, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    views, vps = args.views, 2
    h = make_host_inputs(vps, views, seed=0)
    # the whole run stays within a few minutes whatever --steps says: at ~0.65 s per step, 200 steps is the cap
    steps = max(1, min(args.steps, 200))
    warm = max(1, min(args.warmup, 3))
    times, kind = time_cpu_reference(h, views, vps, budget_s=0, warmup=warm, steps=steps)
    ms = float(np.mean(times)) * 1e3
    val = vps * views / (ms / 1e3)
    cores = os.cpu_count() or 1
    sample = f"{vps} volumes x {views} views (cfg2 content: soft C=8 grad + int64 one-hot label + image) per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(1, args.gpus), note="reference arm: " + CPU_KIND_NOTE[kind] + "; bounded sample: " + sample),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, world=1, note=""):
    strong = args.scaling == "strong"
    total = args.volumes if strong else args.volumes * world
    return {"workload": f"cfg4 throughput sweep with cfg2 content: {total} volumes x {args.views} views "
                        + (f"in total, sharded by volume over {world} GPU(s) (strong scaling)" if strong else
                           f"= {args.volumes} per GPU (weak scaling)")
                        + ", 128^3 -> 128x128x1, soft label C=8 fp32 one-hot (grad wrt volume+params) + int64 one-hot "
                          "label (nearest) + image C=1, raw R6/offset/zoom params composed with augmented p2CH affine",
            "volumes_total": total, "volumes_per_gpu": total / world, "views": args.views, "slice": [S, S, 1], "volume": [S, S, S],
            "l2_policy": "inputs larger than L2 (>= 0.5 GiB per tensor per GPU even at 8 volumes per GPU); no explicit flush",
            "note": note}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period); falls back
    to `nvidia-smi -lms` when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = None
        self._thread = None
        self.proc = None
        self.tmp = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            import threading
            self._stop = threading.Event()
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=self.tmp, stderr=subprocess.DEVNULL)
            time.sleep(0.5)
        except Exception:
            self.proc = None

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "source": "nvml"}
            return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(self.samples), "source": "nvml thread, 5 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in open(self.tmp.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 20"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _init_dist(dev):
    import datetime
    import torch.distributed as dist
    # keep stdout to the single JSON line: NCCL prints its version banner to stdout when NCCL_DEBUG is set in the
    # environment, so fd 1 points at stderr while the communicator is created (init + first collective)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))    # a hang dies in 2 min, not 10
        dist.barrier()
        torch.cuda.synchronize(dev)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)


def _pin_to_local_cpus(local, world):
    """N > 1: bind each rank to host cores of ITS GPU's NUMA node (NVML's CPU affinity of the device), split evenly among the
    ranks that share the node, BEFORE any pinned staging buffer is allocated - so the buffers of the e2e leg are first-touched
    on the socket the GPU hangs off and the H2D copies do not cross the inter-socket link.  Best effort; returns a description."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        near = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = [int(v) for v in vis.split(",")] if vis and all(v.strip().isdigit() for v in vis.split(",")) else list(range(world))
            masks = {}
            for r in range(world):
                h = pynvml.nvmlDeviceGetHandleByIndex(phys[r] if r < len(phys) else r)
                words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
                cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
                masks[r] = tuple(c for c in cpus if c in set(allowed))
            mine = masks[local]
            if mine:
                peers = [r for r in range(world) if masks[r] == mine]          # ranks sharing this NUMA node
                k = peers.index(local)
                per = max(1, len(mine) // len(peers))
                near = list(mine[k * per:(k + 1) * per]) or list(mine)
        except Exception:
            near = None
        if not near:
            per = max(1, len(allowed) // world)
            near = allowed[local * per:(local + 1) * per] or allowed
            how = "even split of the allowed cores (NVML affinity unavailable)"
        else:
            how = "NVML CPU affinity of the GPU, split among the ranks on the same NUMA node"
        os.sched_setaffinity(0, near)
        torch.set_num_threads(max(1, min(len(near), 8)))
        return {"cpus": len(near), "first": near[0], "last": near[-1], "how": how}
    except Exception as e:      # noqa: BLE001
        return {"error": repr(e)}


class Workload:
    """Device-resident inputs of one rank + the step through the public API."""

    def __init__(self, AF, par, dev, nv, V, seed, world):
        self.AF, self.par, self.dev, self.nv, self.V, self.world = AF, par, dev, nv, V, world
        self.h = h = make_host_inputs(nv, V, seed=seed)
        self.host_lab = h["lab"].pin_memory()
        self.host_img = h["img"].pin_memory()
        lab_d = self.host_lab.to(dev)
        self.label, self.soft = one_hot_volumes(lab_d)
        del lab_d
        self.image = self.host_img.to(dev)
        self.nii, self.gpre, self.init = h["nii"].to(dev), h["gpre"].to(dev), h["init"].to(dev)
        self.params = h["params"].to(dev).requires_grad_(True)
        self.soft.requires_grad_(True)
        self.go = torch.cos(torch.arange(nv * V * NUM_CLASSES * S * S, device=dev, dtype=torch.float32) * 0.618).view(nv, V, NUM_CLASSES, S, S, 1)
        self.fov_mm, self.fov_vox = [192.0, 192.0, 1.5], [S, S, 1]
        self.kw = dict(offset_clip=OFFSET_CLIP, zoom_clip=ZOOM_CLIP, spat=S, slice_fov_mm=self.fov_mm, slice_fov_vox=self.fov_vox)

    def step(self):
        """The public call: min passes (reference min-shift semantics; whole-batch under sharding), the shared view prologue,
        the three slicings, the backward (dVolume + dTheta + parameter chain), the all-reduce of the view gradients."""
        self.soft.grad = None
        self.params.grad = None
        ys, yl, yi, ga, nii_o, theta = self.AF.acquire_views(
            self.soft, self.label, self.image, self.nii, self.gpre, self.params, self.init,
            pad_exchange=self.par.exchange_pads if self.world > 1 else None, **self.kw)
        torch.autograd.backward([ys], [self.go])
        g = self.par.reduce_view_grads(self.params.grad)          # [V,NP]; NCCL all-reduce when world > 1
        return g, ga, ys, yl, yi, self.soft.grad

    def free_dense(self):
        self.soft = self.label = None
        torch.cuda.empty_cache()

    def ensure_dense(self):
        if self.soft is None:
            lab_d = self.host_lab.to(self.dev)
            self.label, self.soft = one_hot_volumes(lab_d)
            del lab_d
            self.soft.requires_grad_(True)


def count_own_launches(fn, dev):
    """Kernels of libafb200.so (namespace afb::) launched by one call of `fn`, counted from a CUPTI trace of that call (outside
    every timed region); torch / NCCL helper kernels are counted separately."""
    try:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize(dev)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize(dev)
        own = other = 0
        names = {}
        for ev in prof.events():
            if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.name and "memcpy" not in ev.name.lower() \
                    and "memset" not in ev.name.lower():
                if "afb::" in ev.name:
                    own += 1
                    key = ev.name.split("afb::")[1].split("(")[0].split("<")[0]
                    names[key] = names.get(key, 0) + 1
                else:
                    other += 1
        return own, other, names
    except Exception as e:      # noqa: BLE001
        return None, None, {"error": repr(e)}


def selfcheck_sharded(AF, par, dev, rank, world):
    """N > 1: a small batch (2 volumes per rank, 32^3) sharded over the ranks must give what the unsharded call gives (each
    rank also computes the whole batch locally without collectives): forward bitwise, gradients to 1e-5."""
    import torch.distributed as dist
    Sx, V, R_ = 32, 3, 6
    nvt = 2 * world
    h = make_host_inputs(nvt, V, seed=77, size=Sx)
    h["img"][nvt - 1, 0, 3, 4, 5] = -9.0          # the image minimum lives on the LAST rank's shard only
    lab = h["lab"].to(dev)
    label, soft_full = one_hot_volumes(lab)
    image, nii, gpre, init = h["img"].to(dev), h["nii"].to(dev), h["gpre"].to(dev), h["init"].to(dev)
    go = torch.cos(torch.arange(nvt * V * NUM_CLASSES * Sx * Sx, device=dev, dtype=torch.float32) * 0.618).view(nvt, V, NUM_CLASSES, Sx, Sx, 1)
    kw = dict(offset_clip=OFFSET_CLIP, zoom_clip=ZOOM_CLIP, spat=Sx, slice_fov_mm=[192.0, 192.0, 192.0 / Sx], slice_fov_vox=[Sx, Sx, 1])

    def run(lo, hi, exchange):
        soft = soft_full[lo:hi].detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
        prm = h["params"][lo:hi].to(dev).requires_grad_(True)
        ys, yl, yi, ga, _, _ = AF.acquire_views(soft, label[lo:hi], image[lo:hi], nii[lo:hi], gpre[lo:hi], prm, init,
                                                pad_exchange=exchange, **kw)
        torch.autograd.backward([ys], [go[lo:hi]])
        return ys.detach(), yl, yi, ga.detach(), soft.grad, prm.grad
    lo, hi = par.shard_range(nvt, rank, world)
    sh = run(lo, hi, par.exchange_pads)
    g = par.reduce_view_grads(sh[5])
    full = run(0, nvt, None)
    ok = all(torch.equal(a, b[lo:hi]) for a, b in zip(sh[:4], full[:4]))
    ok = ok and (sh[4] - full[4][lo:hi]).abs().max().item() <= 1e-5 * full[4].abs().max().item()
    want = full[5].sum(0)
    ok = ok and (g - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item() == 1.0)


def time_steps(fn, steps, dev, world, sync_all):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item() / steps


def make_stepper(wl, use_graph, warmup):
    """(callable, mode): the step replayed from a CUDA graph (collectives captured with it), else eager."""
    if use_graph:
        try:
            from acquisition_focus_b200.graphs import GraphedStep
            for _ in range(2):
                wl.step()
            torch.cuda.synchronize(wl.dev)
            g = GraphedStep(wl.step, warmup=max(3, warmup), device=wl.dev)
            return g, "cuda-graph replay of the whole per-rank step (collectives captured)"
        except Exception as e:      # noqa: BLE001
            print(f"[bench] CUDA-graph capture failed ({e!r}); running the step eagerly", file=sys.stderr)
            torch.cuda.synchronize(wl.dev)
    return wl.step, "eager"


def run_ours(args):
    import torch.distributed as dist
    import acquisition_focus_b200 as afb  # noqa: F401
    from acquisition_focus_b200 import functional as AF
    from acquisition_focus_b200 import parallel as par

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus_per_rank = _pin_to_local_cpus(local, world) if world > 1 else None
    if world > 1:
        # self-destruct: a collective mismatch must cost minutes, not the GPU box's whole time limit
        import threading
        killer = threading.Timer(float(os.environ.get("AFB_BENCH_MAX_SECONDS", "420")), lambda: os._exit(3))
        killer.daemon = True
        killer.start()
        _init_dist(dev)
    V = args.views
    if args.scaling == "strong":
        lo, hi = par.shard_range(args.volumes, rank, world)
        nv, total = hi - lo, args.volumes
    else:
        nv, total = args.volumes, args.volumes * world
    if nv <= 0:
        raise SystemExit("fewer volumes than ranks")

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # N > 1: the three tiny collectives of the step as single-CTA kernels over NVLink peer memory (csrc/afb_peer.cu) instead of
    # NCCL calls, unless symmetric memory is unavailable on this box or AFB_PEER=0
    collectives = "none (single GPU)"
    peer = None
    if world > 1:
        collectives = "NCCL (torch.distributed)"
        if os.environ.get("AFB_PEER", "1") != "0":
            ok = torch.ones(1, device=dev)
            try:
                peer = par.enable_peer_collectives(dev)
            except Exception as e:      # noqa: BLE001
                print(f"[bench] rank {rank}: peer-memory collectives unavailable ({e!r}); using NCCL", file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks or none
            if ok.item() == 1.0:
                collectives = "afb_peer_collective: own single-CTA kernels over NVLink peer (symmetric) memory"
            else:
                par.disable_peer_collectives()
                peer = None
    shard_ok = selfcheck_sharded(AF, par, dev, rank, world) if world > 1 else None

    wl = Workload(AF, par, dev, nv, V, seed=1000 + rank, world=world)
    warm = max(3, args.warmup)
    stepper, mode = make_stepper(wl, args.graph == "on", warm)
    for _ in range(warm):
        stepper()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_per_step = time_steps(stepper, args.steps, dev, world, sync_all)
    clk = clocks.stop() if rank == 0 else None
    value = total * V / (ms_per_step / 1e3)
    if mode != "eager":
        for _ in range(12):              # the eager steps after a capture re-populate the regular pool (a few cudaMallocs of GBs)
            wl.step()
        ms_eager = time_steps(wl.step, min(args.steps, 10), dev, world, sync_all)
    else:
        ms_eager = ms_per_step
    del stepper
    torch.cuda.synchronize(dev)

    # ---- e2e: host buffers -> H2D -> one-hot on device -> step -> D2H of reduced grads + grid affines ----
    e2e = run_e2e(args, AF, par, wl, dev, world, total, sync_all)

    # ---- our kernels per step, from a CUPTI trace of one eager step (after the timed regions: an attached profiler slows
    #      eager launches down; on EVERY rank, because the step contains collectives) ----
    wl.ensure_dense()
    own, other, own_names = count_own_launches(wl.step, dev)

    # ---- N > 1: the weak-scaling number as a second key (64 volumes PER GPU) ----
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        wl.free_dense()
        del wl
        torch.cuda.empty_cache()
        wl = Workload(AF, par, dev, args.volumes, V, seed=2000 + rank, world=world)
        st2, mode2 = make_stepper(wl, args.graph == "on", warm)
        for _ in range(warm):
            st2()
        ms_w = time_steps(st2, min(args.steps, 20), dev, world, sync_all)
        weak = {"scaling": "weak", "volumes_per_gpu": args.volumes, "ms_per_step": ms_w, "value": world * args.volumes * V / (ms_w / 1e3),
                "unit": UNIT, "step_mode": mode2}
        del st2

    if peer is not None:
        peer.check()                 # a peer that failed to arrive inside any captured / eager exchange raises here
    # All collectives are over: tear the process group down on EVERY rank before any rank-0-only work, so that nothing
    # below can ever wait on a peer (a stray all_reduce here once hung an 8-GPU run until the NCCL watchdog fired).
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    single = world == 1      # breakdown / variants / baselines are reported by the single-GPU run only
    wl.world = 1

    breakdown, roofline, l2_gbs = ({}, None, None)
    variants = {}
    if single and not args.no_breakdown:
        wl.ensure_dense()
        breakdown, l2_gbs = kernel_breakdown(AF, dev, wl.soft, wl.label, wl.image, wl.nii, wl.gpre, wl.params, wl.init, wl.go,
                                             wl.fov_mm, wl.fov_vox, nv, V)
        roofline = make_roofline(breakdown, l2_gbs, nv)
        wl.free_dense()
        variants["training_case_from_index_labels"] = variant_from_labels(AF, dev, wl.host_lab, wl.host_img, wl.nii, wl.gpre, wl.params,
                                                                          wl.init, wl.go, wl.fov_mm, wl.fov_vox, nv, V, l2_gbs)
    h = wl.h
    host_lab = wl.host_lab
    del wl
    torch.cuda.empty_cache()
    aten = None
    if single and not args.no_variants:
        import bench_extra as BX
        import acquisition_focus_b200 as afb_pkg
        for name, fn in (("cfg1_single_slice", BX.cfg1), ("cfg2_default_batch", BX.cfg2), ("cfg3_embedding", BX.cfg3_summary),
                         ("cfg5_256_stress", BX.cfg5), ("f1_resample_3d", BX.f1_resample_3d),
                         ("f4_clinical_views", BX.f4_clinical_views)):
            try:
                variants[name] = fn(afb_pkg, dev)
            except Exception as e:      # noqa: BLE001
                variants[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
        aten = aten_cuda_baseline(h, V, dev)
    cpu_base = None
    if single and not args.no_cpu_baseline:
        vps = 2
        times, kind = time_cpu_reference(h, V, vps, budget_s=args.cpu_seconds, min_steps=2)
        cval = vps * V / float(np.mean(times))
        cpu_base = {"value": cval, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                    "sample": f"{vps} volumes x {V} views of the same workload per step, {len(times)} steps after 1 warm-up "
                              f"({CPU_KIND_NOTE[kind]}, torch.set_num_threads(all cores))"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clk, "step_mode": mode, "ms_per_step_eager": ms_eager,
            "e2e": e2e, "gpu_launches": (own * args.steps) if own else None,
            "gpu_launches_per_step": {"own_kernels": own, "torch_and_nccl_helper_kernels": other, "by_kernel": own_names,
                                      "how": "CUPTI trace (torch.profiler) of one eager step outside the timed region"},
            "roofline": roofline, "cpu_baseline": cpu_base, "aten_cuda_baseline": aten,
            "kernels": breakdown, "l2_gbs_measured": l2_gbs, "variants": variants}
    if world > 1:
        line["sharded_equals_unsharded"] = shard_ok
        line["weak_scaling"] = weak
        line["cpus_per_rank"] = cpus_per_rank
        line["collectives_impl"] = collectives
        line["collectives_per_step"] = ["all_gather 4 floats (whole-batch pads of soft label + image)", "all_reduce 1 float (d out / d pad)",
                                        f"all_reduce [{V},{NP}] fp32 (view-parameter gradients)"]
    print(json.dumps(line))


def run_e2e(args, AF, par, wl, dev, world, total, sync_all):
    """Same metric through the public host-side entry, H2D and D2H inside the timed region."""
    import torch.distributed as dist
    nv, V = wl.nv, wl.V
    g_host = torch.empty((V, NP), dtype=torch.float32).pin_memory()
    ga_host = torch.empty((nv, V, 4, 4), dtype=torch.float32).pin_memory()
    host_lab, host_img = wl.host_lab, wl.host_img
    h2d = host_lab.numel() * host_lab.element_size() + host_img.numel() * host_img.element_size()
    d2h = g_host.numel() * 4 + ga_host.numel() * 4
    params = wl.params

    from acquisition_focus_b200.running.host_input import HostInputPipeline
    # N > 1: _pin_to_local_cpus gave this rank its OWN share of the NUMA-local cores - the packing pass may use all of them
    nthr = min(16, len(os.sched_getaffinity(0))) if world > 1 and not os.environ.get("AFB_NARROW_THREADS") else None
    depth = max(2, args.e2e_depth)
    pipe = HostInputPipeline(NUM_CLASSES, dev, depth=depth, group_volumes=args.e2e_group, narrow_labels=("auto" if args.e2e_narrow < 0 else "split" if args.e2e_narrow == 2 else bool(args.e2e_narrow)),
                             narrow_threads=nthr)
    host_bytes = h2d

    def consume(db):
        soft_t = db.soft_label.detach().requires_grad_(True)
        pads = par.exchange_pads([db.soft_pad, db.image_pad]) if world > 1 else [db.soft_pad, db.image_pad]
        params.grad = None
        ys, yl, yi, ga, nii_o, theta = AF.acquire_views(soft_t, db.label, db.image, wl.nii, wl.gpre, params, wl.init,
                                                        soft_pad=pads[0], image_pad=pads[1], **wl.kw)
        torch.autograd.backward([ys], [wl.go])
        g = par.reduce_view_grads(params.grad)
        pipe.release(db)
        g_host.copy_(g, non_blocking=True)
        ga_host.copy_(ga.detach(), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    def run(K):
        # public host-side entry, pipelined (`depth` buffer sets): while step k is sliced, the next batches are packed on the host cores and cross PCIe on a copy stream
        # (groups of volumes; the groups that have arrived are expanded to the int64 + fp32 one-hot volumes of run_dl.py:261-264
        # together with the soft volume's min record on an expansion stream).  The upload of EVERY consumed batch - the first
        # one included - is issued inside this function, i.e. inside the timed region: K steps = K full uploads + K fwd/bwd.
        # `depth - 1` batches are in flight ahead of the one being consumed (a slot is reused `depth` submits later)
        ahead = depth - 1
        for j in range(min(ahead, K)):
            pipe.submit(host_lab, host_img)
        for k in range(K):
            if k + ahead < K:
                pipe.submit(host_lab, host_img)
            consume(pipe.get())

    wl.free_dense()
    run(depth)                               # warm-up (allocates the buffer sets)
    K = max(2, args.e2e_steps)
    ms = time_steps(lambda: run(K), 1, dev, world, sync_all) / K
    h2d = pipe.h2d_bytes_last               # what actually crossed PCIe per batch (labels packed to uint8 on the host, or not)
    narrow_threads = pipe.narrow_threads if pipe.narrow else 0
    nthr = pipe.narrow_threads
    pack_ms, enqueue_ms = pipe.pack_seconds_last * 1e3, pipe.enqueue_seconds_last * 1e3
    split_info = ({"packed_volumes_of_last_batch": pipe.packed_volumes_last, "pack_fraction": pipe.pack_fraction,
                   "pack_rate_gbs": (pipe._r_pack or 0) / 1e9, "link_rate_gbs": (pipe._r_link or 0) / 1e9} if pipe.split else None)
    del pipe
    return {"value": total * V / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": K, "ms_per_step": ms, "h2d_gbs_per_rank": h2d / (ms * 1e-3) / 1e9, "pipeline_depth": depth,
            "host_input_bytes_per_step": int(host_bytes), "host_input_gbs_per_rank": host_bytes / (ms * 1e-3) / 1e9,
            "host_cores_for_packing": int(nthr or 0) or None, "split_upload": split_info,
            "host_label_packing": ({"threads": narrow_threads, "pack_ms_per_batch": pack_ms, "worker_ms_per_batch": enqueue_ms, "what": "int64 -> uint8 on the host cores (afb_host_narrow_labels), on a worker "
                                    "thread, group by group, overlapped with the uploads and with the previous step"}
                                   if narrow_threads else None),
            "timed_region": "K steps incl. the pipeline fill: every consumed batch is uploaded inside it (K uploads, K fwd+bwd, K D2H)",
            "bytes_are": "per rank (each rank uploads its own shard)",
            "what": "pinned host index-label int64 + image fp32 -> running.host_input.HostInputPipeline (labels packed to uint8 on the host "
                    "cores when host_label_packing is not null: --e2e-narrow, default auto = when the rank has >= 8 host cores; double "
                    "buffered: H2D in groups of "
                    f"{args.e2e_group} volumes on a copy stream, fused one-hot expansion + min record of the arrived groups on an expansion "
                    "stream, overlapped with the previous step's slicing) -> same acquisition fwd+bwd (dVolume + dTheta) -> D2H reduced "
                    "dTheta + grid affines; one full batch uploaded per step"}


def aten_cuda_baseline(h, views, dev, vps=2):
    """The Blackwell bar: the reference's own op sequence (oracle port: fp64 bookkeeping, whole-volume min-shift,
    affine_grid + checkpointed grid_sample, autograd) with its tensors on the GPU, i.e. through ATen's sm_100 CUDA kernels, on
    a bounded sample of the same workload (same step definition as the cpu_baseline leg)."""
    from oracle import af_oracle as O           # baseline leg only (never on the product path)
    lab = h["lab"][:vps].to(dev)
    label, soft = one_hot_volumes(lab)
    soft = soft.requires_grad_(True)
    img, nii, gpre, init = h["img"][:vps].to(dev), h["nii"][:vps].to(dev), h["gpre"][:vps].to(dev), h["init"].to(dev)
    params = h["params"][:vps].to(dev).requires_grad_(True)
    fov_mm, fov_vox = torch.tensor([192.0, 192.0, 1.5], device=dev), torch.tensor([S, S, 1], device=dev)
    go = torch.cos(torch.arange(vps * views * NUM_CLASSES * S * S, device=dev, dtype=torch.float32) * 0.618).view(vps, views, NUM_CLASSES, S, S, 1)

    def step():
        soft.grad = None
        params.grad = None
        loss = 0
        for v in range(views):
            theta = O.view_theta(params[:, v], init[v:v + 1, :6], init[v, 6:9], init[v:v + 1, 9:], OFFSET_CLIP, ZOOM_CLIP, S)
            ys, yl, yi, ga, _ = O.atm_tail_forward(soft, label, img, nii, gpre[:, v], theta, fov_mm, fov_vox)
            loss = loss + (ys * go[:, v]).sum()
        loss.backward()
        return params.grad.sum(0)
    t = _time(step, dev, reps=3, warm=1)
    return {"value": vps * views / (t * 1e-3), "unit": UNIT, "ms_per_step": t,
            "sample": f"{vps} volumes x {views} views of the same workload per step, 3 steps after 1 warm-up",
            "what": "reference op sequence (oracle port) on device='cuda': ATen affine_grid / grid_sampler_3d(+backward) / min / "
                    "sub / add / linalg sm_100 kernels, eager"}


def variant_from_labels(AF, dev, host_lab, host_img, nii, gpre, params, init, go, fov_mm, fov_vox, nv, V, l2_gbs=None):
    """The reference's TRAINING case (the volume never requires grad, only dTheta is consumed) through the one-hot-from-index
    path: uint8 label map (2 MiB/volume) instead of the fp32 + int64 one-hot volumes (192 MiB/volume); y_soft is bitwise the
    same.  Device-resident and end-to-end (pinned uint8 labels + fp32 image H2D, reduced dTheta + grid affines D2H)."""
    host_u8 = host_lab.to(torch.uint8).pin_memory()
    lab = host_u8.to(dev)
    image = host_img.to(dev)
    g_host = torch.empty((V, NP), dtype=torch.float32).pin_memory()
    ga_host = torch.empty((nv, V, 4, 4), dtype=torch.float32).pin_memory()

    def step(lab_t, img_t):
        params.grad = None
        ys, yl, yi, ga, nii_o, th = AF.acquire_views_from_labels(lab_t, img_t, nii, gpre, params, init, num_classes=NUM_CLASSES,
                                                                 offset_clip=OFFSET_CLIP, zoom_clip=ZOOM_CLIP, spat=S,
                                                                 slice_fov_mm=fov_mm, slice_fov_vox=fov_vox)
        torch.autograd.backward([ys], [go])
        return params.grad.sum(dim=0), ga                 # no collective here: this runs on one rank only

    def e2e():
        g, ga = step(host_u8.to(dev, non_blocking=True), host_img.to(dev, non_blocking=True))
        g_host.copy_(g, non_blocking=True); ga_host.copy_(ga.detach(), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    t = _time(lambda: step(lab, image), dev, reps=10, warm=3)
    te = _time(e2e, dev, reps=3, warm=1)
    # per-kernel roofline of this variant (each kernel alone, CUDA events)
    from acquisition_focus_b200 import _lib as L
    import ctypes as C
    lib, st = L.lib(), L.stream_ptr(dev)
    nS, Npix = nv * V, S * S
    spec = AF.ViewSpec(kind=L.AFFINE_PARAMS, V=V, gpre=gpre.reshape(nS, 4, 4).contiguous(), init=init, R=R, spat=S,
                       offset_clip=OFFSET_CLIP, zoom_clip=ZOOM_CLIP, nii_affine=nii, fov_mm=tuple(fov_mm),
                       params=params.detach().reshape(nS, NP).contiguous())
    spec = AF.prepare_views(spec, nv, (S, S, S), fov_vox, dev)[0]
    lab5 = lab[:, None]
    vd, vs = L.volume_desc(lab5), spec.struct()
    y_soft = torch.empty((nv, V, NUM_CLASSES, S, S, 1), device=dev)
    y_lab = torch.empty((nv, V, NUM_CLASSES, S, S, 1), dtype=torch.int64, device=dev)
    d_aff = torch.zeros(nS, NP, device=dev)
    ws = torch.zeros(int(lib.afb_slice_bwd_workspace_bytes(nS)), dtype=torch.uint8, device=dev)
    kern = {}
    tk = _time(lambda: L.check(lib.afb_slice_onehot_fwd(C.byref(vd), NUM_CLASSES, C.byref(vs), S, S, 1, L.ptr(y_soft), L.ptr(y_lab), 1, st),
                               "afb_slice_onehot_fwd"), dev)
    kern["onehot_fwd(u8 labels -> soft C=8 + int64 one-hot nearest)"] = {"kernel": "onehot_fwd_kernel<unsigned char, 1>", "ms": tk, "bound": "hbm",
                                                                         "bytes": nS * Npix * (NUM_CLASSES * 4 + NUM_CLASSES * 8) + nv * S ** 3}
    tk = _time(lambda: L.check(lib.afb_slice_onehot_bwd(C.byref(vd), NUM_CLASSES, C.byref(vs), S, S, 1, L.ptr(go), None, L.ptr(d_aff), None,
                                                        L.ptr(ws), st), "afb_slice_onehot_bwd"), dev)
    kern["onehot_bwd + view_chain (dTheta)"] = {"kernel": "onehot_bwd_kernel<unsigned char>", "ms": tk, "bound": "hbm",
                                                "bytes": nS * Npix * NUM_CLASSES * 4 + nv * S ** 3}
    pad_i = AF.volume_min(image)
    tk = _time(lambda: AF._slice_forward_raw(image, spec, fov_vox, L.BILINEAR, L.PAD_DEVICE, 0.0, pad_i), dev)
    kern["slice_fwd(image C=1 bilinear)"] = {"kernel": "slice_fwd_kernel<float, 0>", "ms": tk, "bound": "l2", "bytes": nS * Npix * 36}
    tk = _time(lambda: AF.volume_min(image), dev)
    kern["volume_min(image)"] = {"kernel": "volume_min_kernel<float>", "ms": tk, "bound": "hbm", "bytes": image.numel() * 4}
    hbm = _hbm_peak()[0]
    for v in kern.values():
        v["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        v["frac"] = v["gbs"] / (hbm if v["bound"] == "hbm" else (l2_gbs or float("nan")))
    return {"ms_per_step": t, "value": nv * V / (t * 1e-3), "unit": UNIT, "kernels": kern,
            "bytes_note": "compulsory HBM bytes: outputs written once + the uint8 label volumes read once (gathers hit L1/L2); "
                          "the one-hot kernels are bound by their OUTPUT writes (int64 one-hot slices: 64 B per pixel)",
            "e2e": {"value": nv * V / (te * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(host_u8.numel() + host_img.numel() * 4), "d2h_bytes_per_step": int(g_host.numel() * 4 + ga_host.numel() * 4)},
            "what": "uint8 index labels -> y_soft (C=8, bitwise = dense path) + int64 one-hot nearest label + image slices, backward w.r.t. "
                    "the view parameters only (no dVolume): view_prologue, volume_min(image), onehot_fwd, slice_fwd(image), onehot_bwd, view_chain"}


def _time(fn, dev, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def kernel_breakdown(AF, dev, soft, label, image, nii, gpre, params, init, go, fov_mm, fov_vox, nv, V):
    """Time every kernel of the step alone (CUDA events, mean of 5 after 2 warm-ups; operands >> L2)."""
    from acquisition_focus_b200 import _lib as L
    import ctypes as C
    lib = L.lib()
    Npix = S * S
    nS = nv * V
    out = {}
    # L2 read peak: afb_probe_read streams a 32 MiB L2-resident buffer 64 times inside ONE launch (ld.global.cg,
    # 16 B per thread per load), best of 5 - a single 32 MiB torch copy is launch-latency bound and under-reads L2
    buf = torch.empty(32 * 1024 * 1024, dtype=torch.uint8, device=dev).zero_()
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    st = L.stream_ptr(dev)
    passes = 64
    t = min(_time(lambda: L.check(lib.afb_probe_read(L.ptr(buf), buf.numel(), passes, L.ptr(sink), st), "afb_probe_read"),
                  dev, reps=3, warm=1) for _ in range(5))
    l2_gbs = buf.numel() * passes / (t * 1e-3) / 1e9
    del buf
    # min pass
    t = _time(lambda: AF.volume_min(soft, with_mask=True), dev)
    out["volume_min_mask(soft)"] = {"kernel": "volume_min_mask_kernel", "ms": t, "bytes": soft.numel() * 4 + soft.numel() // 8, "bound": "hbm"}
    pad_s, pad_i = AF.volume_min(soft, with_mask=True), AF.volume_min(image)
    t = _time(lambda: AF.volume_min(image), dev)
    out["volume_min(image)"] = {"kernel": "volume_min_kernel<float>", "ms": t, "bytes": image.numel() * 4, "bound": "hbm"}
    spec = AF.ViewSpec(kind=L.AFFINE_PARAMS, V=V, gpre=gpre.reshape(nS, 4, 4).contiguous(), init=init, R=R, spat=S,
                       offset_clip=OFFSET_CLIP, zoom_clip=ZOOM_CLIP, nii_affine=nii, fov_mm=tuple(fov_mm),
                       params=params.detach().reshape(nS, NP).contiguous())
    t = _time(lambda: AF.prepare_views(spec, nv, (S, S, S), fov_vox, dev), dev)
    out["view_prologue"] = {"kernel": "view_prologue_kernel", "ms": t, "bytes": nS * (NP * 4 + 64 + 128), "bound": "latency"}
    spec = AF.prepare_views(spec, nv, (S, S, S), fov_vox, dev)[0]
    sd = soft.detach()
    b_soft, b_lab, b_img = nS * Npix * NUM_CLASSES * 36, nS * Npix * NUM_CLASSES * 16, nS * Npix * 36
    t = _time(lambda: AF._slice_forward3_raw(sd, label, image, spec, fov_vox, (L.PAD_DEVICE, 0.0, pad_s), (L.PAD_DEVICE, 0.0, pad_i)), dev)
    out["slice_fwd3(soft C=8 bilinear + label C=8 int64 nearest + image C=1 bilinear, ONE launch)"] = {
        "kernel": "slice_fwd3_kernel<long>", "ms": t, "bytes": b_soft + b_lab + b_img, "bound": "l2"}
    t = _time(lambda: AF._slice_forward_raw(sd, spec, fov_vox, L.BILINEAR, L.PAD_DEVICE, 0.0, pad_s), dev)
    out["slice_fwd(soft C=8 bilinear) [not in step]"] = {"kernel": "slice_fwd_cl_kernel<float, 0", "ms": t, "bytes": b_soft, "bound": "l2"}
    t = _time(lambda: AF._slice_forward_raw(label, spec, fov_vox, L.NEAREST, L.PAD_ZERO, 0.0, None), dev)
    out["slice_fwd(label C=8 int64 nearest) [not in step]"] = {"kernel": "slice_fwd_cl_kernel<long, 1", "ms": t, "bytes": b_lab, "bound": "l2"}
    t = _time(lambda: AF._slice_forward_raw(image, spec, fov_vox, L.BILINEAR, L.PAD_DEVICE, 0.0, pad_i), dev)
    out["slice_fwd(image C=1 bilinear) [not in step]"] = {"kernel": "slice_fwd_kernel<float, 0>", "ms": t, "bytes": b_img, "bound": "l2"}
    # backward pieces
    d_vol = torch.empty_strided(sd.shape, sd.stride(), dtype=torch.float32, device=dev)
    ws = torch.zeros(int(lib.afb_slice_bwd_workspace_bytes(nS)), dtype=torch.uint8, device=dev)
    d_aff = torch.zeros(nS, NP, device=dev); d_pad = torch.zeros(1, device=dev)
    vd, vs = L.volume_desc(sd), spec.struct()
    t = _time(lambda: L.check(lib.afb_slice_pad_grad(C.byref(vd), C.byref(vs), S, S, 1, L.ptr(go), L.ptr(d_pad), st), "afb_slice_pad_grad"), dev)
    out["slice_pad_grad"] = {"kernel": "slice_pad_grad_kernel", "ms": t, "bytes": nS * Npix * NUM_CLASSES * 4, "bound": "hbm"}
    t = _time(lambda: L.check(lib.afb_min_grad_fill_mask(L.ptr(pad_s._afb_mask), sd.numel(), L.ptr(pad_s), L.ptr(d_pad), L.ptr(d_vol), st), "afb_min_grad_fill_mask"), dev)
    out["min_grad_fill_mask(dVolume)"] = {"kernel": "min_grad_fill_mask_kernel<", "ms": t, "bytes": sd.numel() * 4 + sd.numel() // 8, "bound": "hbm"}
    t = _time(lambda: L.check(lib.afb_min_grad_fill(L.ptr(sd), L.F32, sd.numel(), L.ptr(pad_s), L.ptr(d_pad), L.ptr(d_vol), st), "afb_min_grad_fill"), dev)
    out["min_grad_fill(dVolume, re-reads the volume) [not in step]"] = {"kernel": "min_grad_fill_kernel<float>", "ms": t, "bytes": sd.numel() * 8, "bound": "hbm"}

    def bwd(with_dvol):
        L.check(lib.afb_slice_bwd(C.byref(vd), C.byref(vs), S, S, 1, L.PAD_DEVICE, 0.0, L.ptr(pad_s), L.ptr(go), None,
                                  L.ptr(d_vol) if with_dvol else None, L.ptr(d_aff), None, None, L.ptr(ws), st), "afb_slice_bwd")
    t = _time(lambda: bwd(True), dev)
    out["slice_bwd(soft, dVolume+dTheta)"] = {"kernel": "slice_bwd_cl_kernel<float", "ms": t, "bytes": nS * Npix * NUM_CLASSES * (32 + 4 + 32), "bound": "l2"}
    t = _time(lambda: bwd(False), dev)
    out["slice_bwd(soft, dTheta only) [not in step]"] = {"kernel": "slice_bwd_cl_kernel<float", "ms": t, "bytes": nS * Npix * NUM_CLASSES * (32 + 4), "bound": "l2"}
    t = _time(lambda: L.check(lib.afb_slice_scatter(C.byref(vd), C.byref(vs), S, S, 1, L.ptr(go), L.ptr(d_vol), st), "afb_slice_scatter"), dev)
    out["slice_scatter(dVolume only: REDs, no gather) [not in step]"] = {"kernel": "slice_scatter_kernel<", "ms": t, "bytes": nS * Npix * NUM_CLASSES * (4 + 32), "bound": "l2"}
    for k, v in out.items():
        v["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
    return out, l2_gbs


def ncu_traffic(kernel_prefix, nv):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel, from the committed `ncu --set full` capture of
    this workload at a smaller volume count (profiles/ncu_traffic.json, written by profiles/ncu_summary.py --traffic),
    scaled linearly to `nv` volumes; None if not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path) or not kernel_prefix:
        return None
    d = json.load(open(path))
    for name, val in d.get("kernels", {}).items():
        if name.startswith(kernel_prefix):
            return val * nv / d.get("volumes", nv)
    return None


def _hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "of fallback (B200_PROFILING.md 6.65 TB/s)"


def make_roofline(breakdown, l2_gbs, nv):
    hbm, src = _hbm_peak()
    in_step = {k: v for k, v in breakdown.items() if "not in step" not in k and "[torch]" not in k}
    name = max(in_step, key=lambda k: in_step[k]["ms"])
    k = in_step[name]
    peak = hbm if k["bound"] == "hbm" else l2_gbs
    for v in breakdown.values():
        v["frac"] = None if v["bound"] == "latency" else v["gbs"] / (hbm if v["bound"] == "hbm" else l2_gbs)
    return {"kernel": name, "bound": k["bound"], "achieved": k["gbs"], "peak": peak, "unit": "GB/s", "frac": k["gbs"] / peak,
            "traffic": ncu_traffic(k.get("kernel", ""), nv),
            "traffic_source": "committed `ncu --set full` capture of this workload (profiles/ncu_traffic.json: dram__bytes_read.sum + "
                              "dram__bytes_write.sum per launch at the captured volume count, scaled linearly to this run's volumes); "
                              "DRAM counters cannot be read in-line",
            "cuda_kernel": k.get("kernel"), "peak_source": src if k["bound"] == "hbm" else "L2 read bandwidth measured in this run: afb_probe_read, 64 passes over a 32 MiB L2-resident buffer in one launch",
            "algorithmic_bytes_per_launch": k["bytes"], "ms_per_launch": k["ms"]}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
