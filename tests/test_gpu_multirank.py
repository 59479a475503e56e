"""SURVEY 8(e) on real GPUs: a batch sharded over 2 ranks (NCCL) returns what the unsharded call returns.

Volumes are partitioned by batch with all views of a volume co-located (``parallel.shard_range``); the exchanges are
``parallel.global_pads`` (the whole-batch ``volume.min()`` of utils/nifti_utils.py:200 + the sum of d(out)/d(pad) in the
backward) and ``parallel.reduce_view_grads`` (the [V,P] view-parameter gradients).  Needs >= 2 GPUs: skipped on a one-GPU
box (run with ``gpurun --gpus 2``; ``bench.py --gpus N`` runs the same comparison as a self-check)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cases

pytestmark = pytest.mark.gpu

INIT = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]])


def _step(afb, par, case, lo, hi, via_exchange=False):
    """fwd + bwd over volumes [lo, hi) of the case; returns outputs, dVolume, per-sample parameter gradients."""
    V = case["V"]
    soft = case["soft"][lo:hi].cuda().requires_grad_(True)
    image, label = case["image"][lo:hi].cuda(), case["label"][lo:hi].cuda()
    params = torch.stack(case["params"], dim=1)[lo:hi].cuda().requires_grad_(True)
    gpre = torch.stack(case["gpre"], dim=1)[lo:hi].cuda()
    kw = dict(offset_clip=case["offset_clip"], zoom_clip=case["zoom_clip"], spat=case["S"], slice_fov_mm=case["slice_fov_mm"].tolist(),
              slice_fov_vox=case["slice_fov_vox"].tolist())
    if via_exchange:        # the min passes run inside acquire_views, the exchange hook makes them whole-batch (bench.py's route)
        pads = []

        def hook(local):
            pads.extend(par.exchange_pads(local))
            return pads
        kw["pad_exchange"] = hook
    else:
        pads = par.global_pads([soft, image], [True, False])
        kw.update(soft_pad=pads[0], image_pad=pads[1])
    ys, yl, yi, ga, nii, theta = afb.acquire_views(
        soft, label, image, case["nii"][lo:hi].cuda(), gpre, params, INIT.repeat(V, 1).cuda(), **kw)
    go = cases.pattern((case["B"], V) + tuple(ys.shape[2:]), 1.0)[lo:hi].cuda()
    torch.autograd.backward([ys], [go])
    return ys.detach(), yl, yi, ga.detach(), soft.grad, params.grad, pads


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import acquisition_focus_b200 as afb
        from acquisition_focus_b200 import parallel as par
        case = cases.atm_case(32, 4, 3, seed=61)
        # make the image minimum sit on ONE volume only, so that sharding would change the pad value without the exchange
        case["image"][3, 0, 5, 6, 7] = -7.5
        lo, hi = par.shard_range(case["B"], rank, world)
        ys, yl, yi, ga, dsoft, dparams, pads = _step(afb, par, case, lo, hi)
        g = par.reduce_view_grads(dparams)
        again = _step(afb, par, case, lo, hi, via_exchange=True)
        same_route = all(torch.equal(a, b) for a, b in zip(again[:4], (ys, yl, yi, ga))) and \
            (again[4] - dsoft).abs().max().item() <= 1e-6 * dsoft.abs().max().item()
        dist.barrier()
        torch.cuda.synchronize()
        # numpy arrays are pickled by value (torch tensors would travel as shared-memory handles that die with this process)
        # the same exchanges through the NVLink peer-memory kernels (afb_peer_collective) instead of NCCL
        peer_ok = None
        try:
            pc = par.enable_peer_collectives(torch.device("cuda", rank))
        except Exception as e:      # noqa: BLE001  (symmetric memory unavailable: reported, not failed)
            pc, peer_ok = None, f"unavailable: {e!r}"
        if pc is not None:
            p3 = _step(afb, par, case, lo, hi, via_exchange=True)
            g3 = par.reduce_view_grads(p3[5])
            for _ in range(3):                                    # epochs / slot parity over repeated calls
                g3b = par.reduce_view_grads(p3[5])
            # the three ops of afb_peer_collective against NCCL on arbitrary data
            x = torch.arange(37, dtype=torch.float32, device="cuda") * (rank + 1.5)
            want = x.clone(); dist.all_reduce(want)
            got = pc.all_reduce_sum(torch.stack([x * 0.25, x * 0.75]), 3, pre_sum=2)                     # local rows summed first
            gath = torch.empty(world * 37, device="cuda"); dist.all_gather_into_tensor(gath, x)
            ops_ok = torch.allclose(got, want, rtol=1e-6) and torch.equal(pc.all_gather(x, 3), gath)
            rows = torch.tensor([[1.0 + rank, 3.0], [-2.0, 5.0 + rank]], device="cuda")
            merged = pc.merge_pads(rows)
            ops_ok = ops_ok and merged.tolist() == [[1.0, 3.0], [-2.0, 11.0]]
            pc.check()
            peer_ok = bool(ops_ok and all(torch.equal(a, b) for a, b in zip(p3[:4], (ys, yl, yi, ga)))
                           and (p3[4] - dsoft).abs().max().item() <= 1e-6 * dsoft.abs().max().item()
                           and (g3 - g).abs().max().item() <= 1e-6 * g.abs().max().item() and torch.equal(g3, g3b)
                           and all(torch.equal(a, b) for a, b in zip(p3[6], pads)))
            par.disable_peer_collectives()
            dist.barrier()
            torch.cuda.synchronize()
        res = {"rank": rank, "lo": lo, "hi": hi, "pads": [p.cpu().numpy() for p in pads], "same_route": same_route, "peer_ok": peer_ok}
        if rank == 0:
            dist.destroy_process_group()           # the unsharded run below must not touch a collective
            full = _step(afb, par, case, 0, case["B"])
            res["full"] = [t.cpu().numpy() for t in full[:6]]
            res["full_pads"] = [p.cpu().numpy() for p in full[6]]
        res.update(ys=ys.cpu().numpy(), yl=yl.cpu().numpy(), yi=yi.cpu().numpy(), ga=ga.cpu().numpy(), dsoft=dsoft.cpu().numpy(),
                   g=g.cpu().numpy())
        q.put(res)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_equals_unsharded_nccl():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:                       # back to tensors
        for k, v in list(r.items()):
            if isinstance(v, list):
                r[k] = [torch.from_numpy(x) for x in v]
            elif hasattr(v, "dtype"):
                r[k] = torch.from_numpy(v)
    full_ys, full_yl, full_yi, full_ga, full_dsoft, full_dparams = res[0]["full"]
    for a, b in zip(res[0]["pads"], res[0]["full_pads"]):
        assert torch.equal(a, b)                                        # global (min, multiplicity) on every rank
    assert torch.equal(res[0]["pads"][1], res[1]["pads"][1]) and res[0]["pads"][1][0].item() == -7.5
    for r in res:
        lo, hi = r["lo"], r["hi"]
        assert r["same_route"]
        assert r["peer_ok"] is True or (isinstance(r["peer_ok"], str) and r["peer_ok"].startswith("unavailable")), r["peer_ok"]
        print("peer-memory collectives:", r["peer_ok"])
        assert torch.equal(r["ys"], full_ys[lo:hi]) and torch.equal(r["yl"], full_yl[lo:hi])      # same pad value -> bitwise
        assert torch.equal(r["yi"], full_yi[lo:hi]) and torch.equal(r["ga"], full_ga[lo:hi])
        scale = full_dsoft.abs().max().item()
        assert (r["dsoft"] - full_dsoft[lo:hi]).abs().max().item() <= 1e-6 * scale                 # atomics: order-dependent last bits
        want = full_dparams.sum(0)
        assert (r["g"] - want).abs().max().item() <= 1e-5 * want.abs().max().item()
