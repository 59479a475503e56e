"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes, volumes sharded by batch, the single
all-reduce of the per-view parameter gradients and the pad (min, multiplicity) reduction."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acquisition_focus_b200 import parallel as par


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_vol, V, P = 7, 3, 10
        full = torch.arange(n_vol * V * P, dtype=torch.float32).view(n_vol, V, P)
        lo, hi = par.shard_range(n_vol, rank, world)
        g = par.reduce_view_grads(full[lo:hi])
        ok_g = torch.equal(g, full.sum(0))
        vols = torch.tensor([[3.0, 1.0, 1.0, 5.0], [1.0, 2.0, 9.0, 1.0]])
        mine = vols[rank]
        mc = torch.tensor([mine.min().item(), float((mine == mine.min()).sum())])
        tot = par.allreduce_pad(mc)
        ok_p = tot[0].item() == 1.0 and tot[1].item() == 4.0
        # several tensors at once ([k,2] stack, one collective): second pair has its minimum on rank 1 only
        both = par.allreduce_pad(torch.stack([mc, torch.tensor([5.0 - 3.0 * rank, 2.0 + rank])]))
        ok_p = ok_p and both.tolist() == [[1.0, 4.0], [2.0, 3.0]]
        q.put((rank, lo, hi, ok_g, ok_p))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (1, 2, 7, 64):
        for w in (1, 2, 4, 8):
            spans = [par.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 4) and res[1][1:3] == (4, 7)
    assert all(r[3] and r[4] for r in res)
