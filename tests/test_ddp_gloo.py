"""CPU, world size 2, gloo: the bucketed gradient reducer (b200seg/ddp.py) averages gradients across ranks exactly
like a single process seeing both shards, with buckets filled in gradient-ready order and an async all-reduce per
bucket (N-rank gradients == average of the N single-rank runs on the same shards)."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _model(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1),
                         nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 1, 1))


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(4, 3, 16, 16, generator=g), (torch.rand(4, 1, 16, 16, generator=g) > 0.7).float()


def _worker(rank, world, initfile, out):
    from b200seg.ddp import GradReducer
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    m = _model(seed=rank)                           # ranks seed differently: the reducer must broadcast rank 0's replica
    m[1].running_mean.fill_(float(rank))
    red = GradReducer(m, bucket_mb=0.0005)          # tiny buckets -> several all-reduces, exercising the ordering
    assert len(red.buckets) >= 2
    ref0 = _model(seed=0)
    for (k, a), (_, b) in zip(m.state_dict().items(), ref0.state_dict().items()):
        assert torch.equal(a, b), f"rank {rank}: {k} was not broadcast from rank 0"
    for step in range(2):                           # two steps: buckets must reset correctly
        m.zero_grad(set_to_none=True)
        x, t = _data(rank)
        nn.functional.binary_cross_entropy_with_logits(m(x), t).backward()
        red.finish()
    if rank == 0:
        torch.save({k: p.grad.clone() for k, p in m.named_parameters()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_mean_of_rank_gradients():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "g.pt")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        mp.spawn(_worker, args=(world, os.path.join(d, "init"), out), nprocs=world, join=True)
        got = torch.load(out)
    # reference: per-rank local-BN gradients, averaged
    want = None
    for rank in range(world):
        m = _model()
        x, t = _data(rank)
        nn.functional.binary_cross_entropy_with_logits(m(x), t).backward()
        g = {k: p.grad.clone() for k, p in m.named_parameters()}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for k in want:
        assert torch.allclose(got[k], want[k] / world, rtol=1e-5, atol=1e-7), k
