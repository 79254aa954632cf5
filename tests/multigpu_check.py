"""Multi-GPU correctness of the data-parallel path, run with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/multigpu_check.py

(tests/test_gpu_multi.py launches it when >= 2 GPUs are visible.)  Checks, NCCL backend, b200seg.ddp.GradReducer:
  1. N-rank gradients == mean of the single-rank gradients on the same shards, with the zero-copy
     gradient slots (weight gradients written straight into the flat buckets) and with the weight-gradient side stream;
  2. ranks that seed differently start from rank 0's replica (broadcast) and after K graphed training steps
     (engine.GraphedTrainStep: NCCL all-reduces captured in the CUDA graph) every replica holds bit-identical parameters
     — BatchNorm running statistics stay per rank, as in the reference;
  3. utils.helpers.train(reducer=...) takes the same early-stopping decisions on every rank and only rank 0 writes.
Prints one line 'MULTIGPU_OK ...' on rank 0; any failure raises on the failing rank (non-zero exit).
"""
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "medical-image-segmentation-and-classification_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from b200seg import kernels as K, ops
    from b200seg.ddp import GradReducer
    from b200seg.engine import GraphedTrainStep
    from b200seg.models.segmentation_models import AttentionUNet
    from b200seg.optim import FusedClipAdamW
    from b200seg.utils.synthetic import xray_batch

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    # ---- 1. reduced gradients == mean of per-rank gradients ------------------------------------------------
    torch.manual_seed(123 + rank)                       # deliberately different seeds: the reducer must broadcast
    model = AttentionUNet().to(dev, memory_format=torch.channels_last).eval()     # eval BN: reproducible gradients
    x, t = xray_batch(2, 128, 128, seed=50 + rank, device=dev)
    red = GradReducer(model, bucket_mb=8)
    w0 = [p.detach().clone() for p in model.parameters()]
    chk = [torch.zeros_like(w) for w in w0]
    for w, c in zip(w0, chk):
        c.copy_(w)
        dist.broadcast(c, src=0)
        assert torch.equal(w, c), "parameters were not broadcast from rank 0"

    def local_grads():
        model.zero_grad(set_to_none=True)
        loss, _ = ops.seg_loss(model(x), t, 1.0, 0.0, 1.0)
        loss.backward()

    red.remove()                                        # plain local gradients first
    local_grads()
    mine = [p.grad.detach().clone() for p in model.parameters()]
    want = []
    for g in mine:
        s = g.clone()
        dist.all_reduce(s)
        want.append(s / world)
    for overlap in (False, True):
        red = GradReducer(model, bucket_mb=8, broadcast=False)
        K.set_wgrad_overlap(overlap)
        try:
            local_grads()
            red.finish()
        finally:
            K.set_wgrad_overlap(False)
        torch.cuda.synchronize()
        n_alias = 0
        for p, w in zip(model.parameters(), want):
            b, i = red._where[p]
            n_alias += int(p.grad.data_ptr() == b.flat.data_ptr() + 4 * b.offsets[i])
            err = float((p.grad - w).abs().max())
            assert err <= 1e-6 * (1.0 + float(w.abs().max())), (overlap, tuple(p.shape), err)
        # every tensor-core convolution weight (all but the 3-channel stem, the 1-channel head and the four psi convs)
        n4d = sum(1 for p in model.parameters() if p.dim() == 4 and p.shape[0] >= 32 and p.shape[1] >= 32)
        assert n_alias >= n4d, f"only {n_alias} of {n4d} conv-weight gradients live in their bucket slot"
        red.remove()

    # ---- 2. replicas stay bit-identical through graphed training steps -------------------------------------
    torch.manual_seed(7 + rank)
    model = AttentionUNet().to(dev, memory_format=torch.channels_last).train()
    red = GradReducer(model, bucket_mb=8)
    opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-4, max_norm=1.0)
    st = GraphedTrainStep(model, opt, reducer=red, warmup=3, wgrad_overlap=True)
    for i in range(8):
        xb, tb = xray_batch(2, 128, 128, seed=1000 + 10 * i + rank, device=dev)     # different data on every rank
        st(xb, tb)
    torch.cuda.synchronize()
    assert st.capture_error is None, st.capture_error
    assert st.replays == 5, (st.replays, st.eager_steps)
    spread = red.param_checksum()
    assert spread == 0.0, f"replicas diverged: checksum spread {spread}"
    for p in model.parameters():
        c = p.detach().clone()
        dist.broadcast(c, src=0)
        assert torch.equal(c, p.detach()), "replica parameters are not bit-identical"
    red.remove()

    # ---- 3. train() under data parallelism -----------------------------------------------------------------
    from torch.utils.data import DataLoader, TensorDataset
    from b200seg.utils.helpers import train
    torch.manual_seed(11 + rank)
    model = AttentionUNet()
    xs, ts = xray_batch(10, 64, 64, seed=300 + rank)
    tr = DataLoader(TensorDataset(xs[:8], ts[:8]), batch_size=2)
    va = DataLoader(TensorDataset(xs[8:], ts[8:]), batch_size=2)
    model = model.to(dev, memory_format=torch.channels_last)
    red = GradReducer(model, bucket_mb=8)
    logs = []
    with tempfile.TemporaryDirectory() as d:
        save = os.path.join(d, f"rank{rank}")
        best = train(model, tr, va, dev, epochs=2, lr=1e-3, name="AttentionUNet", save_dir=save, seg=True,
                     reducer=red, log=logs.append)
        wrote = os.path.exists(os.path.join(save, "AttentionUNet_best_loss.pt"))
    assert wrote == (rank == 0), f"rank {rank}: checkpoint written = {wrote}"
    bt = torch.tensor([best], dtype=torch.float64, device=dev)
    b0 = bt.clone()
    dist.broadcast(b0, src=0)
    assert torch.equal(bt, b0), "ranks disagree on the best validation loss"
    red.remove()

    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_OK world={world} zero_copy_slots={n_alias}/{n4d} replicas_in_sync=True", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
