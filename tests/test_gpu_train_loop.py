"""GPU: the train() mirror of utils/helpers.py:231-412 runs end to end on the CUDA path — loss decreases on a tiny
synthetic set, the best checkpoint has the reference's state_dict layout and reloads strictly — and the R2U_Net
inference path (BASELINE.json configs[4]) handles batch 1 and 512x512 inputs."""
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

pytestmark = pytest.mark.gpu


def test_train_mirror_runs_and_saves_reference_layout(tmp_path, monkeypatch):
    import json
    from b200seg.utils.helpers import get_seg_model, train
    monkeypatch.setenv("B200SEG_TRAIN_LOG", str(tmp_path / "train.jsonl"))
    from oracle.synthetic import xray_batch
    torch.manual_seed(0)
    x, t = xray_batch(8, 128, 128, seed=21)
    train_dl = DataLoader(TensorDataset(x[:6], t[:6]), batch_size=2, shuffle=False)
    val_dl = DataLoader(TensorDataset(x[6:], t[6:]), batch_size=2)
    model = get_seg_model("attentionunet")
    logs = []
    best = train(model, train_dl, val_dl, torch.device("cuda"), epochs=3, lr=2e-3, name="AttentionUNet",
                 save_dir=str(tmp_path), seg=True, log=logs.append)
    assert np.isfinite(best)
    ep = [l for l in logs if l.startswith("[AttentionUNet] Ep")]
    assert len(ep) == 3
    losses = [float(l.split("TrainLoss ")[1].split(" ")[0]) for l in ep]
    assert losses[-1] < losses[0], losses
    ck = torch.load(os.path.join(tmp_path, "AttentionUNet_best_loss.pt"), weights_only=True)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "AttentionUNet.npz"))
    assert list(ck.keys()) == [str(k) for k in gold["keys"]]
    get_seg_model("attentionunet").load_state_dict(ck, strict=True)
    recs = [json.loads(ln) for ln in open(tmp_path / "train.jsonl")]       # B200SEG_TRAIN_LOG: one line per epoch
    assert [r["epoch"] for r in recs] == [1, 2, 3] and all(r["images"] == 6 and r["images_per_s"] > 0 for r in recs)
    assert abs(recs[0]["train_loss"] - losses[0]) < 1e-3 and recs[-1]["graph_replays"] + recs[-1]["eager_steps"] == 9


@pytest.mark.parametrize("batch,side", [(1, 256), (3, 256), (1, 512)])
def test_r2u_inference_small_batch_and_512(batch, side):
    from b200seg.models.segmentation_models import R2U_Net
    from oracle import unet_oracle as O
    from oracle.synthetic import xray_batch
    torch.manual_seed(1)
    m = R2U_Net(t=2).cuda().eval()
    x, _ = xray_batch(batch, side, side, seed=3, device="cuda")
    with torch.no_grad():
        y = m(x)
        sd = {k: v.detach().double() if v.is_floating_point() else v for k, v in m.state_dict().items()}
        ref, _ = O.r2u_net_forward(sd, x.double(), t=2, training=False)
    assert y.shape == (batch, 1, side, side)
    err = float((y.double() - ref).norm() / ref.norm())
    print(f"R2U_Net inference b{batch} {side}^2: rel {err:.2e}")
    assert err < 1e-2


@pytest.mark.parametrize("name", ["AttentionUNet", "R2U_Net", "R2AttU_Net"])
def test_inference_bn_folding_matches_unfolded_and_oracle(name):
    """eval + no_grad takes the BN-folded single-launch path; it must agree with the oracle (<= 1e-2) and with the
    unfolded eval path (conv -> BN-apply) of the same module."""
    import os
    from b200seg.models import segmentation_models as M
    from oracle import unet_oracle as O
    from oracle.synthetic import xray_batch
    torch.manual_seed(5)
    kw = {} if name == "AttentionUNet" else {"t": 2}
    m = getattr(M, name)(**kw).cuda()
    x, t = xray_batch(2, 128, 128, seed=9, device="cuda")
    m.train()
    with torch.no_grad():
        for _ in range(3):                        # move the running statistics away from (0, 1)
            m(x)
    m.eval()
    with torch.no_grad():
        y_fold = m(x)
        os.environ["B200SEG_FOLD_BN"] = "0"
        try:
            y_plain = m(x)
        finally:
            os.environ["B200SEG_FOLD_BN"] = "1"
        sd = {k: v.detach().double() if v.is_floating_point() else v for k, v in m.state_dict().items()}
        ref, _ = O.FORWARDS[name](sd, x.double(), training=False, **kw)
        sd32 = {k: v.detach().float() if v.is_floating_point() else v for k, v in m.state_dict().items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):      # the reference's own reduced-precision deviation
            fl, _ = O.FORWARDS[name](sd32, x, training=False, **kw)
    e_fold = float((y_fold.double() - ref).norm() / ref.norm())
    e_plain = float((y_plain.double() - ref).norm() / ref.norm())
    e_floor = float((fl.double() - ref).norm() / ref.norm())
    print(f"{name} eval: folded {e_fold:.2e}  unfolded {e_plain:.2e}  reference-bf16 {e_floor:.2e}")
    tol = max(1e-2, 1.25 * e_floor)
    assert e_fold < tol and e_plain < tol


def test_tester_mirror_matches_oracle_per_sample_metrics():
    """test_segmentation_model (utils/tester.py:249-312 mirror) == average of the oracle's per-sample metrics on the
    same logits; predict_mask == (sigmoid > 0.5) * 255."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import unet_oracle as O
    from oracle.synthetic import xray_batch
    from b200seg.models.segmentation_models import AttentionUNet
    from b200seg.utils import tester as T
    torch.manual_seed(0)
    model = AttentionUNet().cuda()
    x, y = xray_batch(4, 64, 64, seed=3)
    loader = [(x[:2], y[:2]), (x[2:], y[2:])]
    lines = []
    avg = T.test_segmentation_model(model, loader, "cuda", "AttentionUNet", log=lines.append)
    model.eval()
    with torch.no_grad():
        logits = model(x.cuda()).float().cpu()
    want = {k: 0.0 for k in T.METRIC_KEYS}
    for i in range(4):
        m = O.segmentation_metrics(torch.sigmoid(logits[i]), y[i], 0.5)
        for k in want:
            want[k] += m[k] / 4
    for k in T.METRIC_KEYS:
        assert abs(avg[k] - want[k]) < 1e-4 * max(1.0, abs(want[k])), (k, avg[k], want[k])
    assert any("IoU (Jaccard)" in s for s in lines)
    mask = T.predict_mask(model, x[:1])
    assert mask.shape == (64, 64) and mask.dtype.name == "uint8"
    assert (mask == ((logits[0, 0] > 0).numpy() * 255)).all()


@pytest.mark.parametrize("bn_train", [False, True])
@pytest.mark.parametrize("name,kw", [("AttentionUNet", {}), ("R2AttU_Net", {"t": 2})])
def test_wgrad_side_stream_overlap_gives_identical_gradients(name, kw, bn_train):
    """kernels.set_wgrad_overlap(True): weight gradients computed on the side stream are bit-identical to the
    single-stream ones once backward() has returned (the engine callback joins the streams)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle.synthetic import xray_batch
    from b200seg import kernels as K
    from b200seg import ops
    from b200seg.models import segmentation_models as M
    torch.manual_seed(0)
    # eval-mode BatchNorm: gradients are reproducible to ~1e-6, so a race would stand out; train-mode BatchNorm on this
    # tiny batch amplifies the atomic-order noise of the reductions to ~1e-2 and only bounds the error by that noise
    model = getattr(M, name)(**kw).cuda().to(memory_format=torch.channels_last).train(bn_train)
    x, y = xray_batch(4, 64, 64, seed=5)
    x, y = x.cuda(), y.cuda()
    state = {k: v.clone() for k, v in model.state_dict().items()}

    def grads(flag):
        model.load_state_dict(state)
        model.zero_grad(set_to_none=True)
        K.set_wgrad_overlap(flag)
        try:
            loss, _ = ops.seg_loss(model(x), y, 1.0, 0.0, 1.0)
            loss.backward()
        finally:
            K.set_wgrad_overlap(False)
        torch.cuda.synchronize()
        return [p.grad.clone() for p in model.parameters() if p.grad is not None]

    def worst(u, v):
        # global relative error (conv biases in front of a BatchNorm have an exactly-zero gradient whose computed
        # value is rounding noise, so per-tensor ratios are meaningless for them)
        num = sum(float((gu.double() - gv.double()).pow(2).sum()) for gu, gv in zip(u, v))
        den = sum(float(gv.double().pow(2).sum()) for gv in v)
        return (num / den) ** 0.5

    a = grads(False)
    a2 = grads(False)
    b = grads(True)
    b2 = grads(True)
    assert len(a) == len(b) and len(a) > 100
    base = worst(a2, a)              # run-to-run noise of the single-stream path (atomic summation order)
    e1, e2 = worst(b, a), worst(b2, a)
    print(f"{name}: single-stream run-to-run {base:.2e}; overlap vs single-stream {e1:.2e} {e2:.2e}")
    if bn_train:
        assert max(e1, e2) <= max(4.0 * base, 0.1)       # noise-dominated (see above); a race gives O(1)
    else:
        assert max(e1, e2) <= 2.0 * base + 1e-5
    assert not K._OVERLAP["pending"] and not K._OVERLAP["refs"]
    # deterministic-reduction mode removes the atomic-order noise: the comparison becomes BITWISE, train mode included,
    # so a race that corrupts even one element of one gradient is caught
    K.set_deterministic(True)
    try:
        c, c2, d, d2 = grads(False), grads(False), grads(True), grads(True)
    finally:
        K.set_deterministic(False)
    for u, v, w_, z in zip(c, c2, d, d2):
        assert torch.equal(u, v) and torch.equal(u, w_) and torch.equal(u, z)


def test_grad_reducer_on_side_stream_matches_plain_backward():
    """GradReducer (world size 1, NCCL) with the weight gradients on the side stream: the hook packs and reduces on
    that stream; after finish() the gradients equal those of a plain backward."""
    import sys
    from pathlib import Path
    import torch.distributed as dist
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle.synthetic import xray_batch
    from b200seg import kernels as K
    from b200seg import ops
    from b200seg.ddp import GradReducer
    from b200seg.models.segmentation_models import AttentionUNet
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        torch.manual_seed(0)
        model = AttentionUNet().cuda().to(memory_format=torch.channels_last).eval()   # eval BN: reproducible grads
        x, y = xray_batch(2, 64, 64, seed=6)
        x, y = x.cuda(), y.cuda()

        def run(with_reducer):
            model.zero_grad(set_to_none=True)
            red = GradReducer(model, bucket_mb=4) if with_reducer else None
            K.set_wgrad_overlap(with_reducer)
            try:
                loss, _ = ops.seg_loss(model(x), y, 1.0, 0.0, 1.0)
                loss.backward()
                if red is not None:
                    red.finish()
                    red.remove()
            finally:
                K.set_wgrad_overlap(False)
            torch.cuda.synchronize()
            return [p.grad.clone() for p in model.parameters()]

        a, b = run(False), run(True)
        num = sum(float((u.double() - v.double()).pow(2).sum()) for u, v in zip(b, a))
        den = sum(float(v.double().pow(2).sum()) for v in a)
        assert (num / den) ** 0.5 < 1e-5
    finally:
        if created:
            dist.destroy_process_group()
