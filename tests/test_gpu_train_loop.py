"""GPU: the train() mirror of utils/helpers.py:231-412 runs end to end on the CUDA path — loss decreases on a tiny
synthetic set, the best checkpoint has the reference's state_dict layout and reloads strictly — and the R2U_Net
inference path (BASELINE.json configs[4]) handles batch 1 and 512x512 inputs."""
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

pytestmark = pytest.mark.gpu


def test_train_mirror_runs_and_saves_reference_layout(tmp_path):
    from b200seg.utils.helpers import get_seg_model, train
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
