"""CPU: pins the oracle (oracle/unet_oracle.py) against golden vectors produced by the UNMODIFIED reference modules
(oracle/make_golden.py, run in the build container).  Everything is fp64, so agreement is to rounding (1e-9)."""
from collections import OrderedDict
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
from oracle.synthetic import fill_state_dict_, xray_batch

GOLD = Path(__file__).resolve().parent / "golden"
SEED = 7
CASES = {
    "AttentionUNet": ("AttentionUNet", {}, 2, 32),
    "R2U_Net": ("R2U_Net", {"t": 2}, 2, 32),
    "R2AttU_Net": ("R2AttU_Net", {"t": 2}, 2, 32),
    "R2U_Net_t5": ("R2U_Net", {}, 2, 32),
    "ResNetUnet": ("ResNetUnet", {}, 2, 64),
}


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _state_dict_from_golden(g):
    sd = OrderedDict()
    for k, shp, dt in zip(g["keys"], g["shapes"], g["dtypes"]):
        shape = tuple(int(s) for s in str(shp).split(",") if s)
        sd[str(k)] = torch.zeros(shape, dtype=getattr(torch, str(dt)))
    fill_state_dict_(sd, SEED)
    return OrderedDict((k, v.double() if v.is_floating_point() else v) for k, v in sd.items())


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_golden(case):
    name, kw, n, side = CASES[case]
    g = np.load(GOLD / f"{case}.npz")
    sd = _state_dict_from_golden(g)
    x, t = xray_batch(n, side, side, seed=11)
    x, t = x.double(), t.double()
    with torch.no_grad():
        ev, _ = O.FORWARDS[name](sd, x, training=False, **kw)
    assert rel(ev, g["eval_logits"]) < 1e-9
    logits, loss, grads, newb = O.train_step_grads(name, sd, x, t, training=True, **kw)
    assert rel(logits, g["train_logits"]) < 1e-9
    assert abs(float(loss) - float(g["train_loss"])) < 1e-10
    # gradients: every tensor's norm, small tensors in full, the first 256 entries of large ones
    frozen = {str(k) for k, r in zip(g["param_names"], g["requires_grad"]) if not r}
    for k, nrm in zip(g["grad_names"], g["grad_norms"]):
        k = str(k)
        got = grads[k]
        assert abs(float(got.norm()) - float(nrm)) <= 1e-8 * max(float(nrm), 1e-12) + 1e-14, k
        if f"grad::{k}" in g.files:
            assert rel(got, g[f"grad::{k}"]) < 1e-7 or float(nrm) < 1e-9, k
        else:
            assert rel(got.reshape(-1)[:256], g[f"gradhead::{k}"]) < 1e-7 or float(nrm) < 1e-9, k
    assert frozen.isdisjoint(set(map(str, g["grad_names"])))
    # BatchNorm side effects: call counts (t+1 per Recurrent_block forward) and momentum-updated running stats
    for k, v in zip(g["nbt_names"], g["nbt"]):
        assert int(newb[str(k)]) == int(v), k
    for k, nrm in zip(g["running_names"], g["running_norms"]):
        assert abs(float(newb[str(k)].norm()) - float(nrm)) <= 1e-9 * float(nrm), k


def test_recurrent_block_call_count():
    g = np.load(GOLD / "R2U_Net_t5.npz")
    assert set(int(v) for v in g["nbt"]) == {1, 6}          # UpConv BNs once; Recurrent_block BNs t+1 = 6 times
    g2 = np.load(GOLD / "R2U_Net.npz")
    assert set(int(v) for v in g2["nbt"]) == {1, 3}


def test_losses_match_reference():
    g = np.load(GOLD / "losses.npz")
    z = torch.from_numpy(g["z"]).requires_grad_(True)
    t = torch.from_numpy(g["t"])
    assert abs(float(O.bce_with_logits(z, t)) - float(g["bce"])) < 1e-12
    assert abs(float(O.dice_loss(z, t)) - float(g["dice"])) < 1e-12
    c = O.combined_loss(z, t)
    assert abs(float(c) - float(g["combined"])) < 1e-12
    c.backward()
    assert rel(z.grad, g["combined_grad"]) < 1e-10
    assert abs(O.iou(torch.sigmoid(z.detach()), t) - float(g["iou"])) < 1e-9


def test_segmentation_metrics_match_reference():
    """oracle.segmentation_metrics and the host-side metrics_from_counts (b200seg.utils.tester) against the outputs of the
    reference's calculate_segmentation_metrics (tests/golden/metrics.npz, made by oracle/make_golden.py)."""
    from b200seg.utils.tester import METRIC_KEYS, metrics_from_counts
    g = np.load(GOLD / "metrics.npz")
    z, t = torch.from_numpy(g["z"]), torch.from_numpy(g["t"])
    assert tuple(str(k) for k in g["keys"]) == METRIC_KEYS
    for thr, name in ((0.5, "metrics_thr5"), (0.3, "metrics_thr3")):
        want = g[name]
        for i in range(z.shape[0]):
            m = O.segmentation_metrics(torch.sigmoid(z[i]), t[i], thr)
            got = np.array([m[k] for k in METRIC_KEYS])
            assert np.allclose(got, want[i], rtol=2e-6, atol=1e-6), (thr, i, got, want[i])
        # host formula on integer counts (what the CUDA kernel produces)
        pred = torch.sigmoid(z) > thr
        tgt = t > thr
        counts = torch.stack([(pred & tgt).flatten(1).sum(1), pred.flatten(1).sum(1), tgt.flatten(1).sum(1)], 1)
        m = metrics_from_counts(counts, z[0].numel())
        got = np.stack([m[k].numpy() for k in METRIC_KEYS], 1)
        assert np.allclose(got, want, rtol=2e-6, atol=1e-6)
