"""GPU: deterministic-reduction mode (kernels.set_deterministic / B200SEG_DETERMINISTIC=1, C ABI b2_set_deterministic).

The reference's BatchNorm (cuDNN, AttentionUNet.py:7) is run-to-run reproducible; the default CUDA path here
accumulates BatchNorm statistics, backward sums, bias / head gradients, loss sums and the gradient norm with fp64 / fp32
atomics, whose summation order changes from run to run (~1e-2 relative on small train-mode batches, amplified by the
chaotic random-init BatchNorm).  In deterministic mode those reductions go through per-block partials and a fixed-order
second stage: two runs from the same state must be BIT-identical — logits, loss, every gradient, the running
statistics, and the parameters after the fused clip + AdamW step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def deterministic():
    from b200seg import kernels as K
    K.set_deterministic(True)
    try:
        yield
    finally:
        K.set_deterministic(False)


def _snapshot(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def _train_step(model, opt_state, x, t, overlap):
    from b200seg import kernels as K, ops
    from b200seg.optim import FusedClipAdamW
    model.load_state_dict(opt_state)
    model.zero_grad(set_to_none=True)
    opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-4, max_norm=1.0)
    K.set_wgrad_overlap(overlap)
    try:
        K.step_begin()
        logits = model(x)
        loss, sums = ops.seg_loss(logits, t, 0.5, 0.5, 1.0)       # BCE + Dice: every sum of the loss kernel is used
        loss.backward()
    finally:
        K.set_wgrad_overlap(False)
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    opt.step()
    torch.cuda.synchronize()
    return logits.detach().clone(), loss.detach().clone(), grads, _snapshot(model), float(opt.total_norm)


def _assert_bit_equal(a, b, what):
    la, lo, ga, sa, na = a
    lb, lo2, gb, sb, nb = b
    assert torch.equal(la, lb), f"{what}: logits differ"
    assert torch.equal(lo, lo2), f"{what}: loss differs"
    assert ga.keys() == gb.keys()
    bad = [k for k in ga if not torch.equal(ga[k], gb[k])]
    assert not bad, f"{what}: {len(bad)} gradients differ, e.g. {bad[:4]}"
    bad = [k for k in sa if not torch.equal(sa[k], sb[k])]
    assert not bad, f"{what}: state after the optimizer step differs, e.g. {bad[:4]}"
    assert na == nb, f"{what}: gradient norm differs"


@pytest.mark.parametrize("name,kw,batch,side", [("AttentionUNet", {}, 4, 128), ("R2AttU_Net", {"t": 2}, 3, 64),
                                                ("ResNetUnet", {}, 2, 128), ("AttentionUNet", {}, 4, 256)])
def test_train_step_is_bit_reproducible(deterministic, name, kw, batch, side):
    from b200seg.models import segmentation_models as M
    from b200seg.utils.synthetic import xray_batch
    torch.manual_seed(0)
    model = getattr(M, name)(**kw).cuda().to(memory_format=torch.channels_last).train()
    x, t = xray_batch(batch, side, side, seed=2, device="cuda")
    state = _snapshot(model)
    runs = [_train_step(model, state, x, t, overlap=False) for _ in range(3)]
    _assert_bit_equal(runs[0], runs[1], f"{name} run 0 vs 1")
    _assert_bit_equal(runs[0], runs[2], f"{name} run 0 vs 2")
    # weight gradients on the side stream (what train() and bench.py use): still the same bits
    ov = _train_step(model, state, x, t, overlap=True)
    _assert_bit_equal(runs[0], ov, f"{name} side-stream weight gradients")


def test_default_mode_is_close_to_deterministic_mode():
    """the two reduction paths compute the same sums: eval-mode gradients (no chaotic amplification) agree to fp32
    rounding of the reductions"""
    from b200seg import kernels as K, ops
    from b200seg.models.segmentation_models import AttentionUNet
    from b200seg.utils.synthetic import xray_batch
    torch.manual_seed(0)
    model = AttentionUNet().cuda().eval()
    x, t = xray_batch(2, 128, 128, seed=3, device="cuda")

    def grads():
        model.zero_grad(set_to_none=True)
        loss, _ = ops.seg_loss(model(x), t, 1.0, 0.0, 1.0)
        loss.backward()
        torch.cuda.synchronize()
        return [p.grad.clone() for p in model.parameters()]

    a = grads()
    K.set_deterministic(True)
    try:
        b = grads()
    finally:
        K.set_deterministic(False)
    num = sum(float((u.double() - v.double()).pow(2).sum()) for u, v in zip(a, b))
    den = sum(float(v.double().pow(2).sum()) for v in b)
    assert (num / den) ** 0.5 < 1e-5


def test_ops_level_bit_reproducible(deterministic):
    """every reducing entry point on its own, 5 repeats each"""
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    n, h, w, c = 4, 64, 64, 128
    x = torch.randn(n, h, w, c, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(n, h, w, c, device="cuda", generator=g).to(torch.bfloat16)
    wt = torch.randn(c, c, 3, 3, device="cuda", generator=g) / (c * 9) ** 0.5
    wf, _ = K.pack_weights(wt)
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g)

    def once():
        stats = torch.zeros(2, c, dtype=torch.float64, device="cuda")
        z = K.conv_igemm(x, wf, c, 3, stats=stats)
        coef = K.bn_finalize(stats, n * h * w, gamma, beta, 1e-5, 0.0, None, None, None)
        dz, dgamma, dbeta, dbias = K.bn_bwd(dy, z, coef, gamma, relu=True, training=True, want_dbias=True)
        st2 = torch.zeros(2, c, dtype=torch.float64, device="cuda")
        K.channel_stats(z, st2)
        db = K.channel_sum(dy)
        zf = torch.randn(n * h * w, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        tf = (zf > 0.3).float()
        loss, sums = K.loss_fwd(zf, tf, 0.5, 0.5, 1.0)
        return [stats, dz, dgamma, dbeta, dbias, st2, db, loss, sums]

    ref = once()
    for _ in range(4):
        for a, b in zip(ref, once()):
            assert torch.equal(a, b)
