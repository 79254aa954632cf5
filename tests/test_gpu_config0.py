"""GPU parity on BASELINE.json configs[0] EXACTLY — batch 4, 3x256x256 synthetic X-ray + binary mask, the reference's
own random init — for AttU_Net, R2U_Net (t=2) and R2AttU_Net (t=2), against the fp64 oracle run on the same device:

  eval-mode BatchNorm (level iii): logits <= 1e-2 rel-L2, global weight gradient <= 2e-2 (north_star, absolute);
  train-mode BatchNorm (level iv): logits AND global weight gradient <= 1.25x the reference's own bf16-autocast deviation
  measured in the same run (train-mode BN at random init is chaotic: SURVEY.md Appendix C), exact num_batches_tracked.

At 256^2 x 4 the 256^2 / 128^2 levels have >= 4 m-tiles per SM, so the double-M work items of conv_igemm and the
row-pair / X-halo modes of conv_wgrad run INSIDE a whole model here (at 2x128^2 they never trigger).

Also: the constructor-default recurrence depth t=5 (what get_seg_model builds, utils/helpers.py:209-211) on the GPU.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

MODELS = {"AttentionUNet": {}, "R2U_Net": {"t": 2}, "R2AttU_Net": {"t": 2}}


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _sd(m, dtype):
    return {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}


def _build(name, kw, seed):
    from b200seg.models import segmentation_models as M
    torch.manual_seed(seed)
    return getattr(M, name)(**kw).cuda()


def _autocast_step(O, name, sd32, x, t, training, kw):
    """the reference's own reduced-precision path: bf16 autocast forward, fp32 loss (helpers.py:321-329)"""
    params = {k: v.clone().requires_grad_(True) for k, v in sd32.items()
              if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, _ = O.FORWARDS[name]({**sd32, **params}, x, training=training, **kw)
    loss = O.bce_with_logits(logits.float(), t)
    grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    return logits.detach(), dict(zip(params.keys(), grads))


def _global_grad_err(grads, ref_grads):
    num = den = 0.0
    for k, g in ref_grads.items():
        if g is None:
            continue
        assert grads[k] is not None, k
        num += float((grads[k].double() - g).norm() ** 2)
        den += float(g.norm() ** 2)
    return (num / den) ** 0.5


def _run(name, kw, training, batch=4, side=256, seed=0):
    from b200seg import ops
    from b200seg.utils.synthetic import xray_batch
    from oracle import unet_oracle as O
    m = _build(name, kw, seed)
    m.train(training)
    x, t = xray_batch(batch, side, side, seed=7, device="cuda")
    sd64, sd32 = _sd(m, torch.float64), _sd(m, torch.float32)
    logits = m(x)
    loss, _ = ops.seg_loss(logits, t, 1.0, 0.0, 1.0)
    loss.backward()
    mine = {k: p.grad for k, p in m.named_parameters()}
    ref, ref_loss, ref_g, newb = O.train_step_grads(name, sd64, x.double(), t.double(), training=training, **kw)
    del sd64
    fl, fl_g = _autocast_step(O, name, sd32, x, t, training, kw)
    e, e_floor = rel(logits, ref), rel(fl.float(), ref)
    glob, gfloor = _global_grad_err(mine, ref_g), _global_grad_err(fl_g, ref_g)
    mode = "train" if training else "eval"
    print(f"{name}{kw} {batch}x{side}^2 {mode}: logits ours {e:.3e} ref-bf16 {e_floor:.3e}; "
          f"global weight-grad ours {glob:.3e} ref-bf16 {gfloor:.3e}; loss {float(loss):.6f} vs {float(ref_loss):.6f}")
    return m, e, e_floor, glob, gfloor, float(loss), float(ref_loss), newb


@pytest.mark.parametrize("name", list(MODELS))
def test_config0_eval(name):
    m, e, e_floor, glob, gfloor, loss, ref_loss, _ = _run(name, MODELS[name], training=False)
    assert e < 1e-2
    assert glob < 2e-2
    assert abs(loss - ref_loss) < 1e-2 * abs(ref_loss)


@pytest.mark.parametrize("name", list(MODELS))
def test_config0_train(name):
    m, e, e_floor, glob, gfloor, loss, ref_loss, newb = _run(name, MODELS[name], training=True)
    assert e < max(1.25 * e_floor, 2e-2)
    assert glob < max(1.25 * gfloor, 4e-2)
    msd = m.state_dict()
    for k, v in newb.items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
    rv = max(rel(msd[k], v) for k, v in newb.items() if k.endswith("running_var"))
    assert rv < max(1.25 * e_floor, 2e-2)


@pytest.mark.parametrize("name", ["R2U_Net", "R2AttU_Net"])
def test_ctor_default_t5(name):
    """R2U_Net() / R2AttU_Net() as get_seg_model builds them (t=5): eval-mode step within the north_star gates, and a
    train-mode forward advances every Recurrent_block's num_batches_tracked by t+1 = 6."""
    from b200seg.utils.helpers import get_seg_model
    from b200seg.utils.synthetic import xray_batch
    torch.manual_seed(0)
    probe = get_seg_model("r2unet" if name == "R2U_Net" else "r2attunet")
    assert type(probe).__name__ == name and probe.RRCNN1.RCNN[0].t == 5
    del probe
    m, e, e_floor, glob, gfloor, loss, ref_loss, _ = _run(name, {}, training=False, batch=2, side=128)
    assert e < 1e-2 and glob < 2e-2
    m.train()
    x, _ = xray_batch(2, 128, 128, seed=8, device="cuda")
    with torch.no_grad():
        m(x)
    assert int(m.RRCNN3.RCNN[1].conv[1].num_batches_tracked) == 6
    assert int(m.up_RRCNN2.RCNN[0].conv[1].num_batches_tracked) == 6
