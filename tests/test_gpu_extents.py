"""GPU: general image extents.  The reference is fully convolutional (AttentionUNet.py:86-121: any H, W divisible by 16;
32 for ResNetUnet) — 224^2, 320^2, 384^2 ... — so the tcgen05 kernels accept any extent: tiles at the right / bottom /
batch edge are clipped (TMA zero-fills loads and drops stores outside the tensor) and masked out of the BatchNorm
statistics.  Op level vs fp32 torch on the same bf16-rounded operands, then whole models vs the fp64 oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("n,h,w,cin,cout,k", [
    (1, 224, 224, 64, 64, 3), (2, 112, 112, 64, 128, 3), (2, 56, 56, 128, 256, 3), (3, 28, 28, 256, 512, 3),
    (2, 14, 14, 512, 512, 3), (1, 320, 320, 64, 64, 3), (2, 160, 160, 64, 128, 3), (2, 80, 80, 128, 128, 3),
    (3, 40, 40, 128, 256, 3), (5, 20, 20, 256, 256, 3), (2, 24, 40, 64, 64, 3), (1, 12, 20, 64, 128, 3),
    (2, 56, 56, 128, 64, 1), (3, 14, 14, 512, 256, 1), (2, 224, 224, 64, 32, 1), (7, 3, 3, 64, 64, 3)])
def test_conv_any_extent(n, h, w, cin, cout, k):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    dy = torch.randn(n, cout, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    xb, dyb = nhwc(x), nhwc(dy)
    wf, wd = K.pack_weights(wt)
    stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    y = K.conv_igemm(xb, wf, cout, k, bias=b, stats=stats)
    dx = K.conv_igemm(dyb, wd, cin, k, dgrad=True)
    dw = K.conv_wgrad(dyb, xb, k)
    torch.cuda.synchronize()
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    xr = nchw(xb).requires_grad_(True)
    ref = F.conv2d(xr, wr, b, padding=k // 2)
    ref.backward(nchw(dyb))
    e = (rel(nchw(y), ref), rel(nchw(dx), xr.grad), rel(dw.view(cout, k, k, cin).permute(0, 3, 1, 2), wr.grad))
    print(f"{n}x{h}x{w} {cin}->{cout} k{k}: fprop {e[0]:.1e} dgrad {e[1]:.1e} wgrad {e[2]:.1e}")
    assert max(e) < 4e-3, e
    yf = nchw(y).double()
    assert rel(stats[0], yf.sum((0, 2, 3))) < 1e-6 and rel(stats[1], (yf * yf).sum((0, 2, 3))) < 1e-6


@pytest.mark.parametrize("name,kw,batch,side", [("AttentionUNet", {}, 2, 224), ("AttentionUNet", {}, 1, 320),
                                                ("R2AttU_Net", {"t": 2}, 1, 224), ("ResNetUnet", {}, 2, 224),
                                                ("R2U_Net", {"t": 2}, 3, 48), ("AttentionUNet", {}, 1, 384)])
def test_models_at_other_resolutions(name, kw, batch, side):
    """eval-mode forward + backward at extents that are neither powers of two nor multiples of 128 at every level
    (224 -> 112, 56, 28, 14; 320 -> 160, 80, 40, 20; 48 -> 24, 12, 6, 3): north_star gates vs the fp64 oracle, and a
    train-mode step whose BatchNorm statistics count exactly the real pixels (running_var vs the oracle)"""
    import warnings
    from b200seg import ops
    from b200seg.models import segmentation_models as M
    from b200seg.utils.synthetic import xray_batch
    from oracle import unet_oracle as O
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = getattr(M, name)(**kw).cuda().eval()
    x, t = xray_batch(batch, side, side, seed=3, device="cuda")
    logits = m(x)
    loss, _ = ops.seg_loss(logits, t, 1.0, 0.0, 1.0)
    loss.backward()
    sd = {k: (v.detach().double() if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}
    if name == "ResNetUnet":
        trainable = [k for k, p in m.named_parameters() if p.requires_grad]
        params = {k: sd[k].clone().requires_grad_(True) for k in trainable}
        ref, _ = O.resnet_unet_forward({**sd, **params}, x.double(), training=False)
        grads = dict(zip(params, torch.autograd.grad(O.bce_with_logits(ref, t.double()), list(params.values()))))
    else:
        ref, _, grads, _ = O.train_step_grads(name, sd, x.double(), t.double(), training=False, **kw)
    e = rel(logits, ref)
    mine = dict(m.named_parameters())
    den = sum(float(g.norm() ** 2) for g in grads.values() if g is not None)
    glob = (sum(float((mine[k].grad.double() - g).norm() ** 2) for k, g in grads.items() if g is not None) / den) ** 0.5
    print(f"{name}{kw} {batch}x{side}^2 eval: logits {e:.3e}, global weight-grad {glob:.3e}")
    assert logits.shape == (batch, 1, side, side)
    assert e < 1e-2 and glob < 2e-2
    # train mode: clipped tiles must not leak into the batch statistics
    m.train()
    sd64 = {k: (v.detach().double() if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}
    with torch.no_grad():
        m(x)
        _, newb = O.FORWARDS[name](sd64, x.double(), training=True, **kw)
    msd = m.state_dict()
    first_bn = [k for k in newb if k.endswith("running_var")][:3]      # the first layers: before bf16 noise compounds
    for k in first_bn:
        assert rel(msd[k], newb[k]) < 2e-2, (k, rel(msd[k], newb[k]))
    for k, v in newb.items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
