"""GPU parity for the ResNetUnet-specific kernels: strided tcgen05 convs, ConvTranspose2d(2,2) fwd/bwd, 7x7 stem,
3x3/s2 max-pool, residual BN, and the model end to end (eval: absolute gates; train: vs the reference's own bf16)."""
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


@pytest.mark.parametrize("n,hin,cin,cout,k", [(2, 64, 128, 128, 3), (2, 32, 256, 256, 3), (3, 16, 512, 512, 3),
                                              (2, 64, 256, 512, 1), (2, 16, 1024, 2048, 1)])
def test_strided_conv(n, hin, cin, cout, k):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(11)
    x = nhwc(torch.randn(n, cin, hin, hin, device="cuda", generator=g))
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    wf, _ = K.pack_weights(wt, want_dgrad=False)
    stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    y = K.conv_igemm(x, wf, cout, k, stats=stats, stride=2)
    ref = F.conv2d(nchw(x), wt.to(torch.bfloat16).float(), None, stride=2, padding=k // 2)
    assert rel(nchw(y), ref) < 4e-3
    assert rel(stats[0], y.float().reshape(-1, cout).double().sum(0)) < 1e-5


@pytest.mark.parametrize("n,h,cin,cout", [(2, 8, 2048, 2048), (2, 32, 512, 512), (2, 64, 256, 256), (2, 128, 64, 32),
                                          (1, 16, 1024, 1024)])
def test_conv_transpose2x2(n, h, cin, cout):
    from b200seg import ops_resnet as R
    g = torch.Generator(device="cuda").manual_seed(12)
    x = nhwc(torch.randn(n, cin, h, h, device="cuda", generator=g)).requires_grad_(True)
    wt = (torch.randn(cin, cout, 2, 2, device="cuda", generator=g) / cin ** 0.5).requires_grad_(True)
    b = torch.randn(cout, device="cuda", generator=g).requires_grad_(True)
    y = R.conv_transpose2x2(x, wt, b)
    dy = nhwc(torch.randn(n, cout, 2 * h, 2 * h, device="cuda", generator=g))
    y.backward(dy)
    xr = nchw(x.detach()).requires_grad_(True)
    wr = wt.detach().to(torch.bfloat16).float().requires_grad_(True)
    br = b.detach().clone().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, br, stride=2)
    yr.backward(nchw(dy))
    print(f"convT {cin}->{cout} @{h}: y {rel(nchw(y), yr):.2e} dx {rel(nchw(x.grad), xr.grad):.2e} "
          f"dw {rel(wt.grad, wr.grad):.2e} db {rel(b.grad, br.grad):.2e}")
    assert rel(nchw(y), yr) < 4e-3
    assert rel(nchw(x.grad), xr.grad) < 4e-3
    assert rel(wt.grad, wr.grad) < 1e-3
    assert rel(b.grad, br.grad) < 1e-3


def test_stem7x7_and_maxpool3():
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(2, 3, 64, 96, device="cuda", generator=g)
    wt = torch.randn(64, 3, 7, 7, device="cuda", generator=g) * 0.1
    x4 = K.image_to_nhwc4(x)
    z = K.stem7x7_fprop(x4, K.pack_small_weight(wt))
    ref = F.conv2d(nchw(x4[..., :3]), wt.to(torch.bfloat16).float(), None, stride=2, padding=3)
    assert rel(nchw(z), ref) < 4e-3
    p = K.maxpool3x3s2_fwd(z)
    assert torch.equal(nchw(p), F.max_pool2d(nchw(z), 3, 2, 1))


def _setup(seed, init):
    import warnings
    from b200seg.models.segmentation_models import ResNetUnet
    from oracle.synthetic import fill_state_dict_
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ResNetUnet()
    if init == "synthetic":
        fill_state_dict_(m.state_dict(), seed)
    return m.cuda()


def _sd(m, dtype):
    return {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}


def _ref_step(O, sd, x, t, training, trainable, autocast):
    params = {k: sd[k].clone().requires_grad_(True) for k in trainable}
    work = {**sd, **params}
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, newb = O.resnet_unet_forward(work, x, training=training)
    else:
        logits, newb = O.resnet_unet_forward(work, x, training=training)
    loss = O.bce_with_logits(logits.to(t.dtype), t)
    grads = torch.autograd.grad(loss, list(params.values()))
    return logits.detach(), loss.detach(), dict(zip(params.keys(), grads)), newb


@pytest.mark.parametrize("mode,init", [("eval", "default"), ("train", "default"), ("train", "synthetic")])
def test_resnet_unet_end_to_end(mode, init):
    from oracle import unet_oracle as O
    from oracle.synthetic import xray_batch
    from b200seg import ops
    m = _setup(2, init)
    training = mode == "train"
    m.train(training)
    x, t = xray_batch(2, 256, 256, seed=5, device="cuda")
    trainable = [k for k, p in m.named_parameters() if p.requires_grad]
    assert all(not k.startswith("encoder") for k in trainable)
    sd64, sd32 = _sd(m, torch.float64), _sd(m, torch.float32)
    logits = m(x)
    loss, _ = ops.seg_loss(logits, t, 1.0, 0.0, 1.0)
    loss.backward()
    ref, ref_loss, ref_g, newb = _ref_step(O, sd64, x.double(), t.double(), training, trainable, False)
    fl, fl_loss, fl_g, _ = _ref_step(O, sd32, x, t, training, trainable, True)
    e, e_floor = rel(logits, ref), rel(fl.float(), ref)
    params = dict(m.named_parameters())
    den = sum(float(g.norm() ** 2) for g in ref_g.values())
    glob = (sum(float((params[k].grad.double() - g).norm() ** 2) for k, g in ref_g.items()) / den) ** 0.5
    gfloor = (sum(float((fl_g[k].double() - g).norm() ** 2) for k, g in ref_g.items()) / den) ** 0.5
    print(f"ResNetUnet/{mode}/{init}: logits ours {e:.3e} ref-bf16 {e_floor:.3e}; global grad ours {glob:.3e} "
          f"ref-bf16 {gfloor:.3e}")
    if mode == "eval":
        assert e < 1e-2 and glob < 2e-2
    else:
        assert e < max(1.25 * e_floor, 2e-2)
        assert glob < max(1.25 * gfloor, 4e-2)
        msd = m.state_dict()
        for k, v in newb.items():
            if k.endswith("num_batches_tracked"):
                assert int(msd[k]) == int(v), k
