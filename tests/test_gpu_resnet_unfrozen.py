"""GPU: ResNetUnet(freeze=False) (reference ResnetUnet.py:29-30 — the torchvision ResNet-50 encoder trains too): the
encoder's backward (7x7/s2 stem as an im2col GEMM, MaxPool2d(3,2,1) backward, strided 3x3 / 1x1 convolutions through the
zero-inserted dY, Bottleneck residual) against the fp64 oracle, every parameter of the model included.

eval-mode BatchNorm: north_star gates, absolute (logits 1e-2, global weight gradient 2e-2);
train-mode BatchNorm: <= 1.25x the reference's own bf16-autocast deviation measured in the same run."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _sd(m, dtype):
    return {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}


def _ref_step(O, sd, x, t, training, names, autocast):
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    work = {**sd, **params}
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, newb = O.resnet_unet_forward(work, x, training=training)
    else:
        logits, newb = O.resnet_unet_forward(work, x, training=training)
    loss = O.bce_with_logits(logits.to(t.dtype), t)
    grads = torch.autograd.grad(loss, list(params.values()))
    return logits.detach(), dict(zip(params.keys(), grads)), newb


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_resnet_unet_unfrozen_end_to_end(mode):
    import warnings
    from b200seg import ops
    from b200seg.models.segmentation_models import ResNetUnet
    from b200seg.utils.synthetic import xray_batch
    from oracle import unet_oracle as O
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ResNetUnet(freeze=False).cuda()
    names = [k for k, p in m.named_parameters() if p.requires_grad]
    assert len(names) == len(list(m.parameters())) and any(k.startswith("encoder1.0") for k in names)
    training = mode == "train"
    m.train(training)
    x, t = xray_batch(2, 128, 128, seed=5, device="cuda")
    sd64, sd32 = _sd(m, torch.float64), _sd(m, torch.float32)
    logits = m(x)
    loss, _ = ops.seg_loss(logits, t, 1.0, 0.0, 1.0)
    loss.backward()
    ref, ref_g, newb = _ref_step(O, sd64, x.double(), t.double(), training, names, False)
    fl, fl_g, _ = _ref_step(O, sd32, x, t, training, names, True)
    params = dict(m.named_parameters())
    missing = [k for k in names if params[k].grad is None]
    assert not missing, missing[:5]

    def glob(grads, prefix):
        ks = [k for k in ref_g if k.startswith(prefix)]
        den = sum(float(ref_g[k].norm() ** 2) for k in ks)
        return (sum(float((grads[k].double() - ref_g[k]).norm() ** 2) for k in ks) / den) ** 0.5

    mine = {k: params[k].grad for k in names}
    e, e_floor = rel(logits, ref), rel(fl.float(), ref)
    g_all, f_all = glob(mine, ""), glob(fl_g, "")
    g_enc, f_enc = glob(mine, "encoder"), glob(fl_g, "encoder")
    g_stem, f_stem = glob(mine, "encoder1.0"), glob(fl_g, "encoder1.0")
    print(f"ResNetUnet(freeze=False)/{mode}: logits ours {e:.3e} ref-bf16 {e_floor:.3e}; weight-grad global ours "
          f"{g_all:.3e} ref-bf16 {f_all:.3e}; encoder only ours {g_enc:.3e} ref-bf16 {f_enc:.3e}; "
          f"7x7 stem ours {g_stem:.3e} ref-bf16 {f_stem:.3e}")
    if mode == "eval":
        assert e < 1e-2 and g_all < 2e-2 and g_enc < 2e-2 and g_stem < 2e-2
    else:
        assert e < max(1.25 * e_floor, 2e-2)
        assert g_all < max(1.25 * f_all, 4e-2) and g_enc < max(1.25 * f_enc, 4e-2)
        msd = m.state_dict()
        for k, v in newb.items():
            if k.endswith("num_batches_tracked"):
                assert int(msd[k]) == int(v), k


def test_maxpool3x3s2_backward_matches_torch():
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(2, 64, 32, 32, device="cuda", generator=g).to(torch.bfloat16)
    x[:, :, ::3, ::5] = 0.5              # ties inside windows: the first maximum must win, as in ATen
    x[:, :, 1::3, 1::5] = 0.5
    dy = torch.randn(2, 64, 16, 16, device="cuda", generator=g).to(torch.bfloat16)
    xr = x.float().requires_grad_(True)
    torch.nn.functional.max_pool2d(xr, 3, 2, 1).backward(dy.float())
    dx = K.maxpool3x3s2_bwd(dy.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous())
    assert rel(dx.permute(0, 3, 1, 2), xr.grad) < 4e-3


@pytest.mark.parametrize("k,stride,cin,cout", [(3, 2, 128, 128), (1, 2, 256, 512), (3, 1, 64, 64), (1, 1, 64, 256)])
def test_res_conv_bn_block_gradients(k, stride, cin, cout):
    """one encoder layer with identity + ReLU vs fp64 torch autograd on the same bf16-rounded operands (eval-mode BN:
    deterministic, absolute gates)"""
    import torch.nn.functional as F
    from b200seg import ops_resnet as R
    g = torch.Generator(device="cuda").manual_seed(6)
    n, h = 2, 32
    x = torch.randn(n, h, h, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).requires_grad_(True)
    gamma = (torch.rand(cout, device="cuda", generator=g) + 0.5).requires_grad_(True)
    beta = torch.randn(cout, device="cuda", generator=g).requires_grad_(True)
    rm = torch.randn(cout, device="cuda", generator=g) * 0.1
    rv = torch.rand(cout, device="cuda", generator=g) + 0.5
    ho = h // stride
    idt = torch.randn(n, ho, ho, cout, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
    dy = torch.randn(n, ho, ho, cout, device="cuda", generator=g).to(torch.bfloat16)
    xi = x.clone().requires_grad_(True)
    y, _z, _c, _s = R.res_conv_bn(xi, w, gamma, beta, rm, rv, idt, stride, False, 1e-5, True)
    y.backward(dy)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.detach().to(torch.bfloat16).double().requires_grad_(True)
    gr, br = gamma.detach().double().requires_grad_(True), beta.detach().double().requires_grad_(True)
    ir = idt.detach().double().permute(0, 3, 1, 2).requires_grad_(True)
    z = F.conv2d(xr, wr, stride=stride, padding=k // 2)
    z = z.detach().to(torch.bfloat16).double() + (z - z.detach())   # the CUDA path stores z in bf16 (straight-through)
    yr = F.relu(F.batch_norm(z, rm.double(), rv.double(), gr, br, False, 0.1, 1e-5) + ir)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    assert rel(y.permute(0, 3, 1, 2), yr) < 1e-2
    # every gradient of this block passes through the mask of relu(bn(z) + identity); bf16 rounding of bn(z) and of the
    # sum flips ~0.1 % of the masks of this random data, and a flipped mask costs sqrt(f) ~ 3e-2 in relative L2 (the
    # reference's own bf16 autocast sits at 5e-2 .. 7e-2 on such blocks)
    assert rel(xi.grad.permute(0, 3, 1, 2), xr.grad) < 5e-2
    assert rel(w.grad, wr.grad) < 5e-2
    # d(identity) = dy * [out > 0] carries every ReLU-mask flip at full weight (bf16 rounding of bn(z) and of the sum
    # flips ~0.05 % of the masks of this random data => sqrt(f) ~ 2.3e-2); the other gradients average flips out
    assert rel(idt.grad.permute(0, 3, 1, 2), ir.grad) < 5e-2
    assert rel(gamma.grad, gr.grad) < 5e-2 and rel(beta.grad, br.grad) < 5e-2      # (mask flips, as above)
