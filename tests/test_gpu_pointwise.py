"""GPU parity: memory-bound kernels (BatchNorm fwd/bwd, pool, upsample, stem conv, heads, loss) vs torch fp32/fp64
on identical bf16-rounded inputs.  Outputs are bf16 => rel-L2 <= 4e-3; fp32 reductions <= 1e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("n,h,w,c", [(2, 64, 64, 64), (3, 16, 16, 512), (2, 32, 32, 320), (1, 8, 8, 2048),
                                     (2, 128, 128, 32), (2, 16, 16, 8)])
def test_bn_train_fwd_bwd(n, h, w, c):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(5)
    z = nhwc(torch.randn(n, c, h, w, device="cuda", generator=g) * 2 + 0.5)
    dy = nhwc(torch.randn(n, c, h, w, device="cuda", generator=g))
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g) * 0.2
    rm = torch.zeros(c, device="cuda")
    rv = torch.ones(c, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    stats = torch.zeros(2, c, dtype=torch.float64, device="cuda")
    K.channel_stats(z, stats)
    coef = K.bn_finalize(stats, n * h * w, gamma, beta, 1e-5, 0.1, rm, rv, nbt)
    y = K.bn_apply(z, coef, relu=True)
    dz, dgamma, dbeta = K.bn_bwd(dy, z, coef, gamma, relu=True, training=True)

    zr = nchw(z).double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = beta.double().requires_grad_(True)
    rm2 = torch.zeros(c, device="cuda", dtype=torch.float64)
    rv2 = torch.ones(c, device="cuda", dtype=torch.float64)
    yr = torch.relu(F.batch_norm(zr, rm2, rv2, gr, br, True, 0.1, 1e-5))
    yr.backward(nchw(dy).double())
    assert int(nbt) == 1
    assert rel(rm, rm2) < 1e-5 and rel(rv, rv2) < 1e-5
    e = rel(nchw(y), yr)
    # positions whose pre-activation is within rounding of 0 may flip the ReLU mask; they carry ~0 weight in L2
    print(f"bn {n}x{h}x{w}x{c}: y {e:.2e} dz {rel(nchw(dz), zr.grad):.2e} dgamma {rel(dgamma, gr.grad):.2e} "
          f"dbeta {rel(dbeta, br.grad):.2e}")
    assert e < 4e-3
    assert rel(nchw(dz), zr.grad) < 6e-3
    assert rel(dgamma, gr.grad) < 2e-3
    assert rel(dbeta, br.grad) < 2e-3


def test_bn_eval_and_sum_output():
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(6)
    n, h, w, c = 2, 32, 32, 128
    z = nhwc(torch.randn(n, c, h, w, device="cuda", generator=g))
    other = nhwc(torch.randn(n, c, h, w, device="cuda", generator=g))
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g)
    rm = torch.randn(c, device="cuda", generator=g) * 0.1
    rv = torch.rand(c, device="cuda", generator=g) + 0.5
    coef = K.bn_eval_coeffs(gamma, beta, rm, rv, 1e-5)
    y, ys = K.bn_apply(z, coef, relu=True, addend=other, want_sum=True)
    yr = torch.relu(F.batch_norm(nchw(z), rm, rv, gamma, beta, False, 0.1, 1e-5))
    assert rel(nchw(y), yr) < 4e-3
    assert rel(nchw(ys), nchw(y) + nchw(other)) < 4e-3
    dy = nhwc(torch.randn(n, c, h, w, device="cuda", generator=g))
    dz, dgamma, dbeta = K.bn_bwd(dy, z, coef, gamma, relu=True, training=False)
    zr = nchw(z).double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    yr2 = torch.relu(F.batch_norm(zr, rm.double(), rv.double(), gr, beta.double(), False, 0.1, 1e-5))
    yr2.backward(nchw(dy).double())
    assert rel(nchw(dz), zr.grad) < 6e-3
    assert rel(dgamma, gr.grad) < 2e-3


def test_pool_upsample_add_channel_sum():
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(7)
    n, h, w, c = 2, 32, 64, 64
    x = nhwc(torch.randn(n, c, h, w, device="cuda", generator=g))
    y = K.maxpool_fwd(x)
    xr = nchw(x).requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(nchw(y), yr.detach())
    dy = nhwc(torch.randn(n, c, h // 2, w // 2, device="cuda", generator=g))
    yr.backward(nchw(dy))
    dx = K.maxpool_bwd(dy, x)
    assert torch.equal(nchw(dx), xr.grad)
    # ties (post-ReLU zeros): first maximum wins, as in ATen
    xt = torch.relu(nchw(x))
    xt_b = nhwc(xt)
    xr2 = nchw(xt_b).requires_grad_(True)
    F.max_pool2d(xr2, 2, 2).backward(nchw(dy))
    assert torch.equal(nchw(K.maxpool_bwd(dy, xt_b)), xr2.grad)

    u = K.upsample_fwd(x)
    assert torch.equal(nchw(u), F.interpolate(nchw(x), scale_factor=2, mode="nearest"))
    du = nhwc(torch.randn(n, c, 2 * h, 2 * w, device="cuda", generator=g))
    dxu = K.upsample_bwd(du)
    ref = F.avg_pool2d(nchw(du), 2, 2) * 4
    assert rel(nchw(dxu), ref) < 4e-3

    b = nhwc(torch.randn(n, c, h, w, device="cuda", generator=g))
    assert rel(nchw(K.add(x, b)), nchw(x) + nchw(b)) < 4e-3
    assert rel(K.channel_sum(x), nchw(x).sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("k,cin", [(3, 3), (1, 3), (3, 1)])
def test_stem_conv(k, cin):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(8)
    n, h, w, cout = 2, 64, 96, 64
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.3
    b = torch.randn(cout, device="cuda", generator=g)
    x4 = K.image_to_nhwc4(x)
    assert torch.equal(x4[..., :cin], nhwc(x)) and float(x4[..., cin:].abs().sum()) == 0.0
    wk = K.pack_small_weight(wt)
    y = K.conv_smallc_fprop(x4, wk, b, k)
    ref = F.conv2d(nchw(x4[..., :cin]), wt.to(torch.bfloat16).float(), b, padding=k // 2)
    assert rel(nchw(y), ref) < 4e-3
    dy = nhwc(torch.randn(n, cout, h, w, device="cuda", generator=g))
    dw = K.conv_smallc_wgrad(dy, x4, k)
    refw = torch.nn.grad.conv2d_weight(nchw(x4[..., :cin]), (cout, cin, k, k), nchw(dy), padding=k // 2)
    got = dw[:, :, :cin].reshape(cout, k, k, cin).permute(0, 3, 1, 2)
    print(f"stem k{k} cin{cin}: wgrad rel {rel(got, refw):.2e}")
    assert rel(got, refw) < 1e-3


@pytest.mark.parametrize("cin", [3, 1, 2])
def test_stem_conv_as_gemm(cin):
    """3x3 image stem through im2col (K = 32) + the tcgen05 1x1 kernels: the op the models use (ops.stem_conv)."""
    from b200seg import kernels as K
    from b200seg import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    n, h, w, cout = 2, 64, 128, 64
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.3).requires_grad_(True)
    b = torch.randn(cout, device="cuda", generator=g).requires_grad_(True)
    xc = K.stem_im2col3x3(x)
    assert xc.shape == (n, h, w, 32)
    cols = F.unfold(x.to(torch.bfloat16).float(), 3, padding=1).reshape(n, cin, 9, h, w)       # [n, c, tap, h, w]
    want = cols.permute(0, 3, 4, 2, 1).reshape(n, h, w, 9 * cin)                               # column = tap*cin + c
    assert torch.equal(xc[..., :9 * cin].float(), want) and float(xc[..., 9 * cin:].abs().sum()) == 0.0
    y, stats, saved = ops.stem_conv(x, wt, b, True)
    assert saved.shape[-1] == 32
    xb = x.to(torch.bfloat16).float()
    wref = wt.detach().to(torch.bfloat16).float().requires_grad_(True)
    bref = b.detach().clone().requires_grad_(True)
    ref = F.conv2d(xb, wref, bref, padding=1)
    assert rel(nchw(y), ref) < 4e-3
    yf = y.float().reshape(-1, cout)
    assert rel(stats[0], yf.double().sum(0)) < 1e-5 and rel(stats[1], (yf.double() ** 2).sum(0)) < 1e-5
    dy = torch.randn(n, cout, h, w, device="cuda", generator=g)
    y.backward(nhwc(dy))
    ref.backward(nchw(nhwc(dy)))
    print(f"stem-as-gemm cin{cin}: dw rel {rel(wt.grad, wref.grad):.2e} db rel {rel(b.grad, bref.grad):.2e}")
    assert rel(wt.grad, wref.grad) < 1e-3 and rel(b.grad, bref.grad) < 1e-3


@pytest.mark.parametrize("cin,cout", [(64, 1), (32, 1), (64, 3)])
def test_head(cin, cout):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(9)
    n, h, w = 2, 64, 64
    x = nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = torch.randn(cout, cin, device="cuda", generator=g) * 0.2
    b = torch.randn(cout, device="cuda", generator=g)
    y = K.head_fwd(x, wt, b)
    wr = wt.to(torch.bfloat16).float()
    ref = F.conv2d(nchw(x), wr[:, :, None, None], b)
    assert rel(y, ref) < 1e-5
    dy = torch.randn(n, cout, h, w, device="cuda", generator=g)
    dx, dw, db = K.head_bwd(dy, x, wt)
    refdx = F.conv_transpose2d(dy, wr[:, :, None, None])
    refdw = torch.nn.grad.conv2d_weight(nchw(x), (cout, cin, 1, 1), dy)[:, :, 0, 0]
    assert rel(nchw(dx), refdx) < 4e-3
    assert rel(dw, refdw) < 1e-4
    assert rel(db, dy.sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("w_bce,w_dice", [(1.0, 0.0), (0.5, 0.5)])
def test_loss(w_bce, w_dice):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(10)
    z = torch.randn(4, 1, 64, 64, device="cuda", generator=g) * 3
    t = (torch.rand(4, 1, 64, 64, device="cuda", generator=g) > 0.7).float()
    loss, sums = K.loss_fwd(z, t, w_bce, w_dice, 1.0)
    zr = z.double().requires_grad_(True)
    bce = F.binary_cross_entropy_with_logits(zr, t.double())
    p = torch.sigmoid(zr).view(-1)
    tt = t.double().view(-1)
    dice = 1 - (2 * (p * tt).sum() + 1.0) / (p.sum() + tt.sum() + 1.0)
    ref = w_bce * bce + w_dice * dice
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    go = torch.full((), 2.0, device="cuda")
    dz = K.loss_bwd(z, t, sums, go, w_bce, w_dice, 1.0)
    assert rel(dz, 2 * zr.grad) < 1e-4
    pred = z > 0
    assert int(sums[4]) == int((pred & (t > 0.5)).sum()) and int(sums[5]) == int((pred | (t > 0.5)).sum())


def test_seg_counts_and_mask_vs_golden_and_torch():
    """b2_seg_counts / b2_logits_to_mask: integer-exact against torch, metrics against the reference's own outputs."""
    import numpy as np
    from pathlib import Path
    from b200seg import kernels as K
    from b200seg.utils import tester as T
    g = np.load(Path(__file__).parent / "golden" / "metrics.npz")
    z = torch.from_numpy(g["z"]).cuda()
    t = torch.from_numpy(g["t"]).cuda()
    for thr, name in ((0.5, "metrics_thr5"), (0.3, "metrics_thr3")):
        counts = T.segmentation_counts(z, t, thr)
        pred = torch.sigmoid(z) > thr
        tgt = t > thr
        ref = torch.stack([(pred & tgt).flatten(1).sum(1), pred.flatten(1).sum(1), tgt.flatten(1).sum(1)], 1)
        assert torch.equal(counts, ref)
        m = T.metrics_from_counts(counts, z[0].numel())
        got = np.stack([m[k].cpu().numpy() for k in T.METRIC_KEYS], 1)
        assert np.allclose(got, g[name], rtol=2e-6, atol=1e-6)
        one = T.calculate_segmentation_metrics(torch.sigmoid(z[1]), t[1], thr)
        assert np.allclose([one[k] for k in T.METRIC_KEYS], g[name][1], rtol=2e-6, atol=1e-6)
    # ragged size (not a multiple of the vector width), large sample
    gen = torch.Generator(device="cuda").manual_seed(7)
    zz = torch.randn(3, 1, 257, 131, device="cuda", generator=gen)
    tt = (torch.rand(3, 1, 257, 131, device="cuda", generator=gen) > 0.6).float()
    c = K.seg_counts(zz, tt, 0.5)
    p = zz > 0
    ref = torch.stack([(p & (tt > 0.5)).flatten(1).sum(1), p.flatten(1).sum(1), (tt > 0.5).flatten(1).sum(1)], 1)
    assert torch.equal(c, ref)
    mask = K.logits_to_mask(zz, 0.5)
    assert mask.dtype == torch.uint8 and torch.equal(mask, (zz > 0).to(torch.uint8) * 255)
    zr = zz.flatten()[:1001].contiguous()
    assert torch.equal(K.logits_to_mask(zr, 0.5), (zr > 0).to(torch.uint8) * 255)
