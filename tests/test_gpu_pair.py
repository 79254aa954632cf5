"""GPU: CTA-pair mode of the implicit-GEMM convolution (tcgen05.mma.cta_group::2, M = 256 across two SMs, each SM staging
half of the weight rows — B200SEG_PAIR=1: the N = 256 tiles, 2: every eligible layer).  The pair MMA accumulates in
the same K order as the single-CTA MMA, so outputs must be BIT-identical to the default path; the BatchNorm statistics
(summed in a different order) must agree to fp64 rounding."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


class _pair:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        from b200seg import kernels as K
        self.old = os.environ.get("B200SEG_PAIR")
        os.environ["B200SEG_PAIR"] = self.mode
        K.reload_switches()              # the library caches its environment switches

    def __exit__(self, *a):
        from b200seg import kernels as K
        if self.old is None:
            del os.environ["B200SEG_PAIR"]
        else:
            os.environ["B200SEG_PAIR"] = self.old
        K.reload_switches()


SHAPES = [
    # n, h, w, cin, cout, k
    (2, 64, 64, 128, 256, 3),       # N = 256, 64 tiles
    (3, 32, 32, 256, 512, 3),       # two n-tiles
    (2, 16, 16, 512, 1024, 3),      # 4 tiles x 4 n-tiles: fewer pairs than SMs
    (3, 16, 8, 64, 256, 3),         # odd number of m-tiles: the last pair recomputes a tile
    (1, 8, 8, 128, 256, 3),         # a single (padded) tile: pair mode must step aside (m_tiles < 2)
    (2, 256, 256, 64, 64, 3),       # halo mode, N = 64
    (2, 128, 128, 64, 128, 3),      # halo mode, N = 128
    (2, 64, 64, 256, 128, 1),       # 1x1
    (2, 256, 256, 64, 32, 1),       # N = 32 (64 B-swizzled store)
]


@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("n,h,w,cin,cout,k", SHAPES)
def test_pair_mode_bit_identical(n, h, w, cin, cout, k, mode):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1)
    x = nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g))
    dy = nhwc(torch.randn(n, cout, h, w, device="cuda", generator=g))
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    add = nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wf, wd = K.pack_weights(wt)

    def run():
        stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
        y = K.conv_igemm(x, wf, cout, k, bias=b, stats=stats, relu=False)
        dx = K.conv_igemm(dy, wd, cin, k, dgrad=True, addend=add)
        torch.cuda.synchronize()
        return y, stats, dx

    with _pair("0"):                            # single-CTA baseline (pair mode is the default)
        y0, s0, dx0 = run()
    with _pair(mode):
        for _ in range(3):                      # repeated: a protocol race rarely shows on the first launch
            y1, s1, dx1 = run()
            assert torch.equal(y0, y1), "fprop differs"
            assert torch.equal(dx0, dx1), "dgrad (+addend) differs"
            assert torch.allclose(s0, s1, rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize("mode", ["1", "2"])
def test_pair_mode_two_sources_stride_and_placement(mode):
    """virtual concat (two K sources), stride-2 sampling, pixel-shuffle placement (ConvTranspose / folded UpConv)"""
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(2)
    n, h, w = 2, 32, 32
    x0 = nhwc(torch.randn(n, 128, h, w, device="cuda", generator=g))
    x1 = nhwc(torch.randn(n, 128, h, w, device="cuda", generator=g))
    wt = torch.randn(256, 256, 3, 3, device="cuda", generator=g) / 48.0
    wf, _ = K.pack_weights(wt)
    w2 = torch.randn(256, 128, 2, 2, device="cuda", generator=g) / 23.0
    wf2, _ = K.pack_weights(w2)
    w1 = torch.randn(256, 128, 1, 1, device="cuda", generator=g) / 11.0
    wf1, _ = K.pack_weights(w1)

    def run():
        ya = K.conv_igemm(x0, wf, 256, 3, x1=x1, relu=True)
        yb = K.conv_igemm(x0, wf1, 256, 1, stride=2)
        yc = torch.zeros(n, 2 * h, 2 * w, 256, dtype=torch.bfloat16, device="cuda")
        for ph, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            K.conv_igemm(x0, wf2[ph:ph + 1], 256, 1, out=yc, out_mul=2, out_off=(a, b))
        yd = K.conv_igemm(x0, wf2, 256, 2, pad=(1, 0))
        torch.cuda.synchronize()
        return ya, yb, yc, yd

    with _pair("0"):
        ref = run()
    with _pair(mode):
        got = run()
    for r, o in zip(ref, got):
        assert torch.equal(r, o)


def test_pair_mode_whole_model_step():
    """AttentionUNet eval-mode forward + backward with every eligible layer in pair mode == the default path, bitwise
    (eval mode: no batch statistics, so nothing depends on summation order except the bias / BN-affine / head / psi
    gradient sums (atomics), which are compared to fp32 rounding)"""
    from b200seg import ops
    from b200seg.models.segmentation_models import AttentionUNet
    from b200seg.utils.synthetic import xray_batch
    torch.manual_seed(0)
    m = AttentionUNet().cuda().eval()
    x, t = xray_batch(4, 128, 128, seed=3, device="cuda")

    def run():
        m.zero_grad(set_to_none=True)
        logits = m(x)
        loss, _ = ops.seg_loss(logits, t, 1.0, 0.0, 1.0)
        loss.backward()
        torch.cuda.synchronize()
        return logits.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}

    with _pair("0"):
        l0, g0 = run()
    with _pair("2"):
        l1, g1 = run()
    assert torch.equal(l0, l1)
    for k in g0:
        if g0[k].dim() == 4 and g0[k].shape[0] >= 32:
            assert torch.equal(g0[k], g1[k]), k              # tensor-core weight gradients: deterministic on identical inputs
        else:
            assert torch.allclose(g0[k], g1[k], rtol=1e-4, atol=1e-7), k


WG_SHAPES = [
    # n, h, w, cin, cout: Cout a multiple of 256, an even number of 64-channel X blocks
    (2, 64, 64, 128, 256),
    (3, 32, 32, 256, 512),
    (2, 16, 16, 512, 1024),
    (64, 16, 16, 256, 256),       # many pixel chunks: several splits per pair
    (1, 20, 28, 128, 256),        # clipped chunks
    (2, 64, 64, 192, 256),        # an odd number of X blocks: the pair mode must step aside
]


@pytest.mark.parametrize("n,h,w,cin,cout", WG_SHAPES)
def test_wgrad_pair_mode_matches_single_cta(n, h, w, cin, cout):
    """CTA-pair weight gradient (conv_wgrad_kernel<true>: tcgen05.mma.cta_group::2, M = 256 = two Cout tiles, each CTA
    holding one of the two X halo boxes) against the single-CTA kernel (B200SEG_WG_PAIR=0) and the fp32 reference.
    Both accumulate the pixel chunks of a split in the same order, so the results are bit-identical."""
    import torch.nn.functional as F
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    x = nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g))
    dy = nhwc(torch.randn(n, cout, h, w, device="cuda", generator=g))
    pair = K.conv_wgrad(dy, x, 3)
    os.environ["B200SEG_WG_PAIR"] = "0"
    K.reload_switches()
    try:
        single = K.conv_wgrad(dy, x, 3)
    finally:
        del os.environ["B200SEG_WG_PAIR"]
        K.reload_switches()
    torch.cuda.synchronize()
    assert torch.equal(pair, single)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3), dy.float().permute(0, 3, 1, 2),
                                      padding=1)
    got = pair.reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    err = float((got.double() - ref.double()).norm() / ref.double().norm())
    assert err < 1e-3, err
