"""GPU parity of the row-streaming Cout = 64 convolution (csrc/conv_c64.cu: N = 192 MMAs over a ring of TMEM row
accumulators) — the fprop / dgrad of conv_block's 64-channel 3x3 layers (AttentionUNet.py:4-13, R2U_Net.py:4-21):
against torch fp32 on identical bf16-rounded operands and against the generic tile kernel (B200SEG_C64=0)."""
import os

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


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _generic(fn):
    """run fn with the row-streaming kernel switched off"""
    from b200seg import kernels as K
    os.environ["B200SEG_C64"] = "0"
    K.reload_switches()
    try:
        return fn()
    finally:
        del os.environ["B200SEG_C64"]
        K.reload_switches()


# n, h, w, c0, c1: every case has >= 8 * 148 row tiles (the kernel's eligibility bound); odd heights and batch sizes
# make the per-CTA ranges start and end inside images (strips with one or two inner ends, single-row strips)
CASES = [
    (4, 256, 256, 64, 0),
    (5, 251, 128, 64, 0),
    (3, 203, 256, 64, 64),       # elided concat: two K sources
    (7, 60, 384, 128, 0),        # two channel blocks from one tensor
    (37, 33, 128, 64, 0),        # short images: every range spans several of them
    (300, 4, 128, 64, 0),        # four-row images: strips of 1-4 rows, top and bottom edge in every strip
]


@pytest.mark.parametrize("n,h,w,c0,c1", CASES)
@pytest.mark.parametrize("relu", [False, True])
def test_c64_fprop_matches_reference_and_generic_kernel(n, h, w, c0, c1, relu):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(h + w + c0)
    cin = c0 + c1
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(64, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    b = torch.randn(64, device="cuda", generator=g)
    xb = nhwc(x)
    x0 = xb[..., :c0].contiguous()
    x1 = xb[..., c0:].contiguous() if c1 else None
    wf, _ = K.pack_weights(wt)

    def run():
        stats = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
        y = K.conv_igemm(x0, wf, 64, 3, x1=x1, bias=b, stats=stats, relu=relu)
        torch.cuda.synchronize()
        return y, stats

    y, stats = run()
    yg, stats_g = _generic(run)
    ref = F.conv2d(nchw(xb), wt.to(torch.bfloat16).float(), b, padding=1)
    if relu:
        ref = ref.relu()
    e, eg = rel(nchw(y), ref), rel(nchw(yg), ref)
    print(f"c64 fprop {n}x{h}x{w} {cin}->64 relu={relu}: rel {e:.3e} (generic kernel {eg:.3e})")
    assert e < 4e-3 and e < 1.05 * eg + 1e-5
    # same products, another summation order: the two kernels differ by bf16 rounding flips only
    d = (y.float() - yg.float()).abs()
    assert float(d.max()) <= 2.0 ** -7 * float(yg.float().abs().max()) + 1e-6
    assert float((d > 0).float().mean()) < 0.05
    # statistics are checksums of what was stored
    yf = y.double().reshape(-1, 64)
    assert torch.allclose(stats[0], yf.sum(0), rtol=1e-6, atol=1e-6 * float(yf.abs().sum(0).max()))
    assert torch.allclose(stats[1], (yf * yf).sum(0), rtol=1e-6)


@pytest.mark.parametrize("n,h,w,c0,c1", CASES[:3])
def test_c64_dgrad_and_output_views(n, h, w, c0, c1):
    """dgrad of a 64 -> 64 layer is the same kernel on the flipped packing; the result may be a channel slice of a wider
    tensor (ldy > 64), the input a channel slice as well."""
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(7 + h)
    cin = 64
    dy_full = torch.randn(n, h, w, 96, device="cuda", generator=g).to(torch.bfloat16)
    dy = dy_full[..., 16:80]                                   # 64 channels at a 16-channel offset, ld = 96
    wt = torch.randn(64, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    _, wd = K.pack_weights(wt)
    out_full = torch.zeros(n, h, w, 128, device="cuda", dtype=torch.bfloat16)
    out = out_full[..., 64:]
    K.conv_igemm(dy, wd, cin, 3, dgrad=True, out=out)
    ref = F.conv_transpose2d(nchw(dy), wt.to(torch.bfloat16).float(), padding=1)
    e = rel(nchw(out), ref)
    print(f"c64 dgrad {n}x{h}x{w}: rel {e:.3e}")
    assert e < 4e-3
    assert float(out_full[..., :64].abs().max()) == 0.0        # nothing outside the slice was written


def test_c64_kernel_is_the_one_that_runs():
    """the eligible layer launches conv_c64_kernel (and the switch really selects the generic kernel)"""
    from torch.profiler import ProfilerActivity, profile

    from b200seg import kernels as K
    x = torch.randn(4, 256, 256, 64, device="cuda").to(torch.bfloat16)
    wf, _ = K.pack_weights(torch.randn(64, 64, 3, 3, device="cuda") * 0.05)

    def names(fn):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        return [e.key for e in prof.key_averages()]

    on = names(lambda: K.conv_igemm(x, wf, 64, 3))
    off = _generic(lambda: names(lambda: K.conv_igemm(x, wf, 64, 3)))
    assert any("conv_c64_kernel" in k for k in on), on
    assert not any("conv_c64_kernel" in k for k in off) and any("conv_igemm_kernel" in k for k in off), off


def test_c64_deterministic_mode_is_bitwise_reproducible():
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(5, 251, 128, 64, device="cuda", generator=g).to(torch.bfloat16)
    wf, _ = K.pack_weights(torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05)
    K.set_deterministic(True)
    try:
        outs = []
        for _ in range(3):
            stats = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
            y = K.conv_igemm(x, wf, 64, 3, stats=stats, relu=True)
            torch.cuda.synchronize()
            outs.append((y.clone(), stats.clone()))
    finally:
        K.set_deterministic(False)
    for y, s in outs[1:]:
        assert torch.equal(y, outs[0][0]) and torch.equal(s, outs[0][1])


# folded UpConv with Cout = 64 (the Up2 level): n, coarse h, coarse w, cin
FOLD_CASES = [
    (10, 128, 128, 128),      # the AttU_Net Up2 shape, smaller batch
    (5, 251, 128, 64),        # odd height, one channel block
    (3, 150, 256, 128),       # two column segments per row
    (300, 4, 128, 64),        # four-row images: strips with the top and bottom edge
    (37, 33, 128, 128),
]


@pytest.mark.parametrize("n,h,w,cin", FOLD_CASES)
def test_c64_folded_upconv_fprop_matches_reference_and_generic_kernel(n, h, w, cin):
    """The merged folded-UpConv fprop (Upsample x2 -> conv3x3, AttentionUNet.py:15-27) through the row-streaming kernel
    (column phase per CTA, four fine rows per coarse input row, N = 256) against Upsample + conv2d in fp32 and against
    the generic kernel's merged launch (B200SEG_C64_FOLD=0)."""
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(h + w + cin)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(64, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    b = torch.randn(64, device="cuda", generator=g)
    xb = nhwc(x)
    wf, _ = K.pack_weights_upfold(wt, want_dgrad=False)

    def run():
        stats = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
        z = torch.full((n, 2 * h, 2 * w, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
        K.conv_igemm(xb, wf.view(16, 64, cin), 64, 2, bias=b, stats=stats, out=z, fold=1)
        torch.cuda.synchronize()
        return z, stats

    z, stats = run()
    os.environ["B200SEG_C64_FOLD"] = "0"
    K.reload_switches()
    try:
        zg, stats_g = run()
    finally:
        del os.environ["B200SEG_C64_FOLD"]
        K.reload_switches()
    assert torch.isfinite(z.float()).all()
    # both kernels accumulate the same bf16 products in fp32: they differ by summation order only
    assert rel(nchw(z), nchw(zg)) < 2e-3
    # fp32 reference on the folded (fp32-summed, bf16-rounded) weights: phase (a, b) = 2x2 conv of the coarse input
    ref = torch.empty(n, 64, 2 * h, 2 * w, device="cuda")
    xf = nchw(xb)
    for ph, (a, bb) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        wp = wf[ph].float().view(2, 2, 64, cin).permute(2, 3, 0, 1)          # [cout, cin, ty, tx]
        xp = F.pad(xf, (1 - bb, bb, 1 - a, a))
        ref[:, :, a::2, bb::2] = F.conv2d(xp, wp, b)
    assert rel(nchw(z), ref) < 4e-3, rel(nchw(z), ref)
    # the literal op on the unfolded weights (bf16 rounding of the summed taps differs): looser
    lit = F.conv2d(F.interpolate(xf, scale_factor=2, mode="nearest"), wt.to(torch.bfloat16).float(), b, padding=1)
    assert rel(nchw(z), lit) < 1e-2
    # statistics of the rounded outputs
    zf = z.double().reshape(-1, 64)
    assert rel(stats[0], zf.sum(0)) < 1e-6 and rel(stats[1], (zf * zf).sum(0)) < 1e-6
    assert rel(stats, stats_g) < 1e-4
