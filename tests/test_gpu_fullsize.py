"""GPU parity at BASELINE.json's FULL sizes (batch 64, 256x256) through size-independent properties, where the fp64
oracle is too slow to run:
  * adjoint identities of the three tcgen05 convolution kernels,  <dy, conv(x, w)> = <dgrad(dy, w), x> = <wgrad(dy, x), w>,
    on every tile-geometry class of the AttU_Net step (halo rows, multi-row tiles, N = 32/64/128/256, row-pair wgrad,
    1x1 gate GEMMs, folded UpConv phases);
  * epilogue statistics = checksums of the stored output;
  * a full AttentionUNet batch-64 training step: finite loss / gradients, and invariance of loss and gradients under a
    permutation of the batch (BatchNorm statistics and every reduction are order-independent up to rounding)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dot(a, b):
    return float((a.double() * b.double()).sum())


def rel(a, b):
    return abs(a - b) / max(abs(a), abs(b), 1e-30)


FULL = [
    # n, side, cin, cout, k
    (64, 256, 64, 64, 3),       # halo mode, N = 64; row-pair wgrad
    (64, 256, 128, 64, 3),      # two channel blocks
    (64, 128, 128, 128, 3),     # halo mode, N = 128, second epilogue group
    (64, 64, 256, 256, 3),      # 2-row tiles, N = 256
    (64, 32, 512, 512, 3),      # 4-row tiles
    (64, 16, 1024, 1024, 3),    # 8-row tiles
    (64, 256, 64, 32, 1),       # gate 1x1 GEMM, N = 32 (TMA store with the 64B swizzle)
    (64, 128, 128, 64, 1),
]


@pytest.mark.parametrize("n,side,cin,cout,k", FULL)
def test_conv_adjoint_identities_full_size(n, side, cin, cout, k):
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(side + cin)
    x = torch.randn(n, side, side, cin, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(n, side, side, cout, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    wb = w.to(torch.bfloat16).float()                    # what the kernels multiply with
    wf, wd = K.pack_weights(w)
    stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    y = K.conv_igemm(x, wf, cout, k, stats=stats)
    dx = K.conv_igemm(dy, wd, cin, k, dgrad=True)
    dw = K.conv_wgrad(dy, x, k).reshape(cout, k, k, cin).permute(0, 3, 1, 2)
    a, b, c = dot(dy, y), dot(dx, x), dot(dw, wb)
    print(f"{cin}->{cout} @{side} k{k}: <dy,y>={a:.6e} <dx,x>={b:.6e} <dw,w>={c:.6e}")
    # y and dx are rounded to bf16 on store: independent relative errors of <= 2^-9 per element, so an inner product
    # over M elements is off by about 2^-9 * |u| |v| / sqrt(M); dw is fp32.  Gate: 6 sigma of that.
    tol_y = 6 * 2.0 ** -9 * (dot(dy, dy) * dot(y, y)) ** 0.5 / y.numel() ** 0.5
    tol_x = 6 * 2.0 ** -9 * (dot(dx, dx) * dot(x, x)) ** 0.5 / x.numel() ** 0.5
    assert abs(a - c) < tol_y, (a, c, tol_y)
    assert abs(b - c) < tol_x, (b, c, tol_x)
    # epilogue statistics are checksums of what was stored
    yf = y.double().reshape(-1, cout)
    assert rel(float(stats[0].sum()), float(yf.sum())) < 1e-6 or abs(float(yf.sum())) < 1e-3 * float(yf.abs().sum())
    assert rel(float(stats[1].sum()), float((yf * yf).sum())) < 1e-7       # fp32 per tile, fp64 across tiles


def test_upconv_fold_adjoint_full_size():
    """Folded UpConv phase (2x2 taps on the coarse grid, pixel-shuffle store / sub-lattice reads) at 128 -> 64 @ 256^2."""
    from b200seg import kernels as K
    from b200seg import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    n, cin, cout, h = 64, 128, 64, 128
    x = torch.randn(n, h, h, cin, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5).requires_grad_(True)
    gamma = torch.ones(cout, device="cuda", requires_grad=True)
    beta = torch.zeros(cout, device="cuda", requires_grad=True)
    rm, rv = torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")
    y, z, coef, stats = ops.upconv_bn_act(x, w, None, gamma, beta, rm, rv, False, 1e-5, False)   # eval BN == identity
    dy = torch.randn(n, 2 * h, 2 * h, cout, device="cuda", generator=g).to(torch.bfloat16)
    y.backward(dy)
    a, b, c = dot(dy, z), dot(x.grad, x.detach()), dot(w.grad, w.detach())
    print(f"upconv fold: <dy,z>={a:.6e} <dx,x>={b:.6e} <dw,w>={c:.6e}")
    scale = (dot(dy, dy) * dot(z, z)) ** 0.5
    # the phase weights are rounded to bf16 AFTER the fp32 tap sums, so <dw, w> (fp32 weights) carries that rounding
    assert abs(a - b) < 1e-4 * scale
    assert abs(a - c) < 3e-3 * scale


def test_full_batch_training_step_is_permutation_invariant():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle.synthetic import xray_batch
    from b200seg import ops
    from b200seg.models.segmentation_models import AttentionUNet
    torch.manual_seed(0)
    model = AttentionUNet().cuda().to(memory_format=torch.channels_last).eval()    # eval BN: reproducible (see the
    x, y = xray_batch(64, 256, 256, seed=0)                                        # side-stream test for the reason)
    x, y = x.cuda(), y.cuda()
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(1)).cuda()

    def run(xx, yy):
        model.zero_grad(set_to_none=True)
        loss, sums = ops.seg_loss(model(xx), yy, 1.0, 0.0, 1.0)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), [p.grad.clone() for p in model.parameters()], sums.clone()

    l0, g0, s0 = run(x, y)
    l1, g1, s1 = run(x[perm].contiguous(), y[perm].contiguous())
    assert torch.isfinite(torch.tensor(l0)) and all(torch.isfinite(t).all() for t in g0)
    assert abs(l0 - l1) < 1e-6 * abs(l0)
    assert torch.equal(s0[3:], s1[3:])                   # target count and IoU counts are integers
    num = sum(float((a.double() - b.double()).pow(2).sum()) for a, b in zip(g0, g1))
    den = sum(float(b.double().pow(2).sum()) for b in g0)
    print(f"full-size step: loss {l0:.6f}, permuted-batch gradient difference {(num / den) ** 0.5:.2e}")
    assert (num / den) ** 0.5 < 1e-4


def test_batch128_inference_matches_chunks():
    """BASELINE.json configs[4] sweeps the inference batch to 512: a 128-image 256^2 batch has 65536 m-tiles per layer at
    the top level — the tile decode's magic-multiplier divisions must stay exact there (a too-tight host check once
    rejected these launches).  The logits of the big batch must equal those of the same images in chunks of 32."""
    from b200seg.models.segmentation_models import R2U_Net
    torch.manual_seed(0)
    model = R2U_Net(t=1).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(128, 3, 256, 256, device="cuda", generator=g)
    with torch.no_grad():
        big = model(x)
        parts = torch.cat([model(x[i:i + 32]) for i in range(0, 128, 32)])
    torch.cuda.synchronize()
    assert torch.isfinite(big).all()
    err = float((big.double() - parts.double()).norm() / parts.double().norm())
    assert err < 1e-3, err


def test_conv_igemm_accepts_128_image_batches():
    """the generic tile kernel at the largest tile counts of the configs (N = 128 images): statistics = checksum of the
    stored output"""
    from b200seg import kernels as K
    g = torch.Generator(device="cuda").manual_seed(9)
    n, side, cin, cout = 128, 256, 64, 128
    x = torch.randn(n, side, side, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05
    wf, _ = K.pack_weights(w)
    stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    y = K.conv_igemm(x, wf, cout, 3, stats=stats)
    torch.cuda.synchronize()
    yf = y[::16].double().reshape(-1, cout)          # every 16th image against torch on the same bf16 operands
    ref = torch.nn.functional.conv2d(x[::16].float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, cout).double()
    assert float((yf - ref).norm() / ref.norm()) < 4e-3
    s = torch.zeros(cout, dtype=torch.float64, device="cuda")
    for i in range(0, n, 16):
        s += y[i:i + 16].double().reshape(-1, cout).sum(0)
    assert float((stats[0] - s).norm() / s.norm()) < 1e-6
