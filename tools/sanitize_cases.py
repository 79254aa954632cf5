"""Small launches of every protocol-heavy kernel, for compute-sanitizer (tools/sanitize.sh).  Shapes are tiny so that
racecheck / synccheck / memcheck (10-100x slowdown) finish in a minute; each case checks its result, so a tool that
perturbs timing and exposes a protocol race shows up as a wrong answer as well as a report."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "medical-image-segmentation-and-classification_b200"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from b200seg import kernels as K  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def conv_case(n, h, w, cin, cout, k):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    dy = torch.randn(n, cout, h, w, device="cuda", generator=g)
    xb, dyb = nhwc(x), nhwc(dy)
    wf, wd = K.pack_weights(wt)
    stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    y = K.conv_igemm(xb, wf, cout, k, stats=stats)
    dx = K.conv_igemm(dyb, wd, cin, k, dgrad=True)
    dw = K.conv_wgrad(dyb, xb, k)
    torch.cuda.synchronize()
    wr = wt.to(torch.bfloat16).float()
    xr = nchw(xb).requires_grad_(True)
    wr.requires_grad_(True)
    ref = F.conv2d(xr, wr, padding=k // 2)
    ref.backward(nchw(dyb))
    e = (rel(nchw(y), ref), rel(nchw(dx), xr.grad), rel(dw.view(cout, k, k, cin).permute(0, 3, 1, 2), wr.grad))
    assert max(e) < 6e-3, (n, h, w, cin, cout, k, e)
    s_ref = nchw(y).sum((0, 2, 3))
    assert rel(stats[0], s_ref) < 1e-5
    print(f"conv {n}x{h}x{w} {cin}->{cout} k{k}: fprop {e[0]:.1e} dgrad {e[1]:.1e} wgrad {e[2]:.1e}", flush=True)


def pointwise_case():
    g = torch.Generator(device="cuda").manual_seed(2)
    n, h, w, c = 2, 16, 16, 64
    z = torch.randn(n, h, w, c, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(n, h, w, c, device="cuda", generator=g).to(torch.bfloat16)
    stats = torch.zeros(2, c, dtype=torch.float64, device="cuda")
    K.channel_stats(z, stats)
    gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    coef = K.bn_finalize(stats, n * h * w, gamma, beta, 1e-5, 0.0, None, None, None)
    y = K.bn_apply(z, coef, relu=True)
    K.bn_bwd(dy, z, coef, gamma, relu=True, training=True, want_dbias=True)
    K.maxpool_bwd(K.maxpool_fwd(y), y)
    zf = torch.randn(n * h * w, device="cuda", generator=g)
    K.loss_fwd(zf, (zf > 0).float(), 0.5, 0.5, 1.0)
    torch.cuda.synchronize()
    print("pointwise ok", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "conv"):
        conv_case(1, 8, 8, 64, 64, 3)            # Nb > 1 boxes, one work item
        conv_case(2, 16, 16, 64, 128, 3)         # several tiles per CTA, two epilogue groups
        conv_case(1, 128, 128, 64, 64, 3)        # halo mode, row-pair wgrad
        conv_case(2, 16, 16, 128, 256, 1)        # 1x1, N = 256
    if which in ("all", "pointwise"):
        pointwise_case()
    print("sanitize cases ok", flush=True)
