"""CUDA-event timings of the memory-bound BatchNorm kernels at the AttU_Net batch-64 shapes (L2 flushed between
launches): achieved GB/s against the measured copy bandwidth.

    python tools/profile_pointwise.py [--batch 64] [--out gpurun_out/pointwise.json]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "medical-image-segmentation-and-classification_b200"))

import torch  # noqa: E402

from b200seg import kernels as K  # noqa: E402
from b200seg.kernels import _p, _stream, call  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--out", default="gpurun_out/pointwise.json")
a = ap.parse_args()
N = a.batch
dev = torch.device("cuda:0")
SHAPES = [(256, 64), (128, 128), (64, 256), (32, 512), (16, 1024)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


rows = []
for side, c in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(N, side, side, c, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(N, side, side, c, device=dev, generator=g).to(torch.bfloat16)
    npix = N * side * side
    nbytes = npix * c * 2
    coef = [torch.randn(c, device=dev).abs() + 0.5 for _ in range(4)]      # mean, invstd, scale, shift
    gamma = torch.ones(c, device=dev)
    sums = torch.zeros(2, c, dtype=torch.float64, device=dev)
    dz = torch.empty_like(z)
    y = torch.empty_like(z)
    dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
    zero = K.C.c_void_p(0)

    def reduce_():
        call("b2_bn_bwd_reduce", _p(dy), c, _p(z), c, npix, c, _p(coef[2]), _p(coef[3]), _p(coef[0]), _p(coef[1]), 1,
             _p(sums), _stream())

    def apply_():
        call("b2_bn_bwd_apply", _p(dy), c, _p(z), c, npix, c, _p(coef[2]), _p(coef[3]), _p(coef[0]), _p(coef[1]),
             _p(gamma), 1, 1, _p(sums), _p(dz), c, _p(dg), _p(db), zero, _stream())

    def fwd_():
        call("b2_bn_apply", _p(z), c, npix, c, _p(coef[2]), _p(coef[3]), 1, _p(y), c, zero, 0, zero, 0, _stream())

    tr, ta, tf = bench(reduce_), bench(apply_), bench(fwd_)
    row = {"side": side, "c": c, "reduce_ms": tr, "reduce_GBps": 2 * nbytes / tr / 1e6, "apply_ms": ta,
           "apply_GBps": 3 * nbytes / ta / 1e6, "fwd_ms": tf, "fwd_GBps": 2 * nbytes / tf / 1e6}
    rows.append(row)
    print(f"{side:4d}^2 x {c:4d}: bwd reduce {tr:6.3f} ms {row['reduce_GBps']:6.0f} GB/s | bwd apply {ta:6.3f} ms "
          f"{row['apply_GBps']:6.0f} GB/s | fwd apply {tf:6.3f} ms {row['fwd_GBps']:6.0f} GB/s", flush=True)
Path(a.out).parent.mkdir(parents=True, exist_ok=True)
Path(a.out).write_text(json.dumps(rows, indent=1))
