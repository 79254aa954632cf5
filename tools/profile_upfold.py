"""CUDA-event timings of the folded-UpConv launches at the AttU_Net batch-64 shapes (L2 flushed between launches):
the merged weight gradient (b2_wgrad_args::fold, one launch) against the four phase launches, and the merged fprop /
dgrad launches.

    python tools/profile_upfold.py [--batch 64] [--out gpurun_out/upfold.json]
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

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--out", default="gpurun_out/upfold.json")
a = ap.parse_args()
N = a.batch
dev = torch.device("cuda:0")
# coarse side, cin, cout (Up5 ... Up2 of AttentionUNet.py:73-83)
SHAPES = [(16, 1024, 512), (32, 512, 256), (64, 256, 128), (128, 128, 64)]
PHASES = ((0, 0), (0, 1), (1, 0), (1, 1))
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
for side, cin, cout in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, side, side, cin, device=dev, generator=g).to(torch.bfloat16)
    dz = torch.randn(N, 2 * side, 2 * side, cout, device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
    wf, wd = K.pack_weights_upfold(w, want_dgrad=True)
    z = torch.empty_like(dz)
    dx = torch.empty_like(x)
    dweff = torch.empty((4, cout, 4, cin), dtype=torch.float32, device=dev)
    gf = 2.0 * N * side * side * cout * cin * 16 / 1e9

    def merged():
        assert K.conv_wgrad(dz, x, 2, out=dweff, dy_mul=2, fold=True) is not None

    def phases():
        for ph, (pa, pb) in enumerate(PHASES):
            K.conv_wgrad(dz, x, 2, out=dweff[ph], dy_mul=2, dy_off=(pa, pb), pad=(1 - pa, 1 - pb))

    def fprop():
        K.conv_igemm(x, wf.view(16, cout, cin), cout, 2, out=z, fold=1)

    def dgrad():
        K.conv_igemm(dz, wd.view(16, cin, cout), cin, 2, out=dx, dgrad=True, fold=2)

    r = {"coarse": side, "cin": cin, "cout": cout, "gflop": gf}
    for name, fn in (("wgrad_merged", merged), ("wgrad_phases", phases), ("fprop", fprop), ("dgrad", dgrad)):
        ms = bench(fn)
        r[name + "_ms"] = ms
        r[name + "_tflops"] = gf / ms
    rows.append(r)
    print(f"Up {cin:4d}->{cout:3d} @{side:3d}->{2 * side:3d}: wgrad merged {r['wgrad_merged_ms']:.3f} ms "
          f"({r['wgrad_merged_tflops']:.0f} TF/s) | 4 phases {r['wgrad_phases_ms']:.3f} ms "
          f"({r['wgrad_phases_tflops']:.0f}) | fprop {r['fprop_ms']:.3f} ({r['fprop_tflops']:.0f}) | "
          f"dgrad {r['dgrad_ms']:.3f} ({r['dgrad_tflops']:.0f})")
Path(a.out).parent.mkdir(parents=True, exist_ok=True)
json.dump(rows, open(a.out, "w"), indent=1)
