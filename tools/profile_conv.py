"""Per-layer timing table of the tensor-core convolution kernels at the AttU_Net batch-64 shapes (CUDA events), and a
cudaProfilerStart/Stop range around a few representative launches for `ncu --set full --profile-from-start off`.

    python tools/profile_conv.py [--batch 64] [--ncu]   (--ncu: only the representative launches, inside the range)
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
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--out", default="gpurun_out/conv_layers.json")
ap.add_argument("--layers", default="", help="override the layer list: cin,cout,side,k[;cin,cout,side,k...]")
a = ap.parse_args()
N = a.batch
# (cin, cout, side, ksize) of every distinct AttU_Net tensor-core conv
LAYERS = [(64, 64, 256, 3), (128, 64, 256, 3), (64, 128, 128, 3), (128, 128, 128, 3), (256, 128, 128, 3),
          (128, 256, 64, 3), (256, 256, 64, 3), (512, 256, 64, 3), (256, 512, 32, 3), (512, 512, 32, 3),
          (1024, 512, 32, 3), (512, 1024, 16, 3), (1024, 1024, 16, 3),
          (512, 256, 32, 1), (256, 128, 64, 1), (128, 64, 128, 1), (64, 32, 256, 1)]
REPR = [(128, 128, 128, 3), (64, 64, 256, 3), (512, 512, 32, 3)]
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)


def bench(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()                       # evict L2 between timed launches
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


rows = []
layers = REPR if a.ncu else LAYERS
if a.layers:
    layers = [tuple(int(v) for v in item.split(",")) for item in a.layers.split(";")]
for cin, cout, s, k in layers:
    x = torch.randn(N, s, s, cin, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(N, s, s, cout, device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device=dev, generator=g) * 0.05
    wf, wd = K.pack_weights(w)
    stats = torch.zeros(2, cout, dtype=torch.float64, device=dev)
    flops = 2.0 * N * s * s * cin * cout * k * k
    if a.ncu:
        K.conv_igemm(x, wf, cout, k, stats=stats)
        K.conv_wgrad(dy, x, k)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        K.conv_igemm(x, wf, cout, k, stats=stats)
        K.conv_igemm(dy, wd, cin, k, dgrad=True)
        K.conv_wgrad(dy, x, k)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    tf = bench(lambda: K.conv_igemm(x, wf, cout, k, stats=stats))
    td = bench(lambda: K.conv_igemm(dy, wd, cin, k, dgrad=True))
    tw = bench(lambda: K.conv_wgrad(dy, x, k))
    rows.append({"cin": cin, "cout": cout, "side": s, "k": k, "gflop": flops / 1e9,
                 "fprop_ms": tf, "fprop_tflops": flops / tf / 1e9,
                 "dgrad_ms": td, "dgrad_tflops": flops / td / 1e9,
                 "wgrad_ms": tw, "wgrad_tflops": flops / tw / 1e9})
    print(f"{cin:5d}->{cout:5d} @{s:3d} k{k}: fprop {tf:7.3f} ms {flops / tf / 1e9:7.1f} TF | dgrad {td:7.3f} ms "
          f"{flops / td / 1e9:7.1f} TF | wgrad {tw:7.3f} ms {flops / tw / 1e9:7.1f} TF", flush=True)
if rows:
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(rows, indent=1))
