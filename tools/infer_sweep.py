"""BASELINE.json configs[4]: R2U_Net inference throughput sweep (eval mode, no_grad) on one B200.
    python tools/infer_sweep.py [--t 2] [--batches 1,8,64,512] [--sides 256,512]
Prints one JSON line per (side, batch): images/s from CUDA events over `reps` forward passes after warm-up."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "medical-image-segmentation-and-classification_b200"))

import torch  # noqa: E402

from b200seg.models.segmentation_models import R2U_Net  # noqa: E402
from b200seg.utils.synthetic import xray_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--t", type=int, default=2)
ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512")
ap.add_argument("--sides", default="256,512")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--graph", type=int, default=1)
a = ap.parse_args()
torch.manual_seed(0)
dev = torch.device("cuda:0")
model = R2U_Net(t=a.t).to(dev).eval()
for side in map(int, a.sides.split(",")):
    for b in map(int, a.batches.split(",")):
        if side == 512 and b > 128:
            continue
        x, _ = xray_batch(min(b, 8), side, side, seed=0, device=dev)
        x = x.repeat((b + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:b].contiguous()
        with torch.no_grad():
            wstream = torch.cuda.Stream()
            wstream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(wstream):
                for _ in range(2):
                    model(x)
            torch.cuda.current_stream().wait_stream(wstream)
            torch.cuda.synchronize()
            graph = None
            if a.graph:                      # static shapes: replay the whole forward as one CUDA graph
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    y = model(x)
                graph.replay()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                if graph is not None:
                    graph.replay()
                else:
                    y = model(x)
            e1.record()
            torch.cuda.synchronize()
            del graph
        ms = e0.elapsed_time(e1) / a.reps
        print(json.dumps({"model": f"R2U_Net(t={a.t})", "mode": "inference", "cuda_graph": bool(a.graph), "side": side, "batch": b,
                          "ms_per_batch": round(ms, 3), "images_per_s": round(b / ms * 1e3, 1)}), flush=True)
