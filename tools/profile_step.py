"""One training step inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off` (launch list and
--set full captures).  Usage: python tools/profile_step.py [--model AttentionUNet] [--batch 64] [--side 256]"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "medical-image-segmentation-and-classification_b200"))

import torch  # noqa: E402

from b200seg import kernels as K, ops  # noqa: E402
from b200seg.models import segmentation_models as M  # noqa: E402
from b200seg.utils.synthetic import xray_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="AttentionUNet")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--side", type=int, default=256)
ap.add_argument("--t", type=int, default=None)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
kw = {"t": a.t} if a.t is not None else {}
torch.manual_seed(0)
dev = torch.device("cuda:0")
model = getattr(M, a.model)(**kw).to(dev, memory_format=torch.channels_last).train()
from b200seg.optim import FusedClipAdamW  # noqa: E402

opt = FusedClipAdamW(model.parameters(), lr=1e-6, weight_decay=5e-4, max_norm=1.0)     # as utils/helpers.train()
x, t = xray_batch(a.batch, a.side, a.side, seed=0, device=dev)
params = list(model.parameters())


def step():
    opt.zero_grad(set_to_none=True)
    K.step_begin()
    loss, _ = ops.seg_loss(model(x), t, 1.0, 0.0, 1.0)
    loss.backward()
    opt.step()
    return loss


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss.detach()))
