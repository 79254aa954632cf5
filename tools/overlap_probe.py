"""Does a memory-bound kernel overlap with a persistent tensor-core kernel on another stream?  Times wgrad alone, the
memory-bound op alone and both launched together on two streams (CUDA events around the pair).
    python tools/overlap_probe.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "medical-image-segmentation-and-classification_b200"))

import torch  # noqa: E402

from b200seg import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
N = 64


def act(c, s):
    return torch.randn(N, s, s, c, device=dev, generator=g).to(torch.bfloat16)


s_main, s_side = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fa, fb, reps=5):
    """fa on s_main, fb on s_side (either may be None), started together; returns ms until both are done."""
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_main)
        s_side.wait_event(e0)
        if fb is not None:
            with torch.cuda.stream(s_side):
                fb()
        if fa is not None:
            with torch.cuda.stream(s_main):
                fa()
        s_main.wait_stream(s_side)
        e1.record(s_main)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


cases = []
# wgrad layers (dy, x)
wg_layers = {"wgrad 128->128@128": (act(128, 128), act(128, 128)), "wgrad 512->512@32": (act(512, 32), act(512, 32)),
             "wgrad 64->64@256 (row pair)": (act(64, 256), act(64, 256))}
# memory-bound ops on 64ch@256^2
z = act(64, 256)
dy = act(64, 256)
coef = torch.randn(4, 64, device=dev).abs() + 0.1
gamma = torch.ones(64, device=dev)
mem_ops = {
    "bn_bwd (reduce+apply) 64ch@256": lambda: K.bn_bwd(dy, z, coef, gamma, relu=True, training=True, want_dbias=True),
    "bn_apply 64ch@256": lambda: K.bn_apply(z, coef, relu=True),
    "maxpool2x2_fwd 64ch@256": lambda: K.maxpool2x2_fwd(z) if hasattr(K, "maxpool2x2_fwd") else None,
}
wf, wd = K.pack_weights(torch.randn(128, 128, 3, 3, device=dev) * 0.05)
xa = act(128, 128)
mem_ops["dgrad 128->128@128 (tensor kernel)"] = lambda: K.conv_igemm(xa, wd, 128, 3, dgrad=True)
for wn, (wdy, wx) in wg_layers.items():
    fw = lambda: K.conv_wgrad(wdy, wx, 3)
    tw = timed(None, fw)
    for mn, fm in mem_ops.items():
        try:
            fm()
        except Exception as e:      # noqa: BLE001
            print("skip", mn, e)
            continue
        tm = timed(fm, None)
        tb = timed(fm, fw)
        print(f"{wn:32s} {tw:6.3f} ms | {mn:36s} {tm:6.3f} ms | together {tb:6.3f} ms  (sum {tw + tm:6.3f}, "
              f"hidden {100 * (tw + tm - tb) / min(tw, tm):5.1f} % of the shorter)", flush=True)
