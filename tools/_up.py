import sys, torch
sys.path.insert(0, "medical-image-segmentation-and-classification_b200")
from b200seg import kernels as K
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bench(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.zero_(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for cin, cout, h in ((128, 64, 128), (256, 128, 64), (512, 256, 32), (1024, 512, 16)):
    x = torch.randn(64, h, h, cin, device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
    wf, wd = K.pack_weights_upfold(w, want_dgrad=True)
    dz = torch.randn(64, 2 * h, 2 * h, cout, device=dev, generator=g).to(torch.bfloat16)
    tf = bench(lambda: K.conv_igemm(x, wf.view(16, cout, cin), cout, 2, fold=1))
    td = bench(lambda: K.conv_igemm(dz, wd.view(16, cin, cout), cin, 2, fold=2))
    fl = 2.0 * 64 * (2 * h) ** 2 * cin * cout * 4
    print(f"upconv {cin}->{cout} @{h}->{2*h}: fprop {tf:.3f} ms {fl/tf/1e9:7.1f} TF | dgrad {td:.3f} ms {fl/td/1e9:7.1f} TF", flush=True)
