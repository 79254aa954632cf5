"""Summarise one ncu pass over a training step (tools/profile_step.py) taken with
    --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
dram__bytes_read.sum,dram__bytes_write.sum --csv
into a per-kernel table: launches, time, share of the step, DRAM bytes, achieved GB/s (vs the measured copy peak) and
the time-weighted tensor-pipe utilisation.  Usage: summarize_step_profile.py launches.csv out.json [peak_GBps]"""
import csv
import json
import re
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6468.0
lines = [l for l in open(path) if not l.startswith("==")]
per = defaultdict(dict)          # launch id -> {name, metric: value}
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
    d = per[r["ID"]]
    d["name"] = re.sub(r"[<(].*", "", r["Kernel Name"]).replace("void ", "").strip()
    d[m] = v
agg = defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_GB": 0.0, "tensor_ms": 0.0})
for d in per.values():
    a = agg[d["name"]]
    ms = d.get("gpu__time_duration.sum", 0.0)
    a["launches"] += 1
    a["ms"] += ms
    a["dram_GB"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["tensor_ms"] += ms * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0
total = sum(a["ms"] for a in agg.values())
res = {}
print(f"total {total:.2f} ms over {sum(a['launches'] for a in agg.values())} launches (serialised, cold cache)")
print(f"{'ms':>9} {'share':>6} {'n':>5} {'DRAM GB':>8} {'GB/s':>7} {'of peak':>7} {'tensor%':>7}  kernel")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbps = a["dram_GB"] / (a["ms"] * 1e-3) if a["ms"] > 0 else 0.0
    tens = 100.0 * a["tensor_ms"] / a["ms"] if a["ms"] > 0 else 0.0
    res[name] = {"launches": a["launches"], "ms": round(a["ms"], 4), "share": round(a["ms"] / total, 4),
                 "dram_GB": round(a["dram_GB"], 4), "GBps": round(gbps, 1), "hbm_frac": round(gbps / peak, 3),
                 "tensor_pipe_pct": round(tens, 1)}
    if a["ms"] / total >= 0.002:
        print(f"{a['ms']:9.3f} {100 * a['ms'] / total:5.1f}% {a['launches']:5d} {a['dram_GB']:8.3f} {gbps:7.0f} "
              f"{100 * gbps / peak:6.1f}% {tens:6.1f}%  {name}")
conv = [a for n, a in agg.items() if "conv_igemm" in n or "conv_wgrad_kernel" in n]
if conv:
    ms = sum(a["ms"] for a in conv)
    res["_tensor_core_kernels"] = {"ms": round(ms, 3),
                                   "tensor_pipe_pct": round(100.0 * sum(a["tensor_ms"] for a in conv) / ms, 1)}
    print(f"tensor-core conv kernels: {ms:.2f} ms, time-weighted tensor-pipe utilisation "
          f"{res['_tensor_core_kernels']['tensor_pipe_pct']:.1f} %")
json.dump(res, open(out, "w"), indent=1)
