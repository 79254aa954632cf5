"""Per-launch list of one ncu pass over a training step (tools/profile_step.py): duration, DRAM bytes, achieved GB/s,
tensor-pipe activity.  Usage: list_launches.py step_metrics_ncu.csv [substring ...]"""
import csv
import re
import sys
from collections import defaultdict

path, pats = sys.argv[1], sys.argv[2:]
lines = [l for l in open(path) if not l.startswith("==")]
per = defaultdict(dict)
for r in csv.DictReader(lines):
    d = per[int(r["ID"])]
    d["name"] = re.sub(r"[<(].*", "", r["Kernel Name"]).replace("void ", "").replace("b2::", "").strip()
    d["grid"] = r["Grid Size"]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
    if "bytes" in r["Metric Name"]:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r["Metric Name"]] = v
for i in sorted(per):
    d = per[i]
    if pats and not any(p in d["name"] for p in pats):
        continue
    t = d["gpu__time_duration.sum"]
    b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    print(f"{i:4d} {d['name'][:30]:30s} {d['grid']:>14s} {t:8.1f} us {b / 1e6:8.1f} MB {b / t / 1e3:6.0f} GB/s  tensor "
          f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  "
          f"L2->SM {d.get('l1tex__m_xbar2l1tex_read_bytes.sum', 0) / 1e6:8.1f} MB")
