"""Sum DRAM traffic per kernel from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` of one training
step (tools/profile_step.py) and write profiles-style JSON: {kernel: {launches, read_GB, write_GB}}."""
import csv
import json
import re
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
agg = defaultdict(lambda: {"launches": set(), "read_GB": 0.0, "write_GB": 0.0})
for r in csv.DictReader(lines):
    name = re.sub(r"[<(].*", "", r["Kernel Name"]).replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(r["Metric Unit"], 1e-9)
    a = agg[name]
    a["launches"].add(r["ID"])
    if r["Metric Name"] == "dram__bytes_read.sum":
        a["read_GB"] += v * scale
    elif r["Metric Name"] == "dram__bytes_write.sum":
        a["write_GB"] += v * scale
res = {k: {"launches": len(v["launches"]), "read_GB": round(v["read_GB"], 4), "write_GB": round(v["write_GB"], 4)}
       for k, v in sorted(agg.items(), key=lambda kv: -(kv[1]["read_GB"] + kv[1]["write_GB"]))}
json.dump(res, open(out, "w"), indent=1)
for k, v in list(res.items())[:12]:
    print(f"{v['read_GB'] + v['write_GB']:8.3f} GB  r {v['read_GB']:7.3f}  w {v['write_GB']:7.3f}  {v['launches']:4d}  {k}")
