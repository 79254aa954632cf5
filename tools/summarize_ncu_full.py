"""Key metrics of every kernel in an `ncu --set full` report exported with `ncu -i rep --page raw --csv`.
Usage: summarize_ncu_full.py raw.csv [label ...]  (labels are attached to the launches in order)"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__registers_per_thread", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
labels = sys.argv[2:]
ki = hdr.index("Kernel Name")
for i, r in enumerate(body):
    name = r[ki].split("(")[0]
    print(f"{labels[i] if i < len(labels) else ''} : {name}")
    for k in KEYS:
        if k in hdr:
            j = hdr.index(k)
            print(f"   {k} [{units[j]}] = {r[j]}")
    print()
