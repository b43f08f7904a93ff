#!/usr/bin/env python
"""Builds profiles/traffic.json (DRAM bytes per launch of the dominant kernel, from the ncu CSVs that
scripts/gpu_final.sh writes into gpurun_out/traffic_<workload>.csv)."""
import csv, json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes = {"sort_u32": 1 << 30, "scan_i32": 1 << 28, "reduce_i32": 1 << 28}
out = {}
for w, n in sizes.items():
    p = os.path.join(root, "gpurun_out", f"traffic_{w}.csv")
    if not os.path.exists(p):
        continue
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    if len(rows) < 2:
        continue
    hdr = rows[0]
    mi, vi, ui, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd = wr = dur = None
    kname = None
    for r in rows[1:]:
        val = float(r[vi].replace(",", ""))
        if r[mi] == "dram__bytes_read.sum": rd = val * scale.get(r[ui], 1)
        if r[mi] == "dram__bytes_write.sum": wr = val * scale.get(r[ui], 1)
        if r[mi] == "gpu__time_duration.sum": dur = val; kname = r[ki]
    if rd is not None and wr is not None:
        out[w] = {"n_per_gpu": n, "kernel": kname[:80] if kname else None, "dram_bytes_read": rd, "dram_bytes_write": wr,
                  "dram_bytes_per_launch": rd + wr, "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (gpurun_out/traffic_{w}.csv)"}
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
