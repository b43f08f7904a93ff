#!/usr/bin/env python
"""Condenses an Nsight Compute report (.ncu-rep) into a small text summary for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name.txt ["free-form note"]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# ncu summary of {rep}", f"# {note}" if note else "#"]
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines.append(f"\n## kernel: {name[:160]}")
    for i, h in enumerate(hdr):
        if h in WANT:
            lines.append(f"{h:75s} {r[i]:>18s} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    lines.append("warp stall reasons (warps stalled per issue slot): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    idx = {h: i for i, h in enumerate(sh)}
    data = [r for r in srows[2:] if len(r) == len(sh)]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data) or 1
    ops = {}
    for r in data:
        t = r[idx["Source"]].strip().split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        e, s_ = ops.get(op, (0, 0))
        ops[op] = (e + int(r[idx["Instructions Executed"]] or 0), s_ + int(r[idx["# Samples"]] or 0))
    te = sum(e for e, _ in ops.values()) or 1
    lines.append(f"\n## SASS mix (warp-level instructions executed: {te}; pc samples: {tot})")
    for op, (e, s_) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:14]:
        lines.append(f"{op:10s} executed {100 * e / te:5.1f} %   samples {100 * s_ / tot:5.1f} %")
    lines.append("\n## most-sampled SASS instructions (samples, executed, top stall columns)")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:12]:
        st = {k: int(r[idx[k]] or 0) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_wait") if k in idx}
        lines.append(f"{r[idx['# Samples']]:>7s} {r[idx['Instructions Executed']]:>10s}  {st}  {r[idx['Source']].strip()[:70]}")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
