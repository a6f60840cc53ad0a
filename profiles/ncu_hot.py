#!/usr/bin/env python3
"""Summarise an ncu report: key raw metrics + the SASS instructions with most stall samples.
usage: profiles/ncu_hot.py <report.ncu-rep> [top_n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__inst_executed_pipe_lsu.sum"]
print("== raw ==")
for i, h in enumerate(hdr):
    if h in keys:
        print(f"{h:70s} {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ci = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
recs = []
for r in rows[2:]:
    try:
        s = float(r[ci["Warp Stall Sampling (All Samples)"]])
        e = float(r[ci["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    st = {n: float(r[ci[n]] or 0) for n in stall_cols}
    recs.append((s, e, r[ci["Source"]].strip(), st))
S = sum(x[0] for x in recs) or 1
E = sum(x[1] for x in recs) or 1
print(f"== SASS: {len(recs)} instructions, {int(E)} warp-instructions executed, {int(S)} stall samples ==")
agg = {}
for s, e, t, st in recs:
    for n, v in st.items():
        agg[n] = agg.get(n, 0) + v
print("stall mix:", ", ".join(f"{n[6:]} {100 * v / S:.1f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, e, t, st in sorted(recs, key=lambda x: -x[0])[:top]:
    why = max(st.items(), key=lambda kv: kv[1])[0][6:] if s else "-"
    print(f"{100 * s / S:5.1f}% samples {100 * e / E:5.2f}% instr  [{why:10s}] {t[:90]}")
