#!/usr/bin/env python3
"""Per-CUDA-source-line instruction / stall-sample shares of an ncu report.

ncu's CSV export carries metrics only on the SASS page, so the SASS rows are joined (by
instruction order) with `nvdisasm -g` line annotations of the SAME build's cubin.
usage: profiles/ncu_lines.py <report.ncu-rep> <lib.so|cubin> <kernel-name-substring> [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile

rep, binp, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
if binp.endswith(".so"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binp)], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
else:
    cubins = [binp]
lines = []  # (offset, file, line) in order for the kernel
for cb in cubins:
    txt = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
    infn = False; cur = ("?", 0)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            infn = kern in m.group(1); continue
        if not infn: continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m: lines.append((int(m.group(1), 16), cur[0], cur[1]))
    if lines: break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = {n: i for i, n in enumerate(rows[1])}
sass = []
for r in rows[2:]:
    try:
        sass.append((int(r[h["Address"]], 16), float(r[h["Warp Stall Sampling (All Samples)"]] or 0), float(r[h["Instructions Executed"]] or 0)))
    except (ValueError, IndexError):
        pass
base = sass[0][0]
off2line = {o: (f, l) for o, f, l in lines}
agg = {}
miss = 0
for a, s, e in sass:
    key = off2line.get(a - base)
    if key is None: miss += 1; key = ("?", 0)
    v = agg.setdefault(key, [0.0, 0.0]); v[0] += s; v[1] += e
S = sum(v[0] for v in agg.values()) or 1; E = sum(v[1] for v in agg.values()) or 1
print(f"{len(sass)} SASS rows, {len(lines)} disassembled, {miss} unmatched; {int(E)} warp-instr, {int(S)} samples")
srcs = {}
def text(f, l):
    if f not in srcs:
        for d in ("goldpolish_b200/csrc", "."):
            p = os.path.join(d, f)
            if os.path.exists(p): srcs[f] = open(p).read().splitlines(); break
        else: srcs[f] = []
    return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
for (f, l), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][int(os.environ.get("SORT_SAMPLES","0")) ^ 1])[:top]:
    print(f"{100*e/E:5.1f}% instr {100*s/S:5.1f}% samples  {f}:{l}  {text(f, l)}")
