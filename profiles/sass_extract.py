#!/usr/bin/env python3
"""SASS of the hot kernels, trimmed: `cuobjdump -sass` of the built library, one line per instruction (offset + text,
encodings dropped), plus a mnemonic histogram per kernel.
usage: profiles/sass_extract.py <lib.so> <out_dir> <tag> <kernel substring>..."""
import collections
import os
import re
import subprocess
import sys

lib, out_dir, tag, kernels = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, body = None, collections.defaultdict(list)
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;\s*/\*", ln)
    if m and fn:
        body[fn].append((m.group(1), re.sub(r"\s+", " ", m.group(2))))
summary = []
for k in kernels:
    for fn_name, ins in body.items():
        if k not in fn_name:
            continue
        path = os.path.join(out_dir, f"{tag}_{k}_sass.txt")
        hist = collections.Counter()
        for _, t in ins:
            t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
            hist[t2.split()[0].split(".")[0] + ("." + ".".join(t2.split()[0].split(".")[1:3]) if t2.startswith(("RED", "ATOM", "LDG", "STG", "LDS", "MEMBAR", "BAR")) else "")] += 1
        with open(path, "w") as f:
            f.write(f"// {fn_name}\n// {len(ins)} instructions; cuobjdump -sass {os.path.basename(lib)} (sm_100a), encodings dropped\n")
            f.write("// mnemonics: " + ", ".join(f"{n} x{c}" for n, c in hist.most_common(40)) + "\n")
            for off, t in ins:
                f.write(f"{off} {t}\n")
        summary.append((k, len(ins), hist))
for k, n, hist in summary:
    print(k, n, "instructions;", ", ".join(f"{a} x{c}" for a, c in hist.most_common(12)))
