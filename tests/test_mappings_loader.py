"""The multi-threaded mappings loader of the drop-in server (goldpolish_b200/host/gp_host.hpp)
against the reference's OWN AllMappings (src/mappings.cpp, through oracle/_ref/libref_harness.so when it is built, and
always against golden dumps minted from it: tests/golden/mappings_golden.json, make_golden_mappings.py).

No GPU: `gp-host-check mappings` prints what the server would hold, for 1, 3 and 8 loader threads (slices are cut at
line ends; 1 MiB of text per thread at least, so the big case is what exercises the cuts)."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "goldpolish_b200", "bin")
ENV = dict(os.environ, GP_QUIET="1")
GOLDEN = os.path.join(HERE, "golden", "mappings_golden.json")


def make_cases(w):
    """name -> (mappings path, mx_max_per_10kbp); draft.fa(.index) in w.  Seeded, so the golden digests hold."""
    import numpy as np
    rng = np.random.default_rng(20250607)
    n_t = 400
    tl = rng.integers(2000, 50000, n_t)
    with open(os.path.join(w, "draft.fa"), "w") as f:
        for i in range(n_t):
            f.write(f">ctg{i} len={tl[i]}\n" + "A" * int(tl[i]) + "\n")
    subprocess.check_call([os.path.join(BIN, "goldpolish-index"), os.path.join(w, "draft.fa"), os.path.join(w, "draft.fa.index")], env=ENV)
    cases = {}
    # big ntLink-style file (> 4 MiB: several slices), repeated pairs, unknown targets, minimizer counts around the limits
    n = 260000
    r = rng.integers(0, 60000, n)
    t = rng.integers(0, n_t + 20, n)  # (ctg400..419 are not in the index)
    m = rng.integers(0, 45, n)
    with open(os.path.join(w, "big.tsv"), "w") as f:
        f.write("".join(f"read{a} ctg{b} {c}\n" for a, b, c in zip(r.tolist(), t.tolist(), m.tolist())))
    cases["ntlink_big"] = ("big.tsv", 150.0)
    cases["ntlink_big_tight"] = ("big.tsv", 20.0)
    # triples that ignore the line structure (`ifs >> token`, i % 3), odd spacing, CRLF, a dangling last token
    toks = []
    for a, b, c in zip(r[:3000].tolist(), t[:3000].tolist(), m[:3000].tolist()):
        toks += [f"read{a}", f"ctg{b}", f"+{c}" if c % 7 == 0 else str(c)]
    seps = [" ", "\t", "\n", "  ", "\r\n", " \n "]
    body = "".join(tok + seps[int(x)] for tok, x in zip(toks, rng.integers(0, len(seps), len(toks)).tolist()))
    with open(os.path.join(w, "ragged.tsv"), "w", newline="") as f:
        f.write(body + "readX")
    cases["ntlink_ragged"] = ("ragged.tsv", 60.0)
    # PAF: 12 columns; short lines keep the ids of the line before; '@' lines and empty lines; no newline at the end
    lines = []
    for i, (a, b) in enumerate(zip(r[:120000].tolist(), t[:120000].tolist())):
        if i % 997 == 5:
            lines.append(f"lonely{a}")                     # one column: a new read under the PREVIOUS target
        elif i % 1499 == 7:
            lines.append("")                               # empty line
        elif i % 1201 == 3:
            lines.append(f"@read{a}\t100\t0\t100\t+\tctg{b}\t5000\t0\t100\t90\t100\t60")  # skipped
        elif i % 811 == 2:
            lines.append(f"short{a}\t100\t0\t100\t+")      # five columns
        else:
            lines.append(f"read{a}\t100\t0\t100\t+\tctg{b}\t5000\t0\t100\t90\t100\t60" + ("\ttp:A:P" if i % 3 == 0 else ""))
    with open(os.path.join(w, "m.paf"), "w") as f:
        f.write("\n".join(lines))
    cases["paf"] = ("m.paf", 150.0)
    # SAM: header lines, target in column 3
    lines = ["@HD\tVN:1.6", "@SQ\tSN:ctg1\tLN:5"]
    for i, (a, b) in enumerate(zip(r[:50000].tolist(), t[:50000].tolist())):
        lines.append(f"read{a}\t0\tctg{b}\t1\t60\t5M\t*\t0\t0\tACGTA\t*" if i % 501 else f"read{a}\t4")
    with open(os.path.join(w, "m.sam"), "w") as f:
        f.write("\n".join(lines) + "\n")
    cases["sam"] = ("m.sam", 150.0)
    with open(os.path.join(w, "empty.paf"), "w"):
        pass
    cases["paf_empty"] = ("empty.paf", 150.0)
    # an index (src/seqindex.cpp:86-125: four tokens per record, whatever the lines) with duplicate ids -- the first
    # one wins, and its length decides how many reads the minimizer filter keeps -- ragged spacing, > 1 MiB
    recs = []
    for i in range(n_t):
        recs.append((f"ctg{i}", 10 + i, int(tl[i]), 0.0))
        if i % 3 == 0:
            recs.append((f"ctg{i}", 7, int(tl[i]) * 4, 1.5))      # later duplicate: ignored
    recs += [(f"pad{i}", i, 1000 + i, 12.25) for i in range(60000)]
    seps = [" ", "\t", "\n", "\t\t", " \n"]
    with open(os.path.join(w, "dup.index"), "w") as f:
        k = 0
        for rec in recs:
            for tok in rec:
                f.write(str(tok) + (seps[k % len(seps)] if k % 11 == 0 else "\t" if tok is not rec[-1] else "\n"))
                k += 1
    cases["ntlink_dup_index"] = ("big.tsv", 20.0, "dup.index")
    return cases


def ours(w, path, mx_max, threads, index="draft.fa.index"):
    return subprocess.check_output([os.path.join(BIN, "gp-host-check"), "mappings", os.path.join(w, "draft.fa"),
                                    os.path.join(w, index), os.path.join(w, path), str(mx_max), str(threads)],
                                   env=dict(ENV, GP_HOST_THREADS=str(threads)))


def reference(w, path, mx_max, index="draft.fa.index"):
    """The reference's own AllMappings, or None when oracle/_ref is not built."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    lib.ref_mappings_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_char_p]
    lib.ref_mappings_dump.restype = C.c_int
    out = os.path.join(w, "ref_dump.txt")
    assert lib.ref_mappings_dump(os.path.join(w, "draft.fa").encode(), os.path.join(w, index).encode(),
                                 os.path.join(w, path).encode(), mx_max, out.encode()) == 0
    return open(out, "rb").read()


def test_loader_matches_the_reference_mappings(tmp_path):
    w = str(tmp_path)
    golden = json.load(open(GOLDEN))
    for name, (path, mx_max, *index) in make_cases(w).items():
        want = reference(w, path, mx_max, *index)
        for threads in (1, 3, 8):
            got = ours(w, path, mx_max, threads, *index)
            if want is not None:
                assert got == want, (name, threads)
            assert hashlib.sha256(got).hexdigest() == golden[name]["sha256"], (name, threads)
            assert got.count(b"\n") == golden[name]["targets"], name


def test_compressed_and_bam_mappings_are_decoded(tmp_path):
    """The reference reads its mapping file through btllib::DataSource, which decodes compressed files and pipes .bam
    through `samtools view -h` (src/mappings.cpp:136-139); the format is still chosen by the FULL file name
    (:21-33: "x.paf.gz" is ntLink triples).  Same here: .gz through zlib, .bz2 / .xz through their tools, .bam through
    whatever `samtools` is on PATH (a stand-in script here); a missing tool is an error, not an empty mapping."""
    import bz2
    import gzip
    import lzma
    import stat
    w = str(tmp_path)
    cases = make_cases(w)
    raw = open(os.path.join(w, "ragged.tsv"), "rb").read()
    want = ours(w, "ragged.tsv", 60.0, 3)
    for suffix, comp in ((".gz", gzip.compress), (".bz2", bz2.compress), (".xz", lzma.compress)):
        with open(os.path.join(w, "ragged.tsv" + suffix), "wb") as f:
            f.write(comp(raw))
        assert ours(w, "ragged.tsv" + suffix, 60.0, 3) == want, suffix
    # "m.paf.gz": ntLink triples by name, as in the reference -- not PAF
    with open(os.path.join(w, "m.paf.gz"), "wb") as f:
        f.write(gzip.compress(raw))
    assert ours(w, "m.paf.gz", 60.0, 2) == want
    # .bam: SAM text from `samtools view -h <file>`
    sam_want = ours(w, "m.sam", 150.0, 2)
    os.rename(os.path.join(w, "m.sam"), os.path.join(w, "m.bam"))       # (the stand-in just cats it)
    bindir = os.path.join(w, "bin")
    os.makedirs(bindir)
    tool = os.path.join(bindir, "samtools")
    with open(tool, "w") as f:
        f.write('#!/bin/sh\n[ "$1" = view ] && [ "$2" = -h ] && exec cat "$3"\nexit 3\n')
    os.chmod(tool, os.stat(tool).st_mode | stat.S_IEXEC)
    env = dict(ENV, PATH=bindir + os.pathsep + os.environ.get("PATH", ""), GP_HOST_THREADS="2")
    cmd = [os.path.join(BIN, "gp-host-check"), "mappings", os.path.join(w, "draft.fa"), os.path.join(w, "draft.fa.index"),
           os.path.join(w, "m.bam"), "150", "2"]
    assert subprocess.check_output(cmd, env=env) == sam_want
    no_tool = subprocess.run(cmd, env=dict(ENV, PATH="/nonexistent"), capture_output=True)
    assert no_tool.returncode != 0 and b"samtools" in no_tool.stderr
