"""Seeded long-read simulator over GIVEN truth sequences (test tooling for BASELINE.json configs[0]).

The reference's own test (tests/goldpolish_test.sh:6) downloads its reads; there is no network here, so the reads for
the in-tree draft fixture are simulated from the in-tree EXPECTED polished assembly.
Only Generator.integers / Generator.random are used, so the stream is stable across numpy versions.
"""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = os.path.join(HERE, "golden", "fixtures")
_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def read_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    recs, name, parts = [], None, []
    with op(path, "rb") as f:
        for ln in f.read().split(b"\n"):
            if ln.startswith(b">"):
                if name is not None:
                    recs.append((name, b"".join(parts)))
                name, parts = ln[1:].decode().split()[0], []
            elif ln:
                parts.append(ln.strip())
    if name is not None:
        recs.append((name, b"".join(parts)))
    return recs


def load_fixture(which: str):
    """-> (draft records, truth records) of the committed fixture `which` ("config1" | "target").  The draft is the
    reference's in-tree test input; the truth is that draft with the records replaced that the reference's expected
    output changes (config1), or its upper-cased self (target: no expected output of the hot path alone exists)."""
    draft = read_fasta(os.path.join(FIXTURES, f"{which}_draft.fa.gz"))
    truth = dict(draft)
    p = os.path.join(FIXTURES, f"{which}_truth_diff.json.gz")
    if os.path.exists(p):
        with gzip.open(p, "rt") as f:
            truth.update({k: v.encode() for k, v in json.load(f).items()})
    return draft, [(n, truth[n].upper()) for n, _ in draft]


def simulate_reads(truth, coverage=30.0, seed=20250607, err=0.05):
    """-> (reads [(name, seq, quality char)], mappings [(read name, contig name, strand, start, end)]).
    Reads never span contigs; 30 % substitutions / 30 % insertions / 40 % deletions of the errors."""
    rng = np.random.default_rng(seed)
    reads, maps = [], []
    for cname, seq in truth:
        L = len(seq)
        arr = np.frombuffer(seq, dtype=np.uint8)
        covered = 0
        while covered < coverage * L:
            rl = int(min(L, 1500 + rng.integers(0, 12000)))
            start = int(rng.integers(0, L - rl + 1))
            piece = arr[start:start + rl].copy()
            u = rng.random(rl)
            sub = u < err * 0.3
            ins = (u >= err * 0.3) & (u < err * 0.6)
            dele = (u >= err * 0.6) & (u < err)
            code = np.searchsorted(_BASES, np.where(np.isin(piece, _BASES), piece, ord("A")))
            piece[sub] = _BASES[(code[sub] + 1 + rng.integers(0, 3, size=int(sub.sum()))) % 4]
            keep = ~dele
            out = piece[keep]
            ins_pos = np.cumsum(keep)[ins & keep] if ins.any() else np.zeros(0, dtype=np.int64)
            if len(ins_pos):
                out = np.insert(out, ins_pos, _BASES[rng.integers(0, 4, size=len(ins_pos))])
            s = out.tobytes()
            strand = int(rng.integers(0, 2))
            if strand:
                s = s.translate(_COMP)[::-1]
            name = f"r{len(reads)}"
            reads.append((name, s, 33 + 5 + int(rng.integers(0, 20))))
            maps.append((name, cname, strand, start, start + rl))
            covered += rl
    return reads, maps


def write_inputs(workdir, draft, reads, maps):
    """draft.fa, reads.fq (2 + 2 lines per record, as src/seqindex.cpp:21-61 expects), mappings.paf"""
    os.makedirs(workdir, exist_ok=True)
    paths = {k: os.path.join(workdir, v) for k, v in (("draft", "draft.fa"), ("reads", "reads.fq"), ("paf", "mappings.paf"))}
    with open(paths["draft"], "wb") as f:
        for n, s in draft:
            f.write(b">" + n.encode() + b"\n" + s + b"\n")
    rlen = {}
    with open(paths["reads"], "wb") as f:
        for n, s, q in reads:
            rlen[n] = len(s)
            f.write(b"@" + n.encode() + b"\n" + s + b"\n+\n" + bytes([q]) * len(s) + b"\n")
    clen = {n: len(s) for n, s in draft}
    with open(paths["paf"], "w") as f:
        for r, c, strand, a, b in maps:
            f.write(f"{r}\t{rlen[r]}\t0\t{rlen[r]}\t{'-' if strand else '+'}\t{c}\t{clen[c]}\t{a}\t{min(b, clen[c])}\t{b - a}\t{b - a}\t60\n")
    return paths
