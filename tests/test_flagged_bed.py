"""Flagged-region BED.  The reference writes no BED: ntEdit -a1 soft-masks what it could not fix
(subprojects/ntedit/ntedit.cpp:1131-1146) and nothing else.  gp_flagged_bed derives the intervals from the polished
FASTA (maximal lower-case runs); parity = the same derivation applied to the reference's own output gives the same rows."""
import json
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from util import GOLDEN, KS, ROOT, dataset, oracle_build, oracle_polish_contig, plan

import goldpolish_b200 as gp


def _regex_bed(records):
    """Independent statement of the derivation: (name, start, end) of every [a-z]+ run."""
    return [(n, m.start(), m.end()) for n, s in records for m in re.finditer(r"[a-z]+", s)]


def _csr(seqs):
    data = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint64)
    return data, off


def test_derivation_edges():
    seqs = ["", "acgt", "ACGT", "aCgT", "ACgtNNnnACGTac", "NNNN", "a", "Z[`{z"]
    names = [f"r{i}" for i in range(len(seqs))]
    data, off = _csr(seqs)
    got = gp.flagged_bed(data if data.size else np.zeros(1, np.uint8), off, names)
    assert got == _regex_bed(zip(names, seqs))
    assert ("r4", 2, 4) in got and ("r4", 6, 8) in got and ("r4", 12, 14) in got and ("r7", 4, 5) in got


def test_golden_cases_reference_output():
    """The golden ntEdit cases were minted from the reference's own ntedit-gr (tests/golden/make_golden.py): the BED of
    its `_edited.fa` records, derived by the C ABI and by the independent regex, is the same; several cases do flag."""
    g = json.load(open(os.path.join(GOLDEN, "ntedit_cases.json")))
    recs = [(n, c["chain"]) for n, c in g["cases"].items() if c["chain"] is not None]
    data, off = _csr([s for _, s in recs])
    got = gp.flagged_bed(data, off, [n for n, _ in recs])
    assert got == _regex_bed(recs)
    assert len(got) >= 3


@pytest.mark.gpu
def test_bed_of_gpu_output_equals_bed_of_oracle_output():
    d = dataset(genome_len=60000)
    pl = plan(d, bsize=1)
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
        out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
    names = [d.contig_name(c) for c in range(d.n_contigs)]
    got = gp.flagged_bed(out, off, names)
    ref = oracle_build(d, pl)
    ref_recs = []
    for c in range(d.n_contigs):
        s = oracle_polish_contig(d.contig(c), ref[int(pl.contig_batch[c])].bfs)
        ref_recs.append((names[c], "" if s is None else s.decode()))
    assert got == _regex_bed(ref_recs)
    assert len(got) > 10  # 5 % read error at 30x leaves unfixable positions: something is flagged


@pytest.mark.gpu
def test_ntedit_gr_bed_option_matches_reference_cli_output():
    """ntedit-gr --bed on the CLI golden input: rows = the derivation applied to the reference binary's _edited.fa."""
    g = json.load(open(os.path.join(GOLDEN, "ntedit_cli.json")))
    cases = json.load(open(os.path.join(GOLDEN, "ntedit_cases.json")))
    from test_gpu_golden import _truth_filters
    bfs = _truth_filters(gp, cases)
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "in.fa")
        open(fa, "w").write(g["input_fasta"])
        bf = os.path.join(td, "k32.bf")
        hdr = ('[BTLKmerBloomFilter_v6]\nbytes = 524288\nhash_fn = "ntHash_v2"\nhash_num = 4\nk = 32\n[HeaderEnd]\n').encode()
        open(bf, "wb").write(hdr + bfs[0, 0].tobytes())
        bed = os.path.join(td, "flagged.bed")
        subprocess.check_call([os.path.join(ROOT, "goldpolish_b200", "bin", "ntedit-gr"), "-f", fa, "-r", bf, "-b",
                               os.path.join(td, "o"), "-d5", "-i5", "-m1", "-X0.5", "-Y0.5", "-t1", "-a1", "--bed", bed])
        assert open(os.path.join(td, "o_edited.fa")).read() == g["edited_fasta"]
        rows = [tuple(ln.split("\t")) for ln in open(bed).read().splitlines()]
    recs, name = [], None
    for ln in g["edited_fasta"].splitlines():
        if ln.startswith(">"):
            name = ln[1:].split()[0]
        else:
            recs.append((name, ln))
    assert [(n, int(a), int(b)) for n, a, b in rows] == _regex_bed(recs)
