"""goldpolish-mask / goldpolish-to-upper: the oracle restatement against golden vectors
minted from the reference's own script, and the device pass (gp_prep, and fused behind the edit kernel)
against both."""
import json
import os

import numpy as np
import pytest

from oracle import mask_oracle as mo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "mask_golden.json")))


def test_oracle_matches_reference_script():
    for run in GOLD["runs"]:
        for s, want in zip(GOLD["seqs"], run["out"]):
            assert mo.mask(s, run["k"], run["hard"]) == want


def _pack(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer("".join(seqs).encode() or b"\0", dtype=np.uint8).copy()
    return buf, off


def _unpack(out, off):
    return [out[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]


@pytest.mark.gpu
def test_device_mask_matches_golden():
    import goldpolish_b200 as gp
    buf, off = _pack(GOLD["seqs"])
    with gp.Context() as ctx:
        for run in GOLD["runs"]:
            got = _unpack(*ctx.prep(buf, off, 2 if run["hard"] else 1, run["k"]))
            assert got == run["out"], f"k={run['k']} hard={run['hard']}"
        # mask then to-upper, and to-upper alone (scripts/goldpolish-to-upper:15-21)
        got = _unpack(*ctx.prep(buf, off, 1, 32, to_upper=1))
        assert got == [mo.to_upper(mo.mask(s, 32)) for s in GOLD["seqs"]]
        got = _unpack(*ctx.prep(buf, off, 0, 0, to_upper=1))
        assert got == [mo.to_upper(s) for s in GOLD["seqs"]]


@pytest.mark.gpu
def test_mask_fused_behind_the_edit_kernel():
    """prep_mode in the context = `goldpolish-mask -s -k32` of the polished records (goldpolish-make:65-66),
    for the separate calls and for the overlapped pass."""
    import goldpolish_b200 as gp
    from util import dataset, plan
    d = dataset(genome_len=100000)
    pl = plan(d, bsize=2)
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ctx.build_filters(pl.batch_entry_off, pl.entries, fetch=False)
        out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
        plain = _unpack(out, off)
    want = [mo.mask(s, 32) if not dropped[i] else "" for i, s in enumerate(plain)]
    assert any(w != p for w, p in zip(want, plain))  # the pass does something on this data
    with gp.Context(prep_mode=1, prep_k=32) as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ctx.build_filters(pl.batch_entry_off, pl.entries, fetch=False)
        out, off, dropped2 = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
        assert _unpack(out, off) == want and np.array_equal(dropped, dropped2)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
        ctx.pipeline_run()
        out, off, _ = ctx.polish_fetch()
        assert _unpack(out, off) == want


@pytest.mark.gpu
def test_mask_and_to_upper_tools(tmp_path):
    """The CLI drop-ins: `goldpolish-mask -s -k32 file > out` (goldpolish-make:65-66), stdin form, the -n form,
    argparse-style refusals, and `goldpolish-to-upper in out` (goldpolish-make:47-48)."""
    import subprocess
    from util import ROOT
    BIN = os.path.join(ROOT, "goldpolish_b200", "bin")
    seqs = [s for s in GOLD["seqs"] if s][:30]
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for i, s in enumerate(seqs):
            f.write(f">rec{i}" + (" some comment" if i % 3 == 0 else "") + "\n")
            for j in range(0, len(s), 60):  # multi-line input
                f.write(s[j:j + 60] + "\n")

    def parse(text):
        lines = text.splitlines()
        return [(lines[i], lines[i + 1]) for i in range(0, len(lines), 2)]

    def hdr(i):
        return f">rec{i}" + (" some comment" if i % 3 == 0 else "")

    out = subprocess.run([os.path.join(BIN, "goldpolish-mask"), "-s", "-k32", str(fa)], capture_output=True, text=True, check=True).stdout
    assert parse(out) == [(hdr(i), mo.mask(s, 32)) for i, s in enumerate(seqs)]
    out = subprocess.run([os.path.join(BIN, "goldpolish-mask"), "-n", "-k", "20", "-"], stdin=open(fa), capture_output=True, text=True, check=True).stdout
    assert parse(out) == [(hdr(i), mo.mask(s, 20, hard=True)) for i, s in enumerate(seqs)]
    assert subprocess.run([os.path.join(BIN, "goldpolish-mask"), "-k32", str(fa)], capture_output=True).returncode == 2
    assert subprocess.run([os.path.join(BIN, "goldpolish-mask"), "-s", "-n", "-k32", str(fa)], capture_output=True).returncode == 2
    up = tmp_path / "up.fa"
    subprocess.run([os.path.join(BIN, "goldpolish-to-upper"), str(fa), str(up)], check=True)
    assert parse(open(up).read()) == [(hdr(i), s.upper()) for i, s in enumerate(seqs)]
