"""GPU: the CUDA path against the golden vectors minted from the reference's own code."""
import json
import os

import numpy as np
import pytest

from util import GOLDEN, KS, sha

pytestmark = pytest.mark.gpu


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gp():
    import goldpolish_b200
    return goldpolish_b200


def test_filters_golden(gp):
    import sim
    for case in _load("filters.json")["cases"]:
        d = sim.simulate(**case["sim"])
        pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)],
                             [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off),
                             d.map_read, d.map_contig, bsize=case["bsize"])
        with gp.Context(keep_counters=1) as ctx:
            ctx.upload_reads(d.read_seq, d.read_off)
            bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
            for b, rec in enumerate(case["batches"]):
                assert [sha(bfs[b, i]) for i in range(4)] == rec["bf_sha256"], (case["name"], b)
                assert [sha(ctx.fetch_cbf(b, i)) for i in range(4)] == rec["cbf_sha256"], (case["name"], b)
                assert [int(np.unpackbits(bfs[b, i]).sum()) for i in range(4)] == rec["bf_popcount"]


def _truth_filters(gp, g):
    truth = np.frombuffer(g["truth"].encode(), dtype=np.uint8)
    reads = np.tile(truth, 5)
    off = np.arange(6, dtype=np.uint64) * len(truth)
    ent = np.array([(i, 4) for i in range(5)], dtype=[("read_id", np.uint32), ("kmer_threshold", np.uint32)])
    with gp.Context() as ctx:
        ctx.upload_reads(reads, off)
        bfs = ctx.build_filters(np.array([0, 5], dtype=np.uint64), ent)
    assert [sha(bfs[0, i]) for i in range(4)] == g["bf_sha256"]
    return bfs


def _polish(gp, bfs, drafts, ks=KS, **kw):
    seq = np.frombuffer(b"".join(drafts), dtype=np.uint8)
    off = np.cumsum([0] + [len(x) for x in drafts]).astype(np.uint64)
    with gp.Context(ks=ks, **kw) as ctx:
        ctx.load_filters(bfs)
        out, ooff, dropped = ctx.polish(seq, off, np.zeros(len(drafts), dtype=np.uint32))
    return [None if dropped[i] else out[int(ooff[i]):int(ooff[i + 1])].tobytes().decode() for i in range(len(drafts))]


def test_ntedit_golden_cases(gp):
    g = _load("ntedit_cases.json")
    bfs = _truth_filters(gp, g)
    names = list(g["cases"])
    drafts = [g["cases"][n]["draft"].encode() for n in names]
    chain = _polish(gp, bfs, drafts)
    for n, got in zip(names, chain):
        assert got == g["cases"][n]["chain"], f"chain differs for case {n}"
    for i, k in enumerate(KS):
        per_k = _polish(gp, bfs[:, i:i + 1, :], drafts, ks=(k,))
        for n, got in zip(names, per_k):
            assert got == g["cases"][n]["per_k"][str(k)], f"k={k} differs for case {n}"
    for n in names:
        rec = g["cases"][n]
        if "mode0" in rec:
            d1 = [rec["draft"].encode()]
            b32 = bfs[:, 0:1, :]
            assert _polish(gp, b32, d1, ks=(32,), mode=0)[0] == rec["mode0"]
            assert _polish(gp, b32, d1, ks=(32,), mode=2, max_insertions=2, max_deletions=2)[0] == rec["mode2"]
            assert _polish(gp, b32, d1, ks=(32,), mask=0, max_insertions=3, max_deletions=10)[0] == rec["nomask_i3_d10"]


def test_ntedit_cli_golden(gp):
    g = _load("ntedit_cli.json")
    bfs = _truth_filters(gp, _load("ntedit_cases.json"))
    ins, name = {}, None
    for ln in g["input_fasta"].splitlines():
        if ln.startswith(">"):
            name = ln[1:]
            ins[name] = ""
        else:
            ins[name] += ln
    names = list(ins)
    got = _polish(gp, bfs[:, 0:1, :], [ins[n].encode() for n in names], ks=(32,))
    fasta = "".join(f">{n}\n{s}\n" for n, s in zip(names, got) if s is not None)
    assert fasta == g["edited_fasta"]


def test_edge_inputs(gp):
    """Empty and ragged inputs: no batches, batches without reads, reads shorter than k, empty contig list."""
    with gp.Context() as ctx:
        reads = np.frombuffer(b"ACGTACGTAC" + b"ACGT" * 20 + b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", dtype=np.uint8)
        off = np.array([0, 10, 90, 130], dtype=np.uint64)
        ctx.upload_reads(reads, off)
        ent = np.array([(0, 5), (1, 5), (2, 5), (1, 4)], dtype=[("read_id", np.uint32), ("kmer_threshold", np.uint32)])
        bfs = ctx.build_filters(np.array([0, 0, 3, 4], dtype=np.uint64), ent)
        assert bfs.shape[0] == 3 and not bfs[0].any()
        st = ctx.stats()
        # read 1 (80 bp of ACGT repeats) contributes len-k+1 k-mers per k, twice; reads 0 and 2 none
        assert st["kmer_ops"] == 2 * sum(80 - k + 1 for k in KS)
        out, ooff, dropped = ctx.polish(np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(0, np.uint32))
        assert len(dropped) == 0 and ooff.tolist() == [0]
        with pytest.raises(gp.GpError):
            ctx.build_filters(np.array([0, 1], dtype=np.uint64), np.array([(0, 3)], dtype=ent.dtype))  # T < 4
        with pytest.raises(gp.GpError):
            ctx.build_filters(np.array([0, 1], dtype=np.uint64), np.array([(9, 5)], dtype=ent.dtype))  # bad read id
        bfs0 = ctx.build_filters(np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=ent.dtype))
        assert bfs0.shape[0] == 0


def test_nthash_known_answers_on_device(gp):
    """A6 pinned directly: h0..h3 of every valid k-mer as the build kernels' own device code computes them (packed words,
    mask window, byte tables) against the KATs minted from the reference's hashing (tests/golden/nthash_kat.json: random
    32/28/24/20-mers, lower case, N runs)."""
    kats = _load("nthash_kat.json")
    seqs = [r["seq"].encode() for r in kats]
    data = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    off = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint64)
    n = 0
    with gp.Context() as ctx:
        ctx.upload_reads(data, off)
        for i, rec in enumerate(kats):
            pos, hs = ctx.debug_nthash(i, rec["k"])
            assert pos.tolist() == rec["pos"], i
            assert [[str(x) for x in row] for row in hs.tolist()] == rec["hashes"], i
            n += len(rec["pos"])
    assert n > 1000
