"""GPU, BASELINE.json configs[1] size (5 Mbp draft, 30x reads, 543 batches) and a 10 Mbp bsize-8 cut of configs[2]:
the whole workload through the device paths, a stratified sample of its batches against the reference's own code
(oracle/_ref: serve_batch + ntEdit chain + guard; the C restatement where oracle/_ref is not built), and
size-independent properties.  (Small-input bit-exactness: test_gpu_parity.py / test_gpu_golden.py.)"""
import hashlib
import os

import numpy as np
import pytest

from util import KS, dataset, plan

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    import goldpolish_b200 as gp
    d = dataset(genome_len=5_000_000)
    pl = plan(d, bsize=1)
    ctx = gp.Context()
    ctx.upload_reads(d.read_seq, d.read_off)
    yield gp, d, pl, ctx
    ctx.close()


def _valid_kmers_per_read(d, k):
    """Number of k-mers made of ACGT/acgt only, per read (what btllib::NtHash::roll yields)."""
    seq = d.read_seq
    ok = np.isin(seq | 0x20, np.frombuffer(b"acgt", dtype=np.uint8))
    # a k-mer starting at i is valid iff the k flags from i are all set: windowed sum via cumsum
    cs = np.concatenate([[0], np.cumsum(ok, dtype=np.int64)])
    out = np.zeros(d.n_reads, dtype=np.int64)
    for r in range(d.n_reads):
        a, b = int(d.read_off[r]), int(d.read_off[r + 1])
        if b - a >= k:
            w = cs[a + k:b + 1] - cs[a:b + 1 - k]
            out[r] = int(np.count_nonzero(w == k))
    return out


def test_kmer_op_count_and_checksum_of_checksums(full):
    gp, d, pl, ctx = full
    bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
    st = ctx.stats()
    per_read = sum(_valid_kmers_per_read(d, k) for k in KS)
    assert st["kmer_ops"] == int(per_read[pl.entries["read_id"]].sum())
    assert st["serial_kmers"] < st["kmer_ops"] // 50
    digest = hashlib.sha256(b"".join(hashlib.sha256(bfs[b].tobytes()).digest() for b in range(bfs.shape[0]))).hexdigest()
    # waves (limited counting-filter residency) and a second context give the same bits
    ctx2 = gp.Context(max_resident_batches=97)
    ctx2.upload_reads(d.read_seq, d.read_off)
    bfs2 = ctx2.build_filters(pl.batch_entry_off, pl.entries)
    ctx2.close()
    digest2 = hashlib.sha256(b"".join(hashlib.sha256(bfs2[b].tobytes()).digest() for b in range(bfs2.shape[0]))).hexdigest()
    assert digest == digest2
    # every filter of a batch with reads has bits, and never more than 4 per k-mer op
    pop = np.unpackbits(bfs.reshape(bfs.shape[0], -1), axis=1).sum(axis=1)
    has_reads = np.diff(pl.batch_entry_off.astype(np.int64)) > 0
    assert np.all((pop > 0) == has_reads)


def test_polish_subset_consistency_and_identity(full):
    gp, d, pl, ctx = full
    bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
    out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
    st = ctx.stats()
    assert st["edits"] > 1000 and st["masked"] > 1000
    # batches are independent: polishing a handful of contigs alone gives the same records
    pick = [0, 7, 100, 333, d.n_contigs - 1]
    seq = np.concatenate([d.contig_seq[d.contig_off[c]:d.contig_off[c + 1]] for c in pick])
    soff = np.cumsum([0] + [int(d.contig_off[c + 1] - d.contig_off[c]) for c in pick]).astype(np.uint64)
    ctx3 = gp.Context()
    ctx3.load_filters(np.stack([bfs[pl.contig_batch[c]] for c in pick]))
    o3, f3, d3 = ctx3.polish(seq, soff, np.arange(len(pick), dtype=np.uint32))
    for i, c in enumerate(pick):
        assert d3[i] == dropped[c]
        assert o3[int(f3[i]):int(f3[i + 1])].tobytes() == out[int(off[c]):int(off[c + 1])].tobytes()
    # all-ones filters: nothing is absent, nothing changes (records < 100 bp are dropped)
    ctx3.load_filters(np.full((1, 4, gp.BF_BYTES), 0xFF, dtype=np.uint8))
    o4, f4, d4 = ctx3.polish(d.contig_seq, d.contig_off, np.zeros(d.n_contigs, dtype=np.uint32))
    ctx3.close()
    lens = np.diff(d.contig_off)
    assert np.array_equal(d4 != 0, lens < 100)
    kept = np.concatenate([d.contig_seq[d.contig_off[c]:d.contig_off[c + 1]] for c in range(d.n_contigs) if lens[c] >= 100])
    assert np.array_equal(o4[:int(f4[-1])], kept)


@pytest.mark.parametrize("bsize", [1, 8])
def test_three_device_paths_agree(full, bsize, monkeypatch):
    """The oracle is too slow at this size, but three independent device paths must agree bit for bit:
    the in-order one-warp-per-stream kernel (counters in HBM, sequential semantics inside a warp), the
    level-synchronous kernel (order-free rounds over timestamps), and the overlapped pipeline (batches
    built longest-contig first, edit kernel beside the build kernel)."""
    gp, d, pl, ctx = full
    if bsize != 1:  # large streams: heavily loaded counting filters, weighted shares at work
        pl = plan(d, bsize=bsize)
    monkeypatch.setenv("GP_BUILD_KERNEL", "s")
    a = ctx.build_filters(pl.batch_entry_off, pl.entries)
    assert ctx.stats()["build_kernel"] == 1
    monkeypatch.setenv("GP_BUILD_KERNEL", "l")
    b = ctx.build_filters(pl.batch_entry_off, pl.entries)
    assert ctx.stats()["build_kernel"] == 2
    assert np.array_equal(a, b)
    out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
    out = out[:int(off[-1])].copy()
    ctx.build_stage(pl.batch_entry_off, pl.entries)
    ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
    ctx.pipeline_run()
    c = ctx.build_fetch()
    out2, off2, dropped2 = ctx.polish_fetch()
    assert np.array_equal(a, c)
    assert np.array_equal(off, off2) and np.array_equal(dropped, dropped2)
    assert np.array_equal(out, out2[:int(off2[-1])])


@pytest.mark.skipif(not os.environ.get("GP_BIG_TESTS"), reason="BASELINE.json configs[2] size: set GP_BIG_TESTS=1 (about 2 minutes)")
def test_config3_size_paths_agree():
    """configs[2]: 100 Mbp draft, 40x reads, bsize 8 (28 G k-mer ops, heavily loaded counting filters, where an
    order-free update would differ -- SURVEY Appendix C): the in-order kernel, the level-synchronous kernel and
    the overlapped pipeline give the same filters and the same polished records."""
    import goldpolish_b200 as gp
    d = dataset(genome_len=100_000_000, coverage=40.0)
    pl = plan(d, bsize=8)
    digests = {}
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        for algo in ("s", "l"):
            os.environ["GP_BUILD_KERNEL"] = algo
            bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
            digests[algo] = hashlib.sha256(bfs.tobytes()).hexdigest()
        os.environ.pop("GP_BUILD_KERNEL")
        out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
        digests["polish"] = hashlib.sha256(out[:int(off[-1])].tobytes()).hexdigest()
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
        ctx.pipeline_run()
        digests["pipe"] = hashlib.sha256(ctx.build_fetch().tobytes()).hexdigest()
        out2, off2, dropped2 = ctx.polish_fetch()
        digests["pipe_polish"] = hashlib.sha256(out2[:int(off2[-1])].tobytes()).hexdigest()
    assert digests["s"] == digests["l"] == digests["pipe"]
    assert digests["polish"] == digests["pipe_polish"] and np.array_equal(dropped, dropped2)


def _check_batches_against_reference(gp, d, pl, bsize, batches, bfs, out, off, dropped):
    """Filters and polished records (guard applied, as goldpolish-ntedit leaves them) of the given batches against the
    reference's own code run on exactly those batches.  Returns the number of records compared."""
    from oracle.ref_sample import ReferenceSample
    w = dict(bsize=bsize, subsample_max=40.0, mx_max=150.0, mappings="paf")
    rs = ReferenceSample(w, d, batches, threads=os.cpu_count() or 1)
    n = 0
    try:
        rs.run()
        for i, b in enumerate(batches):
            assert np.array_equal(rs.filters(i), bfs[b]), f"filter payloads of batch {b} differ from the reference ({rs.kind})"
            cs = range(b * bsize, min((b + 1) * bsize, d.n_contigs))
            mine = [(d.contig_name(c), out[int(off[c]):int(off[c + 1])].tobytes()) for c in cs if not dropped[c]]
            in_sz = sum(len(d.contig_name(c)) + 3 + int(d.contig_off[c + 1] - d.contig_off[c]) for c in cs)
            out_sz = sum(len(nm) + 3 + len(sq) for nm, sq in mine)
            if gp.guard_rejects(in_sz, out_sz):  # scripts/goldpolish-ntedit:31-40
                mine = [(d.contig_name(c), d.contig(c)) for c in cs]
            assert mine == rs.polished(i), f"polished records of batch {b} differ from the reference ({rs.kind})"
            n += len(mine)
    finally:
        rs.close()
    return n


def _rollbacks_in(d, bfs, bsize, batches):
    """Low-complexity insertion rollbacks (ntedit.cpp:1038-1066) that the C restatement counts in these batches."""
    from oracle import oracle_lib as ol
    n = 0
    for b in batches:
        for c in range(b * bsize, min((b + 1) * bsize, d.n_contigs)):
            cur = d.contig(c)
            for ki, k in enumerate(KS):
                cur, st = ol.ntedit_contig(cur, bfs[b, ki], k)
                if cur is None:
                    break
                n += st["rollbacks"]
    return n


def test_config2_size_sample_matches_reference(full):
    """The benched workload itself (543 batches: ~34 k edits, ~3 M masked positions, ~90 rollbacks per pass), through the
    overlapped pipeline; then the most loaded batch, 30 random ones and -- if the random ones held none -- a batch with
    a low-complexity rollback are compared with what the reference's own code produces for them."""
    gp, d, pl, ctx = full
    ctx.build_stage(pl.batch_entry_off, pl.entries)
    ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
    ctx.pipeline_run()
    bfs = ctx.build_fetch()
    out, off, dropped = ctx.polish_fetch()
    st = ctx.stats()
    assert st["rollbacks"] > 0 and st["edits"] > 10000
    n_batches = len(pl.batch_entry_off) - 1
    load = np.diff(pl.batch_entry_off.astype(np.int64))
    rng = np.random.default_rng(20250607)
    batches = [int(np.argmax(load))] + sorted(int(b) for b in rng.choice(n_batches, size=30, replace=False))
    batches = list(dict.fromkeys(batches))
    if _rollbacks_in(d, bfs, 1, batches) == 0:
        for b in range(n_batches):
            if b not in batches and _rollbacks_in(d, bfs, 1, [b]):
                batches.append(b)
                break
    assert _rollbacks_in(d, bfs, 1, batches) > 0, "the checked sample must exercise the rollback path"
    assert _check_batches_against_reference(gp, d, pl, 1, batches, bfs, out, off, dropped) >= 30


def test_10mbp_bsize8_paths_agree_and_loaded_batches_match_reference(monkeypatch):
    """A 10 Mbp / 40x / bsize 8 cut of configs[2] (streams of ~6 M k-mers into counting filters loaded well beyond one
    touch per counter, where an order-free update differs -- SURVEY Appendix C): the in-order kernel, the
    level-synchronous kernel and the overlapped pipeline agree on every batch, and the 6 most loaded batches match the
    reference's own code."""
    import goldpolish_b200 as gp
    d = dataset(genome_len=10_000_000, coverage=40.0)
    pl = plan(d, bsize=8)
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        monkeypatch.setenv("GP_BUILD_KERNEL", "s")
        a = ctx.build_filters(pl.batch_entry_off, pl.entries)
        monkeypatch.setenv("GP_BUILD_KERNEL", "l")
        b = ctx.build_filters(pl.batch_entry_off, pl.entries)
        assert ctx.stats()["build_kernel"] == 2 and np.array_equal(a, b)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
        ctx.pipeline_run()
        c = ctx.build_fetch()
        out, off, dropped = ctx.polish_fetch()
        assert np.array_equal(a, c)
    rl = np.diff(d.read_off)
    ent_bases = rl[pl.entries["read_id"]]
    eo = pl.batch_entry_off.astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(ent_bases)])
    load = cs[eo[1:]] - cs[eo[:-1]]
    batches = [int(x) for x in np.argsort(-load)[:6]]
    assert _check_batches_against_reference(gp, d, pl, 8, batches, c, out, off, dropped) >= 40
