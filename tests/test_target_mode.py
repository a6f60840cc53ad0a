"""BASELINE.json configs[3] (GoldPolish-Target): BED regions cut out of the draft with 64 bp flanks.  The wrapper
itself (Snakemake, re-mapping) is out of scope; what reaches the hot path is a set of SHORT upper-cased targets
with the few reads that cover them.  This checks the hot path on exactly that shape against the oracle."""
import numpy as np
import pytest

from util import KS, dataset, oracle_build, oracle_polish_contig

pytestmark = pytest.mark.gpu

FLANK = 64


def _extract(d, rnd, n_regions):
    """scripts/goldpolish-target-extract-seq.py:111-148: sort the regions of a contig, merge neighbours closer than
    2 * flank, cut [start - flank, end + flank), upper-case.  Returns (names, seqs, (contig, start, end))."""
    lens = np.diff(d.contig_off)
    big = np.flatnonzero(lens >= 3000)
    coords = {}
    for _ in range(n_regions):
        c = int(rnd.choice(big))
        a = int(rnd.integers(0, lens[c] - 600))
        coords.setdefault(c, []).append((a, a + int(rnd.integers(1, 501))))
    names, seqs, where = [], [], []
    for c in sorted(coords):
        cl = sorted(coords[c])
        merged = [cl[0]]
        for s, e in cl[1:]:
            if s - merged[-1][1] < 2 * FLANK:
                merged[-1] = (merged[-1][0], e)
            else:
                merged.append((s, e))
        seq = d.contig(c)
        for i, (s, e) in enumerate(merged, 1):
            a, b = max(0, s - FLANK), min(len(seq), e + FLANK)
            names.append(f"{d.contig_name(c)}.{i}")
            seqs.append(seq[a:b].upper())
            where.append((c, a, b))
    return names, seqs, where


def test_target_mode_regions_match_oracle():
    import goldpolish_b200 as gp
    d = dataset(genome_len=400000)
    rnd = np.random.default_rng(11)
    names, seqs, where = _extract(d, rnd, 160)
    # reads of a target = reads whose mapping on the source contig overlaps the cut (what re-mapping would report)
    mr, mc = [], []
    for t, (c, a, b) in enumerate(where):
        hit = np.flatnonzero((d.map_contig == c) & (d.map_tstart < b) & (d.map_tend > a))
        mr.extend(d.map_read[hit].tolist())
        mc.extend([t] * len(hit))
    tlens = np.array([len(s) for s in seqs])
    assert tlens.min() >= 2 * FLANK and tlens.max() < 3000
    # the ntLink default of the target pipeline: 100 reads per 10 kbp (scripts/goldpolish:54) -> 1..6 reads per target
    pl = gp.plan_batches(tlens, names, [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off),
                         np.array(mr, dtype=np.uint32), np.array(mc, dtype=np.uint32), bsize=8, subsample_max_per_10kbp=100.0)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(tlens)
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        ctx.polish_stage(buf, off, pl.contig_batch)
        ctx.pipeline_run()
        bfs = ctx.build_fetch()
        out, ooff, dropped = ctx.polish_fetch()
        st = ctx.stats()
    ref = oracle_build(d, pl)
    for b in range(len(pl.batch_entry_off) - 1):
        for ki in range(4):
            assert np.array_equal(bfs[b, ki], ref[b].bfs[ki]), f"filter differs: batch {b} k={KS[ki]}"
    changed = 0
    for t, s in enumerate(seqs):
        want = oracle_polish_contig(s, [bfs[pl.contig_batch[t], ki] for ki in range(4)])
        got = out[int(ooff[t]):int(ooff[t + 1])].tobytes()
        assert (want is None and dropped[t]) or got == want, f"target {names[t]} ({len(s)} bp) differs from the oracle"
        changed += got != s
    assert changed > 0 and st["kmer_ops"] > 0
