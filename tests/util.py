"""Shared test helpers: seeded data sets and the oracle-side pipeline (checker only)."""
from __future__ import annotations

import functools
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import sim  # noqa: E402
from oracle import oracle_lib as ol  # noqa: E402

KS = (32, 28, 24, 20)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes() if not isinstance(a, (bytes, bytearray)) else a).hexdigest()


@functools.lru_cache(maxsize=8)
def dataset(genome_len=60000, seed=20250607, **kw):
    return sim.simulate(genome_len=genome_len, seed=seed, **dict(kw))


def plan(d, bsize=1, subsample=40.0):
    import goldpolish_b200 as gp
    clens = np.diff(d.contig_off)
    rlens = np.diff(d.read_off)
    return gp.plan_batches(clens, [d.contig_name(i) for i in range(d.n_contigs)],
                           [d.read_name(i) for i in range(d.n_reads)], d.read_phred, rlens,
                           d.map_read, d.map_contig, bsize=bsize, subsample_max_per_10kbp=subsample)


def oracle_build(d, pl, batches=None):
    """Oracle filters for the planned batches: {batch: FilterSet}."""
    out = {}
    n_batches = len(pl.batch_entry_off) - 1
    for b in (range(n_batches) if batches is None else batches):
        fs = ol.FilterSet(KS)
        for e in range(int(pl.batch_entry_off[b]), int(pl.batch_entry_off[b + 1])):
            rid, thr = int(pl.entries[e]["read_id"]), int(pl.entries[e]["kmer_threshold"])
            fs.add_read(d.read(rid), thr)
        out[b] = fs
    return out


def oracle_polish_contig(seq: bytes, bfs, ks=KS, **opts):
    """k-chain of scripts/goldpolish-ntedit:20-29 on one contig; None when dropped."""
    cur = seq
    for bf, k in zip(bfs, ks):
        cur, _ = ol.ntedit_contig(cur, bf, k, **opts)
        if cur is None:
            return None
    return cur
