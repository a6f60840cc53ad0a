"""CPU: the level-synchronous reformulation of the counting-Bloom-gated build (csrc/gp_build_levels.cu)
restated in numpy and checked against the sequential oracle (fill_bfs order, src/utils.cpp:96-123) -- filter bits AND
counter bytes -- on streams that are loaded enough for different k-mers to share counters.

    T_{L+1}(x) = min { t : t touches x, thr(t) > L, t > max_j T_L(x_j(t)) }
    occurrence t enters the filter  iff  t > max_j T_{thr(t)-1}(x_j(t))
"""
import numpy as np
import pytest

from util import KS, dataset, ol

CBF = ol.CBF_COUNTERS
BF_BITS = ol.BF_BYTES * 8


def level_build(reads, thrs, k, ki):
    """reads: sequences in the reference's order; thrs: kmer_threshold of each read's target."""
    hs, thr = [], []
    for seq, T in zip(reads, thrs):
        _, h = ol.nthash_all(seq, k)
        hs.append(h)
        thr.append(np.full(len(h), T - 2 + ki, dtype=np.int64))        # utils.cpp:108,121
    h = np.concatenate(hs)
    thr = np.concatenate(thr)
    n = len(h)
    idx = (h % np.uint64(CBF)).astype(np.int64)                          # n x 4 counters
    bit = (h % np.uint64(BF_BITS)).astype(np.int64)
    t = np.arange(n, dtype=np.int64)
    INF = n + 1
    bf = np.zeros(BF_BITS, dtype=bool)
    counters = np.zeros(CBF, dtype=np.uint8)
    alive = t[thr > 0]
    level = 0
    while len(alive):
        # the survivors of level `level` race for T_{level+1} of their counters: the earliest one wins
        T = np.full(CBF, INF, dtype=np.int64)
        np.minimum.at(T, idx[alive].ravel(), np.repeat(alive, 4))
        level += 1
        counters[T < INF] = level
        if level == 1:
            bf[bit[alive[thr[alive] == 1]].ravel()] = True               # count after the update reaches thr = 1
        # level test: all four counters reached `level` strictly before t
        ok = (T[idx[alive]] < alive[:, None]).all(axis=1)
        alive = alive[ok & (thr[alive] > level)]
        bf[bit[alive[thr[alive] == level + 1]].ravel()] = True           # min == thr - 1 before, thr after the update
    return np.packbits(bf, bitorder="little"), counters


@pytest.mark.parametrize("n_reads,seed", [(40, 3), (160, 5)])
def test_level_formulation_equals_sequential_semantics(n_reads, seed):
    d = dataset(genome_len=120000, seed=seed)
    rnd = np.random.default_rng(seed)
    pick = rnd.choice(d.n_reads, size=min(n_reads, d.n_reads), replace=False)
    reads = [d.read(int(i)) for i in pick]
    thrs = [int(rnd.choice([4, 5, 6, 7])) for _ in reads]               # thresholds vary per target inside a batch
    fs = ol.FilterSet(KS)
    for seq, T in zip(reads, thrs):
        fs.add_read(seq, T)
    for ki, k in enumerate(KS):
        bf, counters = level_build(reads, thrs, k, ki)
        assert np.array_equal(bf, fs.bfs[ki]), f"filter bits differ, k={k}"
        assert np.array_equal(counters, fs.cbfs[ki]), f"counter bytes differ, k={k}"
    assert int((fs.cbfs[0] > 0).sum()) > 100000                          # loaded enough for shared counters
