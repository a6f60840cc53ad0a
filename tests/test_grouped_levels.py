"""CPU: what comes next for the build kernel, restated in numpy and proven exact before any CUDA is
written: the level-synchronous formulation of tests/test_level_formulation.py with the occurrences of one k-mer treated
as a GROUP.

All occurrences of a k-mer share their four counters (the three extra hashes are functions of the canonical hash), so
    * only the EARLIEST alive member of a group can win T_{L+1} of those counters: one set of 4 `red.min` per group
      and level instead of one per member;
    * the level test "t > max_j T_L(x_j)" needs the four timestamps once per group: the members later than that
      maximum survive;
    * the filter insert is idempotent: once per group.
The result -- filter bits and counter bytes -- must equal the sequential oracle (fill_bfs order, src/utils.cpp:96-123);
the test also counts the random 32-byte-sector touches both formulations make, which is the number the build kernel's
roofline fraction hangs on (8 per k-mer op is SURVEY §8d's algorithmic figure)."""
import numpy as np
import pytest

from util import KS, dataset, ol

CBF = ol.CBF_COUNTERS
BF_BITS = ol.BF_BYTES * 8


def grouped_level_build(reads, thrs, k, ki):
    hs, thr = [], []
    for seq, T in zip(reads, thrs):
        _, h = ol.nthash_all(seq, k)
        hs.append(h)
        thr.append(np.full(len(h), T - 2 + ki, dtype=np.int64))
    h = np.concatenate(hs)
    thr = np.concatenate(thr)
    n = len(h)
    idx = (h % np.uint64(CBF)).astype(np.int64)
    bit = (h % np.uint64(BF_BITS)).astype(np.int64)
    _, gid = np.unique(h[:, 0], return_inverse=True)                    # group = canonical hash
    ng = int(gid.max()) + 1 if n else 0
    rep = np.full(ng, -1, dtype=np.int64)
    rep[gid[::-1]] = np.arange(n - 1, -1, -1)                           # any member: indices and bits are the group's
    gidx, gbit = idx[rep], bit[rep]
    INF = n + 1
    bf = np.zeros(BF_BITS, dtype=bool)
    counters = np.zeros(CBF, dtype=np.uint8)
    alive = np.flatnonzero(thr > 0)
    level = 0
    touches_member = touches_kernel = touches_group = 0
    touches_kernel += 4 * len(alive)                                     # round 0 of the kernel: 4 red per occurrence
    while len(alive):
        # ---- per group: its earliest alive member races for T_{level+1} ----
        first = np.full(ng, INF, dtype=np.int64)
        np.minimum.at(first, gid[alive], alive)
        g_alive = np.flatnonzero(first < INF)
        T = np.full(CBF, INF, dtype=np.int64)
        np.minimum.at(T, gidx[g_alive].ravel(), np.repeat(first[g_alive], 4))
        level += 1
        counters[T < INF] = level
        if level == 1:
            g1 = np.unique(gid[alive[thr[alive] == 1]])
            bf[gbit[g1].ravel()] = True
        # ---- per group: the four timestamps once; members later than their maximum survive ----
        M = np.full(ng, -1, dtype=np.int64)
        M[g_alive] = T[gidx[g_alive]].max(axis=1)
        ok = alive > M[gid[alive]]
        # touches: per member (plain level formulation), per member as the kernel does it (level 1 looks at one counter
        # first; the red of level L + 1 was counted with round 0 for level 1), per group
        touches_member += 8 * len(alive)
        first_passes = T[idx[alive, 0]] < alive
        touches_kernel += (len(alive) + 3 * int(first_passes.sum())) if level == 1 else 4 * len(alive)
        g_first_passes = T[gidx[g_alive, 0]] < first[g_alive]           # (a group whose earliest member fails may still hold later survivors)
        touches_group += 4 * len(g_alive)                                # the group's red into T_level
        single = np.bincount(gid[alive], minlength=ng)[g_alive] == 1     # level 1's look at one counter first: groups of one
        touches_group += (int(single.sum()) + 3 * int((g_first_passes & single).sum()) + 4 * int((~single).sum())) if level == 1 else 4 * len(g_alive)
        alive = alive[ok & (thr[alive] > level)]
        touches_kernel += 4 * len(alive)                                 # survivors race for the next level
        gi = np.unique(gid[alive[thr[alive] == level + 1]])
        bf[gbit[gi].ravel()] = True
    stats = {"ops": n, "groups": ng, "per_member": touches_member / max(n, 1), "kernel_like": touches_kernel / max(n, 1),
             "per_group": touches_group / max(n, 1)}
    return np.packbits(bf, bitorder="little"), counters, stats


def _streams(d, bsize, rnd):
    """The reads fill_bfs sees for `bsize` neighbouring contigs (the most covered ones), in mapping order, with one
    kmer_threshold per contig: the shape of a real (batch, k) stream -- a genomic k-mer occurs once per covering read."""
    import goldpolish_b200 as gp
    counts = np.bincount(d.map_contig, minlength=d.n_contigs)
    c0 = int(np.argmax(np.convolve(counts, np.ones(bsize), mode="valid")))
    reads, thrs = [], []
    rl = np.diff(d.read_off)
    for c in range(c0, c0 + bsize):
        rs = d.map_read[d.map_contig == c]
        T = gp.kmer_threshold(int(rl[rs].sum()))
        reads += [d.read(int(r)) for r in rs]
        thrs += [T] * len(rs)
    return reads, thrs


@pytest.mark.parametrize("bsize,seed,mixed_thr", [(1, 3, False), (1, 5, False), (4, 9, False), (2, 11, True)])
def test_grouped_level_formulation_is_exact_and_touches_less(bsize, seed, mixed_thr):
    d = dataset(genome_len=120000, seed=seed)
    rnd = np.random.default_rng(seed)
    reads, thrs = _streams(d, bsize, rnd)
    if mixed_thr:  # members of one group with different thresholds (targets of a batch that differ in coverage)
        thrs = [int(rnd.choice([4, 5, 6, 7])) for _ in reads]
    fs = ol.FilterSet(KS)
    for seq, T in zip(reads, thrs):
        fs.add_read(seq, T)
    for ki, k in enumerate(KS):
        bf, counters, st = grouped_level_build(reads, thrs, k, ki)
        assert np.array_equal(bf, fs.bfs[ki]), f"filter bits differ, k={k}"
        assert np.array_equal(counters, fs.cbfs[ki]), f"counter bytes differ, k={k}"
        assert st["per_group"] < st["kernel_like"]
        if True:  # (shown with pytest -s; DESIGN §7 quotes these numbers)
            print(f"\n  bsize {bsize}, {len(reads)} reads, thr {sorted(set(thrs))}, k={k}: {st['ops']} ops in {st['groups']} groups; sector touches per op: "
                  f"plain levels {st['per_member']:.1f}, as the kernel does it {st['kernel_like']:.1f}, per group {st['per_group']:.1f}", end="")
