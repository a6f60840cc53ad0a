"""CPU: the sharding bookkeeping bench.py uses for the strong-scaling run (no GPU, no collective): every batch lands on
exactly one rank, a rank's share is self-contained (entries re-indexed into its own read store, contigs with their
batches), and the guard rule applied per batch is the C ABI's."""
import numpy as np

from util import dataset, plan

import bench
import goldpolish_b200 as gp
from goldpolish_b200 import shard


def test_shares_partition_the_workload_and_are_self_contained():
    d = dataset(genome_len=200000)
    bs = 3
    pl = plan(d, bsize=bs)
    nb = len(pl.batch_entry_off) - 1
    rl = np.diff(d.read_off)
    off = pl.batch_entry_off.astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(rl[pl.entries["read_id"]])])
    work = cs[off[1:]] - cs[off[:-1]] + 1
    world = 3
    assignment = shard.assign_batches(work.tolist(), world)
    assert sorted(b for a in assignment for b in a) == list(range(nb))
    loads = [int(sum(work[b] for b in a)) for a in assignment]
    assert max(loads) - min(loads) <= int(work.max())  # LPT: within one batch of each other
    seen_contigs = []
    for mine in assignment:
        sh = bench.LocalShare(d, pl, mine, bs)
        assert sh.batches.tolist() == sorted(mine)
        seen_contigs += sh.contigs.tolist()
        # entries name the same reads, through the local store
        k = 0
        for i, b in enumerate(mine):
            ents = pl.entries[off[b]:off[b + 1]]
            loc = sh.entries[int(sh.batch_entry_off[i]):int(sh.batch_entry_off[i + 1])]
            assert len(ents) == len(loc) and np.array_equal(ents["kmer_threshold"], loc["kmer_threshold"])
            for e, l in zip(ents[:3], loc[:3]):
                assert sh.read_seq[sh.read_off[l["read_id"]]:sh.read_off[l["read_id"] + 1]].tobytes() == d.read(int(e["read_id"]))
            k += len(ents)
        for i, c in enumerate(sh.contigs.tolist()):
            assert sh.contig_seq[sh.contig_off[i]:sh.contig_off[i + 1]].tobytes() == d.contig(c)
            assert sh.batches[sh.contig_batch[i]] == c // bs
    assert sorted(seen_contigs) == list(range(d.n_contigs))
    whole = bench.LocalShare(d, pl, list(range(nb)), bs)  # the identity share aliases the data set
    assert whole.read_seq is d.read_seq and whole.entries is pl.entries


def test_guard_per_batch_is_the_c_abi_rule():
    d = dataset(genome_len=60000)
    bs = 2
    pl = plan(d, bsize=bs)
    sh = bench.LocalShare(d, pl, list(range(len(pl.batch_entry_off) - 1)), bs)
    n = len(sh.contigs)
    lens = np.diff(sh.contig_off)
    # "polished" records: batch 0 shrinks below 75 % (rejected: keeps its originals), batch 1 loses a record (dropped)
    out_len = lens.copy()
    out_len[0:2] = lens[0:2] // 2
    dropped = np.zeros(n, dtype=np.uint8)
    dropped[2] = 1
    out_len[2] = 0
    off = np.concatenate([[0], np.cumsum(out_len)]).astype(np.uint64)
    out = np.concatenate([sh.contig_seq[sh.contig_off[i]:sh.contig_off[i] + out_len[i]] for i in range(n)])
    new_out, new_off, new_dropped, n_rej = bench.apply_guard(sh, out, off, dropped, gp.guard_rejects)
    in0 = sum(sh.name_len[i] + 3 + lens[i] for i in (0, 1))
    out0 = sum(sh.name_len[i] + 3 + out_len[i] for i in (0, 1))
    assert gp.guard_rejects(int(in0), int(out0))
    rej1 = gp.guard_rejects(int(sum(sh.name_len[i] + 3 + lens[i] for i in (2, 3))), int(sh.name_len[3] + 3 + lens[3]))
    assert n_rej == 1 + int(rej1)
    for i in (0, 1):
        assert new_out[int(new_off[i]):int(new_off[i + 1])].tobytes() == d.contig(i) and not new_dropped[i]
    if not rej1:
        assert new_dropped[2] == 1 and new_off[3] == new_off[2]
    assert new_out[int(new_off[4]):int(new_off[5])].tobytes() == d.contig(4)
