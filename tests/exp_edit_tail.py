"""Experiment (not a test): which contigs make the edit tail of a share of the 8-way config-3 sharding.
usage: python tests/exp_edit_tail.py [share] [world]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import goldpolish_b200 as gp
from goldpolish_b200 import shard
import bench

share = int(sys.argv[1]) if len(sys.argv) > 1 else 3
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
args = argparse.Namespace(config=3)
w = bench.WORKLOADS[3]
d = bench.make_dataset(args, 0)
clens, rlens = np.diff(d.contig_off), np.diff(d.read_off)
pl = gp.plan_batches(clens, [d.contig_name(i) for i in range(d.n_contigs)], [d.read_name(i) for i in range(d.n_reads)],
                     d.read_phred, rlens, d.map_read, d.map_contig, bsize=w["bsize"], subsample_max_per_10kbp=w["subsample_max"])
off = pl.batch_entry_off.astype(np.int64)
csum = np.concatenate([[0], np.cumsum(rlens[pl.entries["read_id"]])])
work = (csum[off[1:]] - csum[off[:-1]]) + 1
mine = shard.assign_batches(work.tolist(), world)[share]
sh = bench.LocalShare(d, pl, mine, w["bsize"])
with gp.Context() as ctx:
    ctx.upload_reads(sh.read_seq, sh.read_off)
    bfs = ctx.build_filters(sh.batch_entry_off, sh.entries, fetch=False)
    out, o, dr = ctx.polish(sh.contig_seq, sh.contig_off, sh.contig_batch)
    st = ctx.stats()
    print(f"share {share}: all contigs: edit kernel {st['edit_kernel_ms']:.1f} ms, triggers {st['triggers']}, edits {st['edits']}, masked {st['masked']}, rollbacks {st['rollbacks']}")
    lens = np.diff(sh.contig_off)
    res = []
    for c in np.argsort(-lens)[:40]:
        seq = sh.contig_seq[sh.contig_off[c]:sh.contig_off[c + 1]]
        ctx.polish(seq, np.array([0, len(seq)], dtype=np.uint64), sh.contig_batch[c:c + 1])
        s1 = ctx.stats()
        res.append((s1["edit_kernel_ms"], int(lens[c]), s1["triggers"], s1["edits"], s1["masked"], s1["rollbacks"], int(sh.contigs[c])))
    for r in sorted(res, reverse=True)[:12]:
        print("  contig %d: %.1f ms alone, len %d, triggers %d, edits %d, masked %d, rollbacks %d" % (r[6], r[0], r[1], r[2], r[3], r[4], r[5]))
