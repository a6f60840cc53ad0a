"""Experiment (not a test): the build kernel alone (gp_build_run) at config 2 / config 3-share size, several repeats --
for A/B runs of two builds of the library (GP_LIB_PATH).  usage: python tests/exp_build_alone.py [2|3s] [repeats]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import goldpolish_b200 as gp
from goldpolish_b200 import shard
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = 2 if which == "2" else 3
w = bench.WORKLOADS[cfg]
d = bench.make_dataset(argparse.Namespace(config=cfg), 0)
clens, rlens = np.diff(d.contig_off), np.diff(d.read_off)
pl = gp.plan_batches(clens, [d.contig_name(i) for i in range(d.n_contigs)], [d.read_name(i) for i in range(d.n_reads)],
                     d.read_phred, rlens, d.map_read, d.map_contig, bsize=w["bsize"], subsample_max_per_10kbp=w["subsample_max"])
mine = list(range(len(pl.batch_entry_off) - 1))
if which == "3s":
    off = pl.batch_entry_off.astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(rlens[pl.entries["read_id"]])])
    mine = shard.assign_batches(((csum[off[1:]] - csum[off[:-1]]) + 1).tolist(), 8)[3]
sh = bench.LocalShare(d, pl, mine, w["bsize"])
with gp.Context() as ctx:
    ctx.upload_reads(sh.read_seq, sh.read_off)
    ctx.build_stage(sh.batch_entry_off, sh.entries)
    ms = []
    for i in range(reps + 3):
        ctx.build_run()
        ctx.synchronize()
        if i >= 3:
            ms.append(ctx.stats()["build_kernel_ms"])
    st = ctx.stats()
    print(f"{os.environ.get('GP_LIB_PATH', 'in-tree')}: build kernel alone {np.mean(ms):.2f} ms (min {min(ms):.2f}, max {max(ms):.2f}), "
          f"{st['kmer_ops'] / np.mean(ms) / 1e6:.2f} G ops/s", flush=True)
