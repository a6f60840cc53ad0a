"""Experiment (not a test): ONE contig of the config-3 data set through the edit kernel (for an ncu capture of the
latency chain that makes the strong-scaling tail).  usage: python tests/exp_one_contig.py [contig] [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import goldpolish_b200 as gp
import bench

contig = int(sys.argv[1]) if len(sys.argv) > 1 else 6051
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
args = argparse.Namespace(config=3)
w = bench.WORKLOADS[3]
d = bench.make_dataset(args, 0)
clens, rlens = np.diff(d.contig_off), np.diff(d.read_off)
pl = gp.plan_batches(clens, [d.contig_name(i) for i in range(d.n_contigs)], [d.read_name(i) for i in range(d.n_reads)],
                     d.read_phred, rlens, d.map_read, d.map_contig, bsize=w["bsize"], subsample_max_per_10kbp=w["subsample_max"])
b = contig // w["bsize"]
sh = bench.LocalShare(d, pl, [b], w["bsize"])
i = int(np.nonzero(sh.contigs == contig)[0][0])
seq = sh.contig_seq[sh.contig_off[i]:sh.contig_off[i + 1]]
with gp.Context() as ctx:
    ctx.upload_reads(sh.read_seq, sh.read_off)
    ctx.build_filters(sh.batch_entry_off, sh.entries, fetch=False)
    for _ in range(reps):
        ctx.polish(seq, np.array([0, len(seq)], dtype=np.uint64), np.zeros(1, dtype=np.uint32))
        st = ctx.stats()
        print(f"contig {contig}: len {len(seq)}, edit kernel {st['edit_kernel_ms']:.1f} ms, triggers {st['triggers']}, edits {st['edits']}, masked {st['masked']}", flush=True)
