"""Experiment (not a test): the config-3 shares of an N-way strong-scaling run, one after the other on ONE GPU.
Shows what LPT on read bases leaves unbalanced (per-share step time, k-mer ops, edit tail).
usage: python tests/exp_shares.py [world] [genome_len]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
import goldpolish_b200 as gp
from goldpolish_b200 import shard
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
args = argparse.Namespace(config=3)
if len(sys.argv) > 2:
    bench.WORKLOADS[3]["genome_len"] = int(sys.argv[2])
w = bench.WORKLOADS[3]
d = bench.make_dataset(args, 0)
clens, rlens = np.diff(d.contig_off), np.diff(d.read_off)
pl = gp.plan_batches(clens, [d.contig_name(i) for i in range(d.n_contigs)], [d.read_name(i) for i in range(d.n_reads)],
                     d.read_phred, rlens, d.map_read, d.map_contig, bsize=w["bsize"], subsample_max_per_10kbp=w["subsample_max"])
off = pl.batch_entry_off.astype(np.int64)
ent_bases = rlens[pl.entries["read_id"]]
csum = np.concatenate([[0], np.cumsum(ent_bases)])
work = (csum[off[1:]] - csum[off[:-1]]) + 1
mode = os.environ.get("SHARD_WEIGHT", "bases")
if mode == "levels":  # weight a batch's read bases by the list rounds its streams need
    maxthr = np.array([pl.entries["kmer_threshold"][off[b]:off[b + 1]].max() if off[b + 1] > off[b] else 4 for b in range(len(off) - 1)])
    work = work * (2 + maxthr)
assignment = shard.assign_batches(work.tolist(), world)
only = os.environ.get("ONLY_SHARE")
for r, mine in enumerate(assignment):
    if only is not None and r != int(only):
        continue
    sh = bench.LocalShare(d, pl, mine, w["bsize"])
    with gp.Context() as ctx:
        ctx.upload_reads(sh.read_seq, sh.read_off)
        ctx.build_stage(sh.batch_entry_off, sh.entries)
        ctx.polish_stage(sh.contig_seq, sh.contig_off, sh.contig_batch)
        for _ in range(2):
            ctx.pipeline_run()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.pipeline_run()
        ctx.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        st = ctx.stats()
    longest = int(np.diff(sh.contig_off).max())
    print(f"share {r}: {len(mine)} batches, {sh.draft_bases} bp, read bases {int(ent_bases[np.concatenate([np.arange(off[b], off[b+1]) for b in mine])].sum())}, "
          f"step {ms:.1f} ms, build span {st['build_kernel_ms']:.1f} ms, edit span {st['edit_kernel_ms']:.1f} ms, "
          f"{st['kmer_ops'] / 1e9:.3f} G ops ({st['kmer_ops'] / st['build_kernel_ms'] / 1e6:.2f} G/s), longest contig {longest}", flush=True)
