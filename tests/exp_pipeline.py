"""Experiment (not a test): separate build_run + polish_run vs the overlapped gp_pipeline_run.
usage: python tests/exp_pipeline.py [genome_len] [bsize]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import goldpolish_b200 as gp
import sim

genome = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
bsize = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = sim.simulate(genome_len=genome, coverage=30.0, seed=20250607)
pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)],
                     [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off),
                     d.map_read, d.map_contig, bsize=bsize, subsample_max_per_10kbp=40.0)
with gp.Context() as ctx:
    ctx.upload_reads(d.read_seq, d.read_off)
    ctx.build_stage(pl.batch_entry_off, pl.entries)
    ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
    for name, fn in (("separate", lambda: (ctx.build_run(), ctx.polish_run())), ("pipeline", ctx.pipeline_run)):
        for _ in range(2):
            fn()
        ctx.stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        st = ctx.stats()
        dt = (time.perf_counter() - t0) / 3
        print(f"{name}: {dt * 1e3:.1f} ms/step  build {st['build_ms']:.1f} (kernel {st['build_kernel_ms']:.1f})  "
              f"polish {st['polish_ms']:.1f} (edit kernel {st['edit_kernel_ms']:.1f})  edits {st['edits']} masked {st['masked']}")
