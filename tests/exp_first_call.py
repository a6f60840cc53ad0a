"""Experiment (not a test): the FIRST overlapped pass of a fresh context (kernels preloaded at gp_ctx_create)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, goldpolish_b200 as gp, sim
d = sim.simulate(genome_len=5_000_000, coverage=30.0, seed=20250607)
pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)], [d.read_name(i) for i in range(d.n_reads)],
                     d.read_phred, np.diff(d.read_off), d.map_read, d.map_contig, bsize=1, subsample_max_per_10kbp=40.0)
with gp.Context() as ctx:
    ctx.upload_reads(d.read_seq, d.read_off)
    ctx.build_stage(pl.batch_entry_off, pl.entries)
    ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
    for i in range(3):
        ctx.pipeline_run(); st = ctx.stats()
        print(f"pass {i}: step {st['build_ms']:.1f} ms, build kernel {st['build_kernel_ms']:.1f}, edit span {st['edit_kernel_ms']:.1f}")
