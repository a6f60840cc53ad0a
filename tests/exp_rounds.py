"""Experiment (not a test): per-round time breakdown of the level-synchronous build kernel.
usage: python tests/exp_rounds.py [genome_len] [bsize]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import goldpolish_b200 as gp
import sim

genome = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
bsize = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = sim.simulate(genome_len=genome, coverage=30.0, seed=20250607)
pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)],
                     [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off),
                     d.map_read, d.map_contig, bsize=bsize, subsample_max_per_10kbp=40.0)
for slots in os.environ.get("OVERLAP", "0 1").split():
    os.environ["GP_LEVEL_OVERLAP"] = slots
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        for _ in range(3):
            ctx.build_run()
        st = ctx.stats()
        rt = ctx.build_round_times()
    n_streams = (len(pl.batch_entry_off) - 1) * 4
    print(f"overlap {slots}: build kernel {st['build_kernel_ms']:.2f} ms, {st['kmer_ops'] / st['build_kernel_ms'] / 1e6:.2f} G ops/s, "
          f"{n_streams} streams, {st['kmer_ops'] / n_streams / 1e3:.0f} k ops/stream")
    le = rt.pop("list_entries")
    print(f"   list entries visited by list rounds: {le} = {le / st['kmer_ops']:.3f} per k-mer op, "
          f"{le / max(sum(v[2] for k, v in rt.items() if k not in ('clear', 'round0')), 1) / 1e3:.0f} k per round")
    for k, (w, r, n) in rt.items():
        if n:
            print(f"   {k:13s} rounds {n:6d}  wait {w:8.2f} ms ({1e3 * w / n:6.2f} us/round)  work {r:8.2f} ms ({1e3 * r / n:6.2f} us/round)")
