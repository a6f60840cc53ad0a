"""Experiment (not a test): speed of a lone stream / lone contig."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sim, goldpolish_b200 as gp
d = sim.simulate(genome_len=50000, contig_min=50000, contig_max=50000, contig_median=50000, tiny_contig_frac=0.0, seed=5)
pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)], [d.read_name(i) for i in range(d.n_reads)],
                     d.read_phred, np.diff(d.read_off), d.map_read, d.map_contig, bsize=1)
ctx = gp.Context(); ctx.upload_reads(d.read_seq, d.read_off)
for it in range(3):
    ctx.build_filters(pl.batch_entry_off, pl.entries, fetch=False)
    out = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
st = ctx.stats()
steps = st["kmer_ops"] / 4 / 32
print(f"contigs {d.n_contigs} entries {len(pl.entries)} kmer_ops {st['kmer_ops']} build_kernel_ms {st['build_kernel_ms']:.2f} -> {st['build_kernel_ms']*1e3/steps:.2f} us/step ({st['build_kernel_ms']*1e-3*1.965e9/steps:.0f} cycles); edit_kernel_ms {st['edit_kernel_ms']:.2f} triggers {st['triggers']} edits {st['edits']}")
