"""Experiment (not a test): goldpolish-mask on the device vs the oracle restatement on the host, 5 Mbp of polished records."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import goldpolish_b200 as gp
import sim
from oracle import mask_oracle as mo

d = sim.simulate(genome_len=5_000_000, coverage=30.0, seed=20250607)
pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)],
                     [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off),
                     d.map_read, d.map_contig, bsize=1, subsample_max_per_10kbp=40.0)
with gp.Context() as ctx:
    ctx.upload_reads(d.read_seq, d.read_off)
    ctx.build_filters(pl.batch_entry_off, pl.entries, fetch=False)
    out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
    recs = [out[int(off[i]):int(off[i + 1])].tobytes() for i in range(d.n_contigs) if not dropped[i]]
    buf = np.frombuffer(b"".join(recs), dtype=np.uint8).copy()
    o2 = np.zeros(len(recs) + 1, dtype=np.uint64); o2[1:] = np.cumsum([len(r) for r in recs])
    ctx.prep(buf, o2, 1, 32)
    t0 = time.perf_counter(); got, goff = ctx.prep(buf, o2, 1, 32); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); want = [mo.mask(r.decode(), 32) for r in recs]; t_cpu = time.perf_counter() - t0
ok = all(got[int(goff[i]):int(goff[i + 1])].tobytes().decode() == want[i] for i in range(len(recs)))
print(f"{len(recs)} records, {len(buf) / 1e6:.2f} Mbp: gp_prep {t_gpu * 1e3:.1f} ms end to end (H2D + kernel + D2H), "
      f"python restatement of the reference script {t_cpu * 1e3:.0f} ms, identical: {ok}")
