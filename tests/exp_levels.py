"""Experiment (not a test): level-synchronous build kernel under different schedules / L2 policies.
usage: python tests/exp_levels.py [genome_len] [bsize] -- variants come from the environment variable VARIANTS,
a ';'-separated list of 'NAME=VALUE,NAME=VALUE' settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import goldpolish_b200 as gp
import sim

genome = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
bsize = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = sim.simulate(genome_len=genome, coverage=30.0, seed=20250607)
pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)],
                     [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off),
                     d.map_read, d.map_contig, bsize=bsize, subsample_max_per_10kbp=40.0)
ref = None
for variant in os.environ.get("VARIANTS", "").split(";"):
    sets = dict(kv.split("=") for kv in variant.split(",") if kv)
    for k, v in sets.items():
        os.environ[k] = v
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        for _ in range(3):
            ctx.build_run()
        st = ctx.stats()
        rt = ctx.build_round_times()
        bf = ctx.build_fetch()
        ct = ctx.build_cta_times() if os.environ.get("GP_LEVEL_CTA_TIMES") and st["build_kernel"] == 2 else None
    for k in sets:
        os.environ.pop(k, None)
    same = "first" if ref is None else ("same bits" if np.array_equal(ref, bf) else "BITS DIFFER")
    if ref is None:
        ref = bf
    print(f"[{variant or 'default'}] bsize {bsize}: build kernel {st['build_kernel_ms']:.2f} ms, "
          f"{st['kmer_ops'] / st['build_kernel_ms'] / 1e6:.2f} G ops/s ({same})")
    rt.pop("list_entries")
    for k, (w, r, n) in rt.items():
        if n > 20:
            print(f"      {k:13s} n {n:6d}  wait {1e3 * w / n:6.2f} us  work {1e3 * r / n:6.2f} us")
    if ct is not None:
        names = gp.Context.INTERVAL_KINDS
        for i, nm in enumerate(names):
            n = ct[:, i, 2].astype(np.float64)
            if n.max() < 20:
                continue
            wait = ct[:, i, 0] / np.maximum(n, 1) / 1e3
            work = ct[:, i, 1] / np.maximum(n, 1) / 1e3
            q = lambda a: " ".join(f"{np.percentile(a, p):6.2f}" for p in (0, 10, 50, 90, 100))
            print(f"      per-CTA {nm:13s} wait us [min p10 p50 p90 max] {q(wait)} | work {q(work)}")
