"""Experiment (not a test): the overlapped pass with the edit kernel on SMs of its own (GP_EDIT_SMS = 2, 4, ... or
unset = chosen from the staged work) against the older scheme where the two kernels share every SM (GP_EDIT_SMS=0),
and against build then polish.  Every variant's filters and polished records are compared with the first one's.
usage: python tests/exp_edit_sms.py [config: 2 | 3 | 3s (share 3 of an 8-way sharding of config 3)] [E values, comma separated]"""
import os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import goldpolish_b200 as gp
from goldpolish_b200 import shard
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "2"
es = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "auto", "2", "4", "8"]
cfg = 2 if which == "2" else 3
args = argparse.Namespace(config=cfg)
w = bench.WORKLOADS[cfg]
d = bench.make_dataset(args, 0)
clens, rlens = np.diff(d.contig_off), np.diff(d.read_off)
pl = gp.plan_batches(clens, [d.contig_name(i) for i in range(d.n_contigs)], [d.read_name(i) for i in range(d.n_reads)],
                     d.read_phred, rlens, d.map_read, d.map_contig, bsize=w["bsize"], subsample_max_per_10kbp=w["subsample_max"])
nb = len(pl.batch_entry_off) - 1
mine = list(range(nb))
if which == "3s":
    off = pl.batch_entry_off.astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(rlens[pl.entries["read_id"]])])
    mine = shard.assign_batches(((csum[off[1:]] - csum[off[:-1]]) + 1).tolist(), 8)[3]
sh = bench.LocalShare(d, pl, mine, w["bsize"])
reps = 5 if sh.draft_bases < 20_000_000 else 2
first = None
with gp.Context() as ctx:
    ctx.upload_reads(sh.read_seq, sh.read_off)
    ctx.build_stage(sh.batch_entry_off, sh.entries)
    ctx.polish_stage(sh.contig_seq, sh.contig_off, sh.contig_batch)
    known = ("GP_EDIT_SMS", "GP_EXP_NO_EDIT", "GP_EXP_PIPE_CTAS", "GP_LEVEL_CTAS", "GP_NO_OVERLAP")
    for e in ["separate"] + es:
        # a variant: "separate", "auto", a number (GP_EDIT_SMS), "n<number>" (that many SMs given away, no edit kernel),
        # or "sep|pipe" followed by :VAR=value settings
        for k in known:
            os.environ.pop(k, None)
        mode = "pipe"
        if e == "separate" or e.startswith("sep"):
            mode = "sep"
        if ":" in e:
            for kv in e.split(":")[1:]:
                k, v = kv.split("=")
                os.environ[k] = v
        elif e.startswith("n"):
            os.environ["GP_EXP_NO_EDIT"] = "1"
            os.environ["GP_EDIT_SMS"] = e[1:]
        elif e not in ("auto", "separate"):
            os.environ["GP_EDIT_SMS"] = e
        fn = (lambda: (ctx.build_run(), ctx.polish_run())) if mode == "sep" else ctx.pipeline_run
        fn()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        st = ctx.stats()
        bf = ctx.build_fetch()
        out, off_o, dropped = ctx.polish_fetch()
        dig = hashlib.sha256(bf.tobytes() + out[:int(off_o[-1])].tobytes() + off_o.tobytes() + dropped.tobytes()).hexdigest()[:16]
        first = first or dig
        print(f"{which} {e:>60s}: step {ms:8.1f} ms, build span {st['build_kernel_ms']:8.1f}, edit span {st['edit_kernel_ms']:8.1f}, "
              f"edit_sms {st['edit_sms']}, reruns {st['polish_reruns']}, {st['kmer_ops'] / max(st['build_kernel_ms'], 1e-3) / 1e6:.2f} G ops/s, "
              f"digest {dig} {'same' if dig == first else 'DIFFERENT'}", flush=True)
