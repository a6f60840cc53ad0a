#!/usr/bin/env python
"""bench.py -- polished Mbp/s of the GoldPolish hot path (filter build + 4 ntEdit rounds + guard).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch set of synthetic input: for every batch
of the workload, build the 4 (counting filter, filter) pairs from its mapped reads and run the
k=32,28,24,20 ntEdit chain over its contigs.  The N=1 workload is BASELINE.json configs[1]: a
synthetic 5 Mbp draft cut into golden-path-sized contigs, 30x simulated ONT reads (5 % error),
PAF-style mappings (subsample cap 40 per 10 kbp), bsize 1.  With N GPUs every rank polishes its
own 5 Mbp shard (independent batches, no collective in the data path): weak scaling.

`value`   : device-resident inputs (packed reads, staged contigs) -> kernels only.
`e2e`     : through the C ABI with HOST buffers: H2D of reads + packing, build, D2H of the
            filter payloads, H2D of contigs, polish, D2H of the polished sequences.
`--impl reference` : the reference's own sources (oracle/_ref, compiled from /root/reference)
            on the host cores, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "polished_mbp_per_s"
UNIT = "Mbp/s"
WORKLOAD = dict(workload="configs[1]: synthetic 5 Mbp draft (contigs ~lognormal median 7 kbp) + 30x simulated ONT reads "
                         "(5% error), PAF mappings s=40, bsize 1, k=32,28,24,20",
                genome_len=5_000_000, coverage=30.0, bsize=1, subsample_max=40.0, ks=[32, 28, 24, 20])
ALGO_BYTES_PER_KMER_OP = 256  # 8 random 32-byte sector touches, SURVEY.md §8(d)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_JSON_FD = None


def own_stdout():
    """Keep stdout for the ONE JSON line: anything a library prints there (NCCL announces its version on stdout)
    goes to stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def make_dataset(rank: int, genome_len: int):
    import sim
    return sim.simulate(genome_len=genome_len, coverage=WORKLOAD["coverage"], seed=20250607 + 1000003 * rank)


def make_plan(d):
    import goldpolish_b200 as gp
    clens = np.diff(d.contig_off)
    rlens = np.diff(d.read_off)
    return gp.plan_batches(clens, [d.contig_name(i) for i in range(d.n_contigs)],
                           [d.read_name(i) for i in range(d.n_reads)], d.read_phred, rlens,
                           d.map_read, d.map_contig, bsize=WORKLOAD["bsize"],
                           subsample_max_per_10kbp=WORKLOAD["subsample_max"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, device_index: int):
        self.path = tempfile.mktemp(prefix="gp_clocks_", suffix=".csv")
        self.proc = None
        self.idx = device_index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kmer_ops_per_launch):
    """DRAM bytes per launch of the build kernel: the committed `ncu --set full` capture
    (profiles/build_kernel_dram.json, taken on the 1 Mbp profiling workload) gives bytes per k-mer
    op; scaled to the k-mer ops of this launch.  None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "build_kernel_dram.json")
    if os.path.exists(p):
        try:
            per_op = json.load(open(p)).get("dram_bytes_per_kmer_op")
            return None if per_op is None else per_op * kmer_ops_per_launch
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own sources on the host cores
# ----------------------------------------------------------------------------------------------
def run_reference_sample(d, pl, n_sample_batches: int, threads: int, workdir: str):
    """Times the reference's serve_batch (filter build) and ntEdit chain + guard on the first
    n_sample_batches batches.  Returns dict(seconds_build, seconds_edit, bases, kind, cores)."""
    import ctypes as C

    from oracle import ref_driver as rd
    nb = min(n_sample_batches, len(pl.batch_entry_off) - 1)
    bs = WORKLOAD["bsize"]
    contigs = list(range(0, min(nb * bs, d.n_contigs)))
    bases = int(sum(len(d.contig(c)) for c in contigs))
    if rd.ref_available():
        h = rd.harness()
        draft, reads, paf = (os.path.join(workdir, f) for f in ("draft.fa", "reads.fq" if d.fastq else "reads.fa", "mappings.paf"))
        if not os.path.exists(draft + ".index"):
            rd.run_index(draft, draft + ".index")
            rd.run_index(reads, reads + ".index")
        bdir = os.path.join(workdir, "bfs")
        shutil.rmtree(bdir, ignore_errors=True)
        os.makedirs(bdir)
        names, ids_files = [], []
        for b in range(nb):
            names.append(str(b).encode())
            p = os.path.join(bdir, f"{b}.ids")
            with open(p, "w") as f:
                for c in range(b * bs, min((b + 1) * bs, d.n_contigs)):
                    f.write(d.contig_name(c) + "\n")
            ids_files.append(p.encode())
            bd = os.path.join(workdir, f"batch{b}")
            os.makedirs(bd, exist_ok=True)
            with open(os.path.join(bd, "batch.fa"), "w") as f:
                for c in range(b * bs, min((b + 1) * bs, d.n_contigs)):
                    f.write(f">{d.contig_name(c)}\n{d.contig(c).decode()}\n")
        ks = (C.c_uint * 4)(*WORKLOAD["ks"])
        cwd = os.getcwd()
        os.chdir(bdir)
        os.environ["GP_ORACLE_QUIET"] = "1"
        try:
            t_build = h.ref_serve_batches(draft.encode(), (draft + ".index").encode(), paf.encode(), reads.encode(),
                                          (reads + ".index").encode(), 150.0, WORKLOAD["subsample_max"], threads, ks, 4,
                                          (C.c_char_p * nb)(*names), (C.c_char_p * nb)(*ids_files), nb)
        finally:
            os.chdir(cwd)
        if t_build < 0:
            raise RuntimeError("reference serve_batches failed")
        bases_arr = (C.c_char_p * nb)(*[os.path.join(workdir, f"batch{b}", "batch").encode() for b in range(nb)])
        bfs_flat = (C.c_char_p * (nb * 4))(*[os.path.join(bdir, f"{b}-k{k}.bf").encode() for b in range(nb) for k in WORKLOAD["ks"]])
        outs = (C.c_char_p * nb)(*[os.path.join(workdir, f"batch{b}", "batch.ntedited.fa").encode() for b in range(nb)])
        t_edit = h.ref_ntedit_chain_many(bases_arr, bfs_flat, ks, 4, outs, nb, threads)
        if t_edit < 0:
            raise RuntimeError("reference ntedit chain failed")
        return dict(seconds_build=t_build, seconds_edit=t_edit, bases=bases, kind="reference", cores=threads,
                    batches=nb)
    # oracle port, one thread
    from oracle import oracle_lib as ol
    t0 = time.perf_counter()
    fsets = {}
    for b in range(nb):
        fs = ol.FilterSet()
        for e in range(int(pl.batch_entry_off[b]), int(pl.batch_entry_off[b + 1])):
            fs.add_read(d.read(int(pl.entries[e]["read_id"])), int(pl.entries[e]["kmer_threshold"]))
        fsets[b] = fs
    t1 = time.perf_counter()
    for c in contigs:
        cur = d.contig(c)
        for ki, k in enumerate(WORKLOAD["ks"]):
            cur, _ = ol.ntedit_contig(cur, fsets[int(pl.contig_batch[c])].bfs[ki], k)
            if cur is None:
                break
    t2 = time.perf_counter()
    return dict(seconds_build=t1 - t0, seconds_edit=t2 - t1, bases=bases, kind="port", cores=1, batches=nb)


def write_files(d, workdir):
    import ctypes as C

    import sim
    # regenerate through the simulator's own writer (same seed -> same data) to get the files
    p = dict(d.params)
    sim.simulate(write_dir=workdir, **{k: v for k, v in p.items()})


def reference_arm(args, rank, world):
    if rank != 0:
        return
    genome = args.genome_len
    d = make_dataset(0, genome)
    pl = make_plan(d)
    threads = os.cpu_count() or 1
    work = tempfile.mkdtemp(prefix="gp_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        write_files(d, work)
        nb_total = len(pl.batch_entry_off) - 1
        # bounded sample: ~3 batches per core per step keeps a K-step run within minutes
        nb = min(nb_total, max(threads * args.ref_batches_per_core, 8))
        times = []
        res = None
        for it in range(args.warmup + args.steps):
            res = run_reference_sample(d, pl, nb, threads, work)
            if it >= args.warmup:
                times.append(res["seconds_build"] + res["seconds_edit"])
        t = float(np.mean(times))
        val = res["bases"] / 1e6 / t
        line = {
            "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u64 integer", "data": "synthetic (seeded simulator, sim/gpsim.c)", "impl": "reference",
            "config": dict(WORKLOAD, genome_len=genome, sample=f"first {res['batches']} of {nb_total} batches per step"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                             "sample": f"first {res['batches']} of {nb_total} batches ({res['bases']} draft bases), "
                                       f"build {res['seconds_build']:.2f}s + edit {res['seconds_edit']:.2f}s"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
    finally:
        shutil.rmtree(work, ignore_errors=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def ours(args, rank, world, local_rank):
    import torch

    import goldpolish_b200 as gp
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; goldpolish_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    genome = args.genome_len
    t0 = time.time()
    d = make_dataset(rank, genome)
    pl = make_plan(d)
    n_batches = len(pl.batch_entry_off) - 1
    draft_bases = int(d.contig_off[-1])
    log(f"[rank {rank}] data: {d.n_contigs} contigs / {draft_bases} bp, {d.n_reads} reads / {int(d.read_off[-1])} bp, "
        f"{n_batches} batches, {len(pl.entries)} read entries ({time.time() - t0:.1f}s)")

    # pinned host buffers (e2e copies come from / go to these)
    reads_h = torch.from_numpy(d.read_seq).pin_memory()
    contigs_h = torch.from_numpy(d.contig_seq).pin_memory()
    bf_h = torch.empty((n_batches, 4, gp.BF_BYTES), dtype=torch.uint8).pin_memory()
    out_h = torch.empty(draft_bases + draft_bases // 4 + 65536, dtype=torch.uint8).pin_memory()

    ctx = gp.Context(device=local_rank)
    # a non-default torch stream: the library's kernels are launched on it so that the
    # torch.cuda.Event pair below brackets exactly the timed work
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_reads(reads_h, d.read_off)
    ctx.build_stage(pl.batch_entry_off, pl.entries)
    ctx.polish_stage(contigs_h, d.contig_off, pl.contig_batch)

    # one step = filter build of every batch + the 4-round ntEdit chain over every contig, as one overlapped
    # pass (gp_pipeline_run: the edit kernel starts on a contig when its batch's filters are final); with
    # --separate the two stages run one after the other (gp_build_run, gp_polish_run)
    def step_resident():
        if args.separate:
            ctx.build_run()
            ctx.polish_run()
        else:
            ctx.pipeline_run()

    def step_e2e():
        ctx.upload_reads(reads_h, d.read_off)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        ctx.polish_stage(contigs_h, d.contig_off, pl.contig_batch)
        step_resident()
        ctx.build_fetch(out=bf_h)
        return ctx.polish_fetch(out=out_h)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    build_kernel_ms = 0.0
    edit_kernel_ms = 0.0
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    st = ctx.stats()  # of the last step
    build_kernel_ms_overlapped = st["build_kernel_ms"]  # device timer when overlapped with the edit kernel
    edit_kernel_ms = st["edit_kernel_ms"]
    # roofline leg: the dominant (build) kernel alone, CUDA events on the launching stream around each launch
    ctx.build_run()
    alone = []
    for _ in range(max(1, min(args.steps, 3))):
        ctx.build_run()
        alone.append(ctx.stats()["build_kernel_ms"])
    build_kernel_ms = float(np.mean(alone))
    ms_step = ms_total / args.steps
    if dist is not None:
        t = torch.tensor([ms_step], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    launches_per_step = st["build_launches"] + st["polish_launches"]
    total_bases = draft_bases
    if dist is not None:
        tb = torch.tensor([draft_bases], device="cuda", dtype=torch.int64)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        total_bases = int(tb.item())

    # ---- end to end through the C ABI with host buffers ----
    # the pinned destination of the filter payloads is named once: the build kernel writes each filter there as
    # it becomes final (gp_build_output_host), build_fetch then only synchronises
    ctx.build_output(bf_h)
    step_e2e()
    barrier()
    t_e2e0 = time.perf_counter()
    for _ in range(args.steps):
        out, off, dropped = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t_e2e0) / args.steps
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    st2 = ctx.stats()
    h2d = int(d.read_off[-1]) + (d.n_reads + 1) * 8 * 2 + d.n_reads * 4 + (n_batches + 1) * 8 + len(pl.entries) * 8 \
        + n_batches * 4 * 4 + draft_bases + (d.n_contigs + 1) * 8 * 3 + d.n_contigs * 8
    d2h = n_batches * 4 * gp.BF_BYTES + int(off[-1]) + d.n_contigs * 5 + 4

    # guard (scripts/goldpolish-ntedit:31-40) on the fetched result: host rule, counted for info
    rejected = 0
    bs = WORKLOAD["bsize"]
    for b in range(n_batches):
        cs = range(b * bs, min((b + 1) * bs, d.n_contigs))
        in_sz = sum(len(d.contig_name(c)) + 3 + int(d.contig_off[c + 1] - d.contig_off[c]) for c in cs)
        out_sz = sum(len(d.contig_name(c)) + 3 + int(off[c + 1] - off[c]) for c in cs if not dropped[c])
        rejected += int(gp.guard_rejects(in_sz, out_sz))

    if rank == 0:
        peaks, peak_src = measured_peaks()
        kops = st["kmer_ops"]
        achieved = kops * ALGO_BYTES_PER_KMER_OP / (build_kernel_ms * 1e-3) / 1e9 if build_kernel_ms > 0 else 0.0
        kname = {1: "gp::build_filters_kernel (one warp per stream, counters in HBM)",
                 2: "gp::build_filters_levels_kernel (level-synchronous rounds, timestamps in L2)"}.get(st["build_kernel"], "?")
        roof = {"bound": "hbm", "achieved": achieved, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                "frac": achieved / float(peaks["hbm_gbs"]), "traffic": ncu_traffic(kops),
                "traffic_source": "profiles/build_kernel_dram.json (ncu dram__bytes_read+write per k-mer op at 1 Mbp) x k-mer ops of this launch",
                "kernel": kname, "kernel_ms": build_kernel_ms, "kmer_ops_per_launch": kops,
                "algorithmic_bytes_per_kmer_op": ALGO_BYTES_PER_KMER_OP, "peak_source": peak_src,
                "kmer_ops_per_s": kops / (build_kernel_ms * 1e-3) if build_kernel_ms > 0 else 0.0,
                "build_slots": st["build_slots"], "kernel_ms_inside_step": build_kernel_ms_overlapped,
                "edit_kernel_ms": edit_kernel_ms,
                "overlap": "edit kernel runs beside the build kernel (its span includes waiting for filters)" if not args.separate
                           else "none (--separate)"}
        if args.roof:
            # measured random-access roofs (sector touches / s), same access shapes as the kernels:
            #   hbm : private 10 MiB counter regions per warp (the one-warp-per-stream kernel's mix)
            #   l2_ld / l2_red : random 4-byte loads / atomicMin over one shared 40 MiB array
            # a k-mer op needs at least 4 loads + 4 atomics -> ops/s roof = 1 / (4/ld + 4/red)
            try:
                r = {}
                for name, mode, region in (("hbm", "0", gp.CBF_BYTES), ("l2_ld", "5", 4096), ("l2_red", "4", 4096)):
                    os.environ["GP_ROOF_MODE"] = mode
                    r[name], _ = ctx.roof_microbench(148 * 24, 4000 if mode != "0" else 2000, region)
                os.environ.pop("GP_ROOF_MODE", None)
                roof["random_access_roof_sectors_per_s"] = r
                if st["build_kernel"] == 2:
                    ops_roof = 1.0 / (4.0 / r["l2_ld"] + 4.0 / r["l2_red"])
                else:
                    ops_roof = r["hbm"] / 8.0
                roof["random_access_roof_kmer_ops_per_s"] = ops_roof
                roof["frac_of_random_access_roof"] = roof["kmer_ops_per_s"] / ops_roof if ops_roof else None
                # against the HBM random-access roof the north star names (8 touches per op at the HBM sector rate):
                # above 1 because the level-synchronous kernel keeps the touched state in L2
                roof["frac_of_hbm_random_access_roof"] = roof["kmer_ops_per_s"] * 8.0 / r["hbm"] if r["hbm"] else None
            except Exception as e:  # measurement aid only
                roof["random_access_roof_error"] = str(e)
        cpu = None
        if args.cpu_baseline:
            work = tempfile.mkdtemp(prefix="gp_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            try:
                threads = os.cpu_count() or 1
                from oracle import ref_driver as rd
                if rd.ref_available():
                    write_files(d, work)
                nb = min(n_batches, max(threads * args.ref_batches_per_core, 8))
                r = run_reference_sample(d, pl, nb, threads, work)
                tt = r["seconds_build"] + r["seconds_edit"]
                cpu = {"value": r["bases"] / 1e6 / tt, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                       "sample": f"first {r['batches']} of {n_batches} batches ({r['bases']} draft bases): "
                                 f"build {r['seconds_build']:.2f}s + edit {r['seconds_edit']:.2f}s"}
            except Exception as e:
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "error", "sample": str(e)}
            finally:
                shutil.rmtree(work, ignore_errors=True)
        line = {
            "metric": METRIC, "value": total_bases / 1e6 / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 integer",
            "data": "synthetic (seeded simulator, sim/gpsim.c)",
            "config": dict(WORKLOAD, genome_len=genome, per_gpu_draft_bases=draft_bases, batches_per_gpu=n_batches,
                           l2="inputs larger than L2: every step reads %.0f MB of packed reads + step anchors and writes %.2f GB of filters "
                              "(126 MB L2); the kernel's own 80 MiB of timestamps are L2-resident by design"
                              % ((int(d.read_off[-1]) * 0.375 + int(d.read_off[-1]) / 32 * 4 * 2) / 1e6, n_batches * 4 * gp.BF_BYTES / 1e9),
                           guard_rejected_batches=rejected, parallelism=f"batches sharded over {world} GPU(s), no collective"),
            "clocks": clocks,
            "e2e": {"value": total_bases / 1e6 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "includes": "H2D reads+pack, H2D contigs, build + polish, D2H filter payloads, D2H polished"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
            "stats": {k: st2[k] for k in ("kmer_ops", "serial_kmers", "triggers", "edits", "masked", "rollbacks",
                                          "build_ms", "polish_ms", "pack_ms", "build_kernel_ms", "edit_kernel_ms")},
        }
        emit(line)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-len", type=int, default=WORKLOAD["genome_len"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-roof", dest="roof", action="store_false")
    ap.add_argument("--ref-batches-per-core", type=int, default=2)
    ap.add_argument("--bsize", type=int, default=WORKLOAD["bsize"], help="contigs per batch (parity/scale experiments)")
    ap.add_argument("--coverage", type=float, default=WORKLOAD["coverage"])
    ap.add_argument("--separate", action="store_true", help="build, then polish (no overlap of the two kernels)")
    args = ap.parse_args()
    WORKLOAD["bsize"] = args.bsize
    WORKLOAD["coverage"] = args.coverage
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        log("bench.py: note: timing rules ask for >= 3 warm-up steps")
    own_stdout()
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
