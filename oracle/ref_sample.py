"""TEST INFRASTRUCTURE ONLY.  The reference's own hot path on a chosen set of batches of a simulated data set.

Used by bench.py (reference arm, cpu_baseline leg and its parity check) and by the GPU parity tests at bench size:
writes the chosen batches' contigs, the reads mapped to them and their mappings as the files the reference's tools
read, runs the reference's serve_batch (src/goldpolish_targeted_bfs.cpp:55-146) and its ntEdit chain + 0.75 guard
(scripts/goldpolish-ntedit:20-40) from oracle/_ref, and hands back what they wrote.  When oracle/_ref was not built
(no /root/reference on this machine) the single-thread C restatement (oracle/gp_oracle.c) stands in.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

KS = [32, 28, 24, 20]
BF_BYTES = 524288


class ReferenceSample:
    """The given batches of a workload as files the reference's tools read, and one timed pass over them:
    the reference's serve_batch (filter build) for every batch on `threads` OpenMP threads, then its ntEdit chain +
    guard as `threads` single-thread workers (scripts/goldpolish runs ntedit-gr with -t1 per batch).
    `w` names the workload: bsize, subsample_max, mx_max, mappings ("paf" | "ntlink").  Batch i of the sample is
    global batch batches[i] = contigs [batches[i] * bsize, (batches[i] + 1) * bsize)."""

    def __init__(self, w: dict, d, batches, threads: int):
        import sim
        from oracle import ref_driver as rd
        self.w = w
        self.batches = [int(b) for b in batches]
        nb = len(self.batches)
        self.d, self.nb, self.threads, self.rd = d, nb, threads, rd
        self.kind = "reference" if rd.ref_available() else "port"
        bs = self.w["bsize"]
        self.contigs = [c for b in self.batches for c in range(b * bs, min((b + 1) * bs, d.n_contigs))]
        self.bases = int(sum(int(d.contig_off[c + 1] - d.contig_off[c]) for c in self.contigs))
        self.work = tempfile.mkdtemp(prefix="gp_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        if self.kind == "reference":
            p = sim.write_subset(d, self.contigs, self.work)
            self.draft, self.reads = p["draft"], p["reads"]
            self.maps = p["paf"] if self.w["mappings"] == "paf" else p["tsv"]
            rd.run_index(self.draft, self.draft + ".index")
            rd.run_index(self.reads, self.reads + ".index")
            self.bdir = os.path.join(self.work, "bfs")
            for i, b in enumerate(self.batches):
                bd = os.path.join(self.work, f"batch{i}")
                os.makedirs(bd, exist_ok=True)
                with open(os.path.join(bd, "batch.fa"), "wb") as f:
                    for c in range(b * bs, min((b + 1) * bs, d.n_contigs)):
                        f.write(b">" + d.contig_name(c).encode() + b"\n" + d.contig(c) + b"\n")

    def close(self):
        shutil.rmtree(self.work, ignore_errors=True)

    def run(self) -> dict:
        if self.kind == "port":
            return self._run_port()
        d, nb, bs, threads = self.d, self.nb, self.w["bsize"], self.threads
        shutil.rmtree(self.bdir, ignore_errors=True)
        os.makedirs(self.bdir)
        names, ids_files = [], []
        for i, b in enumerate(self.batches):
            names.append(str(i).encode())
            p = os.path.join(self.bdir, f"{i}.ids")
            with open(p, "w") as f:
                for c in range(b * bs, min((b + 1) * bs, d.n_contigs)):
                    f.write(d.contig_name(c) + "\n")
            ids_files.append(p.encode())
        # the filter build runs in a process of its own (oracle/ref_worker.py says why); the time is the one the
        # reference-side loop reports for itself, process start-up excluded
        job = dict(cwd=self.bdir, draft=self.draft, draft_index=self.draft + ".index", maps=self.maps, reads=self.reads,
                   reads_index=self.reads + ".index", mx_max=self.w["mx_max"], subsample_max=self.w["subsample_max"],
                   threads=threads, ks=KS, names=[n.decode() for n in names], ids_files=[p.decode() for p in ids_files])
        job_path = os.path.join(self.work, "job.json")
        with open(job_path, "w") as f:
            json.dump(job, f)
        p = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_worker.py"), job_path],
                           capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("reference serve_batches failed: " + p.stderr[-2000:])
        t_build = float(p.stdout.strip().splitlines()[-1])
        job = dict(kind="chain", ks=KS, threads=threads,
                   bases=[os.path.join(self.work, f"batch{b}", "batch") for b in range(nb)],
                   bfs_flat=[os.path.join(self.bdir, f"{b}-k{k}.bf") for b in range(nb) for k in KS],
                   outs=[os.path.join(self.work, f"batch{b}", "batch.ntedited.fa") for b in range(nb)])
        with open(job_path, "w") as f:
            json.dump(job, f)
        p = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_worker.py"), job_path],
                           capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("reference ntedit chain failed: " + p.stderr[-2000:])
        t_edit = float(p.stdout.strip().splitlines()[-1])
        return dict(seconds_build=t_build, seconds_edit=t_edit, bases=self.bases, kind="reference", cores=threads, batches=nb)

    # ---- what the reference produced in the last run(), for the parity check ----
    def filters(self, b: int) -> np.ndarray:
        """[4, BF_BYTES] payloads of the b-th batch of the sample (btllib files: payload = the last `bytes` bytes)."""
        if self.kind == "port":
            return self._port_filters[b]
        out = np.empty((4, BF_BYTES), dtype=np.uint8)
        for i, k in enumerate(KS):
            with open(os.path.join(self.bdir, f"{b}-k{k}.bf"), "rb") as f:
                data = f.read()
            out[i] = np.frombuffer(data[len(data) - BF_BYTES:], dtype=np.uint8)
        return out

    def polished(self, b: int) -> list[tuple[str, bytes]]:
        """(name, sequence) records of batch.ntedited.fa of the b-th batch of the sample (after the reference's 0.75 guard)."""
        if self.kind == "port":
            return self._port_polished[b]
        recs = []
        with open(os.path.join(self.work, f"batch{b}", "batch.ntedited.fa"), "rb") as f:
            lines = f.read().split(b"\n")
        for i in range(0, len(lines) - 1, 2):
            if lines[i].startswith(b">"):
                recs.append((lines[i][1:].decode().split()[0], lines[i + 1]))
        return recs

    def _run_port(self) -> dict:
        """No compiled reference on this machine: the single-thread C restatement (oracle/gp_oracle.c)."""
        from oracle import oracle_lib as ol
        d, bs = self.d, self.w["bsize"]
        rlens = np.diff(d.read_off)
        per_contig = {}
        wanted = set(self.contigs)
        for r, c in zip(d.map_read.tolist(), d.map_contig.tolist()):
            if c in wanted and r not in per_contig.setdefault(c, {}):
                per_contig[c][r] = True
        t0 = time.perf_counter()
        fsets = {}
        self._port_filters, self._port_polished = {}, {}
        for i, b in enumerate(self.batches):
            fs = ol.FilterSet()
            for c in range(b * bs, min((b + 1) * bs, d.n_contigs)):
                ids = list(per_contig.get(c, {}))
                if not ids:
                    continue
                chosen, thr = ol.select_reads([d.read_name(i) for i in ids], [d.read_phred[i] for i in ids],
                                              [int(rlens[i]) for i in ids], int(d.contig_off[c + 1] - d.contig_off[c]),
                                              self.w["subsample_max"])
                for j in chosen:
                    fs.add_read(d.read(ids[j]), thr)
            fsets[b] = fs
            self._port_filters[i] = np.stack(fs.bfs)
        t1 = time.perf_counter()
        for i, b in enumerate(self.batches):
            recs, in_sz, out_sz = [], 0, 0
            cs = range(b * bs, min((b + 1) * bs, d.n_contigs))
            for c in cs:
                cur = d.contig(c)
                in_sz += len(d.contig_name(c)) + 3 + len(cur)
                for ki, k in enumerate(KS):
                    cur, _ = ol.ntedit_contig(cur, fsets[b].bfs[ki], k)
                    if cur is None:
                        break
                if cur is not None:
                    recs.append((d.contig_name(c), cur))
                    out_sz += len(d.contig_name(c)) + 3 + len(cur)
            if ol.lib().gpo_guard_rejects(in_sz, out_sz):  # scripts/goldpolish-ntedit:31-40
                recs = [(d.contig_name(c), d.contig(c)) for c in cs]
            self._port_polished[i] = recs
        t2 = time.perf_counter()
        return dict(seconds_build=t1 - t0, seconds_edit=t2 - t1, bases=self.bases, kind="port", cores=1, batches=self.nb)
