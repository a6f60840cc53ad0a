"""Seeded synthetic drafts / reads / mappings for the GoldPolish hot path.

Thin ctypes wrapper over sim/gpsim.c (built by ``__graft_entry__.build()``).  Test and bench
tooling only; the product never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgpsim.so")


class _Params(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("genome_len", C.c_uint64), ("repeat_frac", C.c_double),
        ("contig_median", C.c_double), ("contig_sigma", C.c_double),
        ("contig_min", C.c_uint32), ("contig_max", C.c_uint32), ("tiny_contig_frac", C.c_double),
        ("draft_err", C.c_double), ("draft_n_run_rate", C.c_double),
        ("draft_lower_rate", C.c_double), ("draft_iupac_rate", C.c_double),
        ("coverage", C.c_double), ("read_mean", C.c_double), ("read_sigma", C.c_double),
        ("read_min", C.c_uint32), ("read_max", C.c_uint32),
        ("read_err", C.c_double), ("read_sub", C.c_double), ("read_ins", C.c_double),
        ("read_n_rate", C.c_double), ("phred_mean", C.c_double), ("phred_sd", C.c_double),
        ("min_overlap", C.c_uint32), ("fastq", C.c_int32),
    ]


class _Sim(C.Structure):
    _fields_ = [
        ("truth", C.c_void_p), ("truth_len", C.c_uint64),
        ("n_contigs", C.c_size_t), ("contig_seq", C.c_void_p), ("contig_bases", C.c_uint64),
        ("contig_off", C.c_void_p), ("contig_tstart", C.c_void_p), ("contig_tend", C.c_void_p),
        ("n_reads", C.c_size_t), ("read_seq", C.c_void_p), ("read_bases", C.c_uint64),
        ("read_off", C.c_void_p), ("read_phred", C.c_void_p), ("read_qchar", C.c_void_p),
        ("read_qlast", C.c_void_p),
        ("n_maps", C.c_size_t), ("map_read", C.c_void_p), ("map_contig", C.c_void_p),
        ("map_overlap", C.c_void_p), ("map_strand", C.c_void_p), ("map_tstart", C.c_void_p),
        ("map_tend", C.c_void_p), ("map_mx", C.c_void_p), ("fastq", C.c_int32),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gpsim.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-std=c11", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.gpsim_default_params.argtypes = [C.POINTER(_Params)]
        _lib.gpsim_generate.argtypes = [C.POINTER(_Params)]
        _lib.gpsim_generate.restype = C.POINTER(_Sim)
        _lib.gpsim_free.argtypes = [C.POINTER(_Sim)]
        _lib.gpsim_write_files.argtypes = [C.POINTER(_Sim), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
        _lib.gpsim_write_files.restype = C.c_int
    return _lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


@dataclass
class SimData:
    """Host copy of one simulated data set.  Sequences are ASCII bytes, CSR offsets."""
    contig_seq: np.ndarray
    contig_off: np.ndarray
    read_seq: np.ndarray
    read_off: np.ndarray
    read_phred: np.ndarray          # value the SeqIndex would report (mean of all but last q char, -33)
    map_read: np.ndarray
    map_contig: np.ndarray
    map_mx: np.ndarray
    map_tstart: np.ndarray          # overlap of the read with the contig, contig coordinates (truth space)
    map_tend: np.ndarray
    truth: np.ndarray
    fastq: bool
    params: dict = field(default_factory=dict)
    read_qchar: np.ndarray = None   # FASTQ: quality character of all but the last base of a read ...
    read_qlast: np.ndarray = None   # ... and of its last base (the index ignores it, src/seqindex.cpp:45)
    map_strand: np.ndarray = None
    map_overlap: np.ndarray = None

    @property
    def n_contigs(self) -> int:
        return len(self.contig_off) - 1

    @property
    def n_reads(self) -> int:
        return len(self.read_off) - 1

    def contig_name(self, i: int) -> str:
        return f"ctg{i}"

    def read_name(self, i: int) -> str:
        return f"read{i}"

    def contig(self, i: int) -> bytes:
        return self.contig_seq[self.contig_off[i]:self.contig_off[i + 1]].tobytes()

    def read(self, i: int) -> bytes:
        return self.read_seq[self.read_off[i]:self.read_off[i + 1]].tobytes()


def simulate(write_dir: str | None = None, **kw) -> SimData:
    """Generate a data set; keyword arguments override gpsim_default_params fields.

    If ``write_dir`` is given, also writes draft.fa, reads.fq|reads.fa, mappings.paf and
    mappings.tsv (ntLink-style triples) there.
    """
    lib = _load()
    p = _Params()
    lib.gpsim_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    g = lib.gpsim_generate(C.byref(p))
    s = g.contents
    try:
        nc, nr, nm = s.n_contigs, s.n_reads, s.n_maps
        d = SimData(
            contig_seq=_arr(s.contig_seq, s.contig_bases, np.uint8),
            contig_off=_arr(s.contig_off, nc + 1, np.uint64).astype(np.int64),
            read_seq=_arr(s.read_seq, s.read_bases, np.uint8),
            read_off=_arr(s.read_off, nr + 1, np.uint64).astype(np.int64),
            read_phred=_arr(s.read_phred, nr, np.float64) if s.fastq else np.zeros(nr),
            map_read=_arr(s.map_read, nm, np.uint32),
            map_contig=_arr(s.map_contig, nm, np.uint32),
            map_mx=_arr(s.map_mx, nm, np.uint32),
            map_tstart=_arr(s.map_tstart, nm, np.uint32),
            map_tend=_arr(s.map_tend, nm, np.uint32),
            truth=_arr(s.truth, s.truth_len, np.uint8),
            fastq=bool(s.fastq),
            params={k: getattr(p, k) for k, _ in _Params._fields_},
            read_qchar=_arr(s.read_qchar, nr, np.uint8), read_qlast=_arr(s.read_qlast, nr, np.uint8),
            map_strand=_arr(s.map_strand, nm, np.uint8), map_overlap=_arr(s.map_overlap, nm, np.uint32),
        )
        if write_dir is not None:
            os.makedirs(write_dir, exist_ok=True)
            reads = os.path.join(write_dir, "reads.fq" if s.fastq else "reads.fa")
            rc = lib.gpsim_write_files(
                g, os.path.join(write_dir, "draft.fa").encode(), reads.encode(),
                os.path.join(write_dir, "mappings.paf").encode(),
                os.path.join(write_dir, "mappings.tsv").encode())
            if rc != 0:
                raise OSError("gpsim_write_files failed")
    finally:
        lib.gpsim_free(g)
    return d


def write_subset(d: SimData, contigs, write_dir: str) -> dict:
    """Write draft.fa, reads.fq|reads.fa, mappings.paf and mappings.tsv for a SUBSET of the contigs (and the reads
    mapped to them), byte-compatible with gpsim_write_files: what the reference's tools see for these contigs is
    what they would see in the full data set (targets are independent).  Returns the paths."""
    os.makedirs(write_dir, exist_ok=True)
    contigs = [int(c) for c in contigs]
    cset = np.zeros(d.n_contigs, dtype=bool)
    cset[contigs] = True
    msel = np.nonzero(cset[d.map_contig])[0]
    reads = np.unique(d.map_read[msel])
    paths = {"draft": os.path.join(write_dir, "draft.fa"),
             "reads": os.path.join(write_dir, "reads.fq" if d.fastq else "reads.fa"),
             "paf": os.path.join(write_dir, "mappings.paf"), "tsv": os.path.join(write_dir, "mappings.tsv")}
    with open(paths["draft"], "wb") as f:
        for c in contigs:
            f.write(b">" + d.contig_name(c).encode() + b"\n" + d.contig(c) + b"\n")
    with open(paths["reads"], "wb") as f:
        for r in reads.tolist():
            seq = d.read(r)
            if d.fastq:
                q = bytes([int(d.read_qchar[r])]) * (len(seq) - 1) + bytes([int(d.read_qlast[r])]) if seq else b""
                f.write(b"@%s len=%d\n" % (d.read_name(r).encode(), len(seq)) + seq + b"\n+\n" + q + b"\n")
            else:
                f.write(b">" + d.read_name(r).encode() + b"\n" + seq + b"\n")
    rlen = np.diff(d.read_off)
    clen = np.diff(d.contig_off)
    with open(paths["paf"], "w") as f, open(paths["tsv"], "w") as g:
        for m in msel.tolist():
            r, c = int(d.map_read[m]), int(d.map_contig[m])
            ov = int(d.map_overlap[m])
            f.write(f"{d.read_name(r)}\t{int(rlen[r])}\t0\t{int(rlen[r])}\t{'-' if d.map_strand[m] else '+'}\t"
                    f"{d.contig_name(c)}\t{int(clen[c])}\t{int(d.map_tstart[m])}\t{int(d.map_tend[m])}\t"
                    f"{int(ov * 0.9)}\t{ov}\t60\n")
            g.write(f"{d.read_name(r)} {d.contig_name(c)} {int(d.map_mx[m])}\n")
    return paths
