"""TEST INFRASTRUCTURE ONLY.  ctypes handle on oracle/liboracle.so (the CPU restatement)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

CBF_COUNTERS = 10485760
BF_BYTES = 524288
KS = (32, 28, 24, 20)


class NteditOpts(C.Structure):
    _fields_ = [("k", C.c_uint), ("hash_num", C.c_uint), ("max_insertions", C.c_uint),
                ("max_deletions", C.c_uint), ("mode", C.c_int), ("mask", C.c_int),
                ("missing_ratio", C.c_float), ("edit_ratio", C.c_float), ("jump", C.c_uint),
                ("min_contig_len", C.c_uint)]


class NteditStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("triggers", "attempts", "subs", "inss", "dels", "masks",
                                          "rollbacks", "indel_calls", "ref_ub")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("gp_oracle.c", "gp_oracle.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "restatement"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.gpo_ntf64.argtypes = [C.c_char_p, C.c_uint]; l.gpo_ntf64.restype = C.c_uint64
        l.gpo_ntr64.argtypes = [C.c_char_p, C.c_uint]; l.gpo_ntr64.restype = C.c_uint64
        l.gpo_nthash_all.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_size_t, C.c_void_p, C.c_void_p]
        l.gpo_nthash_all.restype = C.c_size_t
        l.gpo_kmer_threshold.argtypes = [C.c_uint64]; l.gpo_kmer_threshold.restype = C.c_int
        l.gpo_mappings_cap.argtypes = [C.c_uint64, C.c_double]; l.gpo_mappings_cap.restype = C.c_uint64
        l.gpo_fill_bfs.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint), C.c_int, C.c_uint,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        l.gpo_fill_bfs.restype = C.c_long
        l.gpo_ntedit_default_opts.argtypes = [C.POINTER(NteditOpts), C.c_uint]
        l.gpo_ntedit_contig.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(NteditOpts),
                                        C.c_void_p, C.c_size_t, C.POINTER(NteditStats)]
        l.gpo_ntedit_contig.restype = C.c_long
        l.gpo_insertion_string.argtypes = [C.c_ubyte, C.c_int, C.c_char_p]; l.gpo_insertion_string.restype = C.c_int
        l.gpo_guard_rejects.argtypes = [C.c_uint64, C.c_uint64]; l.gpo_guard_rejects.restype = C.c_int
        _lib = l
    return _lib


def nthash_all(seq: bytes, k: int):
    n = max(len(seq), 1)
    pos = np.zeros(n, dtype=np.uint64)
    hs = np.zeros((n, 4), dtype=np.uint64)
    cnt = lib().gpo_nthash_all(seq, len(seq), k, n, pos.ctypes.data, hs.ctypes.data)
    return pos[:cnt], hs[:cnt]


class FilterSet:
    """The 4 (CBF, BF) pairs of one batch, filled read by read in reference order."""

    def __init__(self, ks=KS):
        self.ks = tuple(ks)
        self.cbfs = [np.zeros(CBF_COUNTERS, dtype=np.uint8) for _ in ks]
        self.bfs = [np.zeros(BF_BYTES, dtype=np.uint8) for _ in ks]
        self._ks = (C.c_uint * len(ks))(*ks)
        self._c = (C.c_void_p * len(ks))(*[a.ctypes.data for a in self.cbfs])
        self._b = (C.c_void_p * len(ks))(*[a.ctypes.data for a in self.bfs])
        self.ops = 0

    def add_read(self, seq: bytes, kmer_threshold: int) -> int:
        n = lib().gpo_fill_bfs(seq, len(seq), self._ks, len(self.ks), kmer_threshold, self._c, self._b)
        if n < 0:
            raise ValueError("kmer_threshold must be >= 4")
        self.ops += n
        return n


def ntedit_opts(k: int, **kw) -> NteditOpts:
    o = NteditOpts()
    lib().gpo_ntedit_default_opts(C.byref(o), k)
    for key, v in kw.items():
        setattr(o, key, v)
    return o


def ntedit_contig(seq: bytes, bf: np.ndarray, k: int, **kw):
    """Returns (edited bytes or None if dropped, stats dict)."""
    o = ntedit_opts(k, **kw)
    cap = 2 * len(seq) + 4096
    out = np.zeros(cap, dtype=np.uint8)
    st = NteditStats()
    bf = np.ascontiguousarray(bf, dtype=np.uint8)
    n = lib().gpo_ntedit_contig(seq, len(seq), bf.ctypes.data, bf.size, C.byref(o), out.ctypes.data, cap, C.byref(st))
    if n == -1:
        return None, st.as_dict()
    if n < 0:
        raise RuntimeError("oracle ntedit output buffer too small")
    return out[:n].tobytes(), st.as_dict()


def select_reads(names: list[str], phred: list[float], lens: list[int], target_len: int, subsample_max: float):
    """goldpolish_targeted_bfs.cpp:95-123: returns (ordered indices into the lists, kmer_threshold)."""
    n_adj = min(len(names), int(lib().gpo_mappings_cap(target_len, subsample_max)))
    order = sorted(range(len(names)), key=lambda i: (-int(phred[i]), names[i].encode()))
    chosen = order[:n_adj]
    bases = sum(lens[i] for i in chosen)
    return chosen, lib().gpo_kmer_threshold(bases)
