"""TEST INFRASTRUCTURE ONLY.  Drives the reference's own binaries built under oracle/_ref/.

Speaks the named-pipe protocol of /root/reference/src/goldpolish_targeted_bfs.cpp:148-244
exactly as /root/reference/scripts/goldpolish:363-426 and goldpolish-polish-batch:62-67 do,
and runs ntedit-gr the way /root/reference/scripts/goldpolish-ntedit:20-40 does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
BIN_INDEX = os.path.join(REF_DIR, "goldpolish-index")
BIN_BFS = os.path.join(REF_DIR, "goldpolish-targeted-bfs")
BIN_NTEDIT = os.path.join(REF_DIR, "ntedit-gr")
LIB_HARNESS = os.path.join(REF_DIR, "libref_harness.so")

_ENV = dict(os.environ, GP_ORACLE_QUIET="1")


def ref_available() -> bool:
    return all(os.path.exists(p) for p in (BIN_INDEX, BIN_BFS, BIN_NTEDIT, LIB_HARNESS))


def run_index(seqs: str, index_out: str) -> None:
    subprocess.check_call([BIN_INDEX, seqs, index_out], env=_ENV, stderr=subprocess.DEVNULL)


def parse_bf(path: str) -> tuple[dict, bytes]:
    """Parse a .bf file: returns (header fields, payload bytes)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"[HeaderEnd]\n") + len(b"[HeaderEnd]\n")
    hdr = {}
    lines = data[:end].decode().splitlines()
    hdr["signature"] = lines[0]
    for ln in lines[1:]:
        if "=" in ln:
            k, v = ln.split("=", 1)
            hdr[k.strip()] = v.strip().strip('"')
    nbytes = int(hdr["bytes"])
    payload = data[len(data) - nbytes:]
    return hdr, payload


class BfServer:
    """Context manager around a running oracle goldpolish-targeted-bfs."""

    def __init__(self, workdir, draft, draft_index, mappings, reads, reads_index,
                 mx_max=150.0, subsample_max=40.0, threads=2, ks=(32, 28, 24, 20), binary=BIN_BFS,
                 env=None):
        self.workdir = workdir
        self.ks = list(ks)
        os.makedirs(workdir, exist_ok=True)
        args = [binary, draft, draft_index, mappings, reads, reads_index, str(mx_max),
                str(subsample_max), str(threads)] + [str(k) for k in ks]
        self.proc = subprocess.Popen(args, cwd=workdir, env=env or _ENV, stderr=subprocess.DEVNULL)
        self._p_name = os.path.join(workdir, "batch_name_input")
        self._p_ready = os.path.join(workdir, "batch_target_ids_input_ready")
        t0 = time.time()
        while not (os.path.exists(self._p_name) and os.path.exists(self._p_ready)):
            if self.proc.poll() is not None:
                raise RuntimeError("BF server exited early")
            if time.time() - t0 > 600:
                raise TimeoutError("BF server did not create its pipes")
            time.sleep(0.01)

    def build(self, batch_name: str, target_ids: list[str]) -> dict[int, str]:
        with open(self._p_name, "w") as f:
            f.write(batch_name + "\n")
        with open(self._p_ready) as f:
            f.read()
        with open(os.path.join(self.workdir, f"{batch_name}-target_ids_input"), "w") as f:
            for t in target_ids:
                f.write(t + "\n")
        with open(os.path.join(self.workdir, f"{batch_name}-bfs_ready")) as f:
            f.read()
        return {k: os.path.join(self.workdir, f"{batch_name}-k{k}.bf") for k in self.ks}

    def close(self):
        if self.proc.poll() is None:
            with open(self._p_name, "w") as f:
                f.write("x\n")
            self.proc.wait(timeout=60)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        try:
            self.close()
        finally:
            if self.proc.poll() is None:
                self.proc.kill()


def run_ntedit(draft_fa: str, bf: str, prefix: str, extra=("-d5", "-i5", "-m1", "-X0.5", "-Y0.5", "-t1", "-a1"),
               binary=BIN_NTEDIT) -> str:
    """One ntedit-gr invocation as in scripts/goldpolish-ntedit:27; returns the _edited.fa path."""
    subprocess.check_call([binary, "-f", draft_fa, "-r", bf, "-b", prefix, *extra], env=_ENV)
    return prefix + "_edited.fa"


def run_ntedit_chain(base: str, bfs: list[str], ks=(32, 28, 24, 20), binary=BIN_NTEDIT) -> tuple[str, bool]:
    """scripts/goldpolish-ntedit:20-40 (bc's scale=4 truncation restated with integers)."""
    in_size = os.path.getsize(base + ".fa")
    prev = None
    for bf, k in zip(bfs, ks):
        inp = base if prev is None else prev
        run_ntedit(inp + ".fa", bf, f"{inp}.k{k}.X0.5.Y0.5", binary=binary)
        prev = f"{inp}.k{k}.X0.5.Y0.5_edited"
    out_size = os.path.getsize(prev + ".fa")
    skipped = (out_size * 10000) // in_size < 7500
    return (base + ".fa" if skipped else prev + ".fa"), skipped


_h = None


def harness():
    """ctypes handle on oracle/_ref/libref_harness.so (in-process reference functions)."""
    global _h
    if _h is None:
        h = C.CDLL(LIB_HARNESS)
        h.ref_kmer_threshold.argtypes = [C.c_ulong]
        h.ref_kmer_threshold.restype = C.c_int
        h.ref_build_open.argtypes = [C.POINTER(C.c_uint), C.c_int, C.c_size_t, C.c_size_t, C.c_uint]
        h.ref_build_open.restype = C.c_void_p
        h.ref_build_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint]
        h.ref_build_get_bf.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        h.ref_build_get_cbf.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        h.ref_build_close.argtypes = [C.c_void_p]
        h.ref_nthash_all.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_size_t, C.c_void_p, C.c_void_p]
        h.ref_nthash_all.restype = C.c_size_t
        h.ref_serve_batches.argtypes = [C.c_char_p] * 5 + [C.c_double, C.c_double, C.c_int,
                                                           C.POINTER(C.c_uint), C.c_int,
                                                           C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int]
        h.ref_serve_batches.restype = C.c_double
        h.ref_ntedit_chain.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_uint), C.c_int, C.c_char_p]
        h.ref_ntedit_chain.restype = C.c_int
        h.ref_ntedit_chain_many.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_uint),
                                            C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int]
        h.ref_ntedit_chain_many.restype = C.c_double
        h.ref_ntedit_contig.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint,
                                        C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_float, C.c_float,
                                        C.c_char_p, C.c_void_p, C.c_size_t]
        h.ref_ntedit_contig.restype = C.c_long
        _h = h
    return _h
