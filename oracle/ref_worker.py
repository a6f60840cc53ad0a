"""TEST INFRASTRUCTURE ONLY.  One call of the reference's serve_batch loop (oracle/_ref/libref_harness.so:
ref_serve_batches) in a process of its own; prints the seconds it reports.

A fresh process per call is the reference's own model: SeqIndex::get_seq keeps its file descriptor in a
`thread_local static` that is opened once per thread (/root/reference/src/seqindex.hpp:64-78), so a process that has
already served batches from one reads file keeps reading THAT file for every later index.

usage: python ref_worker.py <job.json>   (keys: cwd, draft, draft_index, maps, reads, reads_index, mx_max,
                                          subsample_max, threads, ks, names, ids_files; or kind = "chain", bases,
                                          bfs_flat, outs, ks, threads for the ntEdit chain of the same batches)
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_driver as rd  # noqa: E402


def main():
    job = json.load(open(sys.argv[1]))
    os.environ["GP_ORACLE_QUIET"] = "1"
    h = rd.harness()
    ks = (C.c_uint * len(job["ks"]))(*job["ks"])
    if job.get("kind") == "chain":
        # the ntEdit chain + guard of every batch on `threads` forked single-thread workers (ref_ntedit_chain_many).  Also
        # in a process of its own: the caller may hold gigabytes of page-locked memory, which fork() copies eagerly.
        nb = len(job["bases"])
        t = h.ref_ntedit_chain_many((C.c_char_p * nb)(*[x.encode() for x in job["bases"]]),
                                    (C.c_char_p * (nb * len(job["ks"])))(*[x.encode() for x in job["bfs_flat"]]), ks, len(job["ks"]),
                                    (C.c_char_p * nb)(*[x.encode() for x in job["outs"]]), nb, int(job["threads"]))
        print(repr(float(t)))
        return 0 if t >= 0 else 1
    nb = len(job["names"])
    os.chdir(job["cwd"])
    t = h.ref_serve_batches(job["draft"].encode(), job["draft_index"].encode(), job["maps"].encode(), job["reads"].encode(),
                            job["reads_index"].encode(), float(job["mx_max"]), float(job["subsample_max"]), int(job["threads"]),
                            ks, len(job["ks"]), (C.c_char_p * nb)(*[n.encode() for n in job["names"]]),
                            (C.c_char_p * nb)(*[p.encode() for p in job["ids_files"]]), nb)
    print(repr(float(t)))
    return 0 if t >= 0 else 1


if __name__ == "__main__":
    sys.exit(main())
