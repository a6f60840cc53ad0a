"""Mint tests/golden/fixtures/* and tests/golden/fixtures.json from the reference's in-tree test inputs.

Run in the build container (needs /root/reference and oracle/_ref):  python tests/golden/make_golden_fixtures.py

  config1  /root/reference/tests/goldrush_test_golden_path.fa (152 contigs; BASELINE.json configs[0]), reads simulated
           from the in-tree expected output tests/expected_files/goldrush_test_golden_path.goldpolish-polished_expected.fa
           (the real reads are downloaded by tests/goldpolish_test.sh:6)
  target   /root/reference/tests/goldpolish_target_test_golden_path.fa (7 soft-masked records: lower-case input)

Committed per fixture: the draft (gzip; it is the INPUT the GPU box needs, /root/reference does not exist there), the
records of the truth that differ from it, and -- from the reference's OWN server and ntEdit chain (oracle/_ref) on the
simulated reads, bsize 1 -- the SHA-256 of every batch's four filter payloads and of every batch's polished record.
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixture_reads as fr  # noqa: E402
from oracle import ref_driver as rd  # noqa: E402

REF = "/root/reference/tests"
KS = (32, 28, 24, 20)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def mint(which, draft_path, truth_path):
    os.makedirs(fr.FIXTURES, exist_ok=True)
    draft = fr.read_fasta(draft_path)
    with gzip.GzipFile(os.path.join(fr.FIXTURES, f"{which}_draft.fa.gz"), "wb", mtime=0) as f:
        for n, s in draft:
            f.write(b">" + n.encode() + b"\n" + s + b"\n")
    if truth_path:
        truth = dict(fr.read_fasta(truth_path))
        diff = {n: truth[n].decode() for n, s in draft if truth[n] != s}
        with gzip.GzipFile(os.path.join(fr.FIXTURES, f"{which}_truth_diff.json.gz"), "wb", mtime=0) as f:
            f.write(json.dumps(diff).encode())
    draft, truth = fr.load_fixture(which)
    reads, maps = fr.simulate_reads(truth)
    work = tempfile.mkdtemp(prefix="gp_fix_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = {"n_contigs": len(draft), "n_reads": len(reads), "read_bases": sum(len(s) for _, s, _ in reads), "batches": []}
    try:
        p = fr.write_inputs(work, draft, reads, maps)
        rd.run_index(p["draft"], p["draft"] + ".index")
        rd.run_index(p["reads"], p["reads"] + ".index")
        out["reads_index_sha256"] = sha("".join(sorted(open(p["reads"] + ".index").readlines())).encode())
        with rd.BfServer(os.path.join(work, "bfs"), p["draft"], p["draft"] + ".index", p["paf"], p["reads"],
                         p["reads"] + ".index", threads=4) as srv:
            for b, (name, seq) in enumerate(draft):
                paths = srv.build(str(b), [name])
                bd = os.path.join(work, f"b{b}")
                os.makedirs(bd)
                with open(os.path.join(bd, "batch.fa"), "wb") as f:
                    f.write(b">" + name.encode() + b"\n" + seq + b"\n")
                chosen, skipped = rd.run_ntedit_chain(os.path.join(bd, "batch"), [paths[k] for k in KS])
                out["batches"].append({"name": name, "bf_sha256": [sha(rd.parse_bf(paths[k])[1]) for k in KS],
                                       "polished_sha256": sha(open(chosen, "rb").read()), "guard_skipped": bool(skipped),
                                       "changed": open(chosen, "rb").read() != open(os.path.join(bd, "batch.fa"), "rb").read()})
                for k in KS:
                    os.remove(paths[k])
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return out


if __name__ == "__main__":
    g = {"config1": mint("config1", f"{REF}/goldrush_test_golden_path.fa",
                         f"{REF}/expected_files/goldrush_test_golden_path.goldpolish-polished_expected.fa"),
         "target": mint("target", f"{REF}/goldpolish_target_test_golden_path.fa", None)}
    with open(os.path.join(HERE, "fixtures.json"), "w") as f:
        json.dump(g, f, indent=1)
    for k, v in g.items():
        print(k, v["n_contigs"], "contigs", v["n_reads"], "reads", v["read_bases"], "bases;",
              sum(b["changed"] for b in v["batches"]), "records changed by polishing")
