"""GPU, BASELINE.json configs[0]: the reference's in-tree draft fixtures as INPUTS (goldrush_test_golden_path.fa, 152
contigs, and the soft-masked goldpolish_target_test_golden_path.fa), reads simulated from the in-tree expected output
(tests/fixture_reads.py; the reference's test downloads its reads, tests/goldpolish_test.sh:6), bsize 1, k = 32 28 24 20,
through the drop-in TOOLS.  Golden = what the reference's own server and ntedit-gr chain produce on the same files
(tests/golden/make_golden_fixtures.py -> tests/golden/fixtures.json): four filter payloads and the polished record of
every batch."""
import hashlib
import json
import os
import subprocess
import tempfile
import threading
import time

import pytest

from util import GOLDEN, KS, ROOT

import fixture_reads as fr

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "goldpolish_b200", "bin")
ENV = dict(os.environ, GP_QUIET="1")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def _payload(path):
    data = open(path, "rb").read()
    return data[len(data) - 524288:]


@pytest.mark.parametrize("which", ["config1", "target"])
def test_fixture_through_the_dropin_tools(which):
    g = json.load(open(os.path.join(GOLDEN, "fixtures.json")))[which]
    draft, truth = fr.load_fixture(which)
    reads, maps = fr.simulate_reads(truth)
    assert len(draft) == g["n_contigs"] and len(reads) == g["n_reads"] and sum(len(s) for _, s, _ in reads) == g["read_bases"]
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as w:
        p = fr.write_inputs(w, draft, reads, maps)
        for f in (p["draft"], p["reads"]):
            subprocess.check_call([os.path.join(BIN, "goldpolish-index"), f, f + ".index"], env=ENV)
        assert sha("".join(sorted(open(p["reads"] + ".index").readlines())).encode()) == g["reads_index_sha256"]
        from oracle import ref_driver as rd  # (only its FIFO client: the server below is OUR binary)
        bdir = os.path.join(w, "bfs")
        with rd.BfServer(bdir, p["draft"], p["draft"] + ".index", p["paf"], p["reads"], p["reads"] + ".index", threads=4,
                         binary=os.path.join(BIN, "goldpolish-targeted-bfs"), env=ENV) as srv:
            # (a) a handful of batches the way an unmodified driver does it: .bf files, then goldpolish-ntedit on them
            latencies = []
            for b in list(range(min(4, len(draft)))) + [len(draft) - 1]:
                name, seq = draft[b]
                t0 = time.perf_counter()
                paths = srv.build(f"t{b}", [name])
                latencies.append(time.perf_counter() - t0)
                assert [sha(_payload(paths[k])) for k in KS] == g["batches"][b]["bf_sha256"], (which, b)
                bd = os.path.join(w, f"t{b}")
                os.makedirs(bd)
                with open(os.path.join(bd, "batch.fa"), "wb") as f:
                    f.write(b">" + name.encode() + b"\n" + seq + b"\n")
                subprocess.check_call([os.path.join(BIN, "goldpolish-ntedit"), "batch", " ".join(paths[k] for k in KS),
                                       "32 28 24 20", "0.5", "0.5", "1", "batch.ntedited.fa"], cwd=bd, env=ENV,
                                      stdout=subprocess.DEVNULL)
                assert sha(open(os.path.join(bd, "batch.ntedited.fa"), "rb").read()) == g["batches"][b]["polished_sha256"], (which, b)
            # one batch at a time, as an unmodified driver with one client would feed the server: name -> ack -> ids ->
            # four .bf files on disk -> ack.  (Stage + build + payload copy are well under a millisecond of GPU work; the
            # protocol's pipe round trips and 2 MiB of file writes are the rest.)
            print(f"{which}: per-batch server latency (ms): " + " ".join(f"{1e3 * x:.1f}" for x in latencies))
            assert sorted(latencies)[len(latencies) // 2] < 0.5
            # (b) EVERY batch through the fused route (goldpolish-polish-batch's "@polish" request): all clients at once
            names = [f"f{b}" for b in range(len(draft))]
            for b, name in enumerate(names):
                bd = os.path.join(w, name)
                os.makedirs(bd)
                with open(os.path.join(bd, "batch.fa"), "wb") as f:
                    f.write(b">" + draft[b][0].encode() + b"\n" + draft[b][1] + b"\n")
                with open(os.path.join(bdir, "batch_name_input"), "w") as f:
                    f.write(name + "\n")
                with open(os.path.join(bdir, "batch_target_ids_input_ready")) as f:
                    f.read()
            got = {}

            def client(b, name):
                bd = os.path.join(w, name)
                with open(os.path.join(bdir, f"{name}-target_ids_input"), "w") as f:
                    f.write(draft[b][0] + "\n@polish " + os.path.join(bd, "batch.fa") + " " + os.path.join(bd, "batch.ntedited.fa") + "\n")
                with open(os.path.join(bdir, f"{name}-bfs_ready")) as f:
                    f.read()
                got[b] = ([sha(_payload(os.path.join(bdir, f"{name}-k{k}.bf"))) for k in KS],
                          sha(open(os.path.join(bd, "batch.ntedited.fa"), "rb").read()))
                for k in KS:
                    os.remove(os.path.join(bdir, f"{name}-k{k}.bf"))

            ths = [threading.Thread(target=client, args=(b, n)) for b, n in enumerate(names)]
            for t in ths:
                t.start()
            for t in ths:
                t.join(timeout=600)
            assert all(not t.is_alive() for t in ths)
        for b in range(len(draft)):
            assert got[b][0] == g["batches"][b]["bf_sha256"], (which, b, "filters")
            assert got[b][1] == g["batches"][b]["polished_sha256"], (which, b, "polished record")
