"""Mint the golden vectors under tests/golden/ from the reference's OWN code (oracle/_ref).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Everything written here is an OUTPUT of the reference sources compiled by oracle/Makefile:
  * nthash_kat.json      btllib::NtHash-order hashes of k-mers (lib/nthash.hpp through the shim)
  * filters.json         SHA-256 of the Bloom / counting-Bloom bytes that the reference's
                         fill_bfs (src/utils.cpp:96-123) leaves for seeded simulated batches,
                         and of the .bf payloads its goldpolish-targeted-bfs writes over the
                         real FIFO protocol
  * ntedit_cases.json    inputs and `_edited.fa` sequences of the reference's kmerizeAndCorrect
                         (ntedit.cpp:1414-1771) for hand-built edge cases, per k and chained
  * ntedit_cli.json      one run of the reference's ntedit-gr binary on a multi-line, commented,
                         mixed-case FASTA (header handling, <100 bp records dropped)
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sim  # noqa: E402
from oracle import ref_driver as rd  # noqa: E402

KS = (32, 28, 24, 20)
BF_BYTES, CBF = 524288, 10485760


def sha(a) -> str:
    return hashlib.sha256(a if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_build(reads_with_T):
    h = rd.harness()
    ks = (C.c_uint * 4)(*KS)
    hb = h.ref_build_open(ks, 4, CBF, BF_BYTES, 4)
    for s, T in reads_with_T:
        h.ref_build_add_read(hb, s, len(s), T)
    bfs, cbfs = [], []
    for i in range(4):
        b = np.zeros(BF_BYTES, np.uint8); h.ref_build_get_bf(hb, i, b.ctypes.data); bfs.append(b)
        c = np.zeros(CBF, np.uint8); h.ref_build_get_cbf(hb, i, c.ctypes.data); cbfs.append(c)
    h.ref_build_close(hb)
    return bfs, cbfs


def ref_ntedit(seq: bytes, bf: np.ndarray, k: int, **o):
    h = rd.harness()
    cap = 3 * len(seq) + 4096
    buf = np.zeros(cap, np.uint8)
    n = h.ref_ntedit_contig(seq, len(seq), bf.ctypes.data, BF_BYTES, k, 4, o.get("max_insertions", 5),
                            o.get("max_deletions", 5), o.get("mode", 1), o.get("mask", 1), 0.5, 0.5,
                            b"/dev/shm/gp_golden_scratch.fa", buf.ctypes.data, cap)
    return None if n < 0 else buf[:n].tobytes()


def write_sam(d, path):
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:unsorted\n")
        for c in range(d.n_contigs):
            f.write(f"@SQ\tSN:{d.contig_name(c)}\tLN:{int(d.contig_off[c + 1] - d.contig_off[c])}\n")
        for r, c in zip(d.map_read.tolist(), d.map_contig.tolist()):
            f.write(f"{d.read_name(r)}\t0\t{d.contig_name(c)}\t1\t60\t*\t*\t0\t0\t*\t*\n")


def main():
    assert rd.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    rng = random.Random(20250607)
    h = rd.harness()

    # ---- (i) ntHash known answers -------------------------------------------------------
    kat = []
    alph = ["ACGT", "ACGTacgt", "ACGTNacgtnRY"]
    for i in range(60):
        n = rng.choice([20, 24, 28, 32, 33, 40, 75, 150])
        s = "".join(rng.choice(alph[i % 3]) for _ in range(n))
        for k in KS:
            cap = max(n, 1)
            pos = np.zeros(cap, np.uint64); hs = np.zeros((cap, 4), np.uint64)
            cnt = h.ref_nthash_all(s.encode(), n, k, cap, pos.ctypes.data, hs.ctypes.data)
            kat.append({"seq": s, "k": k, "pos": pos[:cnt].tolist(), "hashes": [[str(x) for x in row] for row in hs[:cnt].tolist()]})
    json.dump(kat, open(os.path.join(HERE, "nthash_kat.json"), "w"))

    # ---- (ii) filters -----------------------------------------------------------------
    import goldpolish_b200 as gp
    filt = {"cases": []}
    for name, kw, bsize in (("fastq_bsize1", dict(genome_len=40000, seed=11), 1),
                            ("fasta_bsize3", dict(genome_len=40000, seed=12, fastq=0), 3),
                            ("loaded_bsize8", dict(genome_len=90000, seed=13, coverage=40.0), 8)):
        d = sim.simulate(**kw)
        pl = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)],
                             [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off),
                             d.map_read, d.map_contig, bsize=bsize)
        case = {"name": name, "sim": kw, "bsize": bsize, "batches": []}
        for b in range(len(pl.batch_entry_off) - 1):
            ents = pl.entries[int(pl.batch_entry_off[b]):int(pl.batch_entry_off[b + 1])]
            bfs, cbfs = ref_build([(d.read(int(e["read_id"])), int(e["kmer_threshold"])) for e in ents])
            case["batches"].append({"entries": [[int(e["read_id"]), int(e["kmer_threshold"])] for e in ents],
                                    "bf_sha256": [sha(x) for x in bfs], "cbf_sha256": [sha(x) for x in cbfs],
                                    "bf_popcount": [int(np.unpackbits(x).sum()) for x in bfs]})
        # the same batches through the reference's FIFO server (selection, sort, threshold, file format)
        with tempfile.TemporaryDirectory(dir="/dev/shm") as w:
            sim.simulate(write_dir=w, **kw)
            reads = os.path.join(w, "reads.fq" if d.fastq else "reads.fa")
            rd.run_index(os.path.join(w, "draft.fa"), os.path.join(w, "draft.fa.index"))
            rd.run_index(reads, reads + ".index")
            case["reads_index_sha256"] = sha("".join(sorted(open(reads + ".index").readlines())).encode())
            with rd.BfServer(os.path.join(w, "bfs"), os.path.join(w, "draft.fa"), os.path.join(w, "draft.fa.index"),
                             os.path.join(w, "mappings.paf"), reads, reads + ".index", threads=4) as srv:
                for b in range(len(pl.batch_entry_off) - 1):
                    ids = [d.contig_name(c) for c in range(b * bsize, min((b + 1) * bsize, d.n_contigs))]
                    paths = srv.build(str(b), ids)
                    shas = []
                    for k in KS:
                        hdr, pay = rd.parse_bf(paths[k])
                        shas.append(sha(pay))
                        assert int(hdr["k"]) == k and int(hdr["hash_num"]) == 4 and int(hdr["bytes"]) == BF_BYTES
                    case["batches"][b]["server_bf_sha256"] = shas
                    assert shas == case["batches"][b]["bf_sha256"], "plan_batches disagrees with the reference server"
            if name == "fastq_bsize1":
                # the same mappings as SAM (query column 1, target column 3, '@' header lines
                # skipped, mappings.cpp:112-162) must give the same filters
                sam = os.path.join(w, "mappings.sam")
                write_sam(d, sam)
                with rd.BfServer(os.path.join(w, "bfs_sam"), os.path.join(w, "draft.fa"), os.path.join(w, "draft.fa.index"),
                                 sam, reads, reads + ".index", threads=4) as srv:
                    for b in range(len(pl.batch_entry_off) - 1):
                        paths = srv.build(str(b), [d.contig_name(c) for c in range(b * bsize, min((b + 1) * bsize, d.n_contigs))])
                        assert [sha(rd.parse_bf(paths[k])[1]) for k in KS] == case["batches"][b]["bf_sha256"]
                case["sam_equals_paf"] = True
        filt["cases"].append(case)
    # ntLink-style triples: the minimizer filter (mappings.cpp:230-320) with a tight cap, s = 100
    kw = dict(genome_len=50000, seed=14)
    d = sim.simulate(**kw)
    nl = {"name": "ntlink_mx", "sim": kw, "bsize": 2, "mx_max": 12.0, "subsample_max": 100.0, "batches": []}
    with tempfile.TemporaryDirectory(dir="/dev/shm") as w:
        sim.simulate(write_dir=w, **kw)
        reads = os.path.join(w, "reads.fq")
        rd.run_index(os.path.join(w, "draft.fa"), os.path.join(w, "draft.fa.index"))
        rd.run_index(reads, reads + ".index")
        with rd.BfServer(os.path.join(w, "bfs"), os.path.join(w, "draft.fa"), os.path.join(w, "draft.fa.index"),
                         os.path.join(w, "mappings.tsv"), reads, reads + ".index", mx_max=nl["mx_max"],
                         subsample_max=nl["subsample_max"], threads=4) as srv:
            for b in range((d.n_contigs + 1) // 2):
                ids = [d.contig_name(c) for c in range(b * 2, min(b * 2 + 2, d.n_contigs))]
                paths = srv.build(str(b), ids)
                pays = [rd.parse_bf(paths[k])[1] for k in KS]
                nl["batches"].append({"server_bf_sha256": [sha(x) for x in pays],
                                      "bf_popcount": [int(np.unpackbits(np.frombuffer(x, np.uint8)).sum()) for x in pays]})
    filt["ntlink"] = nl
    json.dump(filt, open(os.path.join(HERE, "filters.json"), "w"))

    # ---- (iii) ntEdit edge cases ----------------------------------------------------------
    def rand_seq(n):
        return "".join(rng.choice("ACGT") for _ in range(n))

    truth = rand_seq(1500) + "AC" * 30 + rand_seq(700) + "A" * 45 + rand_seq(900) + "GATTACA" * 9 + rand_seq(800)
    bfs, _ = ref_build([(truth.encode(), 4)] * 5)  # every truth k-mer reaches every k's threshold

    def mutate(s, edits):
        s = list(s)
        for pos, kind, arg in sorted(edits, reverse=True):
            if kind == "sub":
                s[pos] = arg
            elif kind == "ins":
                s[pos:pos] = list(arg)
            elif kind == "del":
                del s[pos:pos + arg]
        return "".join(s)

    def other(c):
        return {"A": "C", "C": "G", "G": "T", "T": "A"}[c]

    cases = {}
    cases["clean"] = truth
    cases["subs"] = mutate(truth, [(p, "sub", other(truth[p])) for p in (100, 400, 401, 950, 2600, 3300)])
    cases["ins_1_to_5"] = mutate(truth, [(200, "ins", "G"), (500, "ins", "TC"), (800, "ins", "ACG"), (1100, "ins", "TTGA"), (1300, "ins", "CATGC")])
    cases["del_1_to_5"] = mutate(truth, [(250, "del", 1), (550, "del", 2), (850, "del", 3), (1150, "del", 4), (1350, "del", 5)])
    cases["unfixable_blocks"] = mutate(truth, [(600, "sub", "N" * 0 + other(truth[600]))] + [(p, "sub", other(truth[p])) for p in range(1000, 1040)])
    cases["near_ends"] = mutate(truth, [(5, "sub", other(truth[5])), (31, "sub", other(truth[31])), (len(truth) - 10, "sub", other(truth[-10])), (len(truth) - 40, "del", 2)])
    cases["lowercase"] = truth[:300] + truth[300:900].lower() + mutate(truth[900:], [(50, "sub", other(truth[950]))])
    cases["n_runs"] = truth[:700] + "N" * 25 + truth[725:1400] + "N" + truth[1401:]
    cases["iupac"] = mutate(truth, [(350, "sub", "R"), (351, "sub", "Y"), (900, "sub", "K"), (1200, "sub", "B"), (1260, "sub", "N"), (2000, "sub", "W"), (2001, "sub", "S"), (2700, "sub", "D")])
    cases["short_80bp"] = truth[100:180]
    cases["exactly_100bp"] = truth[100:200]
    cases["homopolymer_ins"] = mutate(truth, [(2250, "ins", "AAA"), (1520, "ins", "ACAC"), (3200, "del", 7)])
    cases["tandem_shift"] = mutate(truth, [(1530, "del", 3), (2260, "del", 2), (3190, "ins", "GATTACAGATT")])
    cases["all_n"] = "N" * 400
    cases["mixed_dense"] = mutate(truth, [(p, "sub", other(truth[p])) for p in range(300, 3600, 97)] + [(p, "ins", "T") for p in range(350, 3600, 211)] + [(p, "del", 2) for p in range(420, 3600, 307)])
    out = {"truth": truth, "bf_sha256": [sha(b) for b in bfs], "cases": {}}
    for name, draft in cases.items():
        rec = {"draft": draft, "per_k": {}, "chain": None}
        for i, k in enumerate(KS):
            r = ref_ntedit(draft.encode(), bfs[i], k)
            rec["per_k"][str(k)] = None if r is None else r.decode()
        cur = draft.encode()
        for i, k in enumerate(KS):
            cur = ref_ntedit(cur, bfs[i], k)
            if cur is None:
                break
        rec["chain"] = None if cur is None else cur.decode()
        # other modes / limits on the densest case
        if name in ("mixed_dense", "ins_1_to_5"):
            rec["mode0"] = ref_ntedit(draft.encode(), bfs[0], 32, mode=0).decode()
            rec["mode2"] = ref_ntedit(draft.encode(), bfs[0], 32, mode=2, max_insertions=2, max_deletions=2).decode()
            rec["nomask_i3_d10"] = ref_ntedit(draft.encode(), bfs[0], 32, mask=0, max_insertions=3, max_deletions=10).decode()
        out["cases"][name] = rec
    json.dump(out, open(os.path.join(HERE, "ntedit_cases.json"), "w"))

    # ---- (iv) the ntedit-gr binary: FASTA handling -----------------------------------------
    with tempfile.TemporaryDirectory(dir="/dev/shm") as w:
        fa = os.path.join(w, "in.fa")
        d1 = cases["subs"]
        with open(fa, "w") as f:
            f.write(">c1 first comment here\n")
            for i in range(0, len(d1), 70):
                f.write(d1[i:i + 70] + "\n")
            f.write(">tiny\n" + truth[:80] + "\n")
            f.write(">c3\n" + cases["lowercase"] + "\n")
        # a .bf written by the shim == header + payload
        bfp = os.path.join(w, "k32.bf")
        with open(bfp, "wb") as f:
            f.write(b'[BTLKmerBloomFilter_v6]\nbytes = 524288\nhash_fn = "ntHash_v2"\nhash_num = 4\nk = 32\n[HeaderEnd]\n')
            f.write(b"\n  <binary data>\n" + b"\n" * 48)
            f.write(bfs[0].tobytes())
        res = rd.run_ntedit(fa, bfp, os.path.join(w, "out"))
        cli = {"input_fasta": open(fa).read(), "edited_fasta": open(res).read()}
        # the -x/-y threshold form (no -X/-Y: use_ratio stays false, ntedit.cpp:1519-1520) with
        # ntEdit's own defaults (mode 0, no masking) and with -m2 -i2 -d3 -z200 -a1
        res = rd.run_ntedit(fa, bfp, os.path.join(w, "xy"), extra=("-t1",))
        cli["edited_fasta_defaults_xy"] = open(res).read()
        res = rd.run_ntedit(fa, bfp, os.path.join(w, "xy2"), extra=("-x4", "-y6", "-m2", "-i2", "-d3", "-z200", "-a1", "-t1"))
        cli["edited_fasta_x4_y6_m2_i2_d3_z200_a1"] = open(res).read()
    json.dump(cli, open(os.path.join(HERE, "ntedit_cli.json"), "w"))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
