"""Mint tests/golden/mask_golden.json from the reference's OWN scripts/goldpolish-mask.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_mask.py
The script is executed as it is; `btllib` (absent here, used by the script only to read and write
records) is replaced by a stand-in whose SeqReader yields our records and whose SeqWriter collects
what the script writes.  Cases: hand-built edge cases + seeded random records."""
from __future__ import annotations

import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scripts/goldpolish-mask"


class _Rec:
    def __init__(self, rid, seq):
        self.id, self.comment, self.seq, self.qual = rid, "", seq, ""


def run_reference(seqs, k, hard):
    out = []
    fake = types.ModuleType("btllib")
    ns = {"__name__": "goldpolish_mask_ref", "btllib": fake}
    sys.modules["btllib"] = fake
    try:
        exec(compile(open(REF).read(), REF, "exec"), ns)
        ns["args"] = types.SimpleNamespace(k=k, n=hard, s=not hard)  # extend_gaps reads the global `args.k` (:51)

        class W:
            def write(self, rid, comment, seq, qual):
                out.append(seq)
        ns["extend_gaps"]([_Rec(str(i), s) for i, s in enumerate(seqs)], W(), k, not hard, hard)
    finally:
        del sys.modules["btllib"]
    return out


def cases():
    rnd = random.Random(20250607)
    hand = [
        "", "A", "N", "n", "NNNN", "acgt", "ACGT" * 20, "acgt" * 20,
        "ACGT" * 10 + "acgtacgt" + "ACGT" * 10,                         # short lower-case island
        "ACGT" * 10 + "a" * 40 + "ACGT" * 10,                           # long lower-case island: kept
        "ACGT" * 10 + "NNNN" + "ACG" + "NNNN" + "ACGT" * 10,            # short upper-case island between N runs
        "ACGT" * 10 + "nnnn" + "ACGT" * 10, "ACGT" * 10 + "nNNn" + "ACGT" * 10,
        "ACGT" * 10 + "acgNNNNacg" + "ACGT" * 10,                       # lower-case run swallows N
        "ACGT" * 10 + "acg" + "N" * 40 + "acg" + "ACGT" * 10,           # ... and becomes long
        "ACGT" * 10 + "N" * 5 + "acg" + "ACGT" * 10,                    # N run first: separate runs
        "NNNN" + "ACGT" * 20 + "nnNN", "nnnACGTnnn", "N" * 100, "n" * 100,
        "ACGT" * 10 + "RYKM" + "ACGT" * 10, "ACGT" * 10 + "rykm" * 10 + "ACGT" * 10,
        "ACGT" * 10 + "*-." + "ACGT" * 10, "ACGT" * 3 + "X" + "ACGT" * 20,  # characters outside every class vanish
        "acgt" * 7 + "ACGT" * 20 + "acgt" * 7,                          # the two ends are upper-cased first
        "a" * 31 + "C" * 31 + "g" * 31, "a" * 32 + "C" * 32 + "g" * 32, "U" * 10 + "ACGT" * 20 + "u" * 10,
    ]
    alpha = "ACGT" * 6 + "acgt" * 3 + "NNn" + "RYKMSWryk" + "*X"
    for _ in range(40):
        n = rnd.choice([0, 1, 31, 32, 33, 63, 64, 65, 200, 1500, 5000])
        s = []
        while len(s) < n:  # runs of random lengths so that both sides of k occur
            s.extend(rnd.choice(alpha) * rnd.choice([1, 1, 2, 5, 31, 32, 33, 70]))
        hand.append("".join(s[:n]))
    return hand


def main():
    seqs = cases()
    gold = {"source": "scripts/goldpolish-mask extend_gaps, executed by tests/golden/make_golden_mask.py", "seqs": seqs, "runs": []}
    for k in (32, 20, 5):
        for hard in (False, True):
            gold["runs"].append({"k": k, "hard": hard, "out": run_reference(seqs, k, hard)})
    with open(os.path.join(HERE, "mask_golden.json"), "w") as f:
        json.dump(gold, f)
    print(f"{len(seqs)} records x {len(gold['runs'])} runs")


if __name__ == "__main__":
    main()
