"""Mints tests/golden/mappings_golden.json: what the reference's OWN AllMappings (src/mappings.cpp, compiled unmodified
into oracle/_ref/libref_harness.so) holds for the seeded mapping files of tests/test_mappings_loader.py.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_mappings.py"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_mappings_loader as t  # noqa: E402

with tempfile.TemporaryDirectory() as w:
    out = {}
    for name, (path, mx_max, *index) in t.make_cases(w).items():
        dump = t.reference(w, path, mx_max, *index)
        assert dump is not None, "build oracle/_ref first (make -C oracle ref)"
        out[name] = {"file": path, "mx_max_per_10kbp": mx_max, "targets": dump.count(b"\n"),
                     "mapped_reads": sum(len(line.split(b"\t")[1].split()) for line in dump.splitlines()),
                     "sha256": hashlib.sha256(dump).hexdigest()}
    json.dump(out, open(os.path.join(HERE, "mappings_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
