"""CPU: the reference arm of bench.py (`--impl reference`) prints exactly one JSON line with the keys the driver
reads; it runs the reference's own sources (oracle/_ref) when they were built here, else the oracle port."""
import json
import os
import subprocess
import sys

from util import ROOT


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--config", "2", "--genome-len", "400000", "--ref-bases-per-core", "20000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "polished_mbp_per_s" and d["unit"] == "Mbp/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # both arms describe the workload with the same `config` object (the driver compares them)
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    cfg = bench.config_of(argparse.Namespace(config=2))
    assert set(cfg) == set(d["config"])


def test_reference_arm_source_never_names_the_product():
    import inspect

    import bench
    for obj in (bench.reference_arm, bench.sample_batches, bench.make_dataset, bench.config_of):
        src = "\n".join(ln for ln in inspect.getsource(obj).splitlines() if "not in sys.modules" not in ln)
        assert "goldpolish_b200" not in src and "import gp" not in src
    import oracle.ref_sample
    assert "goldpolish_b200" not in inspect.getsource(oracle.ref_sample)
