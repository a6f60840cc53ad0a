"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import os
import re

import numpy as np
import pytest

from util import ROOT, dataset, plan


def test_abi_exports_every_declared_symbol():
    import ctypes as C
    import goldpolish_b200 as gp
    hdr = open(os.path.join(ROOT, "include", "goldpolish_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gp_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(gp.lib_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/goldpolish_b200.h but not exported"
    from goldpolish_b200.api import EXPORTS
    assert sorted(EXPORTS) == declared


def test_no_cpu_fallback_without_gpu():
    import torch
    import goldpolish_b200 as gp
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gp.GpError) as ei:
        gp.Context()
    assert ei.value.code == -3 and "no CPU path" in str(ei.value)


def test_product_does_not_import_the_oracle():
    """The product package must never route through oracle/ (tests, smoke and bench's baseline only)."""
    pkg = os.path.join(ROOT, "goldpolish_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt, os.path.join(dp, f)


def test_host_rules_match_reference_constants():
    import goldpolish_b200 as gp
    assert [gp.kmer_threshold(b) for b in (0, 1_000_000, 10_000_000, 37_100_000)] == [5, 5, 7, 13]
    assert gp.mappings_cap(7000, 40.0) == 28 and gp.mappings_cap(10000, 100.0) == 100
    assert gp.guard_rejects(10000, 7499) and not gp.guard_rejects(10000, 7500)


def test_selection_order_and_cap():
    import goldpolish_b200 as gp
    ids = [0, 1, 2, 3, 4]
    names = [b"read10", b"read9", b"read100", b"read2", b"read1"]
    phred = [12.9, 12.1, 30.0, 7.0, 12.5]       # truncated to 12, 12, 30, 7, 12 (tuple<SeqId,size_t>)
    lens = [1000, 2000, 3000, 4000, 5000]
    chosen, thr = gp.select_reads_for_target(ids, names, phred, lens, 10000, 3.0)   # cap = 3
    # phred 30 first; then the three phred-12 reads by id (lexicographic): read1 < read10 < read9
    assert chosen == [2, 4, 0]
    assert thr == gp.kmer_threshold(3000 + 5000 + 1000)
    chosen, _ = gp.select_reads_for_target(ids, names, phred, lens, 100, 40.0)      # cap = size_t(0.4) = 0
    assert chosen == []


def test_plan_batches_shapes_and_dedup():
    d = dataset(genome_len=30000, seed=3)
    pl = plan(d, bsize=4)
    nb = (d.n_contigs + 3) // 4
    assert len(pl.batch_entry_off) == nb + 1 and pl.batch_entry_off[0] == 0
    assert np.all(np.diff(pl.batch_entry_off.astype(np.int64)) >= 0)
    assert pl.contig_batch.tolist() == [c // 4 for c in range(d.n_contigs)]
    # duplicated mapping rows do not duplicate reads (mappings.cpp:65-70)
    import goldpolish_b200 as gp
    mr = np.concatenate([d.map_read, d.map_read])
    mc = np.concatenate([d.map_contig, d.map_contig])
    pl2 = gp.plan_batches(np.diff(d.contig_off), [d.contig_name(i) for i in range(d.n_contigs)],
                          [d.read_name(i) for i in range(d.n_reads)], d.read_phred, np.diff(d.read_off), mr, mc, bsize=4)
    assert np.array_equal(pl.entries, pl2.entries)
