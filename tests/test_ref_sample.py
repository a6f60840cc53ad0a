"""CPU: oracle/ref_sample.py drives the reference's own serve_batch + ntEdit chain + guard on chosen batches; its
single-thread fallback (the C restatement) must give the same filters and the same records, batch for batch."""
import numpy as np
import pytest

from util import dataset

from oracle import ref_driver as rd
from oracle.ref_sample import ReferenceSample

W = dict(bsize=3, subsample_max=40.0, mx_max=150.0, mappings="paf")


@pytest.mark.skipif(not rd.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_port_and_reference_agree_on_batches():
    d = dataset(genome_len=60000)
    nb = (d.n_contigs + W["bsize"] - 1) // W["bsize"]
    batches = [nb - 1, 0]  # any order, any subset
    ref = ReferenceSample(W, d, batches, threads=2)
    port = ReferenceSample(W, d, batches, threads=1)
    port.kind = "port"
    try:
        r = ref.run()
        q = port.run()
        assert r["kind"] == "reference" and q["kind"] == "port" and r["bases"] == q["bases"]
        for i in range(len(batches)):
            assert np.array_equal(ref.filters(i), port.filters(i)), f"filters of batch {batches[i]}"
            assert ref.polished(i) == port.polished(i), f"records of batch {batches[i]}"
            assert len(ref.polished(i)) > 0
    finally:
        ref.close()
        port.close()
