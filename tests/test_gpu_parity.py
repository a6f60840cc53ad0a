"""GPU parity tests proper: the CUDA path through the C ABI vs the CPU oracle, bit for bit."""
import numpy as np
import pytest

from util import KS, dataset, oracle_build, oracle_polish_contig, plan

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp():
    import goldpolish_b200
    return goldpolish_b200


@pytest.fixture(scope="module")
def small(gp):
    d = dataset(genome_len=60000)
    ctx = gp.Context(keep_counters=1)
    ctx.upload_reads(d.read_seq, d.read_off)
    yield d, ctx
    ctx.close()


@pytest.mark.parametrize("algo", ["l", "s", "p"])
@pytest.mark.parametrize("bsize", [1, 3])
def test_filters_match_oracle(gp, small, bsize, algo, monkeypatch):
    """Filter bits AND counter bytes, for each build kernel: level-synchronous (default), one warp
    per stream, hash/commit warp pairs."""
    monkeypatch.setenv("GP_BUILD_KERNEL", algo)
    d, ctx = small
    pl = plan(d, bsize=bsize)
    bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
    st = ctx.stats()
    ref = oracle_build(d, pl)
    n_batches = len(pl.batch_entry_off) - 1
    assert bfs.shape == (n_batches, 4, gp.BF_BYTES)
    ops = 0
    for b in range(n_batches):
        ops += ref[b].ops
        for ki in range(4):
            assert np.array_equal(bfs[b, ki], ref[b].bfs[ki]), f"BF payload differs: batch {b} k={KS[ki]}"
            assert np.array_equal(ctx.fetch_cbf(b, ki), ref[b].cbfs[ki]), f"CBF differs: batch {b} k={KS[ki]}"
    assert st["kmer_ops"] == ops
    assert st["build_launches"] > 0


@pytest.mark.parametrize("keep", [0, 1])
@pytest.mark.parametrize("overlap", ["0", "1"])
@pytest.mark.parametrize("bsize", [1, 8])
def test_level_overlap_matches_oracle(gp, bsize, overlap, keep, monkeypatch):
    """The level-synchronous kernel with one stream at a time and with the late list rounds of a stream running
    beside round 0 / the level-1 round of the next one; with and without the extra round that materialises the
    counter bytes (it changes which round is a stream's last, hence the schedule)."""
    monkeypatch.setenv("GP_BUILD_KERNEL", "l")
    monkeypatch.setenv("GP_LEVEL_OVERLAP", overlap)
    d = dataset(genome_len=60000)
    ctx = gp.Context(keep_counters=keep)
    ctx.upload_reads(d.read_seq, d.read_off)
    pl = plan(d, bsize=bsize)
    bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
    assert ctx.stats()["build_slots"] == (2 if overlap == "1" else 1)
    ref = oracle_build(d, pl)
    for b in range(len(pl.batch_entry_off) - 1):
        for ki in range(4):
            assert np.array_equal(bfs[b, ki], ref[b].bfs[ki]), f"BF payload differs: batch {b} k={KS[ki]}"
            if keep and b == len(pl.batch_entry_off) - 2:
                assert np.array_equal(ctx.fetch_cbf(b, ki), ref[b].cbfs[ki]), f"CBF differs: batch {b} k={KS[ki]}"
    ctx.close()


def test_filters_wave_invariance(gp, small):
    """Splitting the batches into waves (limited counting-filter residency) changes nothing."""
    d, ctx = small
    pl = plan(d, bsize=2)
    a = ctx.build_filters(pl.batch_entry_off, pl.entries)
    ctx2 = gp.Context(max_resident_batches=2)
    ctx2.upload_reads(d.read_seq, d.read_off)
    b = ctx2.build_filters(pl.batch_entry_off, pl.entries)
    ctx2.close()
    assert np.array_equal(a, b)


@pytest.mark.parametrize("bsize", [1, 4])
def test_polish_matches_oracle(gp, small, bsize):
    d, ctx = small
    pl = plan(d, bsize=bsize)
    bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
    out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
    st = ctx.stats()
    assert st["polish_launches"] > 0
    for c in range(d.n_contigs):
        want = oracle_polish_contig(d.contig(c), [bfs[pl.contig_batch[c], ki] for ki in range(4)])
        got = out[int(off[c]):int(off[c + 1])].tobytes()
        if want is None:
            assert dropped[c] == 1 and got == b""
        else:
            assert dropped[c] == 0
            assert got == want, f"contig {c} (len {len(d.contig(c))}) differs from the oracle"
    assert st["edits"] > 0 and st["masked"] > 0


@pytest.mark.parametrize("bsize", [1, 4])
def test_pipeline_matches_separate_calls(gp, bsize):
    """gp_pipeline_run (edit kernel beside the build kernel, batches built longest-contig first) gives the
    same filters and the same polished sequences as gp_build_run followed by gp_polish_run, run twice."""
    d = dataset(genome_len=150000)
    pl = plan(d, bsize=bsize)
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
        out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
        out = out[:int(off[-1])].copy()
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
        for _ in range(2):
            ctx.pipeline_run()
            bfs2 = ctx.build_fetch()
            out2, off2, dropped2 = ctx.polish_fetch()
            st = ctx.stats()
            assert np.array_equal(bfs, bfs2)
            assert np.array_equal(off, off2) and np.array_equal(dropped, dropped2)
            assert np.array_equal(out, out2[:int(off2[-1])])
            assert st["build_kernel"] == 2 and st["edits"] > 0


@pytest.mark.parametrize("edit_sms", ["0", "1", "5", "40"])
def test_pipeline_sm_sharing_schemes_agree(gp, edit_sms, monkeypatch):
    """The overlapped pass with the edit kernel on SMs of its own (the build launch hands GP_EDIT_SMS SMs back and
    renumbers its CTAs; whole-SM edit CTAs land there) and with the two kernels sharing every SM (GP_EDIT_SMS=0): the
    same filters and records as the separate calls, through a bounded filter pool (several waves) as well."""
    d = dataset(genome_len=150000)
    pl = plan(d, bsize=2)
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        bfs = ctx.build_filters(pl.batch_entry_off, pl.entries)
        out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
        out = out[:int(off[-1])].copy()
    monkeypatch.setenv("GP_EDIT_SMS", edit_sms)
    for resident in (0, 3):
        with gp.Context(max_resident_filters=resident) as ctx:
            import torch
            ctx.upload_reads(d.read_seq, d.read_off)
            pinned = torch.zeros(bfs.shape, dtype=torch.uint8).pin_memory()
            ctx.build_output(pinned)
            ctx.build_stage(pl.batch_entry_off, pl.entries)
            ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
            ctx.pipeline_run()
            ctx.build_fetch(out=pinned)
            out2, off2, dropped2 = ctx.polish_fetch()
            st = ctx.stats()
            assert st["edit_sms"] == int(edit_sms) and st["polish_reruns"] == 0
            assert np.array_equal(pinned.numpy(), bfs), (edit_sms, resident)
            assert np.array_equal(off, off2) and np.array_equal(dropped, dropped2)
            assert np.array_equal(out, out2[:int(off2[-1])])
            ctx.build_output(None)


@pytest.mark.parametrize("pipeline", [False, True])
def test_filters_streamed_to_pinned_host_memory(gp, pipeline):
    """gp_build_output_host: the build kernel writes every final filter into page-locked host memory itself;
    the bytes equal a plain gp_build_fetch, also for a batch without reads (all-zero filters) and over a dirty
    destination buffer."""
    import torch
    d = dataset(genome_len=120000)
    pl = plan(d, bsize=1)
    off = pl.batch_entry_off.copy()
    keep = np.ones(len(pl.entries), dtype=bool)
    keep[int(off[2]):int(off[3])] = False          # batch 2 loses its reads
    entries = pl.entries[keep]
    off[3:] -= off[3] - off[2]
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        want = ctx.build_filters(off, entries)
        assert not want[2].any()
        pinned = torch.full(want.shape, 0xAB, dtype=torch.uint8).pin_memory()
        ctx.build_output(pinned)
        ctx.build_stage(off, entries)
        if pipeline:
            ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
            ctx.pipeline_run()
        else:
            ctx.build_run()
        ctx.build_fetch(out=pinned)
        assert np.array_equal(pinned.numpy(), want)
        ctx.build_output(None)
        assert np.array_equal(ctx.build_fetch(), want)
        with pytest.raises(gp.GpError):
            ctx.build_output(np.zeros(want.shape, dtype=np.uint8))   # pageable memory is refused


@pytest.mark.parametrize("algo", ["l", "s"])
def test_ragged_reads_and_word_boundaries(gp, algo, monkeypatch):
    """Hand-made reads around every boundary of the packed layout (k-1, k, k+1, 31..33, 63..65 bases, empty),
    N and lower-case bases at word edges, a batch of reads that are all shorter than k, an empty batch:
    filters and counters against the oracle for both build kernels."""
    from oracle import oracle_lib as ol
    monkeypatch.setenv("GP_BUILD_KERNEL", algo)
    rnd = np.random.default_rng(17)
    def rseq(n, alphabet="ACGT"):
        return "".join(rnd.choice(list(alphabet), size=n)) if n else ""
    reads = [rseq(n) for n in (0, 1, 19, 20, 21, 23, 24, 25, 27, 28, 29, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 1000)]
    reads += [rseq(300, "ACGTacgtN") for _ in range(6)]
    for pos in (0, 19, 31, 32, 33, 63, 64, 65, 299):                       # one N at a time, at word edges
        s = list(rseq(300)); s[pos] = "N"; reads.append("".join(s))
    reads += [rseq(5000) for _ in range(4)]
    reads += reads[-4:]                                                    # repeats: counters climb past the thresholds
    reads += reads[-4:]
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    buf = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    short = [i for i, r in enumerate(reads) if len(r) < 20]
    batches = [list(range(len(reads))), short, [], list(range(len(reads) - 12, len(reads))) * 2]
    thr = [4, 5, 4, 6]
    entries = np.zeros(sum(len(b) for b in batches), dtype=np.dtype([("read_id", np.uint32), ("kmer_threshold", np.uint32)]))
    boff = np.zeros(len(batches) + 1, dtype=np.uint64)
    e = 0
    for b, ids in enumerate(batches):
        for i in ids:
            entries[e] = (i, thr[b]); e += 1
        boff[b + 1] = e
    with gp.Context(keep_counters=1) as ctx:
        ctx.upload_reads(buf, off)
        bfs = ctx.build_filters(boff, entries)
        for b, ids in enumerate(batches):
            fs = ol.FilterSet(KS)
            for i in ids:
                fs.add_read(reads[i].encode(), thr[b])
            for ki in range(4):
                assert np.array_equal(bfs[b, ki], fs.bfs[ki]), f"filter differs: batch {b} k={KS[ki]}"
                if b == len(batches) - 1:
                    assert np.array_equal(ctx.fetch_cbf(b, ki), fs.cbfs[ki]), f"counters differ: batch {b} k={KS[ki]}"
        assert bfs[0].any() and not bfs[1].any() and not bfs[2].any()


def test_large_thresholds_fall_back_to_the_in_order_kernel(gp, small):
    """kmer_threshold values beyond what the level kernel's epoch tags can hold (the reference never produces
    them: T <= 13) are built by the in-order kernel, still bit-exact."""
    from oracle import oracle_lib as ol
    d, ctx = small
    pl = plan(d, bsize=2)
    entries = pl.entries.copy()
    entries["kmer_threshold"][: len(entries) // 2] = 60
    bfs = ctx.build_filters(pl.batch_entry_off, entries)
    assert ctx.stats()["build_kernel"] == 1
    for b in (0, len(pl.batch_entry_off) - 2):
        fs = ol.FilterSet(KS)
        for e in range(int(pl.batch_entry_off[b]), int(pl.batch_entry_off[b + 1])):
            fs.add_read(d.read(int(entries[e]["read_id"])), int(entries[e]["kmer_threshold"]))
        for ki in range(4):
            assert np.array_equal(bfs[b, ki], fs.bfs[ki])


def test_polish_identity_filters(gp, small):
    """All-ones filter: every k-mer present, nothing is edited.  All-zero filter: every position that
    passes the look-ahead is soft-masked and nothing else changes (size-independent properties)."""
    d, ctx = small
    nb = 1
    ones = np.full((nb, 4, gp.BF_BYTES), 0xFF, dtype=np.uint8)
    ctx.load_filters(ones)
    cb = np.zeros(d.n_contigs, dtype=np.uint32)
    out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, cb)
    for c in range(d.n_contigs):
        got = out[int(off[c]):int(off[c + 1])].tobytes()
        if len(d.contig(c)) < 100:
            assert dropped[c]
        else:
            assert got == d.contig(c)
    ctx.load_filters(np.zeros_like(ones))
    out, off, dropped = ctx.polish(d.contig_seq, d.contig_off, cb)
    for c in range(d.n_contigs):
        got = out[int(off[c]):int(off[c + 1])].tobytes()
        if len(d.contig(c)) >= 100:
            assert got.upper() == d.contig(c).upper() and len(got) == len(d.contig(c))
            want = oracle_polish_contig(d.contig(c), [np.zeros(gp.BF_BYTES, np.uint8)] * 4)
            assert got == want


def test_read_store_piecewise_and_small_slabs(gp, monkeypatch):
    """gp_reads_begin / _append / _end (any split of the reads into pieces, device staging slabs far smaller than the
    store) builds the same packed store as gp_reads_upload: same filters."""
    d = dataset(genome_len=60000)
    pl = plan(d, bsize=2)
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        a = ctx.build_filters(pl.batch_entry_off, pl.entries)
    monkeypatch.setenv("GP_READ_SLAB_BYTES", "50000")
    lens = np.diff(d.read_off)
    cuts = [0, 1, 2, 7, d.n_reads // 2, d.n_reads - 1, d.n_reads]
    pieces = [(d.read_seq[d.read_off[i]:d.read_off[j]], j - i) for i, j in zip(cuts[:-1], cuts[1:])]
    with gp.Context() as ctx:
        ctx.upload_reads_piecewise(lens, pieces)
        b = ctx.build_filters(pl.batch_entry_off, pl.entries)
        ctx.upload_reads(d.read_seq, d.read_off)  # (slabbed through the same path)
        c = ctx.build_filters(pl.batch_entry_off, pl.entries)
    assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("algo", ["l", "s"])
def test_bounded_filter_pool_waves(gp, algo, monkeypatch):
    """max_resident_filters: the filter pool holds 2 batches at a time and is reused wave after wave; the overlapped
    pipeline polishes every wave before its filters go and the payloads arrive in the named page-locked buffer.  Same
    payloads, same polished records as with every filter resident; the calls that cannot work say so."""
    import torch
    monkeypatch.setenv("GP_BUILD_KERNEL", algo)
    d = dataset(genome_len=60000)
    pl = plan(d, bsize=1)
    nb = len(pl.batch_entry_off) - 1
    assert nb >= 5
    with gp.Context() as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ref_bf = ctx.build_filters(pl.batch_entry_off, pl.entries)
        ref_out, ref_off, ref_dropped = ctx.polish(d.contig_seq, d.contig_off, pl.contig_batch)
    pinned = torch.zeros((nb, 4, gp.BF_BYTES), dtype=torch.uint8).pin_memory()
    with gp.Context(max_resident_filters=2) as ctx:
        ctx.upload_reads(d.read_seq, d.read_off)
        ctx.build_stage(pl.batch_entry_off, pl.entries)
        with pytest.raises(gp.GpError):
            ctx.build_run()                      # nowhere to put the payloads of the earlier waves
        ctx.build_output(pinned)
        ctx.build_run()
        got = ctx.build_fetch(out=pinned)
        assert np.array_equal(got.numpy(), ref_bf)
        with pytest.raises(gp.GpError):
            ctx.build_fetch()                    # a pageable destination cannot be served any more
        ctx.polish_stage(d.contig_seq, d.contig_off, pl.contig_batch)
        with pytest.raises(gp.GpError):
            ctx.polish_run()                     # needs every filter resident
        pinned.zero_()
        ctx.pipeline_run()
        assert ctx.stats()["build_launches"] == (nb + 1) // 2
        assert np.array_equal(ctx.build_fetch(out=pinned).numpy(), ref_bf)
        out, off, dropped = ctx.polish_fetch()
        assert np.array_equal(off, ref_off) and np.array_equal(dropped, ref_dropped)
        assert np.array_equal(out[:int(off[-1])], ref_out[:int(ref_off[-1])])
