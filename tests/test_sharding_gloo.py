"""CPU, world_size 2 over gloo: batches are partitioned across ranks with no data-path
collective and the host gather restores batch order (the N>1 path of bench.py / the drop-in)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from goldpolish_b200.shard import assign_batches, gather_in_batch_order
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    work = [(i * 7919) % 97 + 1 for i in range(37)]
    mine = assign_batches(work, world)[rank]
    # stand-in for "polish my batches": a deterministic function of the batch index
    local = {b: f"batch{b}:{work[b]}".encode() for b in mine}
    res = gather_in_batch_order(local, rank, world, dist)
    t = torch.tensor([float(sum(work[b] for b in mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the timing reduction bench.py uses (max over ranks)
    if rank == 0:
        q.put((res, float(t.item()), mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, tmax, mine0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    work = [(i * 7919) % 97 + 1 for i in range(37)]
    assert res == [f"batch{b}:{work[b]}".encode() for b in range(37)]
    assert 0 < len(mine0) < 37
    # LPT keeps the two ranks within one batch of each other
    assert tmax <= sum(work) / 2 + max(work)


def test_assign_batches_properties():
    import sys
    sys.path.insert(0, ROOT)
    from goldpolish_b200.shard import assign_batches
    rng = np.random.default_rng(1)
    work = rng.integers(1, 1000, size=101).tolist()
    for world in (1, 2, 4, 8):
        parts = assign_batches(work, world)
        assert sorted(b for p in parts for b in p) == list(range(101))
        loads = [sum(work[b] for b in p) for p in parts]
        assert max(loads) - min(loads) <= max(work)


def _gather_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from goldpolish_b200.shard import RecordGather, assign_batches
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    bsize, n_contigs = 3, 41  # the last batch is short
    clens = rng.integers(5, 60, size=n_contigs)
    nb = (n_contigs + bsize - 1) // bsize
    work = [int(clens[b * bsize:(b + 1) * bsize].sum()) for b in range(nb)]
    assignment = assign_batches(work, world)
    g = RecordGather(clens, bsize, assignment, rank, dist, device="cpu")
    res = None
    for step in range(2):  # buffers are reused: a second step with different lengths must not see the first one's bytes
        # this rank's "polished" records: contig c -> bytes of value c (+ step), a few records dropped, a few grown
        mine = g.contigs[rank]
        lens = np.array([0 if (c + step) % 11 == 0 else int(clens[c]) + ((c + step) % 3) for c in mine], dtype=np.int64)
        dropped = np.array([1 if (c + step) % 11 == 0 else 0 for c in mine], dtype=np.uint8)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        out = np.concatenate([np.full(l, (c + step) % 251, dtype=np.uint8) for c, l in zip(mine, lens)] + [np.zeros(0, np.uint8)])
        res = g(out, off, dropped)
    if rank == 0:
        q.put((res[0].tobytes(), res[1].tolist(), clens.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_record_gather_restores_contig_order_over_gloo():
    """The gather bench.py times inside its end-to-end step at N > 1 (goldpolish_b200/shard.py::RecordGather), on the CPU
    with gloo: every contig's bytes land at its place in contig order, dropped records have length 0."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    data, lens_all, clens = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    step = 1
    exp_lens = [0 if (c + step) % 11 == 0 else clens[c] + ((c + step) % 3) for c in range(len(clens))]
    assert lens_all == exp_lens
    exp = b"".join(bytes([(c + step) % 251]) * l for c, l in enumerate(exp_lens))
    assert data == exp
