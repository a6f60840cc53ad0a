"""Multi-GPU sharding of independent batches (one process per GPU, no data-path collective).

Every batch owns its filters (bcgsc/goldpolish src/goldpolish_targeted_bfs.cpp:70-77), so ranks
never exchange data while polishing.  The only communication is the host-side gather of the
polished records, written in batch order exactly as goldpolish-reaper does
(scripts/goldpolish-reaper:51-73).
"""
from __future__ import annotations

import heapq

import numpy as np


def assign_batches(work_per_batch, world_size: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of batches to ranks.  `work_per_batch` is any
    monotone estimate (summed read bases ~ k-mer ops / 4).  Deterministic on every rank."""
    order = sorted(range(len(work_per_batch)), key=lambda b: (-int(work_per_batch[b]), b))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out: list[list[int]] = [[] for _ in range(world_size)]
    for b in order:
        load, r = heapq.heappop(heap)
        out[r].append(b)
        heapq.heappush(heap, (load + int(work_per_batch[b]), r))
    for lst in out:
        lst.sort()
    return out


def gather_in_batch_order(local_results: dict, rank: int, world_size: int, dist=None, dst: int = 0):
    """Gather {batch_index: payload} from every rank on `dst`; returns the payloads ordered by
    batch index on `dst`, None elsewhere.  Uses torch.distributed's object gather (gloo or nccl
    process group); with world_size 1 no communication happens."""
    if world_size == 1 or dist is None:
        return [local_results[b] for b in sorted(local_results)]
    gathered = [None] * world_size if rank == dst else None
    dist.gather_object(local_results, gathered, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for part in gathered:
        for b, v in part.items():
            if b in merged:
                raise ValueError(f"batch {b} was polished by two ranks")
            merged[b] = v
    return [merged[b] for b in sorted(merged)]


class RecordGather:
    """Gather the polished records of sharded batches on rank `dst`, laid out in contig order.

    Batches are `bsize` consecutive contigs; `assignment[r]` lists the batches of rank r (ascending), the same on every
    rank.  A rank hands in its records in the order of its own contigs (CSR: bytes, offsets, dropped flags); every
    gather moves one padded byte buffer and one padded length vector per rank through `dist.gather` -- staged on
    `device` ("cuda" for the NCCL process group of a GPU run, "cpu" for gloo) -- and `dst` copies batch by batch into
    one buffer ordered by contig index, which is the order scripts/goldpolish-reaper:51-73 writes the final FASTA in.
    Buffers are allocated once and reused by every call."""

    def __init__(self, contig_lens, bsize: int, assignment, rank: int, dist, device: str = "cuda", dst: int = 0):
        import torch
        self.torch, self.dist, self.rank, self.dst, self.bsize = torch, dist, rank, dst, bsize
        self.world = len(assignment)
        n_contigs = len(contig_lens)
        self.n_contigs = n_contigs
        clens = np.asarray(contig_lens, dtype=np.int64)
        self.contigs = [np.concatenate([np.arange(b * bsize, min((b + 1) * bsize, n_contigs)) for b in bl])
                        if len(bl) else np.zeros(0, np.int64) for bl in assignment]
        # an edited record may grow: the same bound the per-rank output buffers use
        self.max_bytes = max(int(clens[c].sum()) + int(clens[c].sum()) // 4 + 65536 for c in self.contigs)
        self.max_contigs = max(max(len(c) for c in self.contigs), 1)
        self.dev = torch.empty(self.max_bytes, dtype=torch.uint8, device=device)
        self.lens = torch.zeros(self.max_contigs, dtype=torch.int64, device=device)
        self.recv = self.recv_lens = self.host = None
        if rank == dst:
            self.recv = [torch.empty(self.max_bytes, dtype=torch.uint8, device=device) for _ in range(self.world)]
            self.recv_lens = [torch.zeros(self.max_contigs, dtype=torch.int64, device=device) for _ in range(self.world)]
            self.host = torch.empty((self.world, self.max_bytes), dtype=torch.uint8)
            if device != "cpu":
                self.host = self.host.pin_memory()

    def __call__(self, out, off, dropped):
        """-> (sequence bytes of all contigs in contig order, per-contig lengths with 0 for dropped records) on `dst`,
        None elsewhere."""
        torch = self.torch
        lens_local = np.where(np.asarray(dropped) == 0, np.diff(np.asarray(off).astype(np.int64)), 0).astype(np.int64)
        nbytes = int(off[-1])
        src = out if isinstance(out, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(out))
        self.dev[:nbytes].copy_(src[:nbytes], non_blocking=True)
        self.lens.zero_()
        self.lens[:len(lens_local)].copy_(torch.from_numpy(lens_local), non_blocking=True)
        self.dist.gather(self.dev, self.recv, dst=self.dst)
        self.dist.gather(self.lens, self.recv_lens, dst=self.dst)
        if self.rank != self.dst:
            return None
        lens_all = np.zeros(self.n_contigs, dtype=np.int64)
        rl = [t.cpu().numpy() for t in self.recv_lens]
        for r in range(self.world):
            lens_all[self.contigs[r]] = rl[r][:len(self.contigs[r])]
            used = int(rl[r].sum())
            self.host[r, :used].copy_(self.recv[r][:used], non_blocking=True)
        if self.dev.is_cuda:
            torch.cuda.current_stream().synchronize()
        goff = np.concatenate([[0], np.cumsum(lens_all)])
        final = np.empty(int(goff[-1]), dtype=np.uint8)
        hostnp = self.host.numpy()
        bs = self.bsize
        for r in range(self.world):
            cs = self.contigs[r]
            loff = np.concatenate([[0], np.cumsum(lens_all[cs])])
            i = 0
            while i < len(cs):  # the contigs of one batch are consecutive on both sides: copy batch by batch
                j = min(i + bs - int(cs[i]) % bs, len(cs))
                final[goff[cs[i]]:goff[cs[j - 1] + 1]] = hostnp[r, loff[i]:loff[j]]
                i = j
        return final, lens_all
