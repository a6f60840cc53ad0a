"""Multi-GPU sharding of independent batches (one process per GPU, no data-path collective).

Every batch owns its filters (bcgsc/goldpolish src/goldpolish_targeted_bfs.cpp:70-77), so ranks
never exchange data while polishing.  The only communication is the host-side gather of the
polished records, written in batch order exactly as goldpolish-reaper does
(scripts/goldpolish-reaper:51-73).
"""
from __future__ import annotations

import heapq


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
