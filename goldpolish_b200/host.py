"""Host-side planning that feeds the C ABI: which reads, in which order, with which threshold.

Python mirror of the feeder logic of bcgsc/goldpolish src/goldpolish_targeted_bfs.cpp:86-133
(serve_batch) for callers that already hold names / lengths / qualities in memory; the C++
tools under goldpolish_b200/host/ implement the same rules for the file-based drop-ins.
Pure bookkeeping -- no sequence data is touched here.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .api import READ_ENTRY_DTYPE, kmer_threshold, mappings_cap


def select_reads_for_target(read_ids, read_names, read_phred, read_lens, target_len, subsample_max_per_10kbp):
    """Reads hashed for one target, in order, and the target's kmer_threshold.

    goldpolish_targeted_bfs.cpp:95-125: cap = size_t(len * s / 10000.0); sort by
    (size_t(phred_avg) descending, id ascending); take the first min(n, cap); threshold from the
    summed lengths of the reads taken.
    """
    n_adj = min(len(read_ids), mappings_cap(target_len, subsample_max_per_10kbp))
    order = sorted(range(len(read_ids)), key=lambda i: (-int(read_phred[i]), read_names[i]))
    chosen = [read_ids[i] for i in order[:n_adj]]
    bases = int(sum(int(read_lens[i]) for i in order[:n_adj]))
    return chosen, kmer_threshold(bases)


@dataclass
class BatchPlan:
    batch_entry_off: np.ndarray   # uint64 [n_batches + 1]
    entries: np.ndarray           # READ_ENTRY_DTYPE [n_entries]
    contig_batch: np.ndarray      # uint32 [n_contigs]: batch of every contig
    kmer_ops_bases: int           # sum of read lengths over entries (k-mer ops ~= 4x this)


MX_THRESHOLD_MIN, MX_THRESHOLD_MAX = 1, 30  # src/goldpolish_targeted_bfs.cpp:34-35


def filter_ntlink(reads, mx, target_len, mx_max_per_10kbp, mx_min=MX_THRESHOLD_MIN, mx_max=MX_THRESHOLD_MAX):
    """AllMappings::filter for one target (src/mappings.cpp:230-320): at most ceil(L * x / 10000) reads are wanted;
    the smallest minimizer threshold in (mx_min, mx_max] that leaves no more than that (binary search), reads kept
    in their original order."""
    import math
    if not reads:
        return reads
    max_reads = int(math.ceil(float(target_len) * mx_max_per_10kbp / 10000.0))
    count_ge = lambda t: sum(1 for m in mx if m >= t)
    lo, hi = mx_min, mx_max
    if len(reads) <= max_reads:
        thr = lo
    elif count_ge(hi) > max_reads:
        thr = hi
    else:
        while hi - lo > 1:
            mid = (hi + lo) // 2
            if count_ge(mid) > max_reads:
                lo = mid
            else:
                hi = mid
        thr = hi
    return [r for r, m in zip(reads, mx) if m >= thr]


def plan_batches(contig_lens, contig_names, read_names, read_phred, read_lens, map_read, map_contig,
                 bsize=1, subsample_max_per_10kbp=40.0, name_bytes=True, map_mx=None, mx_max_per_10kbp=150.0) -> BatchPlan:
    """Batches of `bsize` consecutive contigs (scripts/goldpolish:344-354), each contig's
    mapped reads de-duplicated in first-seen order (mappings.cpp:65-70), selected and ordered
    as serve_batch does.  With `map_mx` (ntLink-style triples: minimizers per mapping) rows below MX_THRESHOLD_MIN
    are dropped at load (mappings.cpp:96-99) and every target goes through the minimizer filter first."""
    n_contigs = len(contig_lens)
    per_contig: list[list[int]] = [[] for _ in range(n_contigs)]
    per_contig_mx: list[list[int]] = [[] for _ in range(n_contigs)]
    seen: list[set] = [set() for _ in range(n_contigs)]
    mxs = map_mx.tolist() if map_mx is not None else None
    for i, (r, c) in enumerate(zip(map_read.tolist(), map_contig.tolist())):
        if mxs is not None and mxs[i] < MX_THRESHOLD_MIN:
            continue
        if r not in seen[c]:
            seen[c].add(r)
            per_contig[c].append(r)
            if mxs is not None:
                per_contig_mx[c].append(mxs[i])
    if mxs is not None:
        for c in range(n_contigs):
            per_contig[c] = filter_ntlink(per_contig[c], per_contig_mx[c], int(contig_lens[c]), mx_max_per_10kbp)
    names_key = [n.encode() if name_bytes and isinstance(n, str) else n for n in read_names]
    off = [0]
    ents = []
    contig_batch = np.zeros(n_contigs, dtype=np.uint32)
    bases_total = 0
    for b0 in range(0, n_contigs, bsize):
        b = b0 // bsize
        for c in range(b0, min(b0 + bsize, n_contigs)):
            contig_batch[c] = b
            ids = per_contig[c]
            if not ids:
                continue  # goldpolish_targeted_bfs.cpp:92-94
            chosen, thr = select_reads_for_target(ids, [names_key[i] for i in ids], [read_phred[i] for i in ids],
                                                  [read_lens[i] for i in ids], int(contig_lens[c]),
                                                  subsample_max_per_10kbp)
            for r in chosen:
                ents.append((r, thr))
                bases_total += int(read_lens[r])
        off.append(len(ents))
    entries = np.array(ents, dtype=READ_ENTRY_DTYPE) if ents else np.zeros(0, dtype=READ_ENTRY_DTYPE)
    return BatchPlan(np.asarray(off, dtype=np.uint64), entries, contig_batch, bases_total)
