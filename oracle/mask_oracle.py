"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's record post-pass:

  * mask(seq, k, hard)  -- `extend_gaps` of scripts/goldpolish-mask:44-72 (`goldpolish-mask -s|-n -k K`,
    called as `goldpolish-mask -s -k$(firstword $(K))` by scripts/goldpolish-make:65-66);
  * to_upper(seq)       -- scripts/goldpolish-to-upper:15-21.

Pinned by tests/golden/mask_golden.json, which tests/golden/make_golden_mask.py mints by executing the
reference's own scripts/goldpolish-mask (with a stand-in for the btllib reader/writer, the only thing
the script imports btllib for).  Only tests/ may import this module.
"""
from __future__ import annotations

import re

_GROUPS = re.compile(r"([ACTG]+|[Nn]+|[actgUNMRWSYKVHDBunmrwsykvhdb]+)")  # goldpolish-mask:54


def mask(seq: str, k: int, hard: bool = False) -> str:
    if len(seq) < 2 * k:                                   # :48-51
        seq = seq.upper()
    else:
        seq = seq[:k].upper() + seq[k:-k] + seq[-k:].upper()
    out = []
    for g in _GROUPS.findall(seq):                         # :54-64
        if g[0] == "N" or len(g) >= k:
            out.append(g)
        else:
            out.append("N" * len(g) if hard else g.lower())
    s = "".join(out).strip("Nn")                           # :65-67
    return s if s else "N"


def to_upper(seq: str) -> str:
    return seq.upper()
