"""goldpolish_b200 -- B200-native (sm_100a) GoldPolish hot path.

The product is the CUDA shared library ``libgoldpolish_b200.so`` behind the C ABI declared in
``include/goldpolish_b200.h`` (built by ``__graft_entry__.build()`` / ``csrc/Makefile``).  This
package is the thin ctypes binding used by the tests, ``bench.py`` and Python callers; it holds
no compute of its own and raises if the CUDA library is missing -- there is no CPU path.
"""
from .api import (BF_BYTES, CBF_BYTES, DEFAULT_KS, Context, GpError, flagged_bed, guard_rejects, kmer_threshold,
                  lib_path, load_library, mappings_cap)
from .host import BatchPlan, plan_batches, select_reads_for_target

__all__ = ["BF_BYTES", "CBF_BYTES", "DEFAULT_KS", "Context", "GpError", "flagged_bed", "guard_rejects", "kmer_threshold",
           "lib_path", "load_library", "mappings_cap", "BatchPlan", "plan_batches", "select_reads_for_target"]
