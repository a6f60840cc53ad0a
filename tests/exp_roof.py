"""Experiment (not a test): random-access roof variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import goldpolish_b200 as gp
for mode in ("0", "1", "2"):
    os.environ["GP_ROOF_MODE"] = mode
    ctx = gp.Context()
    for warps in (148 * 8, 148 * 32, 148 * 64):
        sps, ms = ctx.roof_microbench(warps, 2000)
        print(f"mode {mode} warps {warps}: {sps/1e9:.1f} G sector touches/s ({ms:.1f} ms)")
    ctx.close()
