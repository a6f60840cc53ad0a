"""Experiment (not a test): random-access roof variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import goldpolish_b200 as gp
for mode, region in (("3", 4096), ("4", 4096), ("5", 4096)):
    os.environ["GP_ROOF_MODE"] = mode
    ctx = gp.Context()
    for warps in (148 * 24, 148 * 48):
        sps, ms = ctx.roof_microbench(warps, 4000, region)
        print(f"mode {mode} warps {warps}: {sps/1e9:.1f} G sector touches/s ({ms:.2f} ms)")
    ctx.close()
