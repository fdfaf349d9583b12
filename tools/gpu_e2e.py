"""Development aid: time the host-buffer paths on the headline workload: ntgb_eval_host (all outputs
to the host; NTG_B200_HOST_CHUNK_MB selects the library's chunk size) and bench.py's resident
pipeline (only the result table comes back)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ntg_b200 import configs, Problem
spec, P = configs.get("cfg4")
pb = Problem(spec, 0, fast=True)
if "full" in sys.argv[1:] or len(sys.argv) == 1:
    r = bench.time_e2e(torch, pb, spec, "cfg4", P, 10, 3, full=True)
    ceil = bench.d2h_ceiling(torch, 0, r["d2h_bytes_per_step"])
    print(f"chunk_mb={os.environ.get('NTG_B200_HOST_CHUNK_MB', 'default')}: {r['ms_per_step']:.3f} ms per call, {r['value']:.4g} evals/s, "
          f"{r['d2h_bytes_per_step'] / r['ms_per_step'] / 1e6:.1f} GB/s of {ceil:.1f}")
if "resident" in sys.argv[1:]:
    for rep in range(3):
        r = bench.time_e2e(torch, pb, spec, "cfg4", P, 20, 3, full=False)
        print(f"resident (NO_PUSH_KERNEL={os.environ.get('NTG_B200_NO_PUSH_KERNEL')}): {r['ms_per_step']:.3f} ms per step, {r['value']:.4g} evals/s")
pb.close()
