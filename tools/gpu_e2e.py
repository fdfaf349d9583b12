"""Development aid: time ntgb_eval_host (host buffers, copies inside) on the headline workload.
NTG_B200_HOST_CHUNK_MB selects the library's chunk size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ntg_b200 import configs, Problem
spec, P = configs.get("cfg4")
pb = Problem(spec, 0, fast=True)
r = bench.time_e2e(torch, pb, spec, "cfg4", P, 10, 3, full=True)
ceil = bench.d2h_ceiling(torch, 0, r["d2h_bytes_per_step"])
print(f"chunk_mb={os.environ.get('NTG_B200_HOST_CHUNK_MB', 'default')}: {r['ms_per_step']:.3f} ms per call, {r['value']:.4g} evals/s, "
      f"{r['d2h_bytes_per_step'] / r['ms_per_step'] / 1e6:.1f} GB/s of {ceil:.1f}")
pb.close()
