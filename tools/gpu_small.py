"""Graph-timed sweep for the launch-bound configs (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ntg_b200 import configs, Problem
for cfg in ("cfg2", "cfg3"):
    spec, P = configs.get(cfg)
    for fast in (True,):
        pb = Problem(spec, 0, fast=fast)
        r = bench.time_workload(torch, pb, spec, cfg, P, 400, 10, use_graph=True)
        print(cfg, "R=%s" % os.environ.get("NTG_B200_ROUNDS", "auto"), "%.2f us/step" % (r["kernel_ms"] * 1e3),
              "%.3g evals/s" % (P / (r["kernel_ms"] * 1e-3)), flush=True)
        pb.close()
