"""Time ntgb_solve_sqp against ntgb_solve_nlp on the lane change with active bounds (development aid).
Usage: python tools/gpu_sqp.py [P]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

import numpy as np
import torch

from ntg_b200 import Problem, configs
from test_gpu_next import _kincar_active_constraints

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
spec = _kincar_active_constraints()
X = configs.coefficients("cfg3", P, spec, seed=5)
pb = Problem(spec, 0, fast=True)
for name in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("sqp", "nlp")):
    fn = pb.solve_sqp if name == "sqp" else pb.solve_nlp
    fn(torch.from_numpy(X[:256]).cuda())
    Cd = torch.from_numpy(X).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn(Cd)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    f, v, it, st = (t.cpu().numpy() for t in out[:4])
    print(f"{name}: P={P} {dt*1e3:.1f} ms {P/dt:.4g} problems/s iters mean {it.mean():.1f} max {it.max()} status {np.bincount(st, minlength=5)} "
          f"f median {np.median(f):.9g} viol max(ok) {v[st >= 1].max():.2e}", flush=True)
pb.close()
