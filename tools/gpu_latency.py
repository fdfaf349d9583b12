"""Latency of ONE evaluation through the host-buffer call (what an NPSOL callback costs on the
drop-in path): ntgb_eval_host with P = 1, pinned and pageable buffers.  Usage: python tools/gpu_latency.py"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from ntg_b200 import JAC_BAND, JAC_DENSE, Problem, configs

for name, spec in (("vanderpol", configs.vanderpol(20, constraints=True)), ("kincar64", configs.kincar(64, constraints=True))):
    pb = Problem(spec, 0)
    d = pb.dims
    for pinned in (False, True):
        mk = (lambda *s: torch.zeros(*s, dtype=torch.float64).pin_memory()) if pinned else (lambda *s: torch.zeros(*s, dtype=torch.float64))
        X = mk(1, d.nC)
        X += 1.0
        for jac, jn in ((JAC_DENSE, "dense"), (JAC_BAND, "band")):
            out = {"f": mk(1), "g": mk(1, d.nC), "c": mk(1, d.ncnln), "result": mk(1, 2),
                   "J": mk(1, d.nC * d.ncnln if jac == JAC_DENSE else d.ncnln * d.sorder)}
            for _ in range(20):
                pb.eval_host_tensors(X, out, 2, 2, jac)
            n = 500
            t0 = time.perf_counter()
            for _ in range(n):
                pb.eval_host_tensors(X, out, 2, 2, jac)
            dt = (time.perf_counter() - t0) / n
            # funobj and funcon are separate NPSOL callbacks: time them apart too
            t0 = time.perf_counter()
            for _ in range(n):
                pb.eval_host_tensors(X, {"f": out["f"], "g": out["g"]}, 2, -1, jac)
            dto = (time.perf_counter() - t0) / n
            print(f"{name:10s} {'pinned' if pinned else 'pageable':8s} J {jn:5s}: funobj+funcon {dt*1e6:7.1f} us/call   funobj only {dto*1e6:7.1f} us/call", flush=True)
    pb.close()
