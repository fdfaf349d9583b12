"""Time the batched reduced-space BFGS solver (ntgb_solve_eq) on the kincar lane change and the
van der Pol problem.  Usage: python tools/gpu_solve.py [P]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

from ntg_b200 import Problem, configs


def run(name, spec, cfg, P, **kw):
    X = configs.coefficients(cfg, P, spec, seed=3)
    pb = Problem(spec, 0, fast=True)
    Cd = torch.from_numpy(X).cuda()
    pb.solve_eq(Cd.clone(), **kw)  # warm-up (allocations, reduction)
    C2 = Cd.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f, it, st = pb.solve_eq(C2, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it, st = it.cpu().numpy(), st.cpu().numpy()
    print(f"{name}: P={P} {dt*1e3:.1f} ms  {P/dt:.3g} problems/s  iters mean {it.mean():.1f} max {it.max()} "
          f"status1 {np.mean(st==1):.3f} status2 {np.mean(st==2):.3f} status0 {np.mean(st==0):.3f} "
          f"f mean {f.mean().item():.6g}", flush=True)
    pb.close()


if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    run("kincar64", configs.kincar(64, constraints=False, name="solve_kc"), "cfg3", P, max_iter=100)
    run("vanderpol20", configs.vanderpol(20, constraints=False, name="solve_vdp"), "cfg2", P, max_iter=300)
