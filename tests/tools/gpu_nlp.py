"""Development aid: batched augmented-Lagrangian solve (ntgb_solve_nlp) on the kinematic-car lane
change with ACTIVE nonlinear constraints, against scipy SLSQP driven by the CPU oracle.
Usage: python tests/tools/gpu_nlp.py [P]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from scipy.optimize import minimize

from ntg_b200 import Problem, configs
from oracle.oracle import Oracle


def tight_kincar(nbps=40, ninterv=4):
    base = configs.kincar(nbps, constraints=True, name="nlp_kincar")
    import dataclasses
    kw = {f.name: getattr(base, f.name) for f in dataclasses.fields(base)}
    kw.update(ninterv=[ninterv, ninterv], knots=None, bps=None)
    spec = type(base)(**kw)
    lo, up = spec.lowerb.copy(), spec.upperb.copy()
    lo[-2], up[-2] = 0.0, 66.2        # speed^2 (cruise 64; 66.5 at the unconstrained optimum)
    lo[-1], up[-1] = -7.2, 7.2        # curvature numerator (7.66 at the unconstrained optimum)
    spec.lowerb, spec.upperb = lo, up
    return spec


if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    spec = tight_kincar()
    port = Oracle("port")
    nC = spec.nC
    rng = np.random.default_rng(11)
    X = configs.coefficients("cfg3", P, spec, seed=5)
    pb = Problem(spec, 0)
    Cd = torch.from_numpy(X).cuda()
    t0 = time.perf_counter()
    f, v, it, st = pb.solve_nlp(Cd)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    f, v, it, st = (a.cpu().numpy() for a in (f, v, it, st))
    print(f"P={P} {dt*1e3:.1f} ms  status1 {np.mean(st==1):.3f} status2 {np.mean(st==2):.3f} status0 {np.mean(st==0):.3f} "
          f"iters mean {it.mean():.1f} max {it.max()}  f min/mean/max {f.min():.6g} {f.mean():.6g} {f.max():.6g} viol max {v.max():.2e}")
    Cs = Cd.cpu().numpy()
    o = port.eval(spec, Cs[:1], dense=True, band=False, linear=True)
    A = o["A"]; bl, bu = o["bl"], o["bu"]
    lb_l, ub_l = bl[nC:nC + spec.nclin], bu[nC:nC + spec.nclin]
    lb_n, ub_n = bl[nC + spec.nclin:], bu[nC + spec.nclin:]

    def fun(c):
        e = port.eval(spec, c[None, :], mode_obj=2, mode_con=-1, dense=False, band=False)
        return float(e["f"][0]), e["g"][0]

    def con(c):
        e = port.eval(spec, c[None, :], mode_obj=-1, mode_con=2, dense=True, band=False)
        Jd = np.nan_to_num(e["Jdense"][0], nan=0.0)
        return e["c"][0], (Jd.T if Jd.shape[0] == nC else Jd)

    for p in range(min(P, 4)):
        cons = [{"type": "eq", "fun": lambda c: A @ c - lb_l, "jac": lambda c: A},
                {"type": "ineq", "fun": lambda c: con(c)[0] - lb_n, "jac": lambda c: con(c)[1]},
                {"type": "ineq", "fun": lambda c: ub_n - con(c)[0], "jac": lambda c: -con(c)[1]}]
        r = minimize(fun, X[p], jac=True, method="SLSQP", constraints=cons, options={"ftol": 1e-12, "maxiter": 500})
        cv = con(Cs[p])[0]
        print(f"  p={p}: gpu f {f[p]:.10g} viol {v[p]:.2e} st {st[p]} it {it[p]} | slsqp f {r.fun:.10g} ok {r.success} "
              f"| dC {np.abs(r.x - Cs[p]).max():.2e} | active: speed2 max {cv[:spec.nbps].max():.4f} curv max {np.abs(cv[spec.nbps:]).max():.4f}")
    pb.close()
