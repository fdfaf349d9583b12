"""GPU tests of the batched SQP solver ntgb_solve_sqp (the consumer NPSOL is for
the reference, /root/reference/src/ntg.c:250-253).  Not a parity row -- NPSOL's iterates are not
reproduced -- so the bar is: KKT points (feasible by the ORACLE, same cost as scipy SLSQP and as the
numpy restatement of the same algorithm), multipliers and active set consistent with the oracle's
gradient and Jacobian, convergence rate."""
import os
import sys

import numpy as np
import pytest

from common import assert_close
from ntg_b200 import configs

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))


def _bounds(port, spec):
    nC = spec.nC
    o = port.eval(spec, np.zeros((1, nC)), mode_obj=0, mode_con=0, dense=False, band=False, linear=True)
    return o["A"], o["bl"], o["bu"]


def _kkt_residual(port, spec, C, lam):
    """g - A' lam_lin - J' lam_nl at C, by the oracle"""
    nC = spec.nC
    A, _, _ = _bounds(port, spec)
    e = port.eval(spec, C[None, :], mode_obj=2, mode_con=2, dense=True, band=False)
    Jd = np.nan_to_num(e["Jdense"][0], nan=0.0) if e["c"].shape[1] else np.zeros((0, nC))
    Jd = Jd.T if Jd.shape[0] == nC and Jd.shape[1] != nC else Jd
    r = e["g"][0] - A.T @ lam[:spec.nclin] - Jd.T @ lam[spec.nclin:]
    return r, e


def test_sqp_kincar_active_bounds_vs_slsqp_and_restatement(port):
    """512 lane changes with ACTIVE speed / curvature bounds from random starts (the starts scipy SLSQP
    fails from): >= 99 % converge, in < 40 iterations (= evaluations with derivatives) on average"""
    import torch
    from scipy.optimize import minimize
    from ntg_b200 import Problem
    from test_gpu_next import _kincar_active_constraints
    from sqp_reference import ReducedNLP, sqp
    spec = _kincar_active_constraints()
    nC, P = spec.nC, 512
    X = configs.coefficients("cfg3", P, spec, seed=5)
    pb = Problem(spec, 0)
    Cd = torch.from_numpy(X).cuda()
    f, v, it, st, lam, ist = pb.solve_sqp(Cd, multipliers=True)
    Cs, f, v, it, st, lam, ist = (t.cpu().numpy() for t in (Cd, f, v, it, st, lam, ist))
    ok = st == 1
    assert ok.mean() >= 0.99, f"only {ok.mean():.3f} converged; status counts {np.bincount(st)}"
    assert it[ok].mean() < 40, it[ok].mean()
    o = port.eval(spec, Cs, mode_obj=0, mode_con=0, dense=False, band=False, linear=True)
    A, bl, bu = o["A"], o["bl"], o["bu"]
    b = bl[nC:nC + spec.nclin]
    assert np.abs(Cs @ A.T - b).max() <= 1e-8 * (1 + np.abs(b).max()), "linear equalities"
    lb, ub = bl[nC + spec.nclin:], bu[nC + spec.nclin:]
    viol = np.maximum(np.maximum(lb - o["c"], o["c"] - ub), 0).max(axis=1)
    assert (viol[ok] <= 1e-6).all(), viol[ok].max()
    assert_close(f, o["f"], "reported cost equals the oracle's at the returned point")
    assert (o["c"][ok].max(axis=1) > 7.19).all(), "the curvature bound is active at the solution"
    # the problem has one minimiser: every start ends at the same cost
    assert np.ptp(f[ok]) <= 1e-6 * abs(f[ok][0])
    # multipliers / active set: stationarity by the oracle's gradient and Jacobian, complementarity, signs
    for p in np.flatnonzero(ok)[:4]:
        r, e = _kkt_residual(port, spec, Cs[p], lam[p])
        assert np.abs(r).max() <= 1e-4 * max(1.0, np.abs(e["g"][0]).max()), np.abs(r).max()
        ln, sn, cn = lam[p][spec.nclin:], ist[p][spec.nclin:], e["c"][0]
        assert (ist[p][:spec.nclin] == 3).all(), "linear rows of this problem are equalities"
        assert (np.abs(cn - lb)[sn == 1] <= 1e-6).all() and (np.abs(cn - ub)[sn == 2] <= 1e-6).all()
        assert (ln[sn == 1] >= 0).all() and (ln[sn == 2] <= 0).all() and (ln[sn == 0] == 0).all()
        assert (sn != 0).sum() >= 1
    # the numpy restatement of the same algorithm (oracle evaluations) ends at the same point
    nlp = ReducedNLP(port, spec)
    for p in np.flatnonzero(ok)[:3]:
        r = sqp(nlp, nlp.N.T @ (X[p] - nlp.Cpart), max_iter=100)
        assert r["status"] == 1 and abs(r["f"] - f[p]) <= 1e-7 * abs(r["f"])
        assert np.abs(r["C"] - Cs[p]).max() <= 1e-3
        assert abs(r["iters"] - it[p]) <= 10, (r["iters"], it[p])

    # scipy SLSQP started next to the answer agrees
    def fun(c):
        e = port.eval(spec, c[None, :], mode_obj=2, mode_con=-1, dense=False, band=False)
        return float(e["f"][0]), e["g"][0]

    def con(c):
        e = port.eval(spec, c[None, :], mode_obj=-1, mode_con=2, dense=True, band=False)
        Jd = np.nan_to_num(e["Jdense"][0], nan=0.0)
        return e["c"][0], (Jd.T if Jd.shape[0] == nC else Jd)

    cons = [{"type": "eq", "fun": lambda c: A @ c - b, "jac": lambda c: A},
            {"type": "ineq", "fun": lambda c: con(c)[0] - lb, "jac": lambda c: con(c)[1]},
            {"type": "ineq", "fun": lambda c: ub - con(c)[0], "jac": lambda c: -con(c)[1]}]
    p = int(np.flatnonzero(ok)[0])
    r = minimize(fun, Cs[p] + 1e-3, jac=True, method="SLSQP", constraints=cons, options={"ftol": 1e-13, "maxiter": 300})
    assert r.success and abs(r.fun - f[p]) <= 2e-6 * abs(r.fun)
    pb.close()


def test_sqp_all_constraint_kinds(port):
    """packs/endpt.c has every kind of row (nonlinear initial / trajectory / final, linear inequality rows
    of all three kinds, two outputs with different splines): the reduced row gradients built from the
    band Jacobian are only right if the points returned are KKT points by the oracle"""
    import torch
    from ntg_b200 import Problem
    spec = configs.endpoint()
    nC, P = spec.nC, 64
    X = np.random.default_rng(3).uniform(-0.5, 0.5, (P, nC))
    pb = Problem(spec, 0)
    Cd = torch.from_numpy(X).cuda()
    f, v, it, st, lam, ist = pb.solve_sqp(Cd, multipliers=True)
    Cs, f, v, it, st, lam, ist = (t.cpu().numpy() for t in (Cd, f, v, it, st, lam, ist))
    ok = st == 1
    assert ok.mean() >= 0.9, f"only {ok.mean():.2f} converged; status counts {np.bincount(st)}"
    A, bl, bu = _bounds(port, spec)
    lbl, ubl = bl[nC:nC + spec.nclin], bu[nC:nC + spec.nclin]
    lbn, ubn = bl[nC + spec.nclin:], bu[nC + spec.nclin:]
    o = port.eval(spec, Cs, mode_obj=0, mode_con=0, dense=False, band=False)
    lin = Cs @ A.T
    vl = np.maximum(np.maximum(lbl - lin, lin - ubl), 0).max(axis=1)
    vn = np.maximum(np.maximum(lbn - o["c"], o["c"] - ubn), 0).max(axis=1)
    assert (vl[ok] <= 1e-6).all() and (vn[ok] <= 1e-6).all(), (vl[ok].max(), vn[ok].max())
    assert_close(f, o["f"], "cost at the returned point")
    for p in np.flatnonzero(ok)[:6]:
        r, e = _kkt_residual(port, spec, Cs[p], lam[p])
        assert np.abs(r).max() <= 1e-4 * max(1.0, np.abs(e["g"][0]).max()), (p, np.abs(r).max())
        h = np.concatenate([lin[p], e["c"][0]])
        lo, hi = np.concatenate([lbl, lbn]), np.concatenate([ubl, ubn])
        s = ist[p]
        assert (np.abs(h - lo)[s == 1] <= 1e-6).all() and (np.abs(h - hi)[s == 2] <= 1e-6).all()
        assert (lam[p][s == 1] >= 0).all() and (lam[p][s == 2] <= 0).all() and (lam[p][s == 0] == 0).all()
    pb.close()


def test_sqp_edge_cases(port):
    """P = 1 and an equality-only problem (same answer as ntgb_solve_eq); no linear rows at all (N = I);
    a horizon whose dense reduced QP does not fit in shared memory is refused"""
    import torch
    from ntg_b200 import Problem
    from ntg_b200.problem import NtgError
    spec = configs.vanderpol(20, constraints=False, name="sqp_vdp1")
    pb = Problem(spec, 0)
    C1 = torch.ones((1, spec.nC), dtype=torch.float64, device="cuda")
    f, it, st = pb.solve_eq(C1)
    C2 = torch.ones((1, spec.nC), dtype=torch.float64, device="cuda")
    f2, v2, it2, st2 = pb.solve_sqp(C2)
    assert int(st2[0]) == 1 and abs(float(f2[0]) - float(f[0])) <= 1e-6 * abs(float(f[0]))
    pb.close()
    spec = configs.high_order(order=6, mult=3, ninterv=4, nbps=33, name="sqp_hi")
    assert spec.nclin == 0
    pb = Problem(spec, 0)
    X = np.random.default_rng(8).uniform(-0.5, 0.5, (32, spec.nC))
    Cd = torch.from_numpy(X).cuda()
    f, v, it, st = pb.solve_sqp(Cd, max_iter=150)
    Cs = Cd.cpu().numpy()
    o = port.eval(spec, Cs, mode_obj=2, mode_con=0, dense=False, band=False)
    assert_close(f.cpu().numpy(), o["f"], "cost at the returned point")
    ok = st.cpu().numpy() == 1
    assert ok.mean() >= 0.9
    assert (np.abs(o["c"][ok]).max(axis=1) <= 1.0 + 1e-6).all()
    assert (f.cpu().numpy()[ok] <= 1e-8).all()   # positive definite quadratic, C = 0 feasible
    pb.close()
    # 39 free directions: the CTA is two warps (64 threads) instead of one
    spec = configs.high_order(order=6, mult=3, ninterv=12, nbps=49, name="sqp_hi39")
    assert spec.nC == 39 and spec.nclin == 0
    pb = Problem(spec, 0)
    X = np.random.default_rng(8).uniform(-0.5, 0.5, (8, spec.nC))
    Cd = torch.from_numpy(X).cuda()
    f, v, it, st = pb.solve_sqp(Cd, max_iter=400)
    ok = st.cpu().numpy() == 1
    assert ok.mean() >= 0.85, (st.cpu().numpy(), it.cpu().numpy())
    o = port.eval(spec, Cd.cpu().numpy(), mode_obj=2, mode_con=0, dense=False, band=False)
    assert (np.abs(o["c"][ok]).max(axis=1) <= 1.0 + 1e-6).all() and (f.cpu().numpy()[ok] <= 1e-8).all()
    pb.close()
    spec, _ = configs.get("cfg5")
    pb = Problem(spec, 0)
    with pytest.raises(NtgError):
        pb.solve_sqp(torch.zeros((2, spec.nC), dtype=torch.float64, device="cuda"))
    pb.close()
