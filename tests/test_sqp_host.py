"""CPU checks of the SQP solver's algebra: the text of ntg_b200/csrc/ntg_sqp.cuh between its CORE markers
compiled with g++ for a one-thread "CTA" (tests/tools/sqp_host.cpp) against the numpy restatement
(tests/tools/sqp_reference.py), and whole solves driven by the CPU oracle's evaluations."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "tools"))
from sqp_reference import ReducedNLP, gi_qp, sqp, sqp_via_core  # noqa: E402
from ntg_b200 import configs  # noqa: E402


@pytest.fixture(scope="module")
def core_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sqp") / "sqp_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "ntg_b200", "csrc"), "-o", so,
                           os.path.join(HERE, "tools", "sqp_host.cpp")])
    return C.CDLL(so)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_dual_active_set_qp_matches_restatement_and_kkt(core_lib):
    """random strictly convex QPs with one- and two-sided rows, equalities and inconsistent pairs: same
    status, point, multipliers and active set as the numpy restatement, and the KKT conditions hold"""
    rng = np.random.default_rng(0)
    solved = 0
    for trial in range(300):
        n, m = int(rng.integers(1, 14)), int(rng.integers(0, 30))
        M = rng.normal(size=(n, n))
        G, g0 = M @ M.T + 0.1 * np.eye(n), rng.normal(size=n)
        A = np.ascontiguousarray(rng.normal(size=(m, n)))
        ax = A @ rng.normal(size=n)
        bl, bu = ax - rng.uniform(0, 1, m), ax + rng.uniform(0, 1, m)
        k = rng.integers(0, 4, m)
        bl[k == 0], bu[k == 1] = -1e20, 1e20
        eqi = np.flatnonzero(k == 2)[:max(0, n - 1)]
        bl[eqi] = bu[eqi] = ax[eqi]
        if trial % 7 == 0 and m > 2:
            A[1], bl[0], bu[0], bl[1], bu[1] = A[0], 1.0, 1e20, -1e20, 0.0
        x, lam, state, st, _, _ = gi_qp(G, g0, A, bl, bu)
        x2, lam2, st2 = np.zeros(n), np.zeros(max(m, 1)), np.zeros(max(m, 1), dtype=np.int32)
        rc = core_lib.sqp_host_qp(n, m, _dp(G), _dp(g0), _dp(A), _dp(bl), _dp(bu), _dp(x2), _dp(lam2),
                                  st2.ctypes.data_as(C.POINTER(C.c_int)))
        assert rc == st, trial
        if st != 0:
            continue
        solved += 1
        np.testing.assert_allclose(x2, x, rtol=0, atol=1e-9)
        np.testing.assert_allclose(lam2[:m], lam, rtol=0, atol=1e-8)
        assert (st2[:m] == state).all()
        ax2 = A @ x2
        assert np.abs(G @ x2 + g0 - A.T @ lam2[:m]).max() <= 1e-8
        assert (bl - ax2)[bl > -1e19].max(initial=0) <= 1e-8 and (ax2 - bu)[bu < 1e19].max(initial=0) <= 1e-8
        ineq = bl != bu
        assert (np.abs(ax2 - bl)[ineq & (lam2[:m] > 1e-12)] <= 1e-8).all()
        assert (np.abs(ax2 - bu)[ineq & (lam2[:m] < -1e-12)] <= 1e-8).all()
    assert solved > 200


def _problem(name):
    if name == "kincar":
        from test_gpu_next import _kincar_active_constraints
        spec = _kincar_active_constraints()
        return spec, configs.coefficients("cfg3", 512, spec, seed=5)
    if name == "endpt":
        spec = configs.endpoint()
        return spec, np.random.default_rng(3).uniform(-0.5, 0.5, (64, spec.nC))
    spec = configs.high_order(order=6, mult=3, ninterv=4, nbps=33, name="solve_hi")
    return spec, np.random.default_rng(8).uniform(-0.5, 0.5, (32, spec.nC))


@pytest.mark.parametrize("name,nstart,max_mean_iters", [("kincar", 24, 25), ("endpt", 8, 45), ("hi", 4, 70)])
def test_sqp_iteration_core_vs_restatement_with_oracle_evaluations(core_lib, port, name, nstart, max_mean_iters):
    """whole solves from random starts, every evaluation by the CPU oracle: the C++ core and the numpy
    restatement take the same path (same status; iteration counts within a few where rounding moves a
    line-search decision) to the same KKT point"""
    spec, X = _problem(name)
    nlp = ReducedNLP(port, spec)
    its = []
    for p in range(nstart):
        y0 = nlp.N.T @ (X[p] - nlp.Cpart)
        a = sqp(nlp, y0, max_iter=100)
        b = sqp_via_core(core_lib, nlp, y0, max_iter=100)
        assert b["status"] == 1, (name, p, b["status"])
        its.append(b["iters"])
        if a["status"] == 1:
            assert abs(a["f"] - b["f"]) <= 1e-7 * max(1.0, abs(a["f"])), (name, p)
            assert abs(a["iters"] - b["iters"]) <= 10
        assert b["viol"] <= 1e-7
    assert np.mean(its) <= max_mean_iters, np.mean(its)
