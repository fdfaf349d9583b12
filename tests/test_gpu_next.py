"""GPU tests for the rows SURVEY.md section 8(f) marks "next": linear constraints
(A*C, bounds, violation), batched SplineInterp, and the ntg() drop-in driven by
the same fake NPSOL that drives the reference."""
import ctypes as C
import os

import numpy as np
import pytest

from common import assert_bitexact, assert_close, golden_spec, load_golden
from ntg_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg2_vanderpol", "cfg3_kincar", "endpoint"])
def test_linear_constraints_batched(port, name):
    """A*C per problem equals the oracle's A (reference LinearConstraintsMatrix,
    src/constraints.c:198-261) applied on the host; violation against bounds()"""
    import torch
    from ntg_b200 import Problem
    spec, X = golden_spec(name)
    o = port.eval(spec, X, dense=False, band=False, linear=True)
    pb = Problem(spec, 0)
    lin, viol = pb.eval_linear(torch.from_numpy(X).cuda())
    want = X @ o["A"].T
    assert_close(lin.cpu().numpy(), want, "A*C")
    lb, ub = o["bl"][spec.nC:spec.nC + spec.nclin], o["bu"][spec.nC:spec.nC + spec.nclin]
    lin_h = lin.cpu().numpy()
    v = np.maximum(np.maximum(lb - lin_h, lin_h - ub), 0).max(axis=1)
    assert_bitexact(viol.cpu().numpy(), v, "linear violation")
    pb.close()


def test_spline_interp_batched(port):
    """reference SplineInterp (src/colloc.c:449-484) at arbitrary times, incl. both ends,
    interior knots and a point past the last knot"""
    import torch
    from ntg_b200 import Problem
    spec = configs.endpoint()
    rng = np.random.default_rng(9)
    X = rng.uniform(-2, 2, (5, spec.nC))
    t = np.concatenate([[0.0, 0.5, 0.8, 1.25, 2.0, 2.0000000000000004], rng.uniform(0, 2, 20)])
    pb = Problem(spec, 0)
    out = pb.spline_interp(torch.from_numpy(X).cuda(), torch.from_numpy(t).cuda()).cpu().numpy()
    iC = np.concatenate([[0], np.cumsum(spec.ncoef)])
    iz = np.concatenate([[0], np.cumsum(spec.maxderiv)])
    for p in range(X.shape[0]):
        for i, ti in enumerate(t):
            for j in range(spec.nout):
                f = port.spline_interp(ti, spec.knots[j], X[p, iC[j]:iC[j + 1]], spec.order[j],
                                       spec.mult[j], spec.maxderiv[j])
                assert_bitexact(out[p, i, iz[j]:iz[j + 1]], f, f"p={p} t={ti} output {j}")
    pb.close()


def test_single_point_SplineInterp_dropin(port):
    from ntg_b200 import build
    lib = C.CDLL(build.CORE_SO)
    dp = C.POINTER(C.c_double)
    knots = np.array([0.0, 2.5, 5.0])
    coefs = np.random.default_rng(4).uniform(-1, 1, 7)
    for x in (0.0, 1.7, 2.5, 5.0):
        f = np.zeros(3)
        lib.SplineInterp(f.ctypes.data_as(dp), C.c_double(x), knots.ctypes.data_as(dp), 2,
                         coefs.ctypes.data_as(dp), 7, 5, 3, 3)
        assert_bitexact(f, port.spline_interp(x, knots, coefs, 5, 3, 3), f"x={x}")


@pytest.mark.parametrize("example", ["vanderpol", "kincar"])
def test_reference_examples_drop_in(example):
    """examples/vanderpol.c and examples/kincar.c, UNMODIFIED, compiled against
    include/ntg.h with their callbacks as __device__ functions, linked against
    libntg_b200.so: main() runs, calls ntg(), which hands NPSOL (here: the same
    fake npsol_ that drives the reference) GPU-backed funobj/funcon.  Outputs must
    match the fixture produced by the reference's own build of the same program."""
    from ntg_b200 import build, problem
    from oracle import oracle
    so = build.pack_so(f"ref_{example}")
    if not os.path.exists(so):
        pytest.skip("drop-in example packs are built where /root/reference exists")
    lib = problem.load_pack(f"ref_{example}")
    main = C.cast(getattr(lib, f"ntg_example_{example}_main"), C.c_void_p).value
    g = load_golden(f"example_{example}")
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as td:   # vanderpol.c writes its solution to ./coef1
        os.chdir(td)
        try:
            r = oracle.run_product_main(main, g["X"])
        finally:
            os.chdir(cwd)
    assert r["calls"] == 1, "ntg() did not reach npsol_"
    assert [r["n"], r["nclin"], r["ncnln"]] == list(g["dims"])
    assert_bitexact(r["A"], g["A"], "linear constraint matrix handed to NPSOL")
    assert_bitexact(r["bl"], g["bl"], "bl")
    assert_bitexact(r["bu"], g["bu"], "bu")
    # vanderpol's ucf calls pow(); device pow is within 2 ulp of glibc's
    assert_close(r["f"], g["f"], "objective")
    assert_close(r["g"], g["g"], "gradient")


def test_ntg_dropin_with_constraints_dense_jacobian(port):
    """ntg() with nonlinear constraints: NPSOL's funcon gets c and the dense column-major
    Jacobian (ldJ = ncnln) exactly as the reference's NPfuncon fills it."""
    from ntg_b200 import build, problem
    from ntg_b200.abi import BuiltSetup
    from oracle import oracle
    spec, X = golden_spec("endpoint")
    o = port.eval(spec, X, dense=True, band=False, linear=True)
    lib = problem.load_pack("endpt")
    bs = BuiltSetup(spec, lambda role, sym: C.cast(getattr(lib, sym), C.c_void_p).value)
    s = bs.struct
    core = problem.core()
    shim = C.CDLL(oracle.SHIM_SO, mode=C.RTLD_GLOBAL)

    # a main() that forwards the prepared setup to ntg()
    n = spec.nC
    istate = (C.c_int * (n + spec.nclin + spec.ncnln))()
    clambda = (C.c_double * (n + spec.nclin + spec.ncnln))()
    R = (C.c_double * ((n + 1) * (n + 1)))()
    inform, obj = C.c_int(0), C.c_double(0.0)
    x0 = (C.c_double * n)()

    @C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_char_p))
    def main(argc, argv):
        core.ntg.restype = None
        core.ntg(s.nout, s.bps, s.nbps, s.kninterv, s.knots, s.order, s.mult, s.maxderiv, x0,
                 s.nlic, s.lic, s.nltc, s.ltc, s.nlfc, s.lfc,
                 s.nnlic, C.c_void_p(s.nlicf), s.nnltc, C.c_void_p(s.nltcf), s.nnlfc, C.c_void_p(s.nlfcf),
                 s.ninitialconstrav, s.initialconstrav, s.ntrajectoryconstrav, s.trajectoryconstrav,
                 s.nfinalconstrav, s.finalconstrav, s.lowerb, s.upperb,
                 s.nicf, C.c_void_p(s.icf), s.nucf, C.c_void_p(s.ucf), s.nfcf, C.c_void_p(s.fcf),
                 s.ninitialcostav, s.initialcostav, s.ntrajectorycostav, s.trajectorycostav,
                 s.nfinalcostav, s.finalcostav, istate, clambda, R, C.byref(inform), C.byref(obj))
        return 0

    r = oracle.run_product_main(C.cast(main, C.c_void_p).value, X, ncnln=spec.ncnln)
    assert r["calls"] == 1 and (r["n"], r["nclin"], r["ncnln"]) == (spec.nC, spec.nclin, spec.ncnln)
    assert_close(r["f"], o["f"], "f")
    assert_close(r["g"], o["g"], "g")
    assert_close(r["c"], o["c"], "c")
    assert_close(r["Jdense"], np.nan_to_num(o["Jdense"], nan=0.0), "dense Jacobian as NPSOL sees it")
    assert_bitexact(r["A"], o["A"], "A")


def test_ntg_without_npsol_reports_it(capfd):
    """NPSOL absent -> ntg() sets up on the GPU, says so, sets inform, solves nothing."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from ntg_b200 import problem, configs
from ntg_b200.abi import BuiltSetup
spec = configs.vanderpol(20, constraints=False)
lib = problem.load_pack("vdp")
s = BuiltSetup(spec, lambda role, sym: C.cast(getattr(lib, sym), C.c_void_p).value).struct
core = problem.core()
n = spec.nC
inform, obj = C.c_int(0), C.c_double(0.0)
x0 = (C.c_double * n)(*([1.0] * n))
core.ntg.restype = None
core.ntg(s.nout, s.bps, s.nbps, s.kninterv, s.knots, s.order, s.mult, s.maxderiv, x0,
         s.nlic, s.lic, s.nltc, s.ltc, s.nlfc, s.lfc, 0, None, 0, None, 0, None, 0, None, 0, None, 0, None,
         s.lowerb, s.upperb, 0, None, 1, C.c_void_p(s.ucf), 0, None, 0, None,
         s.ntrajectorycostav, s.trajectorycostav, 0, None,
         (C.c_int * (n + 3))(), (C.c_double * (n + 3))(), (C.c_double * ((n + 1) ** 2))(),
         C.byref(inform), C.byref(obj))
print("INFORM", inform.value)
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "INFORM -1000" in p.stdout, p.stdout + p.stderr
    assert "NPSOL" in p.stderr


@pytest.mark.parametrize("name", ["cfg3_kincar", "endpoint"])
def test_batched_merit_linesearch(port, name):
    """ntgb_linesearch: P*nalpha values-only evaluations in one launch + per-problem Armijo /
    argmin selection, against the CPU oracle evaluated at the same trial points."""
    import torch
    from ntg_b200 import Problem
    from common import violation
    spec, X = golden_spec(name)
    rng = np.random.default_rng(21)
    P, nC = X.shape
    D = rng.uniform(-1, 1, (P, nC))
    alphas = np.array([1.0, 0.5, 0.25, 0.125, 0.0625])
    mu, c1 = 3.0, 1e-4
    o = port.eval(spec, X, dense=False, band=False, linear=True)
    def merit(Xe):
        r = port.eval(spec, Xe, mode_obj=0, mode_con=0, dense=False, band=False)
        v = violation(spec, r["c"])
        lin = Xe @ o["A"].T
        lb, ub = o["bl"][nC:nC + spec.nclin], o["bu"][nC:nC + spec.nclin]
        vl = np.maximum(np.maximum(lb - lin, lin - ub), 0).max(axis=1) if spec.nclin else np.zeros(len(Xe))
        return r["f"] + mu * np.maximum(v, vl)
    phi0 = merit(X)
    dphi0 = -np.abs(rng.uniform(0.1, 1.0, P)) * np.abs(phi0)
    trial = np.stack([merit(X + a * D) for a in alphas], axis=1)          # [P][nalpha]
    want_a, want_phi = np.zeros(P), np.zeros(P)
    for p in range(P):
        ok = [i for i, a in enumerate(alphas) if trial[p, i] <= phi0[p] + c1 * a * dphi0[p]]
        i = ok[0] if ok else int(np.argmin(trial[p]))
        want_a[p], want_phi[p] = alphas[i], trial[p, i]
    pb = Problem(spec, 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ab, pbest, Cn = pb.linesearch(t(X), t(D), t(alphas), mu, c1, t(phi0), t(dphi0))
    torch.cuda.synchronize()
    assert_close(pbest.cpu().numpy(), want_phi, "merit at the chosen step")
    assert np.array_equal(ab.cpu().numpy(), want_a), "chosen step sizes"
    assert_bitexact(Cn.cpu().numpy(), X + want_a[:, None] * D, "updated coefficients")
    pb.close()
