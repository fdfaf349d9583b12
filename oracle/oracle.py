"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python (ctypes) front end of the two CPU oracles:

  * `Oracle("ref")`  -> oracle/_ref/libntg_ref.so: the UNMODIFIED reference C
    sources compiled from /root/reference (oracle/Makefile `ref`), driven
    through the fake npsol_ of oracle/ref_driver.c.
  * `Oracle("port")` -> oracle/libntg_oracle.so: our plain-C restatement of the
    same algorithm (oracle/ntg_oracle.c), buildable anywhere with gcc.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.  Nothing under ntg_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from ntg_b200.abi import (AV, BuiltSetup, NtgbDims, NtgbSetup, ProblemSpec, c_double_p, c_int_p)

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libntg_ref.so")
REF_G_SO = os.path.join(HERE, "_ref", "libntg_ref_g.so")
PORT_SO = os.path.join(HERE, "libntg_oracle.so")


def build(which: str = "all") -> None:
    """(Re)build the oracle libraries; `ref` is skipped when /root/reference is absent."""
    subprocess.check_call(["make", "-s", "-C", HERE, which])


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(c_double_p)


class Oracle:
    def __init__(self, kind: str = "port", shipped_flags: bool = False):
        self.kind = kind
        if kind == "ref":
            path = REF_G_SO if shipped_flags else REF_SO
            self.prefix = "ref_"
        elif kind == "port":
            path = PORT_SO
            self.prefix = "port_"
            if not os.path.exists(path):
                build("port")
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            raise FileNotFoundError(f"oracle library missing: {path} (run `make -C oracle`)")
        self.lib = C.CDLL(path)
        self._setups = {}

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # the oracle library carries the gcc build of every pack's callbacks
    def resolve(self, role: str, sym: str) -> int:
        return C.cast(getattr(self.lib, sym), C.c_void_p).value

    def setup(self, spec: ProblemSpec) -> BuiltSetup:
        key = id(spec)
        if key not in self._setups:
            self._setups[key] = BuiltSetup(spec, self.resolve)
        return self._setups[key]

    def dims(self, spec: ProblemSpec) -> NtgbDims:
        d = NtgbDims()
        self.fn("dims")(self.setup(spec).ref(), C.byref(d))
        return d

    def tables(self, spec: ProblemSpec):
        """-> (B list per output [nbps][order][maxderiv], offset [nout][nbps], col0 [ncnln][nout])"""
        d = self.dims(spec)
        tot = sum(spec.nbps * k * m for k, m in zip(spec.order, spec.maxderiv))
        B = np.zeros(tot)
        off = np.zeros((spec.nout, spec.nbps), dtype=np.int32)
        col0 = np.zeros((max(d.ncnln, 1), spec.nout), dtype=np.int32)
        self.fn("tables")(self.setup(spec).ref(), _dp(B), off.ctypes.data_as(c_int_p),
                          col0.ctypes.data_as(c_int_p))
        out, pos = [], 0
        for k, m in zip(spec.order, spec.maxderiv):
            n = spec.nbps * k * m
            out.append(B[pos:pos + n].reshape(spec.nbps, k, m).copy())
            pos += n
        return out, off, col0[:d.ncnln]

    def updateZ(self, spec: ProblemSpec, Cvec: np.ndarray, av, kind: int) -> np.ndarray:
        Z = np.zeros(spec.nZ)
        arr = (AV * max(len(av), 1))(*[AV(o, dd) for o, dd in av])
        Cvec = np.ascontiguousarray(Cvec, dtype=np.float64)
        self.fn("updateZ")(self.setup(spec).ref(), _dp(Cvec), arr, len(av), kind, _dp(Z))
        return Z

    def spline_interp(self, x, knots, coefs, order, mult, maxderiv) -> np.ndarray:
        f = np.zeros(maxderiv)
        knots = np.ascontiguousarray(knots, dtype=np.float64)
        coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        self.fn("spline_interp")(_dp(f), C.c_double(float(x)), _dp(knots), len(knots) - 1,
                                 _dp(coefs), len(coefs), order, mult, maxderiv)
        return f

    def eval(self, spec: ProblemSpec, X: np.ndarray, mode_obj: int = 2, mode_con: int = 2,
             dense: bool = True, band: bool = True, linear: bool = False, reps: int = 1,
             outputs: bool = True):
        """Evaluate the batch X [P][nC].  Returns a dict with f, g, c, Jdense
        ([P][nC][ncnln]: column-major per problem as NPSOL sees it, NaN where
        the reference never writes), Jband ([P][ncnln][S]), pattern_bad,
        seconds (best of `reps` passes, funcon+funobj only)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        P, nC = X.shape
        assert nC == spec.nC
        ncnln, S = spec.ncnln, spec.sorder
        r = {}
        f = np.zeros(P) if outputs else None
        g = np.zeros((P, nC)) if outputs else None
        c = np.zeros((P, max(ncnln, 1))) if outputs else None
        Jd = np.full((P, nC, max(ncnln, 1)), np.nan) if (outputs and dense and ncnln) else None
        Jb = np.full((P, max(ncnln, 1), S), np.nan) if (outputs and band and ncnln) else None
        A = np.zeros((nC, max(spec.nclin, 1))) if linear else None
        nb = nC + spec.nclin + ncnln
        bl = np.zeros(nb) if linear else None
        bu = np.zeros(nb) if linear else None
        bad = C.c_long(0)
        secs = C.c_double(0.0)
        rc = self.fn("eval")(self.setup(spec).ref(), P, _dp(X), mode_obj, mode_con, _dp(f), _dp(g),
                             _dp(c), _dp(Jd), _dp(Jb), C.byref(bad), _dp(A), _dp(bl), _dp(bu),
                             reps, C.byref(secs))
        assert rc == 0
        r.update(f=f, g=g, c=None if c is None else c[:, :ncnln], Jdense=Jd, Jband=Jb,
                 pattern_bad=bad.value, seconds=secs.value)
        if linear:
            r.update(A=A[:, :spec.nclin].T.copy(), bl=bl, bu=bu)  # A as [nclin][nC]
        return r


SHIM_SO = os.path.join(HERE, "libnpsol_shim.so")


def _run_main(shimlib, main_ptr, X: np.ndarray, mode_obj: int, mode_con: int, ncnln: int = 0):
    X = np.ascontiguousarray(X, dtype=np.float64)
    P, n = X.shape
    f = np.zeros(P)
    g = np.zeros((P, n))
    c = np.zeros((P, max(ncnln, 1)))
    Jd = np.zeros((P, n, max(ncnln, 1))) if ncnln else None
    dims = (C.c_int * 3)()
    A = np.zeros((n, 64))
    bl = np.zeros(n + 64 + ncnln)
    bu = np.zeros(n + 64 + ncnln)
    shimlib.shim_run_main.restype = C.c_int
    calls = shimlib.shim_run_main(C.c_void_p(main_ptr), P, _dp(X), mode_obj, mode_con, _dp(f), _dp(g),
                                  _dp(c), _dp(Jd), _dp(A), _dp(bl), _dp(bu), dims)
    n_, nclin, ncn = dims[0], dims[1], dims[2]
    return dict(calls=calls, f=f, g=g, c=c[:, :ncn], Jdense=Jd, n=n_, nclin=nclin, ncnln=ncn,
                A=A.reshape(-1)[:n_ * nclin].reshape(n_, nclin).T.copy() if nclin else None,
                bl=bl[:n_ + nclin + ncn], bu=bu[:n_ + nclin + ncn])


def run_reference_example(name: str, X: np.ndarray, mode_obj: int = 2, mode_con: int = 2):
    """Run the reference's UNMODIFIED examples/<name>.c main() (built into
    oracle/_ref/libref_<name>.so) linked against the unmodified reference
    library; its ntg() call lands in the fake npsol_ which evaluates batch X."""
    ex = C.CDLL(os.path.join(HERE, "_ref", f"libref_{name}.so"))
    ref = C.CDLL(REF_SO)
    main = C.cast(getattr(ex, f"ref_{name}_main"), C.c_void_p).value
    import tempfile
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:   # examples/vanderpol.c:192 writes ./coef1
        os.chdir(td)
        try:
            return _run_main(ref, main, X, mode_obj, mode_con)
        finally:
            os.chdir(cwd)


def run_product_main(main_ptr: int, X: np.ndarray, mode_obj: int = 2, mode_con: int = 2, ncnln: int = 0):
    """Run a user program's main() that calls THIS repo's ntg(): the fake NPSOL
    is made visible process-wide (RTLD_GLOBAL) so ntg()'s dlsym finds it."""
    if not os.path.exists(SHIM_SO):
        build("port")
    shim = C.CDLL(SHIM_SO, mode=C.RTLD_GLOBAL)
    return _run_main(shim, main_ptr, X, mode_obj, mode_con, ncnln)
