"""The C-ABI shared library loads without a GPU, exports every symbol the
public headers declare, registers its callback packs, validates arguments, and
fails LOUDLY (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ntg_b200 import build, configs
from ntg_b200.abi import BuiltSetup, NtgbPack

INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")


def declared_functions(header):
    txt = open(os.path.join(INCLUDE, header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)          # preprocessor lines
    txt = re.sub(r"(__host__ __device__\s*)?static inline[^{;]*\{.*?^\}", "", txt, flags=re.S | re.M)  # inline helpers: not exported
    txt = re.sub(r"typedef[^;]*\(\*[^;]*;", "", txt)       # function-pointer typedefs
    txt = re.sub(r"\w+ \(\*\w+\)\([^)]*\)", "void *cb", txt)  # callback parameters / members
    return sorted(set(re.findall(r"\b(\w+)\s*\([^;{]*\)\s*;", txt)) - {"defined"})


def test_core_exports_every_declared_symbol(built_libs):
    lib = C.CDLL(build.CORE_SO)
    names = declared_functions("ntg_b200.h") + declared_functions("ntg.h")
    assert "ntgb_eval" in names and "ntg" in names and "SplineInterp" in names and len(names) >= 28
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_packs_register_with_exact_shapes(built_libs):
    from ntg_b200 import problem
    want = {"vdp": ([3], 5, (0, 1, 0)), "kincar": ([3, 3], 5, (0, 2, 0)),
            "syn6": ([4] * 6, 8, (0, 4, 0)), "endpt": ([3, 2], 6, (2, 3, 1))}
    for name, (md, mo, cnt) in want.items():
        for suffix, exact in (("", 1), ("_fast", 0)):
            problem.load_pack(name + suffix)
            pk = problem.core().ntgb_find_pack((name + suffix).encode()).contents
            assert pk.max_nout == len(md) and list(pk.maxderiv)[: len(md)] == md
            assert pk.max_order == mo and (pk.max_nnlic, pk.max_nnltc, pk.max_nnlfc) == cnt
            assert pk.exact == exact and pk.launch
    assert not problem.core().ntgb_find_pack(b"no_such_pack")


def _create(spec, packname, device=0):
    from ntg_b200 import problem
    lib = problem.load_pack(packname)
    bs = BuiltSetup(spec, lambda role, sym: C.cast(getattr(lib, sym), C.c_void_p).value)
    h = C.c_void_p()
    rc = problem.core().ntgb_create(C.byref(h), bs.ref(), device)
    return rc, problem.core().ntgb_last_error().decode(), h


def test_argument_validation_happens_before_cuda(built_libs):
    spec = configs.vanderpol()
    spec.order = [25]
    rc, msg, _ = _create(spec, "vdp")
    assert rc == -1 and "order" in msg                      # NTGB_EINVAL: PGS limit order <= 20
    spec = configs.vanderpol()
    spec.bps = spec.bps.copy()
    spec.bps[3] = -0.25
    rc, msg, _ = _create(spec, "vdp")
    assert rc == -1 and "before the first knot" in msg
    spec = configs.vanderpol()
    spec.trajectorycostav = [(0, 7)]
    rc, msg, _ = _create(spec, "vdp")
    assert rc == -1 and "trajectorycostav" in msg
    spec = configs.vanderpol()
    spec.maxderiv = [2]
    spec.trajectorycostav = [(0, 0)]
    spec.trajectoryconstrav = [(0, 0)]
    spec.lic, spec.lfc = None, None
    rc, msg, _ = _create(spec, "vdp")
    assert rc == -4 and "maxderiv" in msg                   # NTGB_ELIMIT: pack compiled for maxderiv 3
    spec = configs.kincar()
    rc, msg, _ = _create(spec, "kincar")
    # same callbacks resolved from a DIFFERENT pack object are unknown to the registry
    from ntg_b200 import problem
    other = problem.load_pack("vdp")
    bs = BuiltSetup(configs.vanderpol(), lambda role, sym: 0x1234)
    h = C.c_void_p()
    rc = problem.core().ntgb_create(C.byref(h), bs.ref(), 0)
    assert rc == -2 and "pack" in problem.core().ntgb_last_error().decode()   # NTGB_ENOPACK


def test_no_gpu_means_loud_failure_not_fallback(built_libs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    rc, msg, _ = _create(configs.vanderpol(), "vdp")
    assert rc == -3 and "no CPU path" in msg                # NTGB_ECUDA
    from ntg_b200 import Problem
    from ntg_b200.problem import NtgError
    with pytest.raises(NtgError):
        Problem(configs.kincar())


def test_missing_library_raises(monkeypatch, tmp_path):
    from ntg_b200 import problem
    monkeypatch.setattr(problem, "_LIBDIR", str(tmp_path))
    monkeypatch.setattr(problem, "_core", None)
    monkeypatch.setattr(problem, "_packs", {})
    with pytest.raises(problem.NtgError, match="no CPU fallback|missing"):
        problem.core()


def test_dropin_host_helpers(built_libs):
    """linspace / DoubleMatrix / MakeMatrix of the drop-in surface (host-only helpers)"""
    lib = C.CDLL(build.CORE_SO)
    v = np.zeros(20)
    lib.linspace(v.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(0.0), C.c_double(5.0), 20)
    from ntg_b200.abi import linspace
    assert np.array_equal(v, linspace(0, 5, 20)) and v[-1] == 5.000000000000001
    lib.DoubleMatrix.restype = C.POINTER(C.POINTER(C.c_double))
    m = lib.DoubleMatrix(3, 4)
    assert all(m[i][j] == 0.0 for i in range(3) for j in range(4))
    m[2][3] = 7.0
    assert m[0][11] == 7.0                                   # one contiguous block, row pointers
    lib.FreeDoubleMatrix(m)


def test_ctypes_structs_match_the_header(tmp_path):
    """the ctypes mirrors in ntg_b200/abi.py must have the C structs' size and field offsets
    (gcc on include/ntg_b200.h says what those are)"""
    import ctypes as C
    import subprocess
    from ntg_b200 import abi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = {"ntgb_eval_args": abi.NtgbEvalArgs, "ntgb_solve_opts": abi.SolveOpts, "ntgb_nlp_opts": abi.NlpOpts,
             "ntgb_dims": abi.NtgbDims, "ntgb_setup": abi.NtgbSetup, "ntgb_pack": abi.NtgbPack}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ntg_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    want = {}
    for line in out.splitlines():
        s, f, v = line.split()
        want[(s, f)] = int(v)
    for cname, cls in pairs.items():
        assert C.sizeof(cls) == want[(cname, "size")], f"sizeof({cname})"
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == want[(cname, fname)], f"offsetof({cname}, {fname})"
