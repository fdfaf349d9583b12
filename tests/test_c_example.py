"""examples/kincar_batch.c: a plain C host program on the batched C ABI, built with gcc."""
import os
import re
import subprocess

import numpy as np
import pytest

from common import assert_close
from ntg_b200 import build, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    if not os.path.exists(build.CORE_SO) or not os.path.exists(build.pack_so("kincar")):
        build.build_all()
    exe = str(tmp_path / "kincar_batch")
    subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "kincar_batch.c"), "-L", build.LIB, "-lntgpack_kincar",
                           "-lntg_b200", f"-Wl,-rpath,{build.LIB}", "-o", exe])
    return exe


def _batch(P):
    """the C program's LCG, reproduced"""
    X = np.empty((P, 14))
    lcg = 12345
    for p in range(P):
        for e in range(14):
            lcg = (lcg * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            u = (lcg >> 11) / 9007199254740992.0
            X[p, e] = 40.0 * u if e < 7 else 4.0 * u - 2.0
    return X


def test_c_program_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = subprocess.run([exe, "8"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stderr   # pack found by host address, then no device


@pytest.mark.gpu
def test_c_program_matches_oracle(tmp_path, port):
    exe = _build(tmp_path)
    P = 1000
    r = subprocess.run([exe, str(P)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    spec = configs.kincar(64)
    X = _batch(P)
    o = port.eval(spec, X, dense=False, band=True)
    rows = re.findall(r"p (\d+) f (\S+) g0 (\S+) c0 (\S+) J0 (\S+) viol (\S+)", r.stdout)
    assert len(rows) >= 4
    from common import violation
    viol = violation(spec, o["c"])
    for p, f, g0, c0, J0, v in rows:
        p = int(p)
        got = np.array([float(f), float(g0), float(c0), float(J0), float(v)])
        # J0: first band value of the problem = trajectory row (m=0, bp=0), slot 0
        want = np.array([o["f"][p], o["g"][p, 0], o["c"][p, 0], o["Jband"][p, 0, 0], viol[p]])
        assert_close(got, want, f"problem {p}")
