"""Size-independent properties of the algorithm, checked on the CPU oracle:
finite differences of f and c against g and J, linearity of the flat outputs
in the coefficients, trapezoid of a constant, and the Z layout contract."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from ntg_b200 import configs

SPECS = {
    "vanderpol": lambda: configs.vanderpol(20),
    "kincar": lambda: configs.kincar(20),
    "endpoint": lambda: configs.endpoint(),
    "syn6_small": lambda: configs.syn6(6, name="syn6_tiny"),
}


@pytest.mark.parametrize("name", list(SPECS))
def test_gradient_and_jacobian_by_central_differences(port, name):
    spec = SPECS[name]()
    rng = np.random.default_rng(42)
    x = rng.uniform(-1, 1, spec.nC)
    base = port.eval(spec, x[None, :])
    h = 1e-6
    Xp = np.tile(x, (spec.nC, 1)) + h * np.eye(spec.nC)
    Xm = np.tile(x, (spec.nC, 1)) - h * np.eye(spec.nC)
    rp, rm = port.eval(spec, Xp, dense=False, band=False), port.eval(spec, Xm, dense=False, band=False)
    g_fd = (rp["f"] - rm["f"]) / (2 * h)
    np.testing.assert_allclose(base["g"][0], g_fd, rtol=2e-6, atol=2e-6 * np.abs(g_fd).max())
    if spec.ncnln:
        J_fd = ((rp["c"] - rm["c"]) / (2 * h))          # [nC][ncnln]
        J = np.nan_to_num(base["Jdense"][0], nan=0.0)   # [nC][ncnln]
        np.testing.assert_allclose(J, J_fd, rtol=2e-6, atol=2e-6 * max(np.abs(J_fd).max(), 1.0))
        # entries the reference never writes are structurally zero
        assert np.all(np.abs(J_fd[np.isnan(base["Jdense"][0])]) < 1e-7)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), a=st.floats(-3, 3), b=st.floats(-3, 3))
def test_flat_outputs_are_linear_in_coefficients(port, seed, a, b):
    spec = configs.endpoint()
    rng = np.random.default_rng(seed)
    x, y = rng.uniform(-2, 2, spec.nC), rng.uniform(-2, 2, spec.nC)
    av = [(j, d) for j in range(spec.nout) for d in range(spec.maxderiv[j])]
    Zx, Zy = port.updateZ(spec, x, av, 1), port.updateZ(spec, y, av, 1)
    Zc = port.updateZ(spec, a * x + b * y, av, 1)
    scale = np.abs(a * Zx).max() + np.abs(b * Zy).max() + 1e-300
    assert np.abs(Zc - (a * Zx + b * Zy)).max() <= 1e-12 * scale


def test_updateZ_only_touches_listed_variables(port):
    """updateZ fills only the listed (output, deriv) pairs and, for the initial /
    final kinds, only the first / last breakpoint (reference src/colloc.c:344-367)."""
    spec = configs.endpoint()
    x = np.random.default_rng(1).uniform(-1, 1, spec.nC)
    iZ = np.concatenate([[0], np.cumsum(spec.maxderiv)]) * spec.nbps
    for kind, bps in ((0, [0]), (1, range(spec.nbps)), (2, [spec.nbps - 1])):
        Z = port.updateZ(spec, x, [(1, 1)], kind)
        touched = {iZ[1] + bp * spec.maxderiv[1] + 1 for bp in bps}
        nz = set(np.nonzero(Z)[0].tolist())
        assert nz <= touched and len(nz) == len(touched)


def test_constant_integrand_integrates_to_horizon(port):
    """C = ones -> z = 1, zd = zdd = 0 -> u = 1 -> integrand (z^2 + zd^2 + u^2)/2 = 1 -> I = t_end - t_0"""
    spec = configs.vanderpol(20, constraints=False)
    r = port.eval(spec, np.ones((1, 7)))
    assert abs(r["f"][0] - (spec.bps[-1] - spec.bps[0])) < 1e-14
