"""Pins the PGS restatement (oracle/pgs_restated.c), where the reference ships
no vectors ("parity unpinned"):

  * exact-rational Cox-de Boor recursion on the exact double inputs,
  * partition of unity / derivative sums,
  * off == left_aug - order,
  * the known answers derived from examples/vanderpol.c defaults,
  * the construction examples/vanderpol.m:15-25 describes (augknt + spcol).
"""
from fractions import Fraction

import numpy as np
import pytest

from ntg_b200 import configs
from ntg_b200.abi import linspace


def aug_knots(knots, order, mult):
    """knot sequence PGS `knots` builds: ends x order, interior x (order-mult)
    (same as MATLAB augknt(knots, order, order-mult), examples/vanderpol.m:17)"""
    t = [knots[0]] * order
    for b in knots[1:-1]:
        t += [b] * (order - mult)
    t += [knots[-1]] * order
    return np.array(t)


def exact_basis(t, k, x, left, nderiv):
    """Textbook Cox-de Boor in exact rational arithmetic, restricted to the
    polynomial piece of interval `left` (1-based) -- also valid past the last
    knot, where de Boor's bsplvd extrapolates that piece.
    Returns D[d][i], i = 0..k-1 for B-splines left-k+1..left (1-based)."""
    T = [Fraction(float(v)) for v in t]
    X = Fraction(float(x))
    memo = {}

    def N(i, kk):  # B_{i,kk}(x), i 1-based index of first knot
        if (i, kk) in memo:
            return memo[(i, kk)]
        if kk == 1:
            v = Fraction(1) if i == left else Fraction(0)
        else:
            v = Fraction(0)
            d1 = T[i + kk - 2] - T[i - 1]
            d2 = T[i + kk - 1] - T[i]
            if d1 != 0:
                v += (X - T[i - 1]) / d1 * N(i, kk - 1)
            if d2 != 0:
                v += (T[i + kk - 1] - X) / d2 * N(i + 1, kk - 1)
        memo[(i, kk)] = v
        return v

    def dN(i, kk, d):  # d-th derivative
        if d == 0:
            return N(i, kk)
        if kk == 1:
            return Fraction(0)
        v = Fraction(0)
        d1 = T[i + kk - 2] - T[i - 1]
        d2 = T[i + kk - 1] - T[i]
        if d1 != 0:
            v += (kk - 1) / d1 * dN(i, kk - 1, d - 1)
        if d2 != 0:
            v -= (kk - 1) / d2 * dN(i + 1, kk - 1, d - 1)
        return v

    return [[dN(left - k + 1 + i, k, d) for i in range(k)] for d in range(nderiv)]


def interv_ref(t, x):
    """de Boor-site interv rule, straight from its specification."""
    t = np.asarray(t)
    if x < t[0]:
        return 1
    last = t[-1]
    cand = [i + 1 for i in range(len(t)) if t[i] < last and t[i] <= x]
    return max(cand) if cand else 1


SHAPES = {
    "vanderpol": lambda: configs.vanderpol(20),
    "kincar64": lambda: configs.kincar(64),
    "syn6_small": lambda: configs.syn6(12, name="syn6_small"),
    "endpoint": lambda: configs.endpoint(),
    "order20": lambda: configs.high_order(20, 3, 3, 17),   # PGS limit: bsplvb jmax = 20
    "order20_mult19": lambda: configs.high_order(20, 19, 5, 23),
}


@pytest.mark.parametrize("shape", list(SHAPES))
def test_tables_against_exact_rational(port, shape):
    spec = SHAPES[shape]()
    B, off, _ = port.tables(spec)
    for j in range(spec.nout):
        k, m, md = spec.order[j], spec.mult[j], spec.maxderiv[j]
        t = aug_knots(spec.knots[j], k, m)
        n = spec.ncoef[j]
        assert len(t) == n + k
        for bp in range(spec.nbps):
            x = spec.bps[bp]
            left = interv_ref(t, x)
            # off = (left_knots-1)*(order-mult) == left_aug - order  
            assert off[j, bp] == left - k, (shape, j, bp)
            lk = interv_ref(spec.knots[j], x)
            assert off[j, bp] == (lk - 1) * (k - m)
            D = exact_basis(t, k, x, left, md)
            for d in range(md):
                row = np.array([float(v) for v in D[d]])
                got = B[j][bp, :, d]
                scale = np.max(np.abs(row))
                assert np.max(np.abs(got - row)) <= 4e-15 * scale, (shape, j, bp, d, got, row)


@pytest.mark.parametrize("shape", list(SHAPES))
def test_partition_of_unity_and_derivative_sums(port, shape):
    spec = SHAPES[shape]()
    B, _, _ = port.tables(spec)
    for j in range(spec.nout):
        for d in range(spec.maxderiv[j]):
            s = B[j][:, :, d].sum(axis=1)
            scale = np.abs(B[j][:, :, d]).max(axis=1)
            assert np.all(np.abs(s - (1.0 if d == 0 else 0.0)) <= 1e-13 * np.maximum(scale, 1.0))


def test_known_answers_vanderpol(port):
    """values derived from examples/vanderpol.c defaults."""
    spec = configs.vanderpol(20)
    B, off, _ = port.tables(spec)
    t = aug_knots(spec.knots[0], 5, 3)
    assert np.array_equal(t, [0] * 5 + [2.5] * 2 + [5] * 5)
    assert np.array_equal(off[0], [0] * 10 + [2] * 10)
    assert np.array_equal(B[0][0].T, [[1, 0, 0, 0, 0], [-1.6, 1.6, 0, 0, 0], [1.92, -3.84, 1.92, 0, 0]])
    last = B[0][19].T
    np.testing.assert_allclose(last, [[0, 0, 0, 0, 1], [0, 0, 0, -1.6, 1.6], [0, 0, 1.92, -3.84, 1.92]],
                               atol=3e-15)
    assert spec.bps[19] == 5.000000000000001  # linspace accumulates (quirk Q1)
    assert linspace(0, 5, 64)[-1] == 4.999999999999999


def test_spcol_construction_scipy(port):
    """examples/vanderpol.m:15-25: augknt + spcol must reproduce the tables
    (scipy BSpline as spcol; coarse tolerance, scipy's derivatives are less accurate)."""
    from scipy.interpolate import BSpline
    spec = configs.kincar(64)
    B, off, _ = port.tables(spec)
    j, k, m, md = 0, 5, 3, 3
    t = aug_knots(spec.knots[j], k, m)
    n = spec.ncoef[j]
    for d in range(md):
        dense = np.zeros((spec.nbps, n))
        for i in range(n):
            coef = np.zeros(n)
            coef[i] = 1.0
            dense[:, i] = BSpline(t, coef, k - 1, extrapolate=True).derivative(d)(spec.bps) if d else \
                BSpline(t, coef, k - 1, extrapolate=True)(spec.bps)
        for bp in range(spec.nbps):
            o = off[j, bp]
            np.testing.assert_allclose(B[j][bp, :, d], dense[bp, o:o + k], rtol=1e-9, atol=1e-9)
            outside = np.delete(dense[bp], np.arange(o, o + k))
            assert np.all(np.abs(outside) < 1e-9)


def test_spline_interp_matches_tables(port):
    """SplineInterp (reference src/colloc.c:449-484) at the breakpoints equals Z"""
    spec = configs.endpoint()
    rng = np.random.default_rng(3)
    Cv = rng.uniform(-1, 1, spec.nC)
    av = [(j, d) for j in range(spec.nout) for d in range(spec.maxderiv[j])]
    Z = port.updateZ(spec, Cv, av, 1)
    iC = np.concatenate([[0], np.cumsum(spec.ncoef)])
    iz = np.concatenate([[0], np.cumsum(spec.maxderiv)])
    for j in range(spec.nout):
        for bp in (0, 3, spec.nbps - 1):
            f = port.spline_interp(spec.bps[bp], spec.knots[j], Cv[iC[j]:iC[j + 1]], spec.order[j],
                                   spec.mult[j], spec.maxderiv[j])
            zz = Z[iz[j] * spec.nbps + bp * spec.maxderiv[j]: iz[j] * spec.nbps + (bp + 1) * spec.maxderiv[j]]
            assert np.array_equal(f, zz)
