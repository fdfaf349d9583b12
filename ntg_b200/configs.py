"""Named problem families: BASELINE.json's configs made concrete (SURVEY.md
section 8, "Configs made concrete") plus small test-only shapes.

Spline parameters of CFG-1..4 are the reference examples' own
(examples/vanderpol.c:17-21,133; examples/kincar.c:133-137,211); both shipped
examples have zero nonlinear constraints, so the constraint packs VDP-C / KC-C
(ntg_b200/packs/*.c) are added to exercise "constr+Jacobian".
"""
from __future__ import annotations

import numpy as np

from .abi import ProblemSpec

_INF = 1.0e20  # "no bound" in the compact bound vectors


def vanderpol(nbps: int = 20, constraints: bool = True, name: str = "cfg2_vanderpol") -> ProblemSpec:
    """CFG-1/CFG-2: 1 output, 2 intervals, order 5, mult 3, maxderiv 3 -> nC = 7.
    Linear constraints are the example's own (examples/vanderpol.c:159-169)."""
    lic = np.zeros((2, 3)); lic[0, 0] = 1.0; lic[1, 1] = 1.0
    lfc = np.zeros((1, 3)); lfc[0, 0] = -1.0; lfc[0, 1] = 1.0
    nn = 1 if constraints else 0
    lower = np.array([1.0, 0.0, 1.0] + [-2.0] * nn)
    upper = np.array([1.0, 0.0, 1.0] + [2.0] * nn)
    return ProblemSpec(
        name=name, pack="vdp", order=[5], mult=[3], maxderiv=[3], ninterv=[2], nbps=nbps,
        callbacks={"ucf": "vdp_ucf", "nltcf": "vdp_nltcf" if constraints else ""},
        nucf=1, nnltc=nn,
        trajectorycostav=[(0, 0), (0, 1), (0, 2)],
        trajectoryconstrav=[(0, 0), (0, 1), (0, 2)] if constraints else [],
        lic=lic, lfc=lfc, lowerb=lower, upperb=upper)


def kincar(nbps: int = 20, constraints: bool = True, name: str = "cfg3_kincar") -> ProblemSpec:
    """CFG-3 (nbps 20) / CFG-4 (nbps 64): 2 outputs, 2/5/3/3 each -> nC = 14.
    Linear constraints pin the full flat flag at both ends like the example's
    lane change (examples/kincar.c:319-339)."""
    lic = np.eye(6)
    lfc = np.eye(6)
    x0 = [0.0, 8.0, 0.0, -2.0, 0.0, 0.0]
    xf = [40.0, 8.0, 0.0, 2.0, 0.0, 0.0]
    nn = 2 if constraints else 0
    lower = np.array(x0 + xf + ([0.0, -50.0] if constraints else []))
    upper = np.array(x0 + xf + ([400.0, 50.0] if constraints else []))
    return ProblemSpec(
        name=name, pack="kincar", order=[5, 5], mult=[3, 3], maxderiv=[3, 3], ninterv=[2, 2],
        nbps=nbps,
        callbacks={"ucf": "kc_ucf", "nltcf": "kc_nltcf" if constraints else ""},
        nucf=1, nnltc=nn,
        trajectorycostav=[(0, 2), (1, 2)],
        trajectoryconstrav=[(0, 1), (0, 2), (1, 1), (1, 2)] if constraints else [],
        lic=lic, lfc=lfc, lowerb=lower, upperb=upper)


def syn6(ninterv: int = 200, nbps: int | None = None, name: str = "cfg5_syn6") -> ProblemSpec:
    """CFG-5: 6 outputs, order 8, mult 4, maxderiv 4, 200 intervals, nbps = 2*ninterv+1
    -> nC = 4824, ncnln = 1604."""
    if nbps is None:
        nbps = 2 * ninterv + 1
    av = [(j, d) for j in range(6) for d in range(4)]
    return ProblemSpec(
        name=name, pack="syn6", order=[8] * 6, mult=[4] * 6, maxderiv=[4] * 6,
        ninterv=[ninterv] * 6, nbps=nbps,
        callbacks={"ucf": "syn6_ucf", "nltcf": "syn6_nltcf"},
        nucf=1, nnltc=4, trajectorycostav=av, trajectoryconstrav=av,
        lowerb=np.array([-1.0, -2.0, -3.0, -4.0]), upperb=np.array([1.0, 2.0, 3.0, 4.0]))


def endpoint(nbps: int = 13, name: str = "test_endpoint") -> ProblemSpec:
    """Test-only: every callback kind, two outputs with DIFFERENT order / mult /
    maxderiv / interval count, non-uniform breakpoints, linear constraints of all
    three kinds."""
    rng = np.random.default_rng(77)
    bps = np.sort(np.concatenate([[0.0, 2.0], rng.uniform(0.0, 2.0, nbps - 2)]))
    knots = [np.array([0.0, 0.5, 1.25, 2.0]), np.array([0.0, 0.8, 2.0])]
    nz = 5
    lic = rng.uniform(-1, 1, (2, nz))
    ltc = rng.uniform(-1, 1, (1, nz))
    lfc = rng.uniform(-1, 1, (2, nz))
    nb = 2 + 1 + 2 + 2 + 3 + 1
    lower = -np.arange(1, nb + 1, dtype=np.float64)
    upper = np.arange(1, nb + 1, dtype=np.float64) * 0.5
    av0 = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)]
    return ProblemSpec(
        name=name, pack="endpt", order=[6, 4], mult=[3, 2], maxderiv=[3, 2], ninterv=[3, 2],
        nbps=nbps, bps=bps, knots=knots, t0=0.0, t1=2.0,
        callbacks={"icf": "ep_icf", "ucf": "ep_ucf", "fcf": "ep_fcf",
                   "nlicf": "ep_nlicf", "nltcf": "ep_nltcf", "nlfcf": "ep_nlfcf"},
        nicf=1, nucf=1, nfcf=1, nnlic=2, nnltc=3, nnlfc=1,
        initialcostav=av0, trajectorycostav=av0, finalcostav=av0,
        initialconstrav=av0, trajectoryconstrav=av0, finalconstrav=av0,
        lic=lic, ltc=ltc, lfc=lfc, lowerb=lower, upperb=upper)


def high_order(order: int = 20, mult: int = 3, ninterv: int = 4, nbps: int = 33, name: str = "test_hi20") -> ProblemSpec:
    """Test-only: the upper limit of the spline machinery (order 20 = PGS bsplvb's jmax)."""
    av = [(0, 0), (0, 1), (0, 2)]
    return ProblemSpec(
        name=name, pack="hi20", order=[order], mult=[mult], maxderiv=[3], ninterv=[ninterv], nbps=nbps,
        t0=0.0, t1=1.0, callbacks={"ucf": "hi20_ucf", "nltcf": "hi20_nltcf"}, nucf=1, nnltc=1,
        trajectorycostav=av, trajectoryconstrav=av, lowerb=np.array([-1.0]), upperb=np.array([1.0]))


def conditional(nbps: int = 32, name: str = "test_cond") -> ProblemSpec:
    """Test-only: callbacks whose derivative sparsity depends on the data (packs/cond.c)."""
    av = [(j, d) for j in range(2) for d in range(3)]
    return ProblemSpec(
        name=name, pack="cond", order=[4, 4], mult=[2, 2], maxderiv=[3, 3], ninterv=[3, 3], nbps=nbps,
        t0=0.0, t1=1.0, callbacks={"ucf": "cond_ucf", "nltcf": "cond_nltcf"}, nucf=1, nnltc=2,
        trajectorycostav=av, trajectoryconstrav=av,
        lowerb=np.array([-1.0, -1.0]), upperb=np.array([1.0, 1.0]))


# BASELINE.json configs -> (spec factory, batch size, coefficient sampler)
def coefficients(cfg: str, P: int, spec: ProblemSpec, seed: int | None = None) -> np.ndarray:
    """Synthetic coefficient batches, seeds and ranges of SURVEY.md section 8(d)."""
    seeds = {"cfg1": 1001, "cfg2": 1002, "cfg3": 1003, "cfg4": 1004, "cfg5": 1005}
    rng = np.random.default_rng(seeds.get(cfg, 999) if seed is None else seed)
    nC = spec.nC
    if cfg in ("cfg3", "cfg4"):
        half = nC // 2
        X = np.empty((P, nC))
        X[:, :half] = rng.uniform(0.0, 40.0, (P, half))
        X[:, half:] = rng.uniform(-2.0, 2.0, (P, nC - half))
        return X
    if cfg == "cfg5":
        return rng.uniform(-1.0, 1.0, (P, nC))
    return rng.uniform(-2.0, 2.0, (P, nC))


CONFIGS = {
    # id: (factory, default batch)
    "cfg1": (lambda: vanderpol(20, constraints=True, name="cfg1_vanderpol_single"), 1),
    "cfg2": (lambda: vanderpol(20, constraints=True, name="cfg2_vanderpol_x4096"), 4096),
    "cfg3": (lambda: kincar(20, constraints=True, name="cfg3_kincar_multistart_x8192"), 8192),
    "cfg4": (lambda: kincar(64, constraints=True, name="cfg4_kincar_mpc_65536x64bps"), 65536),
    "cfg5": (lambda: syn6(200, name="cfg5_syn6_order8_200interv_x16384"), 16384),
}


def get(cfg: str):
    fac, P = CONFIGS[cfg]
    return fac(), P
