"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED
reference (oracle/_ref, built by `make -C oracle ref` from /root/reference).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Each .npz holds the seeded inputs and the reference's own outputs for them:
tables, offsets, Jacobian pattern, f, g, c, band Jacobian, A, bl, bu.  The
reference-example fixtures come from running examples/vanderpol.c and
examples/kincar.c unmodified (their own callbacks, their own setup code).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ntg_b200 import configs  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {
    # name: (spec factory, P, coefficient sampler key)
    "cfg2_vanderpol": (lambda: configs.get("cfg2")[0], 8, "cfg2"),
    "cfg3_kincar": (lambda: configs.get("cfg3")[0], 8, "cfg3"),
    "cfg4_kincar64": (lambda: configs.get("cfg4")[0], 4, "cfg4"),
    "cfg5_syn6": (lambda: configs.get("cfg5")[0], 1, "cfg5"),
    "syn6_small": (lambda: configs.syn6(12, name="syn6_small"), 4, "cfg5"),
    "endpoint": (lambda: configs.endpoint(), 8, "endpoint"),
}


def case_inputs(name):
    fac, P, key = CASES[name]
    spec = fac()
    X = configs.coefficients(key, P, spec)
    if name == "cfg2_vanderpol":  # the two known-answer vectors of SURVEY.md section 8(c)
        X[0] = 1.0
        X[1] = [1, .5, -.25, .75, 2, -1, .125]
    return spec, X


def main():
    os.chdir(HERE)
    ref = oracle.Oracle("ref")
    for name in CASES:
        spec, X = case_inputs(name)
        r = ref.eval(spec, X, dense=False, band=True, linear=True)
        B, off, col0 = ref.tables(spec)
        assert r["pattern_bad"] == 0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X, f=r["f"], g=r["g"], c=r["c"],
                            Jband=r["Jband"], A=r["A"], bl=r["bl"], bu=r["bu"], off=off, col0=col0,
                            **{f"B{j}": b for j, b in enumerate(B)})
        print(name, X.shape, "->", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")
    # the reference's own example programs, unmodified
    rng = np.random.default_rng(2024)
    for ex, n in (("vanderpol", 7), ("kincar", 14)):
        X = rng.uniform(-2, 2, (6, n))
        X[0] = 1.0  # the shipped initial guess
        if ex == "vanderpol":
            X[1] = [1, .5, -.25, .75, 2, -1, .125]
        r = oracle.run_reference_example(ex, X)
        assert r["calls"] == 1
        np.savez_compressed(os.path.join(HERE, f"example_{ex}.npz"), X=X, f=r["f"], g=r["g"], A=r["A"],
                            bl=r["bl"], bu=r["bu"], dims=np.array([r["n"], r["nclin"], r["ncnln"]]))
        print("example", ex, r["f"][:2])
    if os.path.exists("coef1"):
        os.remove("coef1")


if __name__ == "__main__":
    main()
