"""Scratch GPU parity sweep (development aid; the real checks live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from ntg_b200 import configs, Problem, JAC_BAND, JAC_DENSE
from oracle.oracle import Oracle

port = Oracle("port")
def relerr(a, b):
    scale = np.max(np.abs(b)) if b.size else 0.0
    return float(np.max(np.abs(a - b) / (1e-12 * np.abs(b) + 1e-14 * scale + 1e-300))) if b.size else 0.0

def check(tag, spec, X, fast):
    pb = Problem(spec, 0, fast=fast)
    Bd, offd, leftd = pb.tables()
    Bo, offo, col0o = port.tables(spec)
    tab_ok = all(np.array_equal(a, b) for a, b in zip(Bd, Bo)) and np.array_equal(offd, offo)
    col0d, jk0 = pb.pattern()
    pat_ok = np.array_equal(col0d, col0o)
    r = port.eval(spec, X, dense=(spec.ncnln * spec.nC < 200000))
    Xd = torch.from_numpy(X).cuda()
    o = pb.eval(Xd, jac=JAC_BAND, want_Z=True)
    torch.cuda.synchronize()
    f = o["f"].cpu().numpy(); g = o["g"].cpu().numpy(); c = o["c"].cpu().numpy()[:, :spec.ncnln]
    Jb = pb.band_to_rows(o["J"].cpu().numpy()) if o["J"] is not None else None
    res = dict(f=relerr(f, r["f"]), g=relerr(g, r["g"]), c=relerr(c, r["c"]) if spec.ncnln else 0.0,
               J=relerr(Jb, r["Jband"]) if Jb is not None else 0.0)
    exact = dict(f=np.array_equal(f, r["f"]), g=np.array_equal(g, r["g"]), c=np.array_equal(c, r["c"]),
                 J=Jb is None or np.array_equal(Jb, r["Jband"]))
    dn = ""
    if r["Jdense"] is not None:
        o2 = pb.eval(Xd, jac=JAC_DENSE)
        Jd = o2["J"].cpu().numpy()
        ref = np.nan_to_num(r["Jdense"], nan=0.0)
        dn = " dense_err=%.3g pattern=%s" % (relerr(Jd, ref), np.array_equal(Jd != 0, (~np.isnan(r["Jdense"])) & (ref != 0)))
    print(f"{tag:10s} fast={int(fast)} tables_bitexact={tab_ok} pattern={pat_ok} err/tol: "
          + " ".join(f"{k}={v:.3g}" for k, v in res.items()) + " bitexact: "
          + " ".join(f"{k}={int(v)}" for k, v in exact.items()) + dn, flush=True)
    pb.close()

for fast in (False, True):
    for cfg, P in (("cfg2", 257), ("cfg3", 300), ("cfg4", 130), ("cfg5", 5)):
        spec, _ = configs.get(cfg)
        check(cfg, spec, configs.coefficients(cfg, P, spec), fast)
    spec = configs.endpoint()
    check("endpoint", spec, np.random.default_rng(5).uniform(-1.5, 1.5, (77, spec.nC)), fast)
