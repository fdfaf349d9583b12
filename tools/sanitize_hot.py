"""Steady-state launches of K1s for compute-sanitizer memcheck: whole tiles dealt round-robin, the even
split with one / two / four tiles per CTA (tile buffers of exactly the largest tile's rows), the push
instantiation, the dense-layout instantiation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntg_b200 import configs, Problem, JAC_BAND, JAC_DENSE

def run(tag, cfg, P, fast=True, jac=JAC_BAND):
    spec, _ = configs.get(cfg)
    pb = Problem(spec, 0, fast=fast)
    X = torch.from_numpy(configs.coefficients(cfg, P, spec, seed=P)).cuda()
    o = pb.eval(X, jac=jac)
    torch.cuda.synchronize()
    print(tag, P, float(o["f"].sum()), float(o["result"].sum()), flush=True)
    pb.close()

run("cfg4 even, two tiles of 14 per CTA (push)", "cfg4", 8192)
run("cfg4 even, ragged", "cfg4", 8189)
run("cfg4 whole tiles, several per CTA (push)", "cfg4", 20011)
run("cfg3 even, one tile of 28 per CTA", "cfg3", 8192)
run("cfg2 even", "cfg2", 7105)
run("cfg4 dense steady state, even", "cfg4", 8192, jac=JAC_DENSE)
run("cfg4 exact, four tiles per CTA", "cfg4", 16384, fast=False)
print("done")
