"""Tiny end-to-end cases for compute-sanitizer (memcheck / racecheck): every kernel, small batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntg_b200 import configs, Problem, JAC_BAND, JAC_DENSE

def run(tag, spec, P, fast=False, jac=JAC_BAND, env=None):
    if env: os.environ["NTG_B200_KERNEL"] = env
    else: os.environ.pop("NTG_B200_KERNEL", None)
    pb = Problem(spec, 0, fast=fast)
    X = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (P, spec.nC))).cuda()
    o = pb.eval(X, jac=jac, want_Z=True)
    torch.cuda.synchronize()
    print(tag, float(o["f"].sum()), flush=True)
    pb.close()

run("K1s endpoint", configs.endpoint(), 5)
run("K1s kincar dense", configs.kincar(20), 13, jac=JAC_DENSE)
run("K1s kincar64 fast", configs.kincar(64), 9, fast=True)
run("K1 endpoint (forced general)", configs.endpoint(), 5, env="general")
run("K1 syn6 small", configs.syn6(12, name="s"), 3)
run("K1c syn6 301 bps", configs.syn6(150, name="s"), 3)
run("K1c syn6 601 bps (cluster of 4)", configs.syn6(300, name="s"), 2)
spec = configs.endpoint()
pb = Problem(spec, 0)
X = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (4, spec.nC))).cuda()
lin, viol = pb.eval_linear(X)
t = torch.linspace(0, 2, 7, dtype=torch.float64, device="cuda")
s = pb.spline_interp(X, t)
ab, ph, Cn = pb.linesearch(X, torch.ones_like(X), torch.tensor([1.0, 0.5], dtype=torch.float64, device="cuda"), 2.0)
torch.cuda.synchronize()
print("aux ok", float(lin.sum()), float(s.sum()), float(ph.sum()))
