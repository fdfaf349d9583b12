"""Small SQP solves for compute-sanitizer (memcheck / racecheck) -- the QP kernel's shared-memory phases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from ntg_b200 import configs, Problem
from test_gpu_next import _kincar_active_constraints

spec = _kincar_active_constraints()
pb = Problem(spec, 0)
X = torch.from_numpy(configs.coefficients("cfg3", 6, spec, seed=5)).cuda()
f, v, it, st, lam, ist = pb.solve_sqp(X, max_iter=int(os.environ.get("IT", "12")), multipliers=True)
print("kincar", st.tolist(), it.tolist(), flush=True)
pb.close()
spec = configs.endpoint()
pb = Problem(spec, 0)
X = torch.from_numpy(np.random.default_rng(3).uniform(-0.5, 0.5, (4, spec.nC))).cuda()
f, v, it, st = pb.solve_sqp(X, max_iter=int(os.environ.get("IT", "12")))
print("endpt", st.tolist(), it.tolist(), flush=True)
pb.close()
