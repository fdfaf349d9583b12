"""Scratch kernel timing sweep (development aid; bench.py is the contract)."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntg_b200 import configs, Problem, JAC_BAND, JAC_DENSE

ap = argparse.ArgumentParser()
ap.add_argument("--cfgs", default="cfg2,cfg3,cfg4,cfg5")
ap.add_argument("--variants", default="exact,fast")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--p5", type=int, default=4096)
ap.add_argument("--p4", type=int, default=0, help="problems for cfg4 (default: the config's 65536)")
ap.add_argument("--dense", action="store_true")
ap.add_argument("--graph", action="store_true", help="time ONE CUDA graph of --iters launches (what bench.py does)")
ap.add_argument("--mode_obj", type=int, default=2)
ap.add_argument("--mode_con", type=int, default=2)
a = ap.parse_args()
HBM = 6529.1
L2 = 126e6
for cfg in a.cfgs.split(","):
    spec, P = configs.get(cfg)
    if cfg == "cfg5": P = a.p5
    if cfg == "cfg4" and a.p4: P = a.p4
    X = torch.from_numpy(configs.coefficients(cfg, P, spec)).cuda()
    for var in a.variants.split(","):
        pb = Problem(spec, 0, fast=(var == "fast"))
        jac = JAC_DENSE if a.dense else JAC_BAND
        bytes_eval = spec.bytes_per_eval(dense=False)
        per_set = P * spec.bytes_per_eval(dense=a.dense)
        nset = max(1, int(np.ceil(2 * L2 / per_set))) if per_set < 2 * L2 else 1
        nset = min(nset, 64)
        sets = [(X.clone(), pb.alloc_outputs(P, jac)) for _ in range(nset)]
        args = [pb.eval_args(x, o, a.mode_obj, a.mode_con, jac, 0, torch.cuda.current_stream().cuda_stream) for x, o in sets]
        for i in range(3): pb.launch(args[i % nset])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if a.graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cs = torch.cuda.current_stream().cuda_stream
                for i in range(a.iters):
                    x, o = sets[i % nset]
                    pb.launch(pb.eval_args(x, o, a.mode_obj, a.mode_con, jac, 0, cs))
            g.replay(); torch.cuda.synchronize()
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        else:
            e0.record()
            for i in range(a.iters): pb.launch(args[i % nset])
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        gbs = P * bytes_eval / (ms * 1e-3) / 1e9
        print(f"{cfg} {var:5s} P={P} {'dense' if a.dense else 'band'} sets={nset} {ms*1e3:9.1f} us/launch  "
              f"{P/(ms*1e-3):.4g} evals/s  {gbs:7.1f} GB/s algorithmic  frac={gbs/HBM:.3f}", flush=True)
        del sets, args; pb.close(); torch.cuda.empty_cache()
