"""Development aid: device-side time stamps of the small-batch kernel (scratch build under
build/variants/stamps, tools/stamps_build.py; see DESIGN.md "small batches").  Prints, for the last
launches of a stream of back-to-back launches, when three CTAs (first, middle, last of the grid)
passed each phase boundary -- on one time axis, so the gaps BETWEEN launches show too."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NTG_B200_PACK_DIR"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "variants", os.environ.get("NTG_STAMPS_VARIANT", "stamps"))
import numpy as np, torch
from ntg_b200 import configs, Problem, JAC_BAND
from ntg_b200 import problem as _p

NAMES = ["start", "prol", "wait", "C", "A", "chains", "B", "end"]
ORDER = [0, 1, 2, 3, 4, 7, 5, 6]
cases = [("cfg2", 0), ("cfg3", 0)]
for cfg, Pover in cases:
    spec, P = configs.get(cfg)
    P = Pover or P
    pb = Problem(spec, 0, fast=True)
    lib = _p._packs[spec.pack + "_fast"]
    X = torch.from_numpy(configs.coefficients(cfg, P, spec)).cuda()
    nset = 24
    sets = [(X.clone(), pb.alloc_outputs(P, JAC_BAND)) for _ in range(nset)]
    for mode in ("graph",):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cs = torch.cuda.current_stream().cuda_stream
            for i, (x, o) in enumerate(sets):
                pb.launch(pb.eval_args(x, o, 2, 2, JAC_BAND, i, cs))
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        out = (C.c_ulonglong * 192)()
        lib.ntg_read_stamps(out)
        print(f"{cfg} P={P} {mode}: {e0.elapsed_time(e1) / nset * 1e3:.2f} us per launch")
        last = [(nset - 4 + k) for k in range(4)]
        t0 = min(out[(l & 7) * 24 + c * 8] for l in last[:1] for c in range(3))
        for l in last:
            for c, cname in enumerate(("first", "mid", "last")):
                t = [out[(l & 7) * 24 + c * 8 + i] for i in ORDER]
                print(f"  launch {l} {cname:5s} " + " ".join(f"{n}:{(t[i] - t0) / 1e3:7.2f}" for i, n in enumerate(NAMES)))
    pb.close()
