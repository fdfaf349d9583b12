"""Development aid: device-side time stamps of the small-batch kernel (scratch build under
build/variants/stamps, see DESIGN.md "small batches"); prints where one launch spends its time."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NTG_B200_PACK_DIR"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "variants", "stamps")
import numpy as np, torch
from ntg_b200 import configs, Problem, JAC_BAND
from ntg_b200 import problem as _p

for cfg in ("cfg2", "cfg3"):
    spec, P = configs.get(cfg)
    pb = Problem(spec, 0, fast=True)
    lib = _p._packs[spec.pack + "_fast"]
    X = torch.from_numpy(configs.coefficients(cfg, P, spec)).cuda()
    nset = 24
    sets = [(X.clone(), pb.alloc_outputs(P, JAC_BAND)) for _ in range(nset)]
    st = torch.cuda.current_stream().cuda_stream
    for mode in ("stream", "graph"):
        if mode == "graph":
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cs = torch.cuda.current_stream().cuda_stream
                for x, o in sets:
                    pb.launch(pb.eval_args(x, o, 2, 2, JAC_BAND, 0, cs))
            for _ in range(3):
                g.replay()
        else:
            for _ in range(3):
                for x, o in sets:
                    pb.launch(pb.eval_args(x, o, 2, 2, JAC_BAND, 0, st))
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 128)()
        lib.ntg_read_stamps(out)
        for blk, name in ((0, "first CTA"), (16, "last CTA")):
            t = [out[blk + i] for i in range(7)]
            names = ["start", "prologue done", "dep wait done", "C landed", "phase A done", "phase B done", "end"]
            print(cfg, mode, name, " ".join(f"{n}:+{(t[i]-t[0])}ns" for i, n in enumerate(names)))
        print(cfg, mode, "first->last CTA start skew", out[16] - out[0], "ns; kernel span", max(out[6], out[22]) - min(out[0], out[16]), "ns")
    pb.close()
