"""2+ GPU check of the fused multi-GPU gather (run under torchrun, one rank per GPU):
the gathered tables written by peer stores from the evaluator's epilogue must equal, bit for bit,
what an NCCL all_gather of the local (objective, violation) tables gives; then both are timed.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/peer_gather_check.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch
import torch.distributed as dist

from ntg_b200 import JAC_BAND, Problem, configs
from ntg_b200.shard import PeerGather, gather_results, shard_range

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for cfg, Ptot in (("cfg4", 8192 * world), ("cfg2", 1000 * world + 3), ("cfg5", 8 * world)):
    spec, _ = configs.get(cfg)
    lo, hi = shard_range(Ptot, rank, world)
    X = torch.from_numpy(configs.coefficients(cfg, Ptot, spec)[lo:hi]).cuda()
    pb = Problem(spec, local, fast=True)
    pg = PeerGather(pb, Ptot)
    out = pb.alloc_outputs(hi - lo, JAC_BAND)
    st = torch.cuda.current_stream().cuda_stream
    pb.launch(pb.eval_args(X, out, 2, 2, JAC_BAND, 0, st, peers=pg))
    pg.fence()
    want = gather_results(out["result"], Ptot)
    got = pg.table()
    ok = torch.equal(got, want)
    allok = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(allok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{cfg}: P_total {Ptot}, {world} ranks: fused gather == NCCL all_gather on every rank: {bool(allok.item())}", flush=True)
    assert ok, f"rank {rank}: {cfg} tables differ"
    if cfg == "cfg4":
        # the other launch shapes: values only (mode 0, no Jacobian) goes through the general
        # result path of the same kernel; a second steady-state launch overwrites the tables
        pg.table().zero_()
        pg.fence()
        out0 = pb.alloc_outputs(hi - lo, JAC_BAND)
        a0 = pb.eval_args(X, out0, 0, 0, JAC_BAND, 0, st, peers=pg)
        pb.launch(a0)
        pg.fence()
        want0 = gather_results(out0["result"], Ptot)
        ok0 = torch.equal(pg.table(), want0) and torch.equal(want0[:, 0], want[:, 0])
        f0 = torch.tensor([1 if ok0 else 0], device="cuda")
        dist.all_reduce(f0, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{cfg}: values-only launch (mode 0): fused gather == NCCL all_gather: {bool(f0.item())}", flush=True)
        assert ok0, f"rank {rank}: {cfg} mode-0 tables differ"
    pg.close()
    pb.close()
if rank == 0:
    print("PEER GATHER OK", flush=True)
dist.destroy_process_group()
