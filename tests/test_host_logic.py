"""Host-side logic that needs no GPU: problem-size arithmetic, layout
conversion, sharding, and the world_size-2 result gather over gloo."""
import os

import numpy as np
import pytest

from ntg_b200 import configs
from ntg_b200.abi import linspace
from ntg_b200.shard import shard_range, shard_sizes


def test_sizes_match_reference_formulas(port):
    for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
        spec, _ = configs.get(name)
        d = port.dims(spec)
        assert (d.nC, d.nz, d.nZ, d.nclin, d.ncnln, d.sorder) == \
            (spec.nC, spec.nz, spec.nZ, spec.nclin, spec.ncnln, spec.sorder)
    s5 = configs.get("cfg5")[0]
    assert (s5.nC, s5.ncnln, s5.nbps, s5.nZ) == (4824, 1604, 401, 9624)


def test_linspace_is_the_accumulating_recurrence(port):
    import ctypes as C
    for n, d1 in ((20, 5.0), (64, 5.0), (401, 5.0), (3, 5.0), (7, 1.0)):
        v = np.zeros(n)
        port.lib.port_linspace(v.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(0.0), C.c_double(d1), n)
        assert np.array_equal(v, linspace(0.0, d1, n))
    assert np.array_equal(linspace(1, 1, 7), np.ones(7))


def test_shard_ranges_partition_the_batch():
    for P in (0, 1, 7, 4096, 65536, 16384 + 3):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(P, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == P
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sz = shard_sizes(P, world)
            assert sum(sz) == P and max(sz) - min(sz) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port_no, P, q):
    import torch
    import torch.distributed as dist
    from ntg_b200.shard import gather_results, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    spec = configs.vanderpol(20)
    X = configs.coefficients("cfg2", P, spec)
    lo, hi = shard_range(P, rank, world)
    r = Oracle("port").eval(spec, X[lo:hi], dense=False, band=False)
    from common import violation
    local = torch.from_numpy(np.stack([r["f"], violation(spec, r["c"])], axis=1))
    full = gather_results(local, P)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [64, 61])
def test_two_rank_gather_reassembles_the_batch(port, P):
    """world_size 2 over gloo: each rank evaluates its shard (CPU oracle standing in for
    the GPU), the gathered [P][2] table equals the single-process result, in order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = configs.vanderpol(20)
    X = configs.coefficients("cfg2", P, spec)
    r = port.eval(spec, X, dense=False, band=False)
    from common import violation
    assert np.array_equal(full[:, 0], r["f"]) and np.array_equal(full[:, 1], violation(spec, r["c"]))
