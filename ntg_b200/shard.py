"""Multi-GPU: the batch shards by problem index, one process per GPU.

Problems are independent (they share only read-only tables), so there is no
data-path collective: rank r of R owns the contiguous slice
[r*P/R, (r+1)*P/R).  The only exchange is the per-problem
result table (objective, max constraint violation) = 16 B/problem, gathered
with one all_gather over NCCL (NVLink/NVSwitch) -- or gloo in the CPU tests --
or, fused into the evaluator, by peer stores into every rank's table (PeerGather).
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(P: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced, exhaustive: sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} / world {world}")
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(P: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_range(P, r, world) for r in range(world))]


def gather_results(local, P: int, group=None):
    """all-gather the per-problem (objective, violation) rows of every rank into
    the full [P][2] table, in problem order.  `local` is this rank's
    [P_local][2] tensor (cuda for NCCL, cpu for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(P, world)
    assert local.shape[0] == sizes[rank], "local shard does not match shard_range"
    if len(set(sizes)) == 1:
        out = torch.empty((P, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged: pad to the largest shard
    m = max(sizes)
    pad = torch.zeros((m, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * m, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)], dim=0)


class PeerGather:
    """Fused multi-GPU gather: every rank owns one gathered table [P_total][2] in its HBM; the
    tables of the other ranks are mapped into this process with CUDA IPC (ntgb_peer_table_open)
    and the evaluator's epilogue stores each (objective, violation) pair into ALL of them over
    NVLink -- no collective kernel runs beside the persistent evaluator.  Ranks must be on one node
    with peer access (NVSwitch).  `table()` is this rank's copy; read it after `fence()`."""

    def __init__(self, pb, P_total: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from .problem import _check, core
        self.pb, self.P_total, self.group = pb, int(P_total), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerGather: at most 8 ranks (NTGB_MAXPEERS)")
        self.row0 = shard_range(self.P_total, self.rank, self.world)[0]
        # Every rank runs the SAME sequence of collectives whatever fails locally; success is agreed
        # on at the end, so either all ranks get a PeerGather or all of them get the exception
        # (a rank that silently fell back to a collective would deadlock the others).
        self._torch = torch
        self._own, self.tables, self._opened = None, [], []
        ok, why = 1, ""
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        try:
            _check(core().ntgb_peer_table_alloc(pb._h, self.P_total, C.byref(own), handle))
            self._own = own.value
        except Exception as e:
            ok, why = 0, str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw) if ok else b"", group=group)
        if ok and all(len(h) == 64 for h in handles):
            try:
                for r in range(self.world):
                    if r == self.rank:
                        self.tables.append(self._own)
                        continue
                    ptr = C.c_void_p()
                    _check(core().ntgb_peer_table_open(pb._h, handles[r], C.byref(ptr)))
                    self.tables.append(ptr.value)
                    self._opened.append(ptr.value)
            except Exception as e:
                ok, why = 0, str(e)
        else:
            ok, why = 0, why or "another rank could not create its table"
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{pb.device}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            for p in self._opened:
                core().ntgb_peer_table_close(pb._h, p)
            self._opened = []
            dist.barrier(group=group)
            if self._own:
                core().ntgb_peer_table_free(pb._h, self._own)
                self._own = None
            raise RuntimeError("PeerGather: peer tables unavailable on at least one rank" + (f" ({why})" if why else ""))
        return

    def device_pointers(self) -> int:
        """address of a device array holding the table pointers (ntgb_eval_args.peer_result)"""
        if getattr(self, "_ptrs", None) is None:
            self._ptrs = self._torch.tensor(self.tables, dtype=self._torch.int64, device=f"cuda:{self.pb.device}")
        return self._ptrs.data_ptr()

    def table(self):
        """this rank's gathered table as a [P_total][2] float64 cuda tensor (a view, no copy)"""
        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (self.P_total, 2), "typestr": "<f8", "data": (self._own, False),
                                      "version": 2, "strides": None}
        return self._torch.as_tensor(v, device=f"cuda:{self.pb.device}")

    def fence(self):
        """every rank's evaluator launches so far have completed: the tables are consistent"""
        import torch.distributed as dist
        self._torch.cuda.synchronize(self.pb.device)
        dist.barrier(group=self.group)

    def close(self):
        from .problem import core
        import torch.distributed as dist
        dist.barrier(group=self.group)      # nobody may still be writing into a table that goes away
        for p in self._opened:
            core().ntgb_peer_table_close(self.pb._h, p)
        self._opened = []
        dist.barrier(group=self.group)
        if self._own:
            core().ntgb_peer_table_free(self.pb._h, self._own)
            self._own = None
