#!/usr/bin/env python
"""bench.py -- collocation evals/sec (cost + constraints + Jacobian, batched)
and % of the HBM roofline, on 1..8 B200 (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W [--workload cfg4] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic coefficient
vectors: a single launch of the fused evaluator K1 (funobj mode 2 + funcon
mode 2 for every problem of the batch: cost, gradient, constraints, banded
Jacobian) and, at N > 1, the NCCL all-gather of the per-problem
(objective, violation) table.  Weak scaling: every rank evaluates its own
batch of P problems (problems are independent; no data-path collective).

Headline workload (DESIGN.md "Measurement"): CFG-4, the kincar MPC shape with
64 breakpoints x 65536 problems -- the largest of BASELINE.json's configs the
survey names as a 1-GPU roofline config; CFG-2/3/5 are
reported beside it under "other_workloads" at N = 1.

Rank 0 prints ONE JSON line.  `--impl reference` times the reference's own CPU
implementation (oracle/_ref: the unmodified reference C sources) on all host
cores for the same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ntg_b200 import JAC_BAND, configs  # noqa: E402

L2_BYTES = 126e6
KERNEL_OF = {"cfg2": "ntgb::ntg_eval_small_kernel<vdp> (K1s)", "cfg3": "ntgb::ntg_eval_small_kernel<kincar> (K1s)",
             "cfg4": "ntgb::ntg_eval_small_kernel<kincar> (K1s)", "cfg5": "ntgb::ntg_eval_cluster_kernel<syn6> (K1c)"}
METRIC = "collocation_evals_per_sec"
UNIT = "evals/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# --------------------------------------------------------------------------- CPU reference arm
def _ref_worker(args):
    cfg, lo, hi, P, kind, shipped = args
    from oracle.oracle import Oracle
    spec, _ = configs.get(cfg)
    X = configs.coefficients(cfg, P, spec)[lo:hi]
    o = Oracle(kind, shipped_flags=shipped) if kind == "ref" else Oracle(kind)
    t0 = time.perf_counter()
    r = o.eval(spec, X, dense=False, band=False, outputs=False, reps=1)
    return time.perf_counter() - t0, r["seconds"], hi - lo


class CpuReference:
    """The reference's CPU implementation of the path on the host cores: one
    process per core (the reference is non-re-entrant: file-static globals,
    src/ntg.c:17-41), disjoint slices of the sample."""

    def __init__(self, cfg: str, cores: int | None = None):
        from oracle import oracle
        self.cfg = cfg
        self.kind = "ref" if oracle.have_ref() else "port"
        self.cores = cores or len(os.sched_getaffinity(0))
        import multiprocessing as mp
        self.pool = mp.get_context("spawn").Pool(self.cores)

    def sample_size(self, target_seconds: float) -> int:
        spec, P = configs.get(self.cfg)
        # calibrate on one core
        n0 = 4 if self.cfg == "cfg5" else 512
        w, s, n = _ref_worker((self.cfg, 0, n0, n0, self.kind, False))
        per_eval = max(s, 1e-9) / n
        n_target = int(target_seconds * self.cores / per_eval)
        per_core = max(1, min(n_target, P) // self.cores)
        return per_core * self.cores

    def step(self, nsample: int):
        from ntg_b200.shard import shard_range
        jobs = [(self.cfg, *shard_range(nsample, r, self.cores), nsample, self.kind, False)
                for r in range(self.cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_ref_worker, jobs)
        wall = time.perf_counter() - t0
        inner = max(r[1] for r in res)   # slowest worker, funcon+funobj time only
        return wall, inner

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(cfg: str, target_seconds: float = 4.0):
    ref = CpuReference(cfg)
    try:
        n = ref.sample_size(target_seconds)
        ref.step(min(n, ref.cores * 2))  # warm the workers (imports, table build)
        wall, inner = min((ref.step(n) for _ in range(3)), key=lambda t: t[1])  # best of 3 passes
        one = CpuReference(cfg, cores=1)
        n1 = max(1, n // ref.cores)
        one.step(1)
        _, inner1 = one.step(n1)
        one.close()
        spec, P = configs.get(cfg)
        return {"value": n / inner, "unit": UNIT, "cores": ref.cores,
                "kind": "reference" if ref.kind == "ref" else "port",
                "sample": f"{n} of {P} problems of {spec.name} (same seeded coefficients), one process per "
                          f"core, -O2 -ffp-contract=off; funobj mode 2 + funcon mode 2 per problem",
                "value_1core": n1 / inner1, "wall_value": n / wall}
    finally:
        ref.close()


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(a.workload)
    spec, P = configs.get(a.workload)
    n = ref.sample_size(a.ref_step_seconds)
    for _ in range(max(a.warmup, 1)):
        ref.step(n)
    t, tw = [], []
    for _ in range(a.steps):
        wall, inner = ref.step(n)
        t.append(inner)   # slowest worker's time inside funcon+funobj: the path itself,
        tw.append(wall)   # without this harness's process dispatch and input generation
    ref.close()
    per_step = float(np.mean(t))
    value = n / per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a.workload, spec, P, a.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores,
                         "kind": "reference" if ref.kind == "ref" else "port",
                         "sample": f"each step = {n} of {P} problems of {spec.name}, one process per core, "
                                   f"step time = slowest worker's time inside funcon+funobj",
                         "wall_value": n / float(np.mean(tw))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


def workload_config(cfg, spec, P, ngpus):
    return {"workload": spec.name, "cfg": cfg, "problems_per_gpu": P, "total_problems": P * ngpus,
            "nout": spec.nout, "nbps": spec.nbps, "order": spec.order[0], "nC": spec.nC,
            "ncnln": spec.ncnln, "jacobian": "band-compact",
            "bytes_per_eval": spec.bytes_per_eval(), "parallelism": f"problem-sharded x{ngpus}",
            "variant": None, "l2": None}


# --------------------------------------------------------------------------- GPU arm
def time_workload(torch, pb, spec, cfg, P, steps, warmup, dist=None, world=1, use_graph=False, peers=None):
    """-> dict(ms_per_step, kernel_ms, launches).  Ring of buffer sets larger than L2 when one
    set is not."""
    dev = torch.device("cuda", pb.device)
    per_set = P * spec.bytes_per_eval()
    nset = 1 if per_set >= 2 * L2_BYTES else min(64, int(np.ceil(2 * L2_BYTES / per_set)))
    X0 = torch.from_numpy(configs.coefficients(cfg, P, spec)).to(dev)
    sets = [(X0.clone() if i else X0, pb.alloc_outputs(P, JAC_BAND, zero=False)) for i in range(nset)]
    stream = torch.cuda.current_stream(dev)
    args = [pb.eval_args(x, o, 2, 2, JAC_BAND, 0, stream.cuda_stream) for x, o in sets]
    gathered = [torch.empty((world * P, 2), dtype=torch.float64, device=dev) for _ in range(nset)] if dist else None

    graphs = None
    if use_graph and dist is None:
        # launch-bound workloads (a few microseconds of GPU time per step): ONE CUDA graph holds a
        # whole lap over the ring of buffer sets, so the host launch cost is paid once per lap
        for i in range(nset):
            pb.launch(args[i])
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cs = torch.cuda.current_stream(dev).cuda_stream
            for i in range(nset):
                pb.launch(pb.eval_args(sets[i][0], sets[i][1], 2, 2, JAC_BAND, 0, cs))
        laps_w = max(1, warmup // nset + 1)
        laps = max(1, (steps + nset - 1) // nset)
        for _ in range(laps_w):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(laps):
            g.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / (laps * nset)
        chk = float(sets[0][1]["result"][:, 0].sum().item())
        assert np.isfinite(chk), "non-finite objective in the timed run"
        return {"ms_per_step": ms, "kernel_ms": ms, "launches": laps * nset, "nset": nset,
                "footprint_mb": per_set * nset / 1e6}

    # N > 1: the 16 B/problem result gather of step i runs on NCCL's stream while the kernel of
    # step i+1 runs (async_op); two result tables alternate so the gather never reads a table
    # the next kernel is writing.
    NB = 4  # the gather of step i runs after kernel i+1 has drained (persistent CTAs fill every SM),
    #         i.e. during kernel i+2; with 4 tables kernel i+4 never waits for it
    res2 = [torch.empty((P, 2), dtype=torch.float64, device=dev) for _ in range(NB)] if dist else None
    gath2 = [torch.empty((world * P, 2), dtype=torch.float64, device=dev) for _ in range(NB)] if dist else None
    if dist is not None:
        args2 = [[pb.eval_args(x, dict(o, result=res2[b]), 2, 2, JAC_BAND, 0, stream.cuda_stream, peers=peers)
                  for x, o in sets] for b in range(NB)]
    pending = [None] * NB

    def step(i, ev=None):
        s = i % nset
        if ev is not None:
            ev[0].record()
        if dist is not None:
            b = i % NB
            if pending[b] is not None:
                pending[b].wait()          # the gather that last read res2[b] is done
            pb.launch(args2[b][s])
        else:
            pb.launch(args[s])
        if ev is not None:
            ev[1].record()
        if dist is not None and peers is None:
            pending[i % NB] = dist.all_gather_into_tensor(gath2[i % NB], res2[i % NB], async_op=True)

    def drain():
        for b in range(NB):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None
        if peers is not None:
            peers.fence()   # fused gather: the tables are complete once every rank's stream has drained

    for i in range(warmup):
        step(i)
    drain()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(steps):
        step(i, ev[i])
    drain()
    e1.record()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        want = res2[(steps - 1) % NB]
        got = gath2[(steps - 1) % NB] if peers is None else peers.table()
        r0 = dist.get_rank()
        assert torch.equal(got[r0 * P:(r0 + 1) * P], want), "gathered table does not hold this rank's rows"
        if peers is not None:
            other = got[((r0 + 1) % world) * P:((r0 + 1) % world + 1) * P]
            assert bool(torch.isfinite(other).all()) and float(other[:, 0].abs().sum()) > 0.0, \
                "the neighbour's rows never arrived in this rank's gathered table"
    total_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    chk = float((res2[0] if dist is not None else sets[0][1]["result"])[:, 0].sum().item())
    assert np.isfinite(chk), "non-finite objective in the timed run"
    return {"ms_per_step": total_ms / steps, "kernel_ms": kernel_ms, "launches": steps, "nset": nset,
            "footprint_mb": per_set * nset / 1e6}


def time_e2e(torch, pb, spec, cfg, P, steps, warmup, full: bool):
    """Same metric end to end with HOST buffers: every step copies that step's coefficients from
    pinned host memory, evaluates, and copies results back to pinned host memory.
    full=True: ONE call of the C ABI's host-buffer entry point `ntgb_eval_host` per step -- f, g, c
    and the whole Jacobian band come back (what a host-side consumer such as NPSOL needs); the
    library chunks the batch and overlaps copies with compute on two streams.
    full=False ("resident"): only the per-problem (objective, violation) table comes back; g, c, J
    are still computed and written to HBM, where a GPU-side consumer would read them."""
    dev = torch.device("cuda", pb.device)
    d = pb.dims
    Xh = torch.from_numpy(configs.coefficients(cfg, P, spec)).pin_memory()
    if full:
        host = {"f": torch.empty(P, dtype=torch.float64).pin_memory(),
                "result": torch.empty((P, 2), dtype=torch.float64).pin_memory(),
                "g": torch.empty((P, d.nC), dtype=torch.float64).pin_memory(),
                "c": torch.empty((P, d.ncnln), dtype=torch.float64).pin_memory(),
                "J": torch.empty((P, d.ncnln * d.sorder), dtype=torch.float64).pin_memory()}
        h2d = P * d.nC * 8
        d2h = sum(v.numel() * 8 for v in host.values())

        def one_step():
            pb.eval_host_tensors(Xh, host, 2, 2, JAC_BAND)
        nlaunch = None
    else:
        nchunk = 2 if P >= 8192 else 1
        Pc = P // nchunk
        host = {"result": torch.empty((P, 2), dtype=torch.float64).pin_memory()}
        streams = [torch.cuda.Stream(dev) for _ in range(2)]
        bufs = []
        for s in range(2):
            with torch.cuda.stream(streams[s]):
                bufs.append((torch.empty((Pc, d.nC), dtype=torch.float64, device=dev),
                             pb.alloc_outputs(Pc, JAC_BAND, zero=False)))
        h2d = P * d.nC * 8
        d2h = host["result"].numel() * 8
        nlaunch = nchunk

        def one_step():
            for ci in range(nchunk):
                s = ci % 2
                st = streams[s]
                lo, hi = ci * Pc, (ci + 1) * Pc
                with torch.cuda.stream(st):
                    xd, out = bufs[s]
                    xd.copy_(Xh[lo:hi], non_blocking=True)
                    pb.launch(pb.eval_args(xd, out, 2, 2, JAC_BAND, 0, st.cuda_stream))
                    host["result"][lo:hi].copy_(out["result"], non_blocking=True)
            for st in streams:
                st.synchronize()

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    assert np.isfinite(float(host["result"][:, 0].sum())), "non-finite objective in the e2e run"
    return {"value": P / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": dt * 1e3,
            "what": ("one ntgb_eval_host() call per step: all outputs (f, g, c, Jacobian band) to pinned host memory"
                     if full else
                     "(objective, violation) table to host; g, c, J computed and left resident in HBM")}


def time_solvers(torch, device, fast):
    """Wall time of the two batched solvers on the lane-change problem of examples/kincar.c:
    ntgb_solve_eq (linear equalities only, as shipped) and ntgb_solve_nlp (active speed and
    curvature bounds).  Host wall clock around the synchronous C call, after one warm-up solve."""
    import dataclasses
    from ntg_b200 import Problem
    out = {}
    s1 = configs.kincar(64, constraints=False, name="solve_kincar_64bps")
    P1 = 65536
    X1 = torch.from_numpy(configs.coefficients("cfg3", P1, s1, seed=3)).to(f"cuda:{device}")
    pb = Problem(s1, device, fast=fast)
    pb.solve_eq(X1.clone(), max_iter=100)
    C1 = X1.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, it, st = pb.solve_eq(C1, max_iter=100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["solve_eq"] = {"workload": "kincar lane change, 64 breakpoints, linear equalities only", "problems": P1,
                       "ms": dt * 1e3, "problems_per_s": P1 / dt, "iterations_mean": float(it.float().mean()),
                       "converged_frac": float((st >= 1).float().mean())}
    pb.close()
    base = configs.kincar(40, constraints=True, name="solve_kincar_active")
    kw = {f.name: getattr(base, f.name) for f in dataclasses.fields(base)}
    kw.update(ninterv=[4, 4], knots=None, bps=None)
    s2 = type(base)(**kw)
    lo, up = s2.lowerb.copy(), s2.upperb.copy()
    lo[-2], up[-2], lo[-1], up[-1] = 0.0, 66.2, -7.2, 7.2
    s2.lowerb, s2.upperb = lo, up
    P2 = 16384
    X2 = torch.from_numpy(configs.coefficients("cfg3", P2, s2, seed=5)).to(f"cuda:{device}")
    pb = Problem(s2, device, fast=fast)
    pb.solve_nlp(X2.clone())   # warm-up at full size: the scratch is allocated here, not in the timed call
    C2 = X2.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, v, it, st = pb.solve_nlp(C2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["solve_nlp"] = {"workload": "kincar lane change, 40 breakpoints, 4 intervals, active speed^2 / curvature bounds",
                        "problems": P2, "ms": dt * 1e3, "problems_per_s": P2 / dt,
                        "iterations_mean": float(it.float().mean()), "converged_frac": float((st >= 1).float().mean()),
                        "violation_max": float(v.max())}
    pb.close()
    torch.cuda.empty_cache()
    return out


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner on stdout when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the run
    and the line goes to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--variant", default=os.environ.get("NTG_BENCH_VARIANT", "fast"), choices=["exact", "fast"])
    ap.add_argument("--problems", type=int, default=0, help="override problems per GPU")
    ap.add_argument("--no-others", action="store_true", help="skip the other workloads / baselines")
    ap.add_argument("--gather", default=os.environ.get("NTG_BENCH_GATHER", "fused"), choices=["fused", "nccl"],
                    help="N > 1: how the 16 B/problem result table reaches every rank")
    ap.add_argument("--ref-step-seconds", type=float, default=1.0)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    _quiet_stdout()

    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the evaluator has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # --gather fused (default): the evaluator's epilogue stores each rank's (objective, violation)
        # rows into EVERY rank's gathered table over NVLink (CUDA IPC peer tables, ntg_b200/shard.py::
        # PeerGather); no collective kernel runs beside the persistent evaluator.
        # --gather nccl: one asynchronous all_gather per step.  The persistent evaluator fills every
        # SM, so the NCCL kernel can only start when evaluator CTAs retire and then holds back a few
        # CTAs of the NEXT launch, which finish last (N = 2: 144 us per step; with four SMs left
        # free and four NCCL channels 139 us; at N = 4 / 8 the 4 / 8 MB table needs NCCL's default
        # channel count: 138 / 149 us).
        if a.gather == "nccl" and world == 2:
            os.environ.setdefault("NTG_B200_SM_RESERVE", "4")
            os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ntg_b200 import Problem

    spec, P = configs.get(a.workload)
    if a.problems:
        P = a.problems
    fast = a.variant == "fast"
    pb = Problem(spec, local, fast=fast)
    peak, peak_src = measured_peaks()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    peers = None
    gather_how = "none"
    if world > 1:
        gather_how = "nccl all_gather_into_tensor, asynchronous, 4 rotating tables"
        if a.gather == "fused":
            try:
                from ntg_b200.shard import PeerGather
                peers = PeerGather(pb, world * P)
                gather_how = "fused: peer stores from the evaluator's epilogue into every rank's table (CUDA IPC, NVLink)"
            except Exception as e:  # no peer access between these GPUs: the collective still works
                sys.stderr.write(f"bench: fused gather unavailable ({e}); using NCCL\n")
                peers = None
    r = time_workload(torch, pb, spec, a.workload, P, a.steps, a.warmup, dist, world, peers=peers)
    clocks = sampler.stop() if sampler else None

    ms = torch.tensor([r["ms_per_step"], r["kernel_ms"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step, ms_kernel = float(ms[0]), float(ms[1])
    value = world * P / (ms_step * 1e-3)

    e2e = time_e2e(torch, pb, spec, a.workload, P, max(3, min(a.steps, 10)), 3, full=True)
    e2e_res = time_e2e(torch, pb, spec, a.workload, P, max(3, min(a.steps, 10)), 3, full=False)
    if dist is not None:
        t = torch.tensor([e2e["value"], e2e_res["value"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        e2e["value"], e2e_res["value"] = float(t[0]) * world, float(t[1]) * world

    others = {}
    if rank == 0 and world == 1 and not a.no_others:
        for cfg in ("cfg2", "cfg3", "cfg4", "cfg5"):
            if cfg == a.workload:
                continue
            s2, P2 = configs.get(cfg)
            pb2 = Problem(s2, local, fast=fast)
            small = cfg in ("cfg2", "cfg3")
            st = 200 if small else (5 if cfg == "cfg5" else 20)
            r2 = time_workload(torch, pb2, s2, cfg, P2, st, 5, use_graph=small)
            gbs = P2 * s2.bytes_per_eval() / (r2["kernel_ms"] * 1e-3) / 1e9
            others[cfg] = {"workload": s2.name, "problems": P2, "evals_per_s": P2 / (r2["ms_per_step"] * 1e-3),
                           "ms_per_step": r2["ms_per_step"], "kernel_ms": r2["kernel_ms"],
                           "algorithmic_gbs": gbs, "roofline_frac": gbs / peak,
                           "launch": "one CUDA graph per lap over the buffer ring" if small else "stream launch", "kernel": KERNEL_OF[cfg],
                           "l2": f"ring of {r2['nset']} buffer sets, {r2['footprint_mb']:.0f} MB"}
            pb2.close()
            torch.cuda.empty_cache()
        try:   # the batched solvers built on the evaluator (informative, not the metric)
            others["solvers"] = time_solvers(torch, local, fast)
        except Exception as e:  # never let an extra take the bench line down
            others["solvers"] = {"error": str(e)[:200]}

    if rank == 0:
        bytes_launch = P * spec.bytes_per_eval()
        ach = bytes_launch / (ms_kernel * 1e-3) / 1e9
        cfgd = workload_config(a.workload, spec, P, world)
        cfgd["variant"] = ("fast (FMA contraction, node-weight quadrature with 4 partial sums)" if fast else
                           "exact (-fmad=false, reference summation order, bit-identical to the CPU reference)")
        if world > 1:
            cfgd["gather"] = gather_how
        cfgd["l2"] = (f"inputs+outputs of one step = {r['footprint_mb']:.0f} MB over a ring of {r['nset']} "
                      f"buffer set(s), larger than the 126 MB L2")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfgd, "clocks": clocks,
            "e2e": e2e, "e2e_resident": e2e_res, "gpu_launches": r["launches"],
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": KERNEL_OF.get(a.workload, "ntgb::ntg_eval_kernel"),
                         "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": bytes_launch},
            "other_workloads": others,
        }
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                line["roofline"]["traffic"] = json.load(open(tr)).get(a.workload)
            except Exception:
                pass
        if world == 1 and not a.no_others:
            line["cpu_baseline"] = cpu_baseline(a.workload)
        _emit(line)
    pb.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
