"""Parity tests proper: the CUDA path (through the C ABI, libntg_b200.so + a
callback pack) against the CPU oracle and the committed golden vectors.

Bar (BASELINE.json north_star): knot-interval indices and Jacobian sparsity
pattern bit-exact; flat outputs, cost, gradient, constraints, Jacobian within
1e-12 relative / 1e-14 absolute (scaled by the vector's infinity norm).
The "exact" pack variant (-fmad=false, reference summation order) is held to
the stronger bar of bit-identity wherever no libm call is involved.
"""
import os

import numpy as np
import pytest

from common import (GOLDEN_CASES, assert_bitexact, assert_close, golden_spec, load_golden, violation)
from ntg_b200 import JAC_BAND, JAC_DENSE, JAC_NONE, configs

pytestmark = pytest.mark.gpu

NO_LIBM = {"cfg2_vanderpol", "cfg3_kincar", "cfg4_kincar64", "cfg5_syn6", "syn6_small"}


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def gpu_eval(torch, spec, X, fast, jac=JAC_BAND, want_Z=False, mode_obj=2, mode_con=2):
    from ntg_b200 import Problem
    pb = Problem(spec, 0, fast=fast)
    o = pb.eval(torch.from_numpy(np.ascontiguousarray(X)).cuda(), mode_obj, mode_con, jac, want_Z)
    torch.cuda.synchronize()
    r = {k: (None if v is None else v.cpu().numpy()) for k, v in o.items()}
    r["c"] = r["c"][:, :spec.ncnln]
    if r["J"] is not None and jac == JAC_BAND:
        r["Jband"] = pb.band_to_rows(r["J"])
    return pb, r


@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_against_golden_vectors(torch_cuda, name, fast):
    """golden = outputs of the UNMODIFIED reference on seeded inputs"""
    spec, X = golden_spec(name)
    g = load_golden(name)
    pb, r = gpu_eval(torch_cuda, spec, X, fast)
    B, off, _ = pb.tables()
    for j, b in enumerate(B):
        assert_bitexact(b, g["B"][j], f"{name}: K0 table of output {j}")
    assert_bitexact(off, g["off"], f"{name}: knot-interval offsets")
    col0, _ = pb.pattern()
    assert_bitexact(col0, g["col0"], f"{name}: Jacobian sparsity pattern")
    cmp = assert_bitexact if (not fast and name in NO_LIBM) else assert_close
    cmp(r["f"], g["f"], f"{name}: objective")
    cmp(r["g"], g["g"], f"{name}: gradient")
    cmp(r["c"], g["c"], f"{name}: constraints")
    cmp(r["Jband"], g["Jband"], f"{name}: Jacobian band")
    assert_bitexact(pb.linear(), g["A"], f"{name}: linear constraint matrix A")
    bl, bu = pb.bounds()
    assert_bitexact(bl, g["bl"], "bl")
    assert_bitexact(bu, g["bu"], "bu")
    pb.close()


@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
@pytest.mark.parametrize("cfg,P", [("cfg2", 4096), ("cfg3", 8192), ("cfg4", 4096), ("cfg5", 64)])
def test_against_oracle_at_baseline_sizes(torch_cuda, port, cfg, P, fast):
    """full batch for CFG-2/3, 4096-problem subset for CFG-4,
    64-problem subset for CFG-5."""
    spec, _ = configs.get(cfg)
    X = configs.coefficients(cfg, P, spec)
    o = port.eval(spec, X, dense=False, band=True)
    pb, r = gpu_eval(torch_cuda, spec, X, fast)
    cmp = assert_bitexact if not fast else assert_close
    for k, ok in (("f", "f"), ("g", "g"), ("c", "c"), ("Jband", "Jband")):
        cmp(r[k], o[ok], f"{cfg}.{k}")
    assert_bitexact(r["result"][:, 0], r["f"], "result[:,0] is the objective")
    assert_close(r["result"][:, 1], violation(spec, o["c"]), "max nonlinear constraint violation")
    pb.close()


@pytest.mark.parametrize("name", ["cfg2_vanderpol", "cfg4_kincar64", "endpoint"])
def test_dense_npsol_layout_and_unwritten_zeros(torch_cuda, port, name):
    """NPSOL's column-major ncnln x nC Jacobian: band entries equal the oracle's,
    everything the reference never writes stays at the allocation-time zero."""
    spec, X = golden_spec(name)
    o = port.eval(spec, X, dense=True, band=False)
    pb, r = gpu_eval(torch_cuda, spec, X, False, jac=JAC_DENSE)
    written = ~np.isnan(o["Jdense"])
    assert np.all(r["J"][~written] == 0.0), "wrote outside the reference's sparsity pattern"
    cmp = assert_bitexact if name in NO_LIBM else assert_close
    cmp(r["J"][written], o["Jdense"][written], "dense Jacobian band entries")
    pb.close()


@pytest.mark.parametrize("name", ["cfg3_kincar", "endpoint"])
def test_flat_outputs_Z(torch_cuda, port, name):
    """Z[p][iZ_j + bp*maxderiv_j + d] equals updateZ over the union of active variables"""
    spec, X = golden_spec(name)
    pb, r = gpu_eval(torch_cuda, spec, X, False, want_Z=True)
    for p in range(X.shape[0]):
        Z = np.zeros(spec.nZ)
        for kind, lists in ((0, (spec.initialcostav if spec.nicf else [], spec.initialconstrav if spec.nnlic else [])),
                            (1, (spec.trajectorycostav if spec.nucf else [], spec.trajectoryconstrav if spec.nnltc else [])),
                            (2, (spec.finalcostav if spec.nfcf else [], spec.finalconstrav if spec.nnlfc else []))):
            for av in lists:
                if len(av):
                    Zk = port.updateZ(spec, X[p], list(av), kind)
                    Z = np.where(Zk != 0, Zk, Z)
        assert_bitexact(r["Z"][p], Z, f"{name}: flat outputs of problem {p}")
    pb.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_modes(torch_cuda, port, mode):
    """mode 0 values only, 1 derivatives only, 2 both; -1 skips a half"""
    spec, X = golden_spec("endpoint")
    full = port.eval(spec, X, dense=False)
    pb, r = gpu_eval(torch_cuda, spec, X, False, mode_obj=mode, mode_con=mode)
    if mode != 1:
        assert_close(r["f"], full["f"], "f")
        assert_close(r["c"], full["c"], "c")
    else:
        assert np.all(r["f"] == 0) and np.all(r["c"] == 0), "mode 1 must not write values"
    if mode != 0:
        assert_close(r["g"], full["g"], "g")
        assert_close(r["Jband"], full["Jband"], "J")
    else:
        assert np.all(r["g"] == 0) and np.all(r["J"] == 0), "mode 0 must not write derivatives"
    pb.close()
    pb, r = gpu_eval(torch_cuda, spec, X, False, mode_obj=2, mode_con=-1)
    assert_close(r["g"], full["g"], "g with constraints skipped")
    assert np.all(r["c"] == 0) and np.all(r["J"] == 0)
    pb.close()


@pytest.mark.parametrize("name", ["cfg2_vanderpol", "cfg4_kincar64", "endpoint"])
def test_general_kernel_on_small_shapes(torch_cuda, port, name, monkeypatch):
    """Small shapes normally take the register-table kernel K1s; NTG_B200_KERNEL=general
    forces the general kernel K1 so both are held to the same bar on the same inputs."""
    monkeypatch.setenv("NTG_B200_KERNEL", "general")
    spec, X = golden_spec(name)
    o = port.eval(spec, X, dense=False)
    pb, r = gpu_eval(torch_cuda, spec, X, False)
    cmp = assert_bitexact if name in NO_LIBM else assert_close
    for k in ("f", "g", "c", "Jband"):
        cmp(r[k], o[k], f"{name}.{k} (general kernel)")
    pb.close()


@pytest.mark.parametrize("nbps,P", [(257, 5), (320, 37), (512, 3)])
@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
def test_wide_register_table_kernel(torch_cuda, port, nbps, P, fast, monkeypatch):
    """257..512 breakpoints with outputs of DIFFERENT spline setups (no cluster kernel): K1s on CTAs of
    512 threads, one per SM.  Same bar as everywhere (the endpoint callbacks use libm: 1e-12), plus
    agreement with the general kernel K1 on the same inputs, and a ragged batch."""
    spec = configs.endpoint(nbps, name=f"endpoint_{nbps}")
    X = np.random.default_rng(nbps).uniform(-1, 1, (P, spec.nC))
    o = port.eval(spec, X, dense=False)
    pb, r = gpu_eval(torch_cuda, spec, X, fast, want_Z=True)
    for k in ("f", "g", "c", "Jband"):
        assert_close(r[k], o[k], f"endpoint_{nbps}.{k} (K1s, 512 threads)")
    assert_close(r["result"][:, 0], o["f"], "result[:,0]")
    assert_close(r["result"][:, 1], violation(spec, o["c"]), "result[:,1]")
    pb.close()
    monkeypatch.setenv("NTG_B200_KERNEL", "general")
    pb, g = gpu_eval(torch_cuda, spec, X, fast, want_Z=True)
    for k in ("f", "g", "c", "Jband"):
        assert_close(r[k], g[k], f"endpoint_{nbps}.{k} K1s-512 vs K1")
    assert_bitexact(r["Z"], g["Z"], "flat outputs, K1s-512 vs K1")
    pb.close()


@pytest.mark.parametrize("ninterv,P", [(200, 3), (150, 75), (300, 5), (640, 2)])
@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
def test_cluster_kernel_long_horizons(torch_cuda, port, ninterv, P, fast):
    """K1c (thread-block clusters of 2/4/8 CTAs, DSMEM halo) on long horizons: nbps = 401 / 301 /
    601 / 1281, batch sizes below, at and above the number of clusters; all modes' outputs."""
    spec = configs.syn6(ninterv, name=f"syn6_{ninterv}")
    X = configs.coefficients("cfg5", P, spec, seed=ninterv)
    o = port.eval(spec, X, dense=False, band=True)
    pb, r = gpu_eval(torch_cuda, spec, X, fast, want_Z=True)
    cmp = assert_bitexact if not fast else assert_close
    for k in ("f", "g", "c", "Jband"):
        cmp(r[k], o[k], f"syn6/{ninterv}.{k}")
    assert_close(r["result"][:, 1], violation(spec, o["c"]), "violation")
    av = [(j, d) for j in range(6) for d in range(4)]
    cmp(r["Z"][0], port.updateZ(spec, X[0], av, 1), "flat outputs")
    pb.close()
    if ninterv == 150 and not fast:   # values-only and derivatives-only modes through the cluster kernel
        for mode in (0, 1):
            pb, r = gpu_eval(torch_cuda, spec, X, fast, mode_obj=mode, mode_con=mode)
            if mode == 0:
                assert_bitexact(r["f"], o["f"], "f"); assert_bitexact(r["c"], o["c"], "c")
                assert np.all(r["g"] == 0) and np.all(r["J"] == 0)
            else:
                assert_bitexact(r["g"], o["g"], "g"); assert_bitexact(r["Jband"], o["Jband"], "J")
                assert np.all(r["f"] == 0) and np.all(r["c"] == 0)
            pb.close()


@pytest.mark.parametrize("ninterv,nbps,P", [(200, None, 3), (200, None, 150), (150, None, 75), (150, 300, 7), (300, None, 40),
                                            (300, 890, 5), (640, None, 20), (101, 257, 9)])
@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
def test_cluster_hot_kernel(torch_cuda, port, ninterv, nbps, P, fast, monkeypatch):
    """K1c/H, the steady-state cluster kernel (mode 2/2, band layout, no Z): Jacobian rows staged in
    shared memory and drained by cp.async.bulk.  Odd and even nbps (rows that start on odd and even
    elements), clusters of 2 / 4 / 8, CTAs with a ragged last warp, batches below, at and above the
    number of clusters, every ring depth the launcher may choose; and the same launch on the
    general-mode cluster kernel must agree bit for bit."""
    torch = torch_cuda
    spec = configs.syn6(ninterv, nbps=nbps, name=f"syn6_{ninterv}_{nbps}")
    X = configs.coefficients("cfg5", P, spec, seed=ninterv)
    o = port.eval(spec, X, dense=False, band=True)
    cmp = assert_bitexact if not fast else assert_close
    res = {}
    for stages in ("2", "3", "6"):
        monkeypatch.setenv("NTG_B200_HOT_STAGES", stages)
        pb, r = gpu_eval(torch, spec, X, fast)
        for k in ("f", "g", "c", "Jband"):
            cmp(r[k], o[k], f"syn6/{ninterv}/{nbps} stages={stages}: {k}")
        assert_close(r["result"][:, 1], violation(spec, o["c"]), "violation")
        assert np.array_equal(r["result"][:, 0], r["f"])
        res[stages] = r
        pb.close()
    monkeypatch.delenv("NTG_B200_HOT_STAGES")
    monkeypatch.setenv("NTG_B200_KERNEL", "cluster")
    pb, r = gpu_eval(torch, spec, X, fast)
    pb.close()
    for k in ("f", "c", "J", "result") + (() if fast else ("g",)):
        assert np.array_equal(r[k], res["6"][k]), f"K1c/H and K1c disagree in {k}"
    if fast:
        assert_close(r["g"], res["6"]["g"], "g")


def test_cluster_kernel_dense_layout(torch_cuda, port):
    spec = configs.syn6(140, name="syn6_140")
    X = configs.coefficients("cfg5", 2, spec, seed=3)
    o = port.eval(spec, X, dense=True, band=False)
    pb, r = gpu_eval(torch_cuda, spec, X, False, jac=JAC_DENSE)
    written = ~np.isnan(o["Jdense"])
    assert np.all(r["J"][~written] == 0.0)
    assert_bitexact(r["J"][written], o["Jdense"][written], "dense Jacobian through K1c")
    pb.close()


@pytest.mark.parametrize("P", [1, 2, 3, 11, 12, 13, 255, 257, 1000])
def test_ragged_batches(torch_cuda, port, P):
    """batch sizes around the tile size G (12 problems per CTA at nbps = 20), incl. P = 1"""
    spec, _ = configs.get("cfg3")
    X = configs.coefficients("cfg3", P, spec, seed=P)
    o = port.eval(spec, X, dense=False)
    pb, r = gpu_eval(torch_cuda, spec, X, False)
    for k in ("f", "g", "c", "Jband"):
        assert_bitexact(r[k], o[k], f"P={P}: {k}")
    pb.close()


def test_batch_independence_and_full_size_properties(torch_cuda):
    """At BASELINE.json's full CFG-4 size (65536 x 64 breakpoints): results do not depend on
    how the batch is sliced, a permuted batch gives the permuted result, and repeated
    launches are bit-identical (no atomics on the data path)."""
    torch = torch_cuda
    from ntg_b200 import Problem
    spec, P = configs.get("cfg4")
    X = torch.from_numpy(configs.coefficients("cfg4", P, spec)).cuda()
    pb = Problem(spec, 0, fast=True)
    a = pb.eval(X)
    b = pb.eval(X)
    for k in ("f", "g", "c", "J", "result"):
        assert torch.equal(a[k], b[k]), f"{k}: launch-to-launch nondeterminism"
    perm = torch.randperm(P, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    c = pb.eval(X[perm].contiguous())
    for k in ("f", "g", "c", "J"):
        assert torch.equal(c[k], a[k][perm]), f"{k}: depends on position in the batch"
    lo, hi = 12345, 12345 + 1001
    d = pb.eval(X[lo:hi].contiguous())
    for k in ("f", "g", "c", "J"):
        assert torch.equal(d[k], a[k][lo:hi]), f"{k}: depends on batch slicing"
    # z is linear in C and the kincar cost is quadratic: f(2C) = 4 f(C), g(2C) = 2 g(C)
    # (exact in binary floating point: scaling by 2 commutes with every rounding)
    e = pb.eval((2.0 * X).contiguous())
    assert torch.equal(e["f"], 4.0 * a["f"]) and torch.equal(e["g"], 2.0 * a["g"])
    pb.close()


@pytest.mark.parametrize("cfg,P", [("cfg3", 8192), ("cfg3", 3553), ("cfg4", 4096), ("cfg4", 8192), ("cfg4", 8189), ("cfg4", 4737),
                                   ("cfg2", 7105)])
@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
def test_even_split_launch_geometries(torch_cuda, cfg, P, fast):
    """Batches of a few tiles per CTA are split EVENLY over the CTAs (ntg_eval_small.cuh::
    launch_eval_small: one or two tiles per CTA of P / (k * grid) problems, +1 for the first few):
    every output, including the (objective, violation) table, is bit-identical to the same problems
    evaluated in slices that take the other geometry (whole tiles of G*R problems)."""
    torch = torch_cuda
    from ntg_b200 import Problem
    spec, _ = configs.get(cfg)
    X = torch.from_numpy(configs.coefficients(cfg, P, spec, seed=P)).cuda()
    pb = Problem(spec, 0, fast=fast)
    a = pb.eval(X)
    torch.cuda.synchronize()
    step = 1000 if cfg != "cfg4" else 777   # one wave of whole tiles
    for lo in range(0, P, step):
        hi = min(P, lo + step)
        d = pb.eval(X[lo:hi].contiguous())
        for k in ("f", "g", "c", "J", "result"):
            assert torch.equal(d[k], a[k][lo:hi]), f"{cfg} P={P} [{lo},{hi}): {k} depends on the launch geometry"
    pb.close()


def test_eval_host_path_matches_device_path(torch_cuda, port):
    """ntgb_eval_host (host buffers, H2D/D2H inside) == ntgb_eval on device buffers"""
    from ntg_b200 import Problem
    spec, X = golden_spec("endpoint")
    o = port.eval(spec, X, dense=True, band=True)
    pb = Problem(spec, 0)
    for jac, key in ((JAC_BAND, "Jband"), (JAC_DENSE, "Jdense")):
        h = pb.eval_host(X, jac=jac, want_Z=False)
        assert_close(h["f"], o["f"], "f")
        assert_close(h["g"], o["g"], "g")
        assert_close(h["c"], o["c"], "c")
        if jac == JAC_BAND:
            assert_close(pb.band_to_rows(h["J"]), o["Jband"], "band J")
        else:
            assert_close(h["J"], np.nan_to_num(o["Jdense"], nan=0.0), "dense J")
        assert_close(h["result"][:, 1], violation(spec, o["c"]), "violation")
    # growing and shrinking batches reuse / regrow the scratch
    for P in (1, 40, 3):
        Xp = np.random.default_rng(P).uniform(-1, 1, (P, spec.nC))
        assert_close(pb.eval_host(Xp)["g"], port.eval(spec, Xp, dense=False, band=False)["g"], f"P={P}")
    pb.close()


def test_error_paths_on_gpu(torch_cuda):
    from ntg_b200 import Problem
    from ntg_b200.problem import NtgError
    import torch
    spec, _ = configs.get("cfg2")
    pb = Problem(spec, 0)
    X = torch.zeros((4, spec.nC), dtype=torch.float64, device="cuda")
    out = pb.alloc_outputs(4)
    with pytest.raises(NtgError, match="mode"):
        pb.launch(pb.eval_args(X, out, mode_obj=5))
    a = pb.eval_args(X, out)
    a.C = None
    with pytest.raises(NtgError, match="C == NULL"):
        pb.launch(a)
    with pytest.raises(NtgError, match="device"):
        Problem(spec, 99)
    pb.close()


def _random_spec(rng, family, unsorted=False):
    """Random spline setups around the packs' compile-time shapes: random order / mult /
    interval count, non-uniform knots, non-uniform breakpoints (first and last on the ends)."""
    import copy
    if family == "endpt":
        base = configs.endpoint()
        order = [int(rng.integers(3, 7)), int(rng.integers(2, 5))]
        nbps = int(rng.choice([2, 3, 17, 64, 200, 300]))
    else:  # kincar: both outputs share one setup (one_table) -> K1s below 257 breakpoints, K1c above
        base = configs.kincar(20)
        o = int(rng.integers(3, 6))
        order = [o, o]
        nbps = int(rng.choice([40, 257, 300, 449, 700]))
    spec = copy.deepcopy(base)
    spec.order = order
    spec.mult = [int(rng.integers(max(1, md - 1), k)) for k, md in zip(order, spec.maxderiv)]
    if family == "kincar":
        spec.mult = [spec.mult[0]] * 2
        ni = int(rng.integers(1, 40))
        spec.ninterv = [ni, ni]
        kn = np.sort(np.concatenate([[0.0, 5.0], rng.uniform(0.0, 5.0, ni - 1)]))
        spec.knots = [kn, kn.copy()]
        t1 = 5.0
    else:
        spec.ninterv = [int(rng.integers(1, 9)), int(rng.integers(1, 9))]
        spec.knots = [np.sort(np.concatenate([[0.0, 2.0], rng.uniform(0.0, 2.0, ni - 1)])) for ni in spec.ninterv]
        t1 = 2.0
    spec.bps = np.sort(np.concatenate([[0.0, t1], rng.uniform(0.0, t1, nbps - 2)])) if nbps > 2 else np.array([0.0, t1])
    if unsorted and nbps > 4:
        # the reference does not require ascending breakpoints: shuffle the interior, repeat one,
        # put one past the last knot (extrapolated, like linspace's accumulated end point).  A
        # breakpoint BEFORE the first knot is undefined in the reference itself: interv returns
        # left = 1 and bsplvb then reads t(left+1-j) in front of the knot array.
        inner = spec.bps[1:-1].copy()
        rng.shuffle(inner)
        inner[0] = inner[1]
        inner[3] = t1 + 0.05
        spec.bps = np.concatenate([[0.0], inner, [t1]])
    spec.nbps = nbps
    spec.name = f"random_{family}"
    return spec


@pytest.mark.parametrize("seed", list(range(22)))
def test_random_shapes(torch_cuda, port, seed, monkeypatch):
    rng = np.random.default_rng(1000 + seed)
    family = "endpt" if seed % 2 == 0 else "kincar"
    spec = _random_spec(rng, family, unsorted=seed >= 14)
    if seed % 4 == 2:
        monkeypatch.setenv("NTG_B200_KERNEL", "general")
    P = int(rng.choice([1, 5, 37]))
    X = rng.uniform(-2.0, 2.0, (P, spec.nC))
    o = port.eval(spec, X, dense=False, band=True)
    pb, r = gpu_eval(torch_cuda, spec, X, False, want_Z=False)
    B, off, _ = pb.tables()
    Bo, offo, col0o = port.tables(spec)
    for a, b in zip(B, Bo):
        assert_bitexact(a, b, "K0 table")
    assert_bitexact(off, offo, "offsets")
    assert_bitexact(pb.pattern()[0], col0o, "pattern")
    cmp = assert_close if family == "endpt" else assert_bitexact   # endpt calls libm
    for k in ("f", "g", "c", "Jband"):
        cmp(r[k], o[k], f"seed {seed} ({family}, order {spec.order}, mult {spec.mult}, "
                        f"ninterv {spec.ninterv}, nbps {spec.nbps}, P {P}): {k}")
    pb.close()


def test_callback_abort_request(torch_cuda):
    """A callback that writes *mode = -1 asks the solver to stop (reference src/ntg.c:369): the
    device path raises the abort flag, ntgb_eval_host returns NTGB_EABORT."""
    import torch
    from ntg_b200 import Problem
    from ntg_b200.problem import NtgError
    spec, X = golden_spec("endpoint")
    pb = Problem(spec, 0)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    Xd = torch.from_numpy(X).cuda()
    out = pb.alloc_outputs(X.shape[0])
    st = torch.cuda.current_stream().cuda_stream
    pb.launch(pb.eval_args(Xd, out, stream=st, abort_flag=flag))
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    Xbad = X.copy()
    Xbad[3, :] = 1.0e7          # ep_ucf sets *mode = -1 when zp[0][0] > 1e6
    pb.launch(pb.eval_args(torch.from_numpy(Xbad).cuda(), out, stream=st, abort_flag=flag))
    torch.cuda.synchronize()
    assert int(flag.item()) == 1
    with pytest.raises(NtgError, match="-6"):
        pb.eval_host(Xbad)
    pb.eval_host(X)              # and the next clean call succeeds again
    pb.close()


@pytest.mark.parametrize("case", ["k1s_endpoint", "k1s_kincar64", "k1s_endpoint320", "k1_endpoint", "k1_syn6", "k1c_301", "k1c_601", "k1ch_301",
                                  "k1ch_300", "k1ch_601"])
@pytest.mark.parametrize("jac", [JAC_BAND, JAC_DENSE], ids=["band", "dense"])
def test_no_out_of_bounds_writes(torch_cuda, case, jac, monkeypatch):
    """compute-sanitizer is closed on this pool: every output lives inside a larger buffer of
    sentinels; after a launch the guard zones on both sides must be untouched and the interior of
    the fully-written outputs must hold no sentinel."""
    torch = torch_cuda
    from ntg_b200 import Problem
    from ntg_b200.abi import NtgbEvalArgs
    spec, P = {"k1s_endpoint": (configs.endpoint(), 7), "k1s_kincar64": (configs.kincar(64), 9),
               "k1s_endpoint320": (configs.endpoint(320, name="e320"), 5),
               "k1_endpoint": (configs.endpoint(), 7), "k1_syn6": (configs.syn6(12, name="s12"), 3),
               "k1c_301": (configs.syn6(150, name="s150"), 3), "k1c_601": (configs.syn6(300, name="s300"), 2),
               "k1ch_301": (configs.syn6(150, name="s150"), 3), "k1ch_300": (configs.syn6(150, nbps=300, name="s150e"), 4),
               "k1ch_601": (configs.syn6(300, name="s300"), 2)}[case]
    hot = case.startswith("k1ch")   # no Z requested: the steady-state kernels
    if hot and jac == JAC_DENSE:
        pytest.skip("steady-state kernels are band-layout only")
    if case == "k1_endpoint":
        monkeypatch.setenv("NTG_B200_KERNEL", "general")
    if jac == JAC_DENSE and spec.ncnln * spec.nC * P > 4e7:
        pytest.skip("dense Jacobian too large for this case")
    pb = Problem(spec, 0)
    d = pb.dims
    PAD, S = 4096, -7.25e300
    sizes = {"f": P, "g": P * d.nC, "c": P * d.ncnln, "result": P * 2, "Z": P * d.nZ,
             "J": P * d.ncnln * (d.sorder if jac == JAC_BAND else d.nC)}
    big = {k: torch.full((n + 2 * PAD,), S, dtype=torch.float64, device="cuda") for k, n in sizes.items()}
    X = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, (P, spec.nC))).cuda()
    a = NtgbEvalArgs()
    a.P, a.C, a.mode_obj, a.mode_con, a.nstate = P, X.data_ptr(), 2, 2, 0
    for k in ("f", "g", "c", "J", "Z", "result"):
        if not (hot and k == "Z"):
            setattr(a, k, big[k].data_ptr() + PAD * 8)
    a.jac_layout = jac
    a.stream = torch.cuda.current_stream().cuda_stream
    pb.launch(a)
    torch.cuda.synchronize()
    for k, n in sizes.items():
        b = big[k].cpu().numpy()
        assert np.all(b[:PAD] == S) and np.all(b[PAD + n:] == S), f"{case}: wrote outside {k}"
        if hot and k == "Z":
            assert np.all(b == S), "Z was not requested"
        elif k != "J" or jac == JAC_BAND:
            assert not np.any(b[PAD:PAD + n] == S), f"{case}: {k} not fully written"
    pb.close()


def test_alternating_batch_sizes_keep_launching(torch_cuda, port):
    """Regression: the shared-memory opt-in is a per-kernel LIMIT; alternating between batch sizes
    that need more and less dynamic shared memory (what ntgb_eval_host's chunking does) must not
    lower it under a configuration that is used again."""
    spec, _ = configs.get("cfg4")
    from ntg_b200 import Problem
    import torch
    pb = Problem(spec, 0, fast=True)
    X = configs.coefficients("cfg4", 70000, spec)
    ref = port.eval(spec, X[:64], dense=False, band=False)
    for P in (65536, 5829, 1417, 5829, 65536, 300, 65536):
        o = pb.eval(torch.from_numpy(X[:P]).cuda())
        torch.cuda.synchronize()
        assert_close(o["f"][:64].cpu().numpy(), ref["f"], f"P={P}")
    h = pb.eval_host(X[:20000])
    h = pb.eval_host(X[:20000])
    assert_close(h["f"][:64], ref["f"], "chunked host path, second call")
    # two chunks (~192 MB of results each): every row equals the device-buffer path, bit for bit
    o = pb.eval(torch.from_numpy(X[:20000]).cuda())
    torch.cuda.synchronize()
    for k in ("f", "g", "result"):
        assert np.array_equal(h[k], o[k].cpu().numpy()), f"chunked host path: {k}"
    assert np.array_equal(h["J"].reshape(20000, -1), o["J"].cpu().numpy().reshape(20000, -1)), "chunked host path: J"
    pb.close()


def test_cfg5_full_batch_properties(torch_cuda, port):
    """BASELINE.json's synthetic config at its full size on one GPU (16384 problems, 6 outputs,
    order 8, 401 breakpoints: 11.6 GB of results): a slice evaluated on its own is bit-identical to
    the same rows of the full batch, eight scattered problems match the CPU oracle bit for bit
    (exact variant), and the (objective, violation) table agrees with f and c."""
    torch = torch_cuda
    from ntg_b200 import Problem
    spec, P = configs.get("cfg5")
    Xh = configs.coefficients("cfg5", P, spec)
    X = torch.from_numpy(Xh).cuda()
    pb = Problem(spec, 0, fast=False)
    a = pb.eval(X)
    lo, hi = 9000, 9000 + 77
    b = pb.eval(X[lo:hi].contiguous())
    for k in ("f", "g", "c", "J", "result"):
        assert torch.equal(b[k], a[k][lo:hi]), f"{k}: depends on batch slicing"
    pick = np.array([0, 1, 4095, 8191, 9001, 12345, 16000, 16383])
    o = port.eval(spec, Xh[pick], dense=False, band=True)
    idx = torch.from_numpy(pick).cuda()
    assert_bitexact(a["f"][idx].cpu().numpy(), o["f"], "f")
    assert_bitexact(a["g"][idx].cpu().numpy(), o["g"], "g")
    assert_bitexact(a["c"][idx].cpu().numpy()[:, :spec.ncnln], o["c"], "c")
    assert_bitexact(pb.band_to_rows(a["J"][idx].cpu().numpy()), o["Jband"], "J")
    assert torch.equal(a["result"][:, 0], a["f"])
    assert_close(a["result"][idx, 1].cpu().numpy(), violation(spec, o["c"]), "violation")
    del a, b
    torch.cuda.empty_cache()
    pb.close()


@pytest.mark.parametrize("order,mult,ninterv,nbps,P", [(20, 3, 4, 33, 9), (20, 19, 7, 64, 5), (12, 5, 3, 300, 3), (1 + 2, 0, 5, 9, 4)])
def test_order_limits(torch_cuda, port, order, mult, ninterv, nbps, P):
    """Maximum sizes: order 20 is PGS bsplvb's limit (jmax = 20); also
    multiplicity order-1 (smoothest), multiplicity 0, and a long horizon on the general kernel."""
    spec = configs.high_order(order, mult, ninterv, nbps)
    X = np.random.default_rng(order * 100 + mult).uniform(-1, 1, (P, spec.nC))
    o = port.eval(spec, X, dense=False, band=True)
    pb, r = gpu_eval(torch_cuda, spec, X, False)
    B, off, _ = pb.tables()
    Bo, offo, _ = port.tables(spec)
    assert_bitexact(B[0], Bo[0], "K0 table at the order limit")
    assert_bitexact(off, offo, "offsets")
    for k in ("f", "g", "c", "Jband"):
        assert_bitexact(r[k], o[k], f"order {order} mult {mult}: {k}")
    pb.close()


def test_empty_batch_is_a_no_op(torch_cuda):
    import torch
    from ntg_b200 import Problem
    spec, _ = configs.get("cfg2")
    pb = Problem(spec, 0)
    X = torch.zeros((0, spec.nC), dtype=torch.float64, device="cuda")
    out = pb.alloc_outputs(1)
    a = pb.eval_args(torch.zeros((1, spec.nC), dtype=torch.float64, device="cuda"), out)
    a.P = 0
    pb.launch(a)
    torch.cuda.synchronize()
    assert float(out["f"].abs().sum()) == 0.0
    pb.close()


@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
def test_data_dependent_sparsity_falls_back_to_dense(torch_cuda, port, fast):
    """The build-time probe declares derivative entries that packs/cond.c writes only for |z| > 50
    structurally zero (checked in tests/test_build.py); coefficients that reach those branches must
    still give the reference's results: the kernel verifies the probed zeros at every evaluation and
    takes the dense chain rule when one of them is not +0.0."""
    spec = configs.conditional()
    rng = np.random.default_rng(4242)
    X = rng.uniform(-2.0, 2.0, (96, spec.nC))
    X[32:64] = rng.uniform(-120.0, 120.0, (32, spec.nC))      # both conditional branches fire
    X[64:] = rng.uniform(40.0, 60.0, (32, spec.nC))           # around the threshold: mixed within a warp
    o = port.eval(spec, X, dense=False, band=True)
    zfired = (X[32:64].max() > 50.0) and (X[32:64].min() < -50.0)
    assert zfired
    pb, r = gpu_eval(torch_cuda, spec, X, fast)
    cmp = assert_close if fast else assert_bitexact
    for k in ("f", "g", "c", "Jband"):
        cmp(r[k], o[k], f"conditional sparsity ({'fast' if fast else 'exact'}): {k}")
    # the conditional entries really are non-zero somewhere (otherwise this test proves nothing)
    J = o["Jband"]
    assert np.count_nonzero(o["g"]) > 0 and np.isfinite(J).all()
    pb.close()


def test_eval_host_small_calls_like_npsol(torch_cuda, port, monkeypatch):
    """Small ntgb_eval_host calls go through one device-mapped staging block (no copies).  NPSOL's
    pattern -- funobj (f, g) and funcon (c, dense J) alternating at P = 1 -- plus changes of batch
    size, Jacobian layout and a Z request in between must keep every result equal to the copying
    path's, INCLUDING the zeros outside the Jacobian band that are only written when a region is
    first used."""
    from ntg_b200 import Problem
    spec, X = golden_spec("endpoint")
    o = port.eval(spec, X, dense=True, band=True)
    pb = Problem(spec, 0)

    def check(h, idx, jac, tag):
        assert_bitexact(h["f"], o["f"][idx], f"{tag}: f")
        assert_close(h["g"], o["g"][idx], f"{tag}: g")
        assert_close(h["c"], o["c"][idx], f"{tag}: c")
        if jac == JAC_DENSE:
            assert_close(h["J"], np.nan_to_num(o["Jdense"][idx], nan=0.0), f"{tag}: dense J incl. out-of-band zeros")
        elif jac == JAC_BAND:
            assert_close(pb.band_to_rows(h["J"]), o["Jband"][idx], f"{tag}: band J")

    seq = [(slice(0, 1), JAC_DENSE), (slice(1, 2), JAC_DENSE), (slice(0, 3), JAC_DENSE), (slice(2, 3), JAC_BAND),
           (slice(0, 1), JAC_DENSE), (slice(0, 5), JAC_BAND), (slice(3, 4), JAC_DENSE)]
    for step, (idx, jac) in enumerate(seq):
        ho = pb.eval_host(X[idx], mode_obj=2, mode_con=-1, jac=JAC_NONE)                 # funobj
        assert_bitexact(ho["f"], o["f"][idx], f"step {step}: funobj f")
        h = pb.eval_host(X[idx], mode_obj=2, mode_con=2, jac=jac, want_Z=(step == 3))  # funobj + funcon
        check(h, idx, jac, f"step {step}")
    # J without Z, then Z WITHOUT J (Z is carved where that J region started), then J again: the
    # out-of-band zeros of the third call must not be the second call's flat outputs
    for jac in (JAC_DENSE, JAC_BAND):
        check(pb.eval_host(X[0:1], jac=jac), slice(0, 1), jac, "J, no Z")
        hz = pb.eval_host(X[0:1], jac=JAC_NONE, want_Z=True)
        assert np.abs(hz["Z"]).max() > 0
        check(pb.eval_host(X[0:1], jac=jac), slice(0, 1), jac, "J again after a Z-only call")
    # the copying path gives the same bits
    monkeypatch.setenv("NTG_B200_NO_ZEROCOPY", "1")
    h2 = pb.eval_host(X[0:1], jac=JAC_DENSE)
    monkeypatch.delenv("NTG_B200_NO_ZEROCOPY")
    h1 = pb.eval_host(X[0:1], jac=JAC_DENSE)
    for k in ("f", "g", "c", "J"):
        assert_bitexact(h1[k], h2[k], f"staging block vs copies: {k}")
    pb.close()


def test_fused_peer_gather_matches_nccl_all_gather(torch_cuda):
    """Two ranks (one per GPU, torchrun): the gathered (objective, violation) tables the evaluators
    write with peer stores equal an NCCL all_gather bit for bit -- K1s steady state and values-only,
    a ragged shard, K1c.  Needs two GPUs with peer access; skipped on a one-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "peer_gather_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", tool],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PEER GATHER OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
