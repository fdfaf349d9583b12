"""CPU checks of K1s's launch geometry (ntg_b200/csrc/ntg_small_plan.h, compiled with g++ through
tests/tools/small_plan_host.cpp): for every batch size and shape, the tiles the CTAs walk -- whole
tiles dealt round-robin, or the even split of batches of a few tiles per CTA -- cover every problem
exactly once, no tile exceeds its buffers, and the shared memory stays within the cap the kernel
is launched with.  The kernel's own walk is checked on the GPU (tests/test_gpu_parity.py::
test_even_split_launch_geometries, ::test_ragged_batches); this is the host half of that logic."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def plan_bin(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("plan") / "small_plan_host")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "ntg_b200", "csrc"), "-o", exe,
                           os.path.join(HERE, "tools", "small_plan_host.cpp")])
    return exe


def run_plan(exe, cases):
    """cases: (P, nbps, S, nout, nC, segtot, sm_count) -> dict rows"""
    inp = "".join(" ".join(str(int(v)) for v in c) + "\n" for c in cases)
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    keys = ("block", "G", "R", "rows", "ktiles", "even_grid", "smem", "grid", "ok", "max_tiles", "max_tile")
    rows = [dict(zip(keys, map(int, l.split()))) for l in out if l.strip()]
    assert len(rows) == len(cases)
    return rows


# shape = (nbps, S, nout, nC, segtot): CFG-2/3 (20 breakpoints), CFG-4 (64), the 320-breakpoint endpoint shape
SHAPES = {"cfg2": (20, 5, 1, 7, 3), "cfg3": (20, 10, 2, 14, 6), "cfg4": (64, 10, 2, 14, 6),
          "endpoint13": (13, 10, 2, 17, 7), "endpoint320": (320, 10, 2, 17, 7), "wide256": (256, 8, 1, 40, 10)}


def test_named_configurations_take_the_expected_geometry(plan_bin):
    """the cases DESIGN.md quotes: CFG-3 one tile of 28 per CTA on all 296 slots, CFG-2 whole tiles of 24
    on 171 CTAs (the tie goes to fewer CTAs), 8192 / 16384 lane changes two / four tiles of 14 per CTA,
    the full 65 536 whole tiles of 12 dealt round-robin (19 for the busiest CTA)"""
    r = run_plan(plan_bin, [(8192,) + SHAPES["cfg3"] + (148,), (4096,) + SHAPES["cfg2"] + (148,),
                            (8192,) + SHAPES["cfg4"] + (148,), (16384,) + SHAPES["cfg4"] + (148,),
                            (65536,) + SHAPES["cfg4"] + (148,)])
    assert (r[0]["ktiles"], r[0]["rows"], r[0]["grid"], r[0]["R"]) == (1, 28, 296, 3)
    assert (r[1]["ktiles"], r[1]["rows"], r[1]["grid"], r[1]["R"]) == (0, 24, 171, 2)
    assert (r[2]["ktiles"], r[2]["rows"], r[2]["grid"], r[2]["R"]) == (2, 14, 296, 4)
    assert (r[3]["ktiles"], r[3]["rows"], r[3]["grid"], r[3]["R"]) == (4, 14, 296, 4)
    assert (r[4]["ktiles"], r[4]["rows"], r[4]["grid"], r[4]["max_tiles"]) == (0, 12, 296, 19)
    assert all(x["ok"] == 1 for x in r)


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_every_problem_lands_in_exactly_one_tile(plan_bin, shape):
    rng = np.random.default_rng(abs(hash(shape)) % 2**31)
    sizes = list(range(1, 700)) + [int(v) for v in rng.integers(700, 200000, 400)] + \
        [296 * k + d for k in (1, 4, 12, 28, 36, 112) for d in (-1, 0, 1)] + [4096, 8192, 16384, 32768, 65536, 131072]
    cases = [(P,) + SHAPES[shape] + (sm,) for P in sizes for sm in ((148,) if P > 3000 else (148, 132, 8))]
    rows = run_plan(plan_bin, cases)
    nbps = SHAPES[shape][0]
    for c, r in zip(cases, rows):
        assert r["ok"] == 1, f"{shape} P={c[0]} sm={c[-1]}: {r}"
        assert r["rows"] <= r["G"] * r["R"] and 1 <= r["R"] <= 8
        assert r["block"] in (256, 512) or (r["block"] % 32 == 0 and r["block"] >= 64 and c[0] < (512 if nbps > 256 else 256) // nbps)
        assert r["G"] * nbps <= r["block"]
        cap_kb = 200 if nbps > 256 else 100
        assert r["smem"] <= cap_kb * 1024 or r["R"] == 1, f"{shape} P={c[0]}: {r}"
        assert r["smem"] <= 227 * 1024
        if r["ktiles"]:
            assert r["max_tiles"] <= r["ktiles"] <= 63 and r["even_grid"] == r["grid"]


def test_even_split_is_only_taken_for_a_few_tiles_per_cta(plan_bin):
    """beyond 8 tiles per CTA the round-robin tiles stay (measured: no gain, and contiguous shares are
    2 % slower at the full CFG-4 batch); below, the even split never gives a CTA more tiles than whole
    tiles would"""
    sizes = [1184 * k + d for k in range(1, 60) for d in (0, 7, 600)]
    rows = run_plan(plan_bin, [(P,) + SHAPES["cfg4"] + (148,) for P in sizes])
    for P, r in zip(sizes, rows):
        whole = -(-P // 12)
        per_cta = -(-whole // min(296, whole))
        if per_cta > 8:
            assert r["ktiles"] == 0, (P, r)
        if r["ktiles"]:
            assert r["ktiles"] * -(-r["rows"] // r["G"]) <= per_cta * 3 + 1, (P, r)
