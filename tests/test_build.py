"""The pack generator: wrapper text and a real nvcc build of a pack from a user C file."""
import ctypes as C
import os
import textwrap

import pytest

from ntg_b200 import build


def test_wrapper_for_static_callback_and_renamed_main(tmp_path):
    src = tmp_path / "user.c"
    src.write_text(textwrap.dedent('''
        #include <stdlib.h>
        #include <math.h>
        #include "ntg.h"
        #define NOUT 1
        #define MAXDERIV 2
        #define ORDER 4
        static void cost(int *mode,int *nstate,int *i,double *f,double *df,double **zp);
        int main(void) { double *p = calloc(3, sizeof(double)); free(p); return 0; }
        void cost(int *mode,int *nstate,int *i,double *f,double *df,double **zp)
        { if (*mode != 1) *f = zp[0][0]*zp[0][0] + exp(zp[0][1]);
          if (*mode != 0) { df[0] = 2*zp[0][0]; df[1] = exp(zp[0][1]); } }
    '''))
    m = build.PackManifest("t_user", str(src), ["MAXDERIV"], "ORDER", {"ucf": "cost"}, static=["cost"],
                           rename_main="t_user_main", c_compat=True)
    w = open(build.generate_wrapper(m)).read()
    assert "static __host__ __device__ void cost(int *, int *, int *, double *, double *, double **);" in w
    assert '#define main t_user_main' in w and 'extern "C" int t_user_main(void);' in w
    assert "#define calloc(n, s) ntg_calloc_((n), (s))" in w
    assert "static constexpr int kMaxOrd = (ORDER);" in w and "constexpr int t[] = {(MAXDERIV)};" in w
    so = build.build_pack(m, force=True)
    try:
        from ntg_b200 import problem
        lib = problem.load_pack("t_user")           # registers itself on load
        pk = problem.core().ntgb_find_pack(b"t_user").contents
        assert pk.max_nout == 1 and pk.maxderiv[0] == 2 and pk.max_order == 4 and pk.ucf and not pk.nltcf
        assert lib.t_user_main() == 0               # the user's main(), C linkage, host side intact
    finally:
        os.remove(so)
        problem._packs.pop("t_user", None)


def test_reference_examples_compile_unmodified():
    """examples/vanderpol.c and examples/kincar.c from the reference tree build as packs as they are
    (only where /root/reference exists; elsewhere the prebuilt packs must be present)."""
    ms = build.reference_example_packs()
    if not ms:
        for n in ("ref_vanderpol", "ref_kincar"):
            if not os.path.exists(build.pack_so(n)):
                pytest.skip("reference tree absent and no prebuilt example packs")
        return
    for m in ms:
        so = build.build_pack(m)
        lib = C.CDLL(so)
        assert hasattr(lib, m.rename_main)


def test_sparsity_probe_masks():
    """probe_sparsity(): bit iz_j + l of a row's mask is set iff the callback ever produced a
    non-zero there at the probe points.  kincar (packs/kincar.c): the cost touches only the two
    second derivatives; cond.c hides two entries behind |z| > 50 that the probe cannot see -- the
    kernels' run-time check covers those (tests/test_gpu_parity.py)."""
    by = {m.name: m for m in build.repo_packs()}
    k = build.probe_sparsity(by["kincar"])
    assert k["ucf"] == [0b100100]
    assert k["nltcf"] == [0b010010, 0b110110]
    c = build.probe_sparsity(by["cond"])
    assert c["ucf"] == [0b100010] and c["nltcf"] == [0b000010, 0b100001]
    e = build.probe_sparsity(by["endpt"])
    assert set(e) == {"icf", "ucf", "fcf", "nlicf", "nltcf", "nlfcf"} and len(e["nltcf"]) == 3
    w = open(build.generate_wrapper(by["kincar"])).read()
    assert "sp_ucf() { return 0x24ull; }" in w and "0x12ull, 0x36ull" in w


def test_sparsity_probe_failure_means_dense(tmp_path):
    """a file the probe cannot compile must leave every mask dense (all ones), never fail the build"""
    src = tmp_path / "broken.c"
    src.write_text("void f(int *mode,int *nstate,int *i,double *f,double *df,double **zp) { this is not C }\n")
    m = build.PackManifest("t_broken", str(src), ["2"], "3", {"ucf": "f"})
    assert build.probe_sparsity(m) == {}
