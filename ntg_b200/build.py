"""Build recipe: nvcc cross-compiles everything for sm_100a, in-tree.

  ntg_b200/lib/libntg_b200.so         core: C ABI, K0, registry, ntg() drop-in
  ntg_b200/lib/libntgpack_<name>.so   one per callback pack: the user's C file
                                      compiled a second time as __device__ code
                                      + the fused evaluator instantiated on it

A pack is described by a small manifest (PACKS below, or tools/ntg_pack.py on
the command line).  The generated wrapper translation unit
  1. pre-includes the libc headers the user file includes,
  2. forward-declares the named callbacks `__host__ __device__` (the later
     plain C definition then compiles for both sides),
  3. #includes the UNMODIFIED user file (main() renamed if present),
  4. emits the traits struct and NTGB_DEFINE_PACK(), whose static initialiser
     registers {host callback address -> device launcher} with the core.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ntg_b200")
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
GEN = os.path.join(ROOT, "build", "gen")
INCLUDE = os.path.join(ROOT, "include")
REFERENCE = os.environ.get("NTG_REFERENCE", "/root/reference")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
          "-I", INCLUDE, "-I", CSRC, "-diag-suppress", "2464", "-diag-suppress", "177",
          "-diag-suppress", "550"]

SIGS = {
    "icf": "void {n}(int *, int *, double *, double *, double **)",
    "ucf": "void {n}(int *, int *, int *, double *, double *, double **)",
    "fcf": "void {n}(int *, int *, double *, double *, double **)",
    "nlicf": "void {n}(int *, int *, double *, double **, double **)",
    "nltcf": "void {n}(int *, int *, int *, double *, double **, double **)",
    "nlfcf": "void {n}(int *, int *, double *, double **, double **)",
}
ROLE_T = {"icf": "ntg_icf_t", "ucf": "ntg_ucf_t", "fcf": "ntg_fcf_t",
          "nlicf": "ntg_nlicf_t", "nltcf": "ntg_nltcf_t", "nlfcf": "ntg_nlfcf_t"}


@dataclass
class PackManifest:
    name: str
    src: str                                  # path of the user's C file
    maxderiv: List[str]                       # per output; C expressions allowed (e.g. "MAXDERIV")
    max_order: str                            # C expression
    callbacks: Dict[str, str] = field(default_factory=dict)   # role -> function name
    counts: Dict[str, str] = field(default_factory=dict)      # nlicf/nltcf/nlfcf -> count expr
    static: List[str] = field(default_factory=list)           # callbacks declared `static` in the file
    device_helpers: List[str] = field(default_factory=list)   # extra prototypes to mark __host__ __device__
    rename_main: Optional[str] = None         # rename main() to this symbol (exported with C linkage)
    main_args: str = "void"                   # parameter list of the user's main()
    c_compat: bool = False                    # file uses C-only idioms (implicit void* conversions)
    exact: bool = True                        # -fmad=false: reference operation order, bit-comparable


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA evaluator cannot be built (there is no CPU fallback)")
    return exe


def _newer(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd: List[str], verbose: bool) -> str:
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]) + " ...")
    return r.stdout + r.stderr


CORE_SO = os.path.join(LIB, "libntg_b200.so")


def build_core(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    srcs = [os.path.join(CSRC, "ntg_core.cu"), os.path.join(CSRC, "ntg_dropin.cu")]
    deps = srcs + [os.path.join(CSRC, h) for h in ("ntg_kernel_args.h", "pgs_device.cuh")] + \
        [os.path.join(INCLUDE, h) for h in ("ntg_b200.h", "ntg.h")]
    if force or _newer(CORE_SO, deps):
        _run([nvcc()] + ARCH + COMMON + ["-fmad=false", "-o", CORE_SO] + srcs + ["-ldl"], verbose)
    return CORE_SO


def pack_so(name: str) -> str:
    return os.path.join(LIB, f"libntgpack_{name}.so")


CON_ROLES = ("nlicf", "nltcf", "nlfcf")


def probe_sparsity(m: PackManifest, verbose: bool = False) -> Dict[str, List[int]]:
    """Which entries of its derivative vector does each callback ever set to something other
    than +0.0?  The user's file is compiled once more with gcc (plain C, main() renamed,
    unresolved library symbols left alone -- nothing but the callbacks is called) together with
    a small generated main() that calls every callback at 48 points in modes 1 and 2 and prints,
    per derivative row, a bit mask over iz_j + l.  The kernels use the masks to skip the chain-rule
    terms of entries that are structurally zero and CHECK the assumption on the device at every
    evaluation (ntg_eval_small.cuh::sp_clean), so a wrong probe costs speed, never correctness.
    Returns {} when the probe cannot be built or run (the kernels then treat everything as dense)."""
    base = m.name[:-5] if m.name.endswith("_fast") else m.name
    os.makedirs(GEN, exist_ok=True)
    cache = os.path.join(GEN, f"probe_{base}.txt")
    src = os.path.abspath(m.src)
    roles = {r: fn for r, fn in m.callbacks.items() if fn}
    key = "key " + repr((sorted(roles.items()), m.maxderiv, sorted(m.counts.items())))

    def parse(text: str) -> Dict[str, List[int]]:
        out: Dict[str, List[int]] = {}
        for line in text.splitlines():
            w = line.split()
            if len(w) >= 2 and w[0] in SIGS:
                out[w[0]] = [int(x, 16) for x in w[1:]]
        return out

    if os.path.exists(cache) and not _newer(cache, [src, os.path.abspath(__file__)]):
        text = open(cache).read()
        if text.startswith(key + "\n"):
            return parse(text)
    nout = len(m.maxderiv)
    L = []
    for h in ("assert.h", "float.h", "math.h", "stdio.h", "stdlib.h", "string.h", "unistd.h"):
        L.append(f"#include <{h}>")
    L.append('#include "ntg.h"')
    L.append("#define main ntg_probe_user_main_")
    L.append(f'#include "{src}"')
    L.append("#undef main")
    L.append("static unsigned long long s_ = 88172645463325252ULL;")
    L.append("static double rnd_(void) { s_ ^= s_ << 13; s_ ^= s_ >> 7; s_ ^= s_ << 17; "
             "return (double)(s_ >> 11) / 9007199254740992.0 * 6.0 - 3.0; }")
    L.append("int main(void) {")
    L.append(f"    enum {{ NOUT_ = {nout}, MAXC_ = 64 }};")
    L.append("    int md_[NOUT_] = {" + ", ".join(f"({e})" for e in m.maxderiv) + "};")
    L.append("    int nz_ = 0, j_, l_, t_, mode_, r_;")
    L.append("    double z_[512], *zp_[NOUT_], f_[MAXC_], dfs_[MAXC_][512], *dfp_[MAXC_];")
    L.append("    unsigned long long mask_[6][MAXC_];")
    L.append("    memset(mask_, 0, sizeof mask_);")
    L.append("    for (j_ = 0; j_ < NOUT_; j_++) { zp_[j_] = z_ + nz_; nz_ += md_[j_]; }")
    L.append("    if (nz_ > 64) return 2;")
    L.append("    for (r_ = 0; r_ < MAXC_; r_++) dfp_[r_] = dfs_[r_];")
    L.append("    for (t_ = 0; t_ < 48; t_++) {")
    L.append("        for (l_ = 0; l_ < nz_; l_++) z_[l_] = t_ == 0 ? 0.0 : (t_ == 1 ? 1.0 : (t_ == 2 ? -1.0 : rnd_()));")
    L.append("        for (mode_ = 1; mode_ <= 2; mode_++) {")
    L.append("            int nstate_ = t_ == 3 ? 1 : 0, i_ = t_ % 7, mo_;")
    order = list(SIGS)
    for ri, role in enumerate(order):
        fn = roles.get(role)
        if not fn:
            continue
        if role in CON_ROLES:
            cnt = m.counts.get(role, "0")
            L.append(f"            if (({cnt}) > MAXC_) return 3;")
            L.append("            memset(dfs_, 0, sizeof dfs_); memset(f_, 0, sizeof f_); mo_ = mode_;")
            args = "&mo_, &nstate_, " + ("&i_, " if role == "nltcf" else "") + "f_, dfp_, zp_"
            L.append(f"            {fn}({args});")
            L.append(f"            for (r_ = 0; r_ < ({cnt}); r_++) for (l_ = 0; l_ < nz_; l_++) {{")
            L.append("                unsigned long long b_; memcpy(&b_, &dfs_[r_][l_], 8);")
            L.append(f"                if (b_ != 0) mask_[{ri}][r_] |= 1ULL << l_; }}")
        else:
            L.append("            memset(dfs_, 0, sizeof dfs_); memset(f_, 0, sizeof f_); mo_ = mode_;")
            args = "&mo_, &nstate_, " + ("&i_, " if role == "ucf" else "") + "f_, dfs_[0], zp_"
            L.append(f"            {fn}({args});")
            L.append("            for (l_ = 0; l_ < nz_; l_++) {")
            L.append("                unsigned long long b_; memcpy(&b_, &dfs_[0][l_], 8);")
            L.append(f"                if (b_ != 0) mask_[{ri}][0] |= 1ULL << l_; }}")
    L.append("        }")
    L.append("    }")
    for ri, role in enumerate(order):
        if not roles.get(role):
            continue
        cnt = m.counts.get(role, "0") if role in CON_ROLES else "1"
        L.append(f'    printf("{role}"); for (r_ = 0; r_ < ({cnt}); r_++) printf(" %llx", mask_[{ri}][r_]); printf("\\n");')
    L.append("    return 0;")
    L.append("}")
    csrc = os.path.join(GEN, f"probe_{base}.c")
    exe = os.path.join(GEN, f"probe_{base}")
    with open(csrc, "w") as fh:
        fh.write("\n".join(L) + "\n")
    # the NTG entry points a user program calls from main() (include/ntg.h): never called by the
    # probe, defined here only so that it links without the CUDA library
    stubs = os.path.join(GEN, "probe_stubs.c")
    with open(stubs, "w") as fh:
        fh.write("#include <stdlib.h>\n" + "".join(
            f"void {n}() {{ abort(); }}\n" for n in
            ("ntg", "npsoloption", "linspace", "printNTGBanner", "MakeMatrix", "FreeMatrix", "DoubleMatrix",
             "FreeDoubleMatrix", "PrintMatrix", "PrintVector", "PrintiVector", "SplineInterp")))
    gcc = shutil.which("gcc")
    text = ""
    try:
        if gcc is None:
            raise RuntimeError("gcc not found")
        cmd = [gcc, "-O0", "-w", "-I", INCLUDE, "-o", exe, csrc, stubs, "-lm"]
        if verbose:
            print("+", " ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-400:])
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        if r.returncode != 0:
            raise RuntimeError(f"probe exited {r.returncode}")
        text = r.stdout
    except Exception as e:  # dense masks are always correct
        if verbose:
            print(f"sparsity probe of pack {m.name} unavailable ({e}); treating derivatives as dense", flush=True)
        text = ""
    with open(cache, "w") as fh:
        fh.write(key + "\n" + text)
    return parse(text)


def generate_wrapper(m: PackManifest, verbose: bool = False) -> str:
    os.makedirs(GEN, exist_ok=True)
    out = os.path.join(GEN, f"pack_{m.name}.cu")
    nout = len(m.maxderiv)
    L = []
    L.append(f"/* generated by ntg_b200/build.py for pack '{m.name}' -- do not edit */")
    for h in ("assert.h", "float.h", "math.h", "stdio.h", "stdlib.h", "string.h", "unistd.h"):
        L.append(f"#include <{h}>")
    L.append('#include "ntg.h"')
    L.append('#include "ntg_eval_cluster_hot.cuh"')
    if m.c_compat:
        L.append("/* C-only idiom support: `double *p = calloc(...)` must parse as C++ */")
        L.append("namespace { struct ntg_voidp { void *p; template <class T> operator T *() const "
                 "{ return static_cast<T *>(p); } };")
        L.append("inline ntg_voidp ntg_calloc_(size_t n, size_t s) { return ntg_voidp{calloc(n, s)}; }")
        L.append("inline ntg_voidp ntg_malloc_(size_t n) { return ntg_voidp{malloc(n)}; } }")
        L.append("#define calloc(n, s) ntg_calloc_((n), (s))")
        L.append("#define malloc(n) ntg_malloc_((n))")
    L.append('extern "C" {')
    for role, fn in m.callbacks.items():
        if not fn:
            continue
        st = "static " if fn in m.static else ""
        L.append(f"{st}__host__ __device__ " + SIGS[role].format(n=fn) + ";")
    for proto in m.device_helpers:
        L.append(f"__host__ __device__ {proto};")
    L.append("}")
    if m.rename_main:
        L.append(f'extern "C" int {m.rename_main}({m.main_args});')
        L.append(f"#define main {m.rename_main}")
    L.append(f'#include "{os.path.abspath(m.src)}"')
    if m.rename_main:
        L.append("#undef main")
    if m.c_compat:
        L.append("#undef calloc")
        L.append("#undef malloc")
    tn = f"ntgb_traits_{m.name}"
    L.append(f"struct {tn} {{")
    L.append(f"    static constexpr int kNout = {nout};")
    L.append("    __host__ __device__ static constexpr int md(int j) {")
    L.append("        constexpr int t[] = {" + ", ".join(f"({e})" for e in m.maxderiv) + "};")
    L.append("        return t[j];")
    L.append("    }")
    L.append(f"    static constexpr int kMaxOrd = ({m.max_order});")
    L.append(f"    static constexpr bool kExact = {'true' if m.exact else 'false'};")
    for role, key in (("nlicf", "kNnlic"), ("nltcf", "kNnltc"), ("nlfcf", "kNnlfc")):
        cnt = m.counts.get(role, "0") if m.callbacks.get(role) else "0"
        L.append(f"    static constexpr int {key} = ({cnt});")
    for role in SIGS:
        fn = m.callbacks.get(role)
        fn = f"::{fn}" if fn else "nullptr"
        L.append(f"    static constexpr {ROLE_T[role]} cb_{role} = {fn};")
        L.append(f"    static {ROLE_T[role]} cb_{role}_host() {{ return {fn}; }}")
    # structural sparsity of the derivative vectors (probe_sparsity): bit iz_j + l
    sp = probe_sparsity(m, verbose)
    for role in SIGS:
        masks = sp.get(role) if m.callbacks.get(role) else None
        if role in CON_ROLES:
            body = ("constexpr unsigned long long t[] = {" + ", ".join(f"0x{v:x}ull" for v in masks) + "}; return t[m];"
                    if masks else "return ~0ull;")
            L.append(f"    __host__ __device__ static constexpr unsigned long long sp_{role}(int m) {{ {body} }}")
        else:
            v = f"0x{masks[0]:x}ull" if masks else "~0ull"
            L.append(f"    __host__ __device__ static constexpr unsigned long long sp_{role}() {{ return {v}; }}")
    L.append("};")
    L.append(f"NTGB_DEFINE_PACK({m.name}, {tn}, {1 if m.exact else 0})")
    text = "\n".join(L) + "\n"
    if not os.path.exists(out) or open(out).read() != text:
        with open(out, "w") as fh:
            fh.write(text)
    return out


def build_pack(m: PackManifest, verbose: bool = False, force: bool = False, ptxas_v: bool = False) -> str:
    core = build_core(verbose)
    wrapper = generate_wrapper(m, verbose)
    so = pack_so(m.name)
    deps = [wrapper, os.path.abspath(m.src), core, os.path.join(CSRC, "ntg_eval_kernel.cuh"), os.path.join(CSRC, "ntg_eval_small.cuh"), os.path.join(CSRC, "ntg_small_plan.h"), os.path.join(CSRC, "ntg_eval_cluster.cuh"), os.path.join(CSRC, "ntg_eval_cluster_hot.cuh"),
            os.path.join(CSRC, "ntg_kernel_args.h"), os.path.join(INCLUDE, "ntg_b200.h"),
            os.path.join(INCLUDE, "ntg.h")]
    if force or ptxas_v or _newer(so, deps):
        cmd = [nvcc()] + ARCH + COMMON + ["-fmad=false" if m.exact else "-fmad=true"]
        cmd += os.environ.get("NTG_B200_NVCC_EXTRA", "").split()
        if ptxas_v:
            cmd += ["-Xptxas", "-v"]
        cmd += ["-o", so, wrapper, "-L", LIB, "-lntg_b200", "-Xlinker", "-rpath=$ORIGIN", "-Xlinker", "-Bsymbolic"]
        log = _run(cmd, verbose)
        if ptxas_v:
            print(log)
    return so


def _both(m: PackManifest) -> List[PackManifest]:
    """every pack is built twice: exact (reference operation order, -fmad=false)
    and fast (FMA contraction allowed)"""
    import copy
    f = copy.deepcopy(m)
    f.name = m.name + "_fast"
    f.exact = False
    return [m, f]


def repo_packs() -> List[PackManifest]:
    P = os.path.join(PKG, "packs")
    packs = [
        PackManifest("vdp", os.path.join(P, "vdp.c"), ["3"], "5",
                     {"ucf": "vdp_ucf", "nltcf": "vdp_nltcf"}, {"nltcf": "1"}),
        PackManifest("kincar", os.path.join(P, "kincar.c"), ["3", "3"], "5",
                     {"ucf": "kc_ucf", "nltcf": "kc_nltcf"}, {"nltcf": "2"}),
        PackManifest("syn6", os.path.join(P, "syn6.c"), ["4"] * 6, "8",
                     {"ucf": "syn6_ucf", "nltcf": "syn6_nltcf"}, {"nltcf": "4"}),
        PackManifest("endpt", os.path.join(P, "endpt.c"), ["3", "2"], "6",
                     {"icf": "ep_icf", "ucf": "ep_ucf", "fcf": "ep_fcf",
                      "nlicf": "ep_nlicf", "nltcf": "ep_nltcf", "nlfcf": "ep_nlfcf"},
                     {"nlicf": "2", "nltcf": "3", "nlfcf": "1"}),
        PackManifest("hi20", os.path.join(P, "hi20.c"), ["3"], "20",
                     {"ucf": "hi20_ucf", "nltcf": "hi20_nltcf"}, {"nltcf": "1"}),
        PackManifest("cond", os.path.join(P, "cond.c"), ["3", "3"], "4",
                     {"ucf": "cond_ucf", "nltcf": "cond_nltcf"}, {"nltcf": "2"}),
    ]
    out: List[PackManifest] = []
    for m in packs:
        out += _both(m)
    return out


def reference_example_packs() -> List[PackManifest]:
    """The reference's own example programs, compiled UNMODIFIED from where they
    lie (only possible where /root/reference exists; the built .so travels)."""
    ex = os.path.join(REFERENCE, "examples")
    if not os.path.isdir(ex):
        return []
    return [
        PackManifest("ref_vanderpol", os.path.join(ex, "vanderpol.c"), ["MAXDERIV"], "ORDER",
                     {"ucf": "ucf"}, static=["ucf"], rename_main="ntg_example_vanderpol_main",
                     c_compat=True),
        PackManifest("ref_kincar", os.path.join(ex, "kincar.c"), ["MAXDERIV", "MAXDERIV"], "ORDER",
                     {"ucf": "tcf"}, rename_main="ntg_example_kincar_main", main_args="int, char **",
                     c_compat=True),
    ]


def build_all(verbose: bool = False, force: bool = False) -> List[str]:
    built = [build_core(verbose, force)]
    packs = repo_packs() + reference_example_packs()
    # the packs are independent translation units: one nvcc per host core, up to 8 at a time
    from concurrent.futures import ThreadPoolExecutor
    jobs = int(os.environ.get("NTG_B200_BUILD_JOBS", "0")) or min(8, len(os.sched_getaffinity(0)))
    for m in packs:           # wrapper generation and the sparsity probes share files: serial
        generate_wrapper(m, verbose)
    with ThreadPoolExecutor(max_workers=max(1, jobs)) as ex:
        built += list(ex.map(lambda m: build_pack(m, verbose, force), packs))
    return built


if __name__ == "__main__":
    for so in build_all(verbose="-v" in sys.argv, force="-f" in sys.argv):
        print(so)
