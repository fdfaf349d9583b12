"""Development aid: build time-stamped copies of the vdp_fast / kincar_fast packs under
build/variants/stamps (read back by tools/gpu_stamps.py).  The product sources are not touched: a
copy of ntg_b200/csrc gets %globaltimer stamps at the phase boundaries of K1s."""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntg_b200 import build

scr = os.path.join(ROOT, "build", "scratch")
shutil.rmtree(os.path.join(scr, "csrc"), ignore_errors=True)
os.makedirs(scr, exist_ok=True)
shutil.copytree(build.CSRC, os.path.join(scr, "csrc"))
p = os.path.join(scr, "csrc", "ntg_eval_small.cuh")
s = open(p).read()

def rep(old, new):
    global s
    assert old in s, old
    s = s.replace(old, new, 1)

rep('namespace ntgb {\n', 'namespace ntgb {\n__device__ unsigned long long g_stamps[8 * 3 * 8];\n'
    '__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }\n'
    '#define STAMP(i) do { if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1 || blockIdx.x == gridDim.x / 2)) g_stamps[(A.nstate & 7) * 24 + (blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 1 ? 16 : 8)) + (i)] = gtime(); } while (0)\n')
rep('    extern __shared__ double smem[];\n', '    extern __shared__ double smem[];\n    STAMP(0);\n')
rep('        /* maximum constraint violation per problem: eight lanes per problem', '        STAMP(7);\n        /* maximum constraint violation per problem: eight lanes per problem')
rep('    if (flags & 1) {\n        asm volatile("griddepcontrol.wait;" ::: "memory");',
    '    __syncthreads(); STAMP(1);\n    if (flags & 1) {\n        asm volatile("griddepcontrol.wait;" ::: "memory");\n        STAMP(2);')
rep('        cp_async_wait_all();\n        __syncthreads(); /* coefficients of this tile landed',
    '        cp_async_wait_all();\n        __syncthreads(); STAMP(3); /* coefficients of this tile landed')
rep('        __syncthreads();\n\n        /* ------- phase B:', '        __syncthreads(); STAMP(4);\n\n        /* ------- phase B:')
rep('        /* the barrier at the top of the next iteration separates this phase B from the next phase A */\n    }',
    '        __syncthreads(); STAMP(5);\n    }')
rep('    cp_async_wait_all();\n    if constexpr (PUSH) { /* the last tile', '    STAMP(6);\n    cp_async_wait_all();\n    if constexpr (PUSH) { /* the last tile')
open(p, "w").write(s)
for vdir, extra in (("stamps", []),):
    out = os.path.join(ROOT, "build", "variants", vdir)
    os.makedirs(out, exist_ok=True)
    for name in ("vdp_fast", "kincar_fast"):
        m = [x for x in build.repo_packs() if x.name == name][0]
        w = build.generate_wrapper(m)
        txt = open(w).read() + ('\nextern "C" void ntg_read_stamps(unsigned long long *o) '
                                '{ cudaMemcpyFromSymbol(o, ntgb::g_stamps, sizeof(unsigned long long) * 192); }\n')
        w2 = os.path.join(scr, f"pack_{name}.cu")
        open(w2, "w").write(txt)
        common = list(build.COMMON)
        common[common.index(build.CSRC)] = os.path.join(scr, "csrc")
        cmd = [build.nvcc()] + build.ARCH + common + extra + ["-fmad=true", "-o", os.path.join(out, f"libntgpack_{name}.so"), w2,
                                                      "-L", build.LIB, "-lntg_b200", "-Xlinker", "-rpath=" + build.LIB, "-Xlinker", "-Bsymbolic"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(vdir, name, "rc", r.returncode, r.stderr[-400:])
