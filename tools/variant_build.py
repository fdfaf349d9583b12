"""Development aid: build an experimental copy of a pack from a PATCHED copy of ntg_b200/csrc
(product sources untouched) into build/variants/<name>/ for A/B timing with NTG_B200_PACK_DIR.
usage: python tools/variant_build.py NAME PACK 'old1=>new1' ['old2=>new2' ...] [--flags "-DX=1"]"""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntg_b200 import build

name, pack = sys.argv[1], sys.argv[2]
reps = [a for a in sys.argv[3:] if "=>" in a]
flags = []
if "--flags" in sys.argv:
    flags = sys.argv[sys.argv.index("--flags") + 1].split()
scr = os.path.join(ROOT, "build", "scratch", name)
shutil.rmtree(scr, ignore_errors=True)
shutil.copytree(build.CSRC, os.path.join(scr, "csrc"))
for r in reps:
    old, new = r.split("=>", 1)
    hit = 0
    for fn in os.listdir(os.path.join(scr, "csrc")):
        p = os.path.join(scr, "csrc", fn)
        s = open(p).read()
        if old in s:
            open(p, "w").write(s.replace(old, new))
            hit += 1
    assert hit, f"pattern not found: {old}"
out = os.path.join(ROOT, "build", "variants", name)
os.makedirs(out, exist_ok=True)
m = [x for x in build.repo_packs() if x.name == pack][0]
w = build.generate_wrapper(m)
common = list(build.COMMON)
common[common.index(build.CSRC)] = os.path.join(scr, "csrc")
cmd = [build.nvcc()] + build.ARCH + common + ["-fmad=false" if m.exact else "-fmad=true"] + flags + \
      ["-o", os.path.join(out, f"libntgpack_{pack}.so"), w, "-L", build.LIB, "-lntg_b200", "-Xlinker", "-rpath=" + build.LIB,
       "-Xlinker", "-Bsymbolic"]
r = subprocess.run(cmd, capture_output=True, text=True)
print(name, pack, "rc", r.returncode, r.stderr[-600:])
