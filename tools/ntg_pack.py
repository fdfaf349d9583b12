#!/usr/bin/env python
"""Compile a user's UNMODIFIED NTG callback file into a device callback pack.

    python tools/ntg_pack.py --name vanderpol --src examples/vanderpol.c \
        --maxderiv MAXDERIV --max-order ORDER --ucf ucf --static ucf \
        --rename-main ntg_example_vanderpol_main --c-compat

    python tools/ntg_pack.py --name mycar --src mycar.c --maxderiv 3 3 --max-order 5 \
        --ucf tcf --nltcf my_constraints:2

--maxderiv / --max-order take C expressions (the file's own #defines work).
Constraint callbacks are given as function[:count].  The result is
ntg_b200/lib/libntgpack_<name>.so; load it (dlopen / ctypes / link) next to
libntg_b200.so and ntg() / ntgb_create() find the callbacks by host address.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntg_b200.build import PackManifest, build_pack  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--name", required=True)
    ap.add_argument("--src", required=True)
    ap.add_argument("--maxderiv", nargs="+", required=True, help="per output, C expressions")
    ap.add_argument("--max-order", required=True)
    for role in ("icf", "ucf", "fcf", "nlicf", "nltcf", "nlfcf"):
        ap.add_argument(f"--{role}", default="")
    ap.add_argument("--static", nargs="*", default=[], help="callbacks declared static in the file")
    ap.add_argument("--device-helper", action="append", default=[],
                    help="prototype of a helper the callbacks call, e.g. 'double sq(double)'")
    ap.add_argument("--rename-main", default=None)
    ap.add_argument("--main-args", default="void")
    ap.add_argument("--c-compat", action="store_true", help="file relies on C's implicit void* conversions")
    ap.add_argument("--fast", action="store_true", help="allow FMA contraction (default: exact, -fmad=false)")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    callbacks, counts = {}, {}
    for role in ("icf", "ucf", "fcf", "nlicf", "nltcf", "nlfcf"):
        v = getattr(a, role)
        if not v:
            continue
        fn, _, cnt = v.partition(":")
        callbacks[role] = fn
        if role.startswith("nl"):
            counts[role] = cnt or "1"
    m = PackManifest(a.name, a.src, a.maxderiv, a.max_order, callbacks, counts, static=a.static,
                     device_helpers=a.device_helper, rename_main=a.rename_main, main_args=a.main_args,
                     c_compat=a.c_compat, exact=not a.fast)
    print(build_pack(m, verbose=a.verbose, force=True, ptxas_v=a.verbose))


if __name__ == "__main__":
    main()
