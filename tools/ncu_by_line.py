"""Stall samples of an .ncu-rep (captured with --import-source on, built with -lineinfo) by SOURCE line.
usage: python tools/ncu_by_line.py report.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
f, cur = None, None
agg, inst, src, why = collections.Counter(), collections.Counter(), {}, collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not r:
        continue
    if r[0] != "":
        try:
            cur = (f, int(r[0]))
            src[cur] = ",".join(r[1:len(r) - len(hdr) + 2])[:100]
        except ValueError:
            pass
        continue
    if len(r) != len(hdr) or cur is None:
        continue
    try:
        s, e = int(r[isamp]), int(r[iex])
    except ValueError:
        continue
    agg[cur] += s
    inst[cur] += e
    for i, h in stall:
        if r[i].isdigit():
            why[cur][h] += int(r[i])
tot = sum(agg.values()) or 1
print(f"total samples {tot}, instructions {sum(inst.values())}")
for k, v in agg.most_common(top):
    w = ", ".join(f"{h[6:]} {c}" for h, c in why[k].most_common(2))
    print(f"{100*v/tot:5.1f}% inst={inst[k]:>10d} {k[0]}:{k[1]:<5d} {src[k].strip()[:90]}  [{w}]")
