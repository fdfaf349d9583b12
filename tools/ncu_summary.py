"""Summarise an .ncu-rep (details page + top stall instructions) into text.
usage: python tools/ncu_summary.py report.ncu-rep [out.txt]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout

def run(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout

rows = list(csv.reader(io.StringIO(run("details"))))
hdr = rows[0]
si, mi, vi, ui = (hdr.index(x) for x in ("Section Name", "Metric Name", "Metric Value", "Metric Unit"))
ki = hdr.index("Kernel Name")
print("kernel:", rows[1][ki], file=out)
keep = ("GPU Speed Of Light Throughput", "Memory Workload Analysis", "Warp State Statistics", "Scheduler Statistics",
        "Instruction Statistics", "Compute Workload Analysis")
names = ("Registers Per Thread", "Achieved Occupancy", "Theoretical Occupancy", "Grid Size", "Block Size",
         "Dynamic Shared Memory Per Block", "Waves Per SM")
for r in rows[1:]:
    if r[si] in keep or r[mi] in names:
        print(f"{r[si][:30]:30s} {r[mi]:52s} {r[vi]:>16s} {r[ui]}", file=out)
raw = list(csv.reader(io.StringIO(run("raw"))))
h, u, v = raw[0], raw[1], raw[2]
for w in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
    if w in h:
        i = h.index(w)
        print(f"{'raw':30s} {w:52s} {v[i]:>16s} {u[i]}", file=out)
src = list(csv.reader(io.StringIO(run("source"))))
if len(src) > 2:
    hdr, data = src[1], src[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot = sum(int(r[isamp]) for r in data) or 1
    print(f"\nSASS instructions: {len(data)}; stall samples: {tot}; top instructions by samples:", file=out)
    for n in sorted(sorted(range(len(data)), key=lambda n: -int(data[n][isamp]))[:18]):
        r = data[n]
        reasons = sorted(((hdr[i], int(r[i])) for i in range(len(hdr)) if hdr[i].startswith("stall_")
                          and "Not Issued" not in hdr[i] and r[i].isdigit() and int(r[i]) > 0), key=lambda x: -x[1])[:2]
        print(f"  #{n:5d} {100*int(r[isamp])/tot:5.1f}% exec={r[iex]:>9s} {r[isrc].strip()[:58]:58s} {reasons}", file=out)
