"""Stamp profiles/traffic.json from an ncu --set full report of the workload's dominant kernel.
usage: python tools/traffic_stamp.py cfg4 gpurun_out/k1s.ncu-rep profiles/r02_cfg4_K1s_ncu.txt [kernel-name-substring]
The entry carries bench.py::kernel_source_hash(cfg) of the tree the capture was taken with; bench.py
drops roofline.traffic to null when the kernel sources have changed since."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

cfg, rep, summary = sys.argv[1], sys.argv[2], sys.argv[3]
want = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
units = rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
vals = []
for r in rows[2:]:
    if want and want not in r[ik]:
        continue
    vals.append(float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]])
assert vals, "no matching kernel in the report"
from ntg_b200 import configs  # noqa: E402
spec, P = configs.get(cfg)
path = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(path))
t[cfg] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_in_report": len(vals),
          "kernel_source_hash": bench.kernel_source_hash(cfg), "from": summary}
json.dump(t, open(path, "w"), indent=1)
print(t[cfg])
