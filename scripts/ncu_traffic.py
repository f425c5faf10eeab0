#!/usr/bin/env python
"""profiles/r2/traffic.json from an `ncu --set full` capture of scripts/profile_scan.py (the three C2 scans, once):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and duration of every kernel of every scan, tagged with the
hash of the kernel sources the capture was taken from -- bench.py quotes `roofline.traffic` only while that hash matches.

    python scripts/ncu_traffic.py gpurun_out/scan_full.ncu-rep [rows]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SELECTIVITIES, kernel_source_hash

rep = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
table = list(csv.reader(out.splitlines()))
hdr, units = table[0], table[1]
ci = {n: i for i, n in enumerate(hdr)}


def val(r, name):
    v = float(r[ci[name]].replace(",", ""))
    u = units[ci[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


launches = []
for r in table[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("mbc::", "")
    launches.append({"kernel": name, "seconds": val(r, "gpu__time_duration.sum"),
                     "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                     "registers": int(float(r[ci["launch__registers_per_thread"]]))})
# one scan = the launches up to and including agg_finish_kernel
scans, cur = [], []
for l in launches:
    cur.append(l)
    if l["kernel"].startswith("agg_finish") or l["kernel"].startswith("agg_count_only"):
        scans.append(cur)
        cur = []
scans = scans[-len(SELECTIVITIES):]
per_sel, per_kernel = {}, {}
for s, sc in zip(SELECTIVITIES, scans):
    per_sel[str(s)] = sum(l["dram_read"] + l["dram_write"] for l in sc)
    per_kernel[str(s)] = [{k: (round(v, 9) if isinstance(v, float) else v) for k, v in l.items()} for l in sc]
doc = {"workload": "c2", "rows": rows, "kernel_source_hash": kernel_source_hash(),
       "source": f"{os.path.basename(rep)}: ncu --set full --clock-control none of scripts/profile_scan.py {rows} 1",
       "unit": "bytes", "note": "dram__bytes_read.sum + dram__bytes_write.sum over the kernels of one scan",
       "per_selectivity": per_sel, "per_scan_mean": sum(per_sel.values()) / max(len(per_sel), 1), "kernels": per_kernel}
os.makedirs(os.path.join(ROOT, "profiles", "r2"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "r2", "traffic.json"), "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps({k: doc[k] for k in ("kernel_source_hash", "per_selectivity", "per_scan_mean")}))
