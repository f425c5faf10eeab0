#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: share of stall samples and of executed instructions.
usage: scripts/ncu_hot_lines.py report.ncu-rep kernel_regex [launch_index] [top]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split into launches at 'Function Name' rows; group file sections of the same launch
launches, cur = [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1]
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        if cur is None or (fpath, "seen") in cur["files"]:
            cur = {"files": set(), "lines": collections.OrderedDict()}
            launches.append(cur)
        cur["files"].add((fpath, "seen"))
        cur["hdr"] = hdr
        cur["file"] = fpath
        continue
    if cur is None or len(r) != len(cur["hdr"]):
        continue
    h = cur["hdr"]
    ci = {n: i for i, n in enumerate(h)}
    if r[0]:      # a CUDA line row with aggregated metrics
        def f(name):
            try: return float(r[ci[name]])
            except Exception: return 0.0
        key = (cur["file"].split("/")[-1], r[0])
        cur["lines"][key] = (r[1].strip(), f("# Samples"), f("Instructions Executed"), f("stall_long_sb"), f("stall_barrier"))
L = launches[launch]
ts = sum(v[1] for v in L["lines"].values()) or 1
ti = sum(v[2] for v in L["lines"].values()) or 1
print(f"launch {launch} of {len(launches)}: samples {ts:.0f}, warp instructions {ti:.0f}")
for k, v in sorted(L["lines"].items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:>4}  samp {100*v[1]/ts:5.1f}%  inst {100*v[2]/ti:5.1f}%  long_sb {v[3]:6.0f} bar {v[4]:6.0f}  {v[0][:100]}")
