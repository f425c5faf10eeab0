#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libmbcol.so (cuobjdump -sass; no GPU needed): instruction count and the opcodes that
show which hardware path a kernel uses -- UBLKCP (cp.async.bulk, the TMA 1-D bulk copy), SYNCS (mbarrier), LDG/STG widths,
LDS/STS, REDUX / SHFL / VOTE / MATCH, ATOM / RED, BAR.  Writes profiles/r2/sass_opcodes.json + .txt.

    python scripts/sass_summary.py [path/to/libmbcol.so]"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "minibase-columnar-database_b200", "csrc", "libmbcol.so")
KEEP = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E", "LDS", "STS", "REDUX",
        "SHFL", "VOTE", "MATCH", "ATOMG", "ATOMS", "RED", "BAR", "NANOSLEEP", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "STTM", "CS2R", "POPC", "DADD", "BRA")


def summarise(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEEP:
                if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                    cur[k] += 1
                    break
    return {k: dict(v) for k, v in sorted(kernels.items())}


def main():
    s = summarise(LIB)
    d = os.path.join(ROOT, "profiles", "r2")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "sass_opcodes.json"), "w") as f:
        json.dump(s, f, indent=1, sort_keys=True)
    with open(os.path.join(d, "sass_opcodes.txt"), "w") as f:
        f.write("# cuobjdump -sass libmbcol.so (sm_100a), per kernel: total instructions and selected opcodes (scripts/sass_summary.py)\n")
        for k, c in s.items():
            ops = " ".join(f"{o}={n}" for o, n in sorted(c.items()) if o != "_total")
            f.write(f"{k}: total={c.get('_total', 0)} {ops}\n")
    for k in ("mbc::filter_kernel", "mbc::fused_scan_kernel", "void mbc::write_kernel<false>", "mbc::shard_push_kernel", "mbc::bitmap_build_kernel"):
        print(k, {o: n for o, n in s.get(k, {}).items() if o in ("_total", "UBLKCP", "SYNCS", "LDG.E.128", "STG.E.128", "REDUX")})


if __name__ == "__main__":
    main()
