#!/usr/bin/env python
"""Extract known-answer fixtures from the reference's shipped session transcript.

Reads (never at test time -- only when this script is run in the authoring container):
    /root/reference/phase3_output   a 35 709-line typescript of a real ColumnarMain session
    /root/reference/minidata.txt    the 500-row table every command in it was run on
Writes:
    tests/golden/minidata.tsv       the data fixture (header `name:type`, tab separated)
    tests/golden/phase3_golden.json one entry per index / indexes_query / bmj / nlj / sort command:
        the command line, the result count the reference printed, the printed rows (or their
        sha256 when there are more than 400), the side-filter bitsets `bmj` prints, and the
        per-value `BitSet.toByteArray().length` list `index ... bitmap` prints.
"""
import hashlib
import json
import os
import re
import shutil

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    shutil.copyfile(os.path.join(REF, "minidata.txt"), os.path.join(OUT, "minidata.tsv"))
    lines = open(os.path.join(REF, "phase3_output"), errors="replace").read().split("\n")
    cmd_idx = [i for i, ln in enumerate(lines) if ln.startswith("> ")]
    cmd_idx.append(len(lines))
    entries = []
    for a, b in zip(cmd_idx[:-1], cmd_idx[1:]):
        cmd = lines[a][2:].strip()
        body = [ln.rstrip("\r") for ln in lines[a + 1:b]]
        kind = cmd.split(" ")[0] if cmd else ""
        e = {"line": a + 1, "cmd": cmd, "kind": kind}
        if kind in ("batchinsert", "index"):
            # "Wrote Pages: {0=2, 1=2, 129=1, ...}": page id -> times written, as the buffer manager flushed them
            for ln in body:
                if ln.startswith("Wrote Pages: {"):
                    e["wrote_pages"] = {k: int(v) for k, v in re.findall(r"(\d+)=(\d+)", ln)}
                    break
        if kind == "batchinsert":
            m = [re.match(r"Record count: (\d+)", ln) for ln in body]
            m = [x for x in m if x]
            if m:
                e["record_count"] = int(m[0].group(1))
        elif kind == "index" and cmd.endswith("bitmap") or (kind == "index" and "bitmap" in cmd):
            e["bitmap_bytes"] = [int(x) for ln in body for x in re.findall(r"Size while writing:(\d+)Pages", ln)]
        elif kind in ("bmj", "nlj", "indexes_query"):
            cnt = [re.match(r"Total Results Count By Query: (\d+)", ln) for ln in body]
            cnt = [x for x in cnt if x]
            if not cnt:
                e["failed"] = True          # e.g. BufferPoolExceededException runs
            else:
                e["count"] = int(cnt[0].group(1))
                # rows sit between the header line (column names) and the first blank line before the stars
                star = next(i for i, ln in enumerate(body) if ln.startswith("Total Results Count")) - 1
                hdr = None
                for i, ln in enumerate(body[:star]):
                    if re.fullmatch(r"[A-Za-z0-9_.]+(, [A-Za-z0-9_.]+)*", ln) and not ln.startswith("Replacer"):
                        hdr = i
                        break
                # nlj interleaves "Next Pass Over Inner Table" banners with the rows
                rows = [ln for ln in body[hdr + 1:star]
                        if ln.strip() and not ln.startswith("****") and not ln.startswith("Next Pass Over")
                        and not re.match(r"(Tuple Size|Number of Tuples Buffer|Total Outer Tuples)", ln)] if hdr is not None else []
                e["header"] = body[hdr] if hdr is not None else None
                if len(rows) != e["count"]:
                    e["row_parse_mismatch"] = len(rows)
                text = "\n".join(rows)
                e["rows_sha256"] = hashlib.sha256(text.encode()).hexdigest()
                e["rows_sorted_sha256"] = hashlib.sha256("\n".join(sorted(rows)).encode()).hexdigest()
                if len(rows) <= 400:
                    e["rows"] = rows
                else:
                    e["rows_head"] = rows[:5]
            if kind == "bmj":
                for i, ln in enumerate(body):
                    if ln.startswith("OuterConstraint Bitset"):
                        e["outer_bitset"] = [int(x) for x in re.findall(r"\d+", body[i + 1])]
                    if ln.startswith("InnerConstraint Bitset"):
                        e["inner_bitset"] = [int(x) for x in re.findall(r"\d+", body[i + 1])]
        elif kind == "sort":
            # `sort DB CF [sort columns] [projected columns] ASC|DSC NUMBUF SORTBUF`: after "SORTED COLUMNS" one line per row,
            # "<projected values separated by blanks> :<position>", then the row count
            if "SORTED COLUMNS" not in body:
                e["failed"] = True
            else:
                start = body.index("SORTED COLUMNS") + 1
                rows = []
                for ln in body[start:]:
                    if re.fullmatch(r"\d+", ln.strip()):
                        e["count"] = int(ln.strip())
                        break
                    if ln.strip():
                        rows.append(ln.rstrip())
                e["rows"] = rows
        else:
            continue
        rp = [re.match(r"Read Page Count: (\d+)", ln) for ln in body]
        rp = [x for x in rp if x]
        if rp:
            e["read_pages"] = int(rp[0].group(1))
        entries.append(e)
    with open(os.path.join(OUT, "phase3_golden.json"), "w") as f:
        json.dump({"source": "phase3_output of Neehaarika/MiniBase-Columnar-Database", "entries": entries}, f, indent=0)
    kinds = {}
    for e in entries:
        kinds[e["kind"]] = kinds.get(e["kind"], 0) + 1
    print("wrote", len(entries), "entries", kinds)


if __name__ == "__main__":
    main()
