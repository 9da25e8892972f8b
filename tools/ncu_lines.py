#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per source line and per routine.

usage: ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python tools/ncu_lines.py src.csv
Routines are found by scanning the .cuh sources for `MMZ_DI|__device__ ... name(` definitions.
"""
import csv
import re
import sys
from collections import defaultdict


def routines(path):
    out = []
    try:
        lines = open(path).read().split("\n")
    except OSError:
        return out
    pat = re.compile(r"^\s*(?:template.*>\s*)?(?:MMZ_DI|__device__|__global__|static\s+MMZ_DI)[^;=]*?\b([A-Za-z_0-9]+)\s*\(")
    for i, l in enumerate(lines, 1):
        m = pat.match(l)
        if m and not l.strip().endswith(";"):
            out.append((i, m.group(1)))
    return out


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur_file = None
    per_line = defaultdict(lambda: [0, 0, 0])  # samples, inst, thread inst
    hdr = None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1]
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            hdr = row
            i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or row[0] == "" or not row[0].isdigit():
            continue
        try:
            per_line[(cur_file, int(row[0]))][0] += int(row[i_s])
            per_line[(cur_file, int(row[0]))][1] += int(row[i_i])
            per_line[(cur_file, int(row[0]))][2] += int(row[i_t])
        except ValueError:
            pass
    tot = [sum(v[k] for v in per_line.values()) for k in range(3)]
    print(f"total samples {tot[0]}  warp-inst {tot[1]:.4g}  thread-inst {tot[2]:.4g}  avg threads/inst {tot[2] / max(1, tot[1]):.1f}")
    files = sorted({f for f, _ in per_line})
    per_fn = defaultdict(lambda: [0, 0, 0])
    for f in files:
        rs = routines(f)
        for (ff, ln), v in per_line.items():
            if ff != f:
                continue
            name = "?"
            for start, nm in rs:
                if start <= ln:
                    name = nm
            key = f.split("/")[-1] + ":" + name
            for k in range(3):
                per_fn[key][k] += v[k]
    print("\nper routine (share of samples | share of warp instructions | threads/inst)")
    for key, v in sorted(per_fn.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {key:48s} {100 * v[0] / tot[0]:6.2f}%  {100 * v[1] / tot[1]:6.2f}%  {v[2] / max(1, v[1]):5.1f}")
    print("\ntop lines by samples")
    for (f, ln), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {f.split('/')[-1]}:{ln:<5d} {100 * v[0] / tot[0]:6.2f}%  {100 * v[1] / tot[1]:6.2f}%  {v[2] / max(1, v[1]):5.1f}")


if __name__ == "__main__":
    main()
