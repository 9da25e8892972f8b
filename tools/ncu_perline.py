#!/usr/bin/env python
"""Per-line / per-routine shares of warp instructions and stall samples from an .ncu-rep captured with --import-source on.

usage: python tools/ncu_perline.py prof.ncu-rep [top]
(exports `ncu -i ... --page source --print-source cuda,sass --csv` itself; the source text comes from the report, so
the line numbers are those of the profiled build)."""
import csv
import subprocess
import sys
from collections import defaultdict


def num(x):
    try:
        return int(x)
    except (TypeError, ValueError):
        return 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode("utf-8", "replace")
    per = {}
    stall_cols = None
    stalls = defaultdict(lambda: defaultdict(int))
    hdr, cur = None, None
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr, r))
        k = (cur, ln)
        if k not in per:
            per[k] = [r[1], 0, 0]
        per[k][1] += num(d.get("# Samples"))
        per[k][2] += num(d.get("Instructions Executed"))
        for i in stall_cols:
            stalls[k][hdr[i]] += num(r[i])
    tot_s = sum(v[1] for v in per.values()) or 1
    tot_i = sum(v[2] for v in per.values()) or 1
    print(f"samples {tot_s}  warp instructions {tot_i:.4g}")
    print("\ntop lines by warp instructions")
    for (f, ln), (src, s, i) in sorted(per.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"{f}:{ln:5d} inst {100 * i / tot_i:5.2f}% samp {100 * s / tot_s:5.2f}%  {src.strip()[:110]}")
    print("\ntop lines by stall samples")
    for (f, ln), (src, s, i) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top // 2]:
        st = stalls[(f, ln)]
        main_st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{f}:{ln:5d} samp {100 * s / tot_s:5.2f}% inst {100 * i / tot_i:5.2f}%  [{main_st}]  {src.strip()[:80]}")


if __name__ == "__main__":
    main()
