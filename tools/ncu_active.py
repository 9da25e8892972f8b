#!/usr/bin/env python
"""Where the WORKING warps spend their time: stall samples without the `barrier` reason, per routine and per line, from
an .ncu-rep captured with --import-source on (a kernel that is bound by its critical path shows mostly barrier samples
on the waiting warps; what is left is the critical path).  usage: python tools/ncu_active.py prof.ncu-rep [top]"""
import bisect
import csv
import os
import subprocess
import sys
from collections import defaultdict

from ncu_lines import routines


def num(x):
    try:
        return int(x)
    except (TypeError, ValueError):
        return 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode("utf-8", "replace")
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mujoco-maze_b200", "csrc")
    rts = {}
    hdr, cur = None, None
    per = defaultdict(lambda: [0, 0, ""])
    tot_all = 0
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr, r))
        s = num(d.get("# Samples"))
        b = sum(num(v) for k, v in d.items() if k.startswith("stall_barrier") and "Not Issued" not in k)
        tot_all += s
        per[(cur, ln)][0] += s - b
        per[(cur, ln)][1] += num(d.get("Instructions Executed"))
        per[(cur, ln)][2] = r[1]
    tot = sum(v[0] for v in per.values()) or 1
    print(f"samples {tot_all}, without barrier waits {tot} ({100 * tot / max(tot_all, 1):.1f} %)")
    byr = defaultdict(lambda: [0, 0])
    for (f, ln), (s, i, _) in per.items():
        if f not in rts:
            rts[f] = routines(os.path.join(here, f))
        names = rts[f]
        k = bisect.bisect_right([a for a, _ in names], ln) - 1
        name = names[k][1] if k >= 0 else "?"
        byr[(f, name)][0] += s
        byr[(f, name)][1] += i
    toti = sum(v[1] for v in byr.values()) or 1
    print("\nroutines by active samples")
    for (f, name), (s, i) in sorted(byr.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / tot:6.2f}% active  {100 * i / toti:6.2f}% inst  {f}:{name}")
    print("\nlines by active samples")
    for (f, ln), (s, i, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / tot:6.2f}% active  {100 * i / toti:6.2f}% inst  {f}:{ln}  {src.strip()[:100]}")


if __name__ == "__main__":
    main()
