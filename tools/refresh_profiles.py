#!/usr/bin/env python
"""Regenerates profiles/r1_bench.md (head table + JSON lines), profiles/r1_ncu_step_hybrid.md, the launch list, traffic.json and
the DESIGN.md table from a capture run:  python tools/refresh_profiles.py r1r   (files gpurun_out/<run>_bench*.log, <run>_prof.ncu-rep,
<run>_launches.csv). Development aid: keeps the committed evidence in step with the committed kernels."""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
run = sys.argv[1]


def last(f):
    return json.loads(open(f).read().strip().split("\n")[-1])


rows = [("AntUMaze-v0, 65536 envs (BASELINE configs[2], north-star)", f"gpurun_out/{run}_bench.log"),
        ("Ant4Rooms-v0, 65536 envs / GPU (configs[3] per-GPU shard)", f"gpurun_out/{run}_bench_Ant4Rooms-v0_65536.log"),
        ("AntPush-v0, 32768 envs (configs[4])", f"gpurun_out/{run}_bench_AntPush-v0_32768.log"),
        ("PointUMaze-v0, 4096 envs (configs[1])", f"gpurun_out/{run}_bench_PointUMaze-v0_4096.log"),
        ("SwimmerUMaze-v0, 65536 envs", f"gpurun_out/{run}_bench_SwimmerUMaze-v0_65536.log"),
        ("reference arm (--impl reference): fp64 CPU port, 16 host threads", f"gpurun_out/{run}_bench_ref.log")]
L = {n: last(f) for n, f in rows}
b = open("profiles/r1_bench.md").read()
head = [f"# Round 1 final bench lines (1 x B200, `python bench.py [--workload ...]`, run {run} = the committed kernels, defaults: 40 timed steps after 10 warm-up steps)\n",
        "| workload | kernel | ms/step | env steps/s (device-resident) | e2e env steps/s (host buffers) | HBM achieved GB/s | frac of measured 6457.4 |",
        "|---|---|---|---|---|---|---|"]
for name, _ in rows:
    d = L[name]
    k = d["config"].get("kernel") or {}
    r = d.get("roofline") or {}
    frac = r.get("frac")
    head.append(f"| {name} | {k.get('kernel', '-')} ({k.get('envs_per_sm', '-')} envs/SM, {k.get('floats_per_env', '-')} floats/env) | {d['ms_per_step']:.3f} | "
                f"{d['value']:.4g} | {d['e2e']['value']:.4g} | {r.get('achieved') and round(r['achieved'], 3)} | {frac and f'{frac:.2e}'} |")
d = L[rows[0][0]]
head.append(f"\nCPU baseline in the same run (rank 0): `{json.dumps(d['cpu_baseline'])}`\n")
head.append(f"Clocks during the timed region: `{json.dumps(d['clocks'])}`\n")
i = b.index("2 x B200 (`gpurun --gpus 2`")
b = "\n".join(head) + "\n" + b[i:]
j = b.index("## Full JSON lines")
b = b[:j] + "## Full JSON lines\n\n" + "\n".join("```\n" + json.dumps(L[n]) + "\n```" for n, _ in rows) + "\n"
open("profiles/r1_bench.md", "w").write(b)

summ = subprocess.run([sys.executable, "tools/ncu_summary.py", f"gpurun_out/{run}_prof.ncu-rep", "22"], capture_output=True, text=True).stdout
h = open("profiles/r1_ncu_step_hybrid.md").read()
h = h[: h.index("```\n")] + "```\n" + summ + "```\n"
h = re.sub(r"r1[a-z]_prof", f"{run}_prof", h)
open("profiles/r1_ncu_step_hybrid.md", "w").write(h)
shutil.copy(f"gpurun_out/{run}_launches.csv", "profiles/r1_launches_hybrid_AntUMaze65536.csv")
raw = subprocess.run(["ncu", "-i", f"gpurun_out/{run}_prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hd, un, da = r[0], r[1], r[2]


def val(k):
    i = hd.index(k)
    return float(da[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un[i]]


tr = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
json.dump({"AntUMaze-v0:65536": tr,
           "_source": f"ncu --set full, profiles/r1_ncu_step_hybrid.md (dram__bytes_read.sum + dram__bytes_write.sum of one maze_hkernel<14,0> launch, run {run})"},
          open("profiles/traffic.json", "w"), indent=1)
D = open("DESIGN.md").read()
a = D.index("| Workload (1 × B200")
e = D.index("| 2 × B200, AntUMaze-v0")
A, R4, P, Pt, Sw, Rf = [L[n] for n, _ in rows]


def line(title, x, bold=False):
    v = f"**{x['value']:.3g}**" if bold else f"{x['value']:.3g}"
    return (f"| {title} | {x['ms_per_step']:.3g} | {v} | {x['e2e']['value']:.3g} | {x['roofline']['achieved']:.2f} GB/s | {x['roofline']['frac']:.1e} |\n")


tbl = (f"| Workload (1 × B200, run {run}, `profiles/r1_bench.md`) | ms / step | env steps/s | e2e (host buffers) | HBM achieved | frac of 6457 GB/s (measured) |\n"
       "|---|---|---|---|---|---|\n"
       + line("AntUMaze-v0, 65 536 envs (north-star config), hybrid kernel", A, True)
       + line("Ant4Rooms-v0, 65 536 envs per GPU, hybrid kernel", R4)
       + line("AntPush-v0, 32 768 envs, hybrid kernel with box geoms", P)
       + line("PointUMaze-v0, 4096 envs, 8 lanes per environment", Pt)
       + line("SwimmerUMaze-v0, 65 536 envs, 8 lanes per environment", Sw))
D = D[:a] + tbl + D[e:]
D = re.sub(r"\| CPU restatement, 16 host threads \(same box; `--impl reference`\) \| — \| [0-9.e+]+ \|", f"| CPU restatement, 16 host threads (same box; `--impl reference`) | — | {Rf['value']:.3g} |", D)
open("DESIGN.md", "w").write(D)
print("traffic", tr, "Ant", A["value"], A["ms_per_step"])
