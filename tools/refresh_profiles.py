#!/usr/bin/env python
"""Regenerates profiles/<round>_bench.md, <round>_ncu_step_hybrid.md, the launch list and kernel_counters.json from a
capture run:  python tools/refresh_profiles.py r2e [r2]   (files gpurun_out/<run>_bench*.log, <run>_prof*.ncu-rep,
<run>_launches.csv written by tools/capture_run.sh). Development aid: keeps the committed evidence in step with the
committed kernels."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
run = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r2"


def last(f):
    return json.loads(open(f).read().strip().split("\n")[-1])


rows = [("AntUMaze-v0, 65536 envs (BASELINE configs[2], north star), the driver's command `--gpus 1 --steps 20 --warmup 5`", f"gpurun_out/{run}_bench.log"),
        ("AntUMaze-v0, 65536 envs, K = 1000 (`--steps 1000`: a whole TimeLimit episode, every environment truncates and restarts inside the timed loop)", f"gpurun_out/{run}_bench_k1000.log"),
        ("Ant4Rooms-v0, 65536 envs / GPU (configs[3] per-GPU shard)", f"gpurun_out/{run}_bench_Ant4Rooms-v0_65536.log"),
        ("AntPush-v0, 32768 envs (configs[4])", f"gpurun_out/{run}_bench_AntPush-v0_32768.log"),
        ("PointUMaze-v0, 4096 envs (configs[1])", f"gpurun_out/{run}_bench_PointUMaze-v0_4096.log"),
        ("PointPush-v0, 65536 envs", f"gpurun_out/{run}_bench_PointPush-v0_65536.log"),
        ("SwimmerUMaze-v0, 65536 envs", f"gpurun_out/{run}_bench_SwimmerUMaze-v0_65536.log"),
        ("reference arm (`--impl reference`): fp64 CPU port (NOT mujoco-py), all host threads", f"gpurun_out/{run}_bench_ref.log")]
L = {n: last(f) for n, f in rows if os.path.exists(f)}
out = [f"# Round {rnd[1:]} bench lines (1 x B200, run {run} = the committed kernels; 20 timed steps after 5 warm-up steps unless stated)\n",
       "| workload | kernel | ms/step | env steps/s (device-resident) | e2e env steps/s (host buffers) | issue slots used | fp32 of peak | HBM frac (533 B/env-step) |",
       "|---|---|---|---|---|---|---|---|"]
for name, _ in rows:
    if name not in L:
        continue
    d = L[name]
    k = (d.get("details") or {}).get("kernel") or {}
    r = d.get("roofline") or {}
    iss, fp = (r.get("issue") or {}).get("frac"), (r.get("fp32") or {}).get("frac")
    out.append(f"| {name} | {k.get('kernel', '-')} ({k.get('envs_per_sm', '-')} envs/block, {k.get('floats_per_env', '-')} floats/env) | {d['ms_per_step']:.3f} | "
               f"{d['value']:.4g} | {d['e2e']['value']:.4g} | {iss and f'{iss:.3f}'} | {fp and f'{fp:.3f}'} | {r.get('frac') and format(r['frac'], '.2e')} |")
d = L[rows[0][0]]
out.append(f"\nCPU baseline in the same run (rank 0): `{json.dumps(d.get('cpu_baseline'))}`\n")
out.append(f"Clocks during the timed region: `{json.dumps(d['clocks'])}`\n")
extra = f"profiles/{rnd}_bench_notes.md"
if os.path.exists(extra):
    out.append(open(extra).read())
out.append("## Full JSON lines\n")
out += ["```\n" + json.dumps(L[n]) + "\n```" for n, _ in rows if n in L]
open(f"profiles/{rnd}_bench.md", "w").write("\n".join(out) + "\n")

for rep, key in ((f"gpurun_out/{run}_prof.ncu-rep", "AntUMaze-v0:65536"), (f"gpurun_out/{run}_prof_point.ncu-rep", "PointUMaze-v0:4096")):
    if os.path.exists(rep):
        subprocess.run([sys.executable, "tools/ncu_counters.py", rep, key, f"run {run}"], stdout=subprocess.DEVNULL, check=True)
summ = subprocess.run([sys.executable, "tools/ncu_summary.py", f"gpurun_out/{run}_prof.ncu-rep", "22"], capture_output=True, text=True).stdout
lines = subprocess.run([sys.executable, "tools/ncu_perline.py", f"gpurun_out/{run}_prof.ncu-rep", "40"], capture_output=True, text=True).stdout
head = f"profiles/{rnd}_ncu_step_hybrid_head.md"
text = (open(head).read() if os.path.exists(head) else f"# ncu summary of the step kernel, round {rnd[1:]} (run {run})\n\n")
text = text.replace("{run}", run)
open(f"profiles/{rnd}_ncu_step_hybrid.md", "w").write(text + "```\n" + summ + "```\n\nPer source line (tools/ncu_perline.py):\n\n```\n" + lines + "```\n")
shutil.copy(f"gpurun_out/{run}_launches.csv", f"profiles/{rnd}_launches_hybrid_AntUMaze65536.csv")
print("ok")
