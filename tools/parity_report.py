#!/usr/bin/env python
"""Regenerates profiles/<round>_parity.md from the JSON files the GPU tests write to gpurun_out/parity/ (development aid).

    python tools/parity_report.py [number of GPU tests in that run] [round, default r2]
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
P = "gpurun_out/parity"
ntests = sys.argv[1] if len(sys.argv) > 1 else "all"
rnd = sys.argv[2] if len(sys.argv) > 2 else "r2"


def e(x):
    return "-" if x is None else f"{x:.1e}"


out = [
    f"# GPU parity evidence, round {rnd[1:]} (B200, `pytest tests -m gpu`, JSON written by the tests to gpurun_out/parity/)",
    "",
    "Comparator: the fp64 CPU restatement `oracle/mmz_oracle.c` (physics parity-UNPINNED against MuJoCo, see DESIGN.md section 4); "
    "clamp and top-down-view goldens: the REAL reference Python code. Model convention of every run: `legacy_capsule_volume=True` "
    "(MuJoCo 2.0 capsule volume pi r^2 L + pi r^3, pinned by tests/test_abi_and_host.py), welded bodies merged, line-search tolerance 1e-2 (MuJoCo's default ls_tolerance).",
    "Teacher-forced single evaluations / steps from identical (qpos, qvel, t, action); relative errors are |gpu - oracle| / (1 + |oracle|). "
    f"{ntests} GPU tests, all green (last run of the round, final kernels; regenerate with `python tools/parity_report.py`).",
    "",
    "## One forward-dynamics evaluation (`mmz_forward`): qacc, contact and row counts",
    "",
    "| env id | kernel | envs | same contact+row counts | max rel err (same rows) | median | Newton iters gpu / oracle | contacts mean / max |",
    "|---|---|---|---|---|---|---|---|",
]
for f in sorted(glob.glob(f"{P}/forward_*.json")):
    d = json.load(open(f))
    k = d["kernel"]
    out.append(f"| {d['env']} | {k['kernel']} ({k['lanes_per_env']} lanes, {k['floats_per_env']} floats/env) | {d['n']} | {d['frac_same_rows']:.3f} | "
               f"{e(d['max_rel_err_same_rows'])} | {e(d['median_rel_err'])} | {d['gpu_iters_mean']:.2f} / {d['oracle_iters_mean']:.2f} | "
               f"{d['ncon_mean']:.1f} / {d['ncon_max']} |")
out += ["", "## One `MazeEnv.step` (`mmz_step`): state, observation, reward, done", "",
        "| env id | envs | qpos err max / p99 | qvel err max / p99 / median | reward err max | done mismatches | flips | unstable gpu / oracle |",
        "|---|---|---|---|---|---|---|---|"]
for f in sorted(glob.glob(f"{P}/step_*.json")):
    d = json.load(open(f))
    out.append(f"| {d['env']} | {d['n']} | {e(d['qpos_err_max'])} / {e(d['qpos_err_p99'])} | {e(d['qvel_err_max'])} / {e(d['qvel_err_p99'])} / "
               f"{e(d['qvel_err_median'])} | {e(d['reward_err_max'])} | {d['done_mismatch']} | {d.get('flips', '-')} | {d['unstable_gpu']} / {d['unstable_oracle']} |")
out += ["", "Large maxima with tiny p99 are environments whose active contact set flipped between fp32 and fp64 (a contact sitting at its margin); "
        "`flips` counts them (velocity error > 1e-3); the tests allow at most one per env id, require exact `done` bits and positions within 1e-2 for EVERY "
        "environment, and the tight tolerances for the rest.", ""]
views = sorted(glob.glob(f"{P}/view_*.json"))
if views:
    out += ["## Top-down view after one physics step (`maze_view_kernel` vs the oracle; the oracle's raster is pinned to the reference method to 1e-9)", "",
            "| case | envs | max abs error of the 75 view entries | median |", "|---|---|---|---|"]
    for f in views:
        d = json.load(open(f))
        out.append(f"| {d['case']} | {d['n']} | {e(d['view_err_max'])} | {e(d['view_err_median'])} |")
    out.append("")
if os.path.exists(f"{P}/rollout_drift.json"):
    out += ["## 10-step free-running drift (reported, not asserted)", "", "```", open(f"{P}/rollout_drift.json").read().strip(), "```", ""]
if os.path.exists(f"{P}/clamp_goldens.json"):
    out += ["## Wall clamp vs the real reference Python (`tests/golden/reference_python_half.json`)", "", "```",
            open(f"{P}/clamp_goldens.json").read().strip(), "```", ""]
if os.path.exists(f"{P}/coupled_AntPush-v0.json"):
    d = json.load(open(f"{P}/coupled_AntPush-v0.json"))
    out += ["## AntPush-v0 with the ants pressed against the movable block (`test_forward_parity_ant_against_block`)", "",
            f"{d['n']} environments, the torso 0.35 - 0.9 in front of the block's face: {100 * d['touching']:.1f} % of the ants touch the block (more "
            f"contacts than the same pose 3 units away), {d['ncon_mean']:.1f} contacts per environment on average (max {d['ncon_max']}). Ant-block "
            "contacts move dofs of both dof trees, so solver v3 leaves the side-by-side elimination of the two trees for the dense one while the "
            "block's own contacts still run through the lane = contact passes: same contact and row counts as the fp64 restatement in "
            f"{100 * d['frac_same_rows']:.0f} % of the environments, qacc max relative error {e(d['max_rel_err_same_rows'])}, median {e(d['median_rel_err'])}.", ""]
open(f"profiles/{rnd}_parity.md", "w").write("\n".join(out))
print(f"wrote profiles/{rnd}_parity.md")
