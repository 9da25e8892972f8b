#!/usr/bin/env python
"""Cross-check of the fp64 restatement (oracle/) against REAL MuJoCo, for whoever has MuJoCo.

The physics half of the oracle is "parity unpinned" (DESIGN.md section 4): MuJoCo 2.0 / mujoco-py / gym cannot be
installed in the build image. On a box where they import, this tool

  1. runs the UNMODIFIED reference (baseline/_ref, through tools/reference_loop.py --dump: gym.make, reset, step) on
     the BASELINE.json configurations and records the reset state, the sampled actions and qpos / qvel / number of
     contacts / obs / reward / done after every step under tests/golden/mujoco_crosscheck/<env id>.json;
  2. replays every episode in the oracle from the recorded reset state with the recorded actions and prints the
     errors (the same comparison runs as tests/test_mujoco_crosscheck.py whenever such a file is committed).

    python tools/mujoco_crosscheck.py [--n 8] [--steps 25] [--seed 0] [--compare-only]

Without MuJoCo it prints {"unavailable": "<why>"} and exits 0.
"""
import argparse
import glob
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "tests", "golden", "mujoco_crosscheck")
# the env ids of BASELINE.json's configs (configs[3] differs from configs[2] only by the maze)
ENV_IDS = ("PointUMaze-v0", "AntUMaze-v0", "Ant4Rooms-v0", "AntPush-v0")


def probe():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_loop.py"), "--probe"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        return json.loads(r.stdout.strip().splitlines()[-1])
    except (IndexError, ValueError):
        return {"available": False, "why": f"probe failed: {r.stderr.strip()[-300:]}"}


def compare(dump_path):
    """Replay a dump in the oracle. Returns per-quantity maximum errors (relative to 1 + |reference value|)."""
    for p in (os.path.join(ROOT, "mujoco-maze_b200"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import mujoco_maze  # noqa: F401  (this repository's package)
    from mujoco_maze import gym
    from oracle import mmz_oracle

    with open(dump_path) as f:
        rec = json.load(f)
    model = gym.make(rec["env_id"]).unwrapped.model
    nq, nv = int(model.nq), int(model.nv)
    o = mmz_oracle.OracleEnv(model)
    stats = {"env_id": rec["env_id"], "episodes": len(rec["episodes"]), "steps": 0, "qpos": 0.0, "qvel": 0.0, "obs": 0.0,
             "reward": 0.0, "done_mismatch": 0, "ncon_mismatch": 0, "first_step_qvel": 0.0}
    for ep in rec["episodes"]:
        q0, v0 = np.asarray(ep["qpos0"], float), np.asarray(ep["qvel0"], float)
        if q0.size != nq or v0.size != nv:
            raise ValueError(f"{rec['env_id']}: the reference has nq, nv = {q0.size}, {v0.size}, the compiled model {nq}, {nv}")
        o.set_state(q0, v0, 0)
        for k, st in enumerate(ep["steps"]):
            obs, rew, bits, _ = o.step(np.asarray(st["action"], float))
            q, v, _ = o.get_state()
            rq, rv, robs = (np.asarray(st[key], float) for key in ("qpos", "qvel", "obs"))
            eq = float((np.abs(q - rq) / (1 + np.abs(rq))).max())
            ev = float((np.abs(v - rv) / (1 + np.abs(rv))).max())
            stats["qpos"], stats["qvel"] = max(stats["qpos"], eq), max(stats["qvel"], ev)
            if k == 0:
                stats["first_step_qvel"] = max(stats["first_step_qvel"], ev)
            stats["obs"] = max(stats["obs"], float((np.abs(obs - robs) / (1 + np.abs(robs))).max()))
            stats["reward"] = max(stats["reward"], abs(float(rew) - st["reward"]))
            stats["done_mismatch"] += int(bool(bits & 1) != bool(st["done"]))
            stats["ncon_mismatch"] += int(o.counts()["ncon"] != st["ncon"])
            stats["steps"] += 1
            o.set_state(rq, rv, k + 1)  # teacher forcing: every step starts from the reference's state
    return stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--compare-only", action="store_true", help="only replay the dumps already under tests/golden/mujoco_crosscheck")
    a = ap.parse_args()
    if not a.compare_only:
        p = probe()
        if not p.get("available"):
            print(json.dumps({"unavailable": p.get("why", "?"), "ref_path": p.get("ref_path")}))
            return 0
        os.makedirs(OUT_DIR, exist_ok=True)
        for env_id in ENV_IDS:
            out = os.path.join(OUT_DIR, f"{env_id}.json")
            subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "reference_loop.py"), "--dump", env_id,
                                   "--n", str(a.n), "--steps", str(a.steps), "--seed", str(a.seed), "--out", out])
    dumps = sorted(glob.glob(os.path.join(OUT_DIR, "*.json")))
    if not dumps:
        print(json.dumps({"unavailable": "no dumps under tests/golden/mujoco_crosscheck"}))
        return 0
    for d in dumps:
        print(json.dumps(compare(d)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
