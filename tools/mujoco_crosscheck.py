#!/usr/bin/env python
"""Cross-check of the fp64 restatement (oracle/) against REAL MuJoCo, for whoever has MuJoCo.

The physics half of the oracle is "parity unpinned" (DESIGN.md section 4): MuJoCo 2.0 / mujoco-py / gym cannot be
installed in the build image. On a box where they import, this tool

  1. runs the UNMODIFIED reference (baseline/_ref, through tools/reference_loop.py --dump: gym.make, reset, step) on
     the BASELINE.json configurations and records the reset state, the sampled actions and qpos / qvel / number of
     contacts / obs / reward / done after every step under tests/golden/mujoco_crosscheck/<env id>.json;
  2. runs tests/test_reference_probe.py::test_oracle_against_committed_mujoco_dumps, which replays every episode in the
     oracle from the recorded reset state with the recorded actions (tests/mujoco_compare.py) and asserts the errors.

    python tools/mujoco_crosscheck.py [--n 8] [--steps 25] [--seed 0] [--compare-only]

Without MuJoCo it prints {"unavailable": "<why>"} and exits 0.
"""
import argparse
import glob
import json
import os
import subprocess
import sys

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


def compare_all():
    """The comparison itself lives with the tests (tests/mujoco_compare.py: only tests may use the oracle)."""
    return subprocess.call([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_reference_probe.py"), "-q", "-s",
                            "-k", "committed_mujoco_dumps"])


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
    return compare_all()


if __name__ == "__main__":
    sys.exit(main())
