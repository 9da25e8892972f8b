"""Replay of a MuJoCo dump (tools/reference_loop.py --dump, written by tools/mujoco_crosscheck.py on a box that has MuJoCo)
in the fp64 oracle: test infrastructure (the oracle may only be used from tests/, smoke() and bench.py's CPU legs)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
