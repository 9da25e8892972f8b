#!/usr/bin/env python
"""The reference's OWN loop, timed or dumped - for boxes where MuJoCo is importable (it is not in the build image).

This script never imports this repository's package: it puts the offline install of the UNMODIFIED reference
(`baseline/_ref`) first on sys.path and talks to it through `gym.make`, exactly like the reference's
tests do (`/root/reference/tests/test_envs.py:7-17`: make, reset, step(action_space.sample())).

    python tools/reference_loop.py --probe
        -> one JSON line {"available": bool, "why": "...", "gym": ver, "mujoco": "mujoco_py x.y" | "mujoco x.y"}
    python tools/reference_loop.py --time ENV_ID --steps K --warmup W [--procs P]
        -> {"env_steps_per_sec": aggregate, "per_proc": [...], "procs": P, "steps": K}
    python tools/reference_loop.py --dump ENV_ID --n N --steps K --seed S --out FILE.json
        -> per environment: the state the reference reset to, the sampled actions, qpos / qvel / ncon / obs / reward /
           done after every step (consumed by tools/mujoco_crosscheck.py and tests/test_mujoco_crosscheck.py)

MMZ_REF_PATH overrides the location of the reference install (the CPU tests point it at a fake).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_path():
    return os.environ.get("MMZ_REF_PATH", os.path.join(ROOT, "baseline", "_ref"))


def _isolate():
    """sys.path for the reference: its install first, and nothing of this repository (same package name)."""
    mine = {os.path.join(ROOT, "mujoco-maze_b200"), ROOT, os.path.join(ROOT, "tools")}
    sys.path[:] = [ref_path()] + [p for p in sys.path if os.path.abspath(p or ".") not in mine]


def probe():
    _isolate()
    out = {"available": False, "why": "", "ref_path": ref_path()}
    if not os.path.isdir(os.path.join(ref_path(), "mujoco_maze")):
        out["why"] = f"no reference install at {ref_path()} (pip install --no-deps --target baseline/_ref /root/reference)"
        return out
    try:
        import gym  # noqa: F401

        out["gym"] = getattr(gym, "__version__", "?")
    except Exception as e:  # noqa: BLE001
        out["why"] = f"import gym failed: {type(e).__name__}: {e}"
        return out
    sim = None
    for name in ("mujoco_py", "mujoco"):
        try:
            mod = __import__(name)
            sim = f"{name} {getattr(mod, '__version__', '?')}"
            break
        except Exception as e:  # noqa: BLE001
            out["why"] = f"import {name} failed: {type(e).__name__}: {e}"
    if sim is None:
        return out
    out["mujoco"] = sim
    try:
        import mujoco_maze  # noqa: F401  (the reference's: registers its env ids with gym)

        out["reference"] = os.path.dirname(mujoco_maze.__file__)
        env = gym.make("PointUMaze-v0")
        env.reset()
        env.step(env.action_space.sample())
    except Exception as e:  # noqa: BLE001
        out["why"] = f"the reference does not run: {type(e).__name__}: {e}"
        return out
    out["available"], out["why"] = True, "gym + MuJoCo + the reference import and step"
    return out


def _worker(args):
    env_id, steps, warmup, seed = args
    _isolate()
    import gym
    import mujoco_maze  # noqa: F401

    env = gym.make(env_id)
    try:
        env.seed(seed)
        env.action_space.seed(seed)
    except Exception:  # noqa: BLE001
        pass
    env.reset()
    for _ in range(warmup):
        _, _, done, _ = env.step(env.action_space.sample())
        if done:
            env.reset()
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, done, _ = env.step(env.action_space.sample())
        if done:
            env.reset()
    return steps / (time.perf_counter() - t0)


def time_loop(env_id, steps, warmup, procs):
    with mp.get_context("spawn").Pool(procs) as pool:
        rates = pool.map(_worker, [(env_id, steps, warmup, s) for s in range(procs)])
    return {"env_steps_per_sec": float(sum(rates)), "per_proc": [float(r) for r in rates], "procs": procs, "steps": steps,
            "warmup": warmup, "env_id": env_id}


def dump(env_id, n, steps, seed, out_path):
    """States and outputs of the reference for n independent episodes: everything a CPU restatement needs to replay them."""
    _isolate()
    import gym
    import numpy as np

    import mujoco_maze  # noqa: F401

    rng = np.random.default_rng(seed)
    episodes = []
    for i in range(n):
        env = gym.make(env_id)
        try:
            env.seed(seed + i)
        except Exception:  # noqa: BLE001
            pass
        r = env.reset()
        obs0 = r[0] if isinstance(r, tuple) else r  # reference quirk Q2: reset() returns (obs, {})
        agent = env.unwrapped.wrapped_env
        data = agent.sim.data if hasattr(agent, "sim") else agent.data
        lo, hi = env.action_space.low, env.action_space.high
        ep = {"qpos0": np.array(data.qpos).ravel().tolist(), "qvel0": np.array(data.qvel).ravel().tolist(),
              "obs0": np.asarray(obs0).tolist(), "steps": []}
        for _ in range(steps):
            a = rng.uniform(lo, hi)
            obs, rew, done, info = env.step(a)
            ep["steps"].append({"action": a.tolist(), "qpos": np.array(data.qpos).ravel().tolist(),
                                "qvel": np.array(data.qvel).ravel().tolist(), "ncon": int(data.ncon),
                                "obs": np.asarray(obs).tolist(), "reward": float(rew), "done": bool(done)})
            if done:
                break
        episodes.append(ep)
    rec = {"env_id": env_id, "n": n, "steps": steps, "seed": seed, "probe": probe(), "episodes": episodes}
    with open(out_path, "w") as f:
        json.dump(rec, f)
    return {"written": out_path, "episodes": len(episodes)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe", action="store_true")
    ap.add_argument("--time", metavar="ENV_ID")
    ap.add_argument("--dump", metavar="ENV_ID")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="mujoco_dump.json")
    a = ap.parse_args()
    if a.time:
        procs = a.procs or len(os.sched_getaffinity(0))
        print(json.dumps(time_loop(a.time, a.steps, a.warmup, procs)))
    elif a.dump:
        print(json.dumps(dump(a.dump, a.n, a.steps, a.seed, a.out)))
    else:
        print(json.dumps(probe()))


if __name__ == "__main__":
    main()
