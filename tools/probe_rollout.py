#!/usr/bin/env python
"""Per-step timing and solver diagnostics of a free-running rollout (development aid, GPU only).

    python tools/probe_rollout.py [ENV_ID] [N_ENVS] [STEPS]
Writes gpurun_out/probe_<ENV_ID>.json.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mujoco-maze_b200"), ROOT):
    sys.path.insert(0, p)
import mujoco_maze  # noqa: E402,F401
from mujoco_maze import gym  # noqa: E402
from mujoco_maze.backend import BatchedSim  # noqa: E402


def main():
    env_id = sys.argv[1] if len(sys.argv) > 1 else "AntUMaze-v0"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    model = gym.make(env_id, num_envs=1).unwrapped.model
    sim = BatchedSim(model, n, auto_reset=True)
    diag = sim.enable_step_diag()
    r = np.asarray(model.meta["act_ctrlrange"], float)
    lo, hi = (torch.tensor(x, device="cuda", dtype=torch.float32) for x in (r[:, 0], r[:, 1]))
    sim.reset(seed=0)
    rows = []
    for s in range(steps):
        a = lo + (hi - lo) * torch.rand((n, sim.nu), device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        obs, rew, done, info = sim.step(a)
        e1.record()
        torch.cuda.synchronize()
        d = diag.cpu().numpy()
        rows.append(dict(step=s, ms=e0.elapsed_time(e1), newton_mean=float(d[:, 0].mean()), newton_max=int(d[:, 0].max()),
                         ls_mean=float(d[:, 1].mean()), ls_max=int(d[:, 1].max()), ncon_max_mean=float(d[:, 2].mean()),
                         ncon_max=int(d[:, 2].max()), capped_envs=int((d[:, 3] > 0).sum()),
                         done=int((done & 1).sum().item()), unstable=int(((done & 4) != 0).sum().item()),
                         z_mean=float(obs[:, 2].mean().item())))
        print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"probe_{env_id}.json"), "w") as f:
        json.dump(dict(env=env_id, n=n, kernel=sim.kernel_config, rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
