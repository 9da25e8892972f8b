"""Solver statistics of the benchmark workload (development aid): Newton / line-search iterations per env-step.

    python tools/solver_stats.py AntUMaze-v0 16384 [steps]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, "mujoco-maze_b200")
import mujoco_maze  # noqa: E402,F401
from mujoco_maze import gym  # noqa: E402

env_id, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
env = gym.make(env_id, num_envs=n, device="cuda:0", auto_reset=True)
env.reset(seed=0)
sim = env.unwrapped.sim
diag = sim.enable_step_diag(True)
lo, hi = (torch.as_tensor(x, device="cuda:0") for x in (env.action_space.low, env.action_space.high))
g = torch.Generator(device="cuda:0").manual_seed(1)
rows = []
for k in range(steps):
    a = lo + (hi - lo) * torch.rand((n, lo.numel()), device="cuda:0", generator=g)
    env.step(a)
    d = diag.cpu().numpy().astype(float)
    rows.append(d.mean(0))
    if k >= steps - 3:
        it = d[:, 0]
        print(f"step {k}: newton/step mean {it.mean():.2f} p50 {np.median(it):.0f} p99 {np.quantile(it, .99):.0f} max {it.max():.0f} | "
              f"line-search/step {d[:, 1].mean():.2f} | max contacts mean {d[:, 2].mean():.2f} | capped {d[:, 3].sum():.0f}")
    if k == steps - 1:
        cm = d[:, 2].astype(int)
        print("max contacts per env-step, histogram:", {int(v): int((cm == v).sum()) for v in np.unique(cm)})
