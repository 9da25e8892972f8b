#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): a few MazeEnv.step launches of a small batch,
with the TimeLimit firing (in-kernel auto-reset) on the second step.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py AntUMaze-v0 [N] [steps]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, "mujoco-maze_b200")
import mujoco_maze  # noqa: E402,F401
from mujoco_maze import gym  # noqa: E402

env_id = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
env = gym.make(env_id, num_envs=n, device="cuda:0", auto_reset=True)
env.reset(seed=0)
sim = env.unwrapped.sim
q, v, t = sim.get_state()
sim.set_state(q, v, torch.full((n,), 998, dtype=torch.int32, device="cuda:0"))
lo, hi = (torch.as_tensor(x, device="cuda:0") for x in (env.action_space.low, env.action_space.high))
g = torch.Generator(device="cuda:0").manual_seed(1)
ndone = 0
for k in range(steps):
    a = lo + (hi - lo) * torch.rand((n, lo.numel()), device="cuda:0", generator=g)
    obs, rew, done, info = env.step(a)
    ndone += int((done.to(torch.uint8) & 1).sum().item()) if torch.is_tensor(done) else int(np.sum(done))
# the host call: pinned buffers mapped into the device address space, read / written by the step kernel itself
h = [torch.empty(s, dtype=d).pin_memory() for s, d in (((n, sim.nu), torch.float32), ((n, sim.obs_dim), torch.float32),
                                                       ((n,), torch.float32), ((n,), torch.uint8), ((n, 4), torch.float32))]
h[0].copy_(a.cpu())
sim.step_host(*h)
assert bool(torch.isfinite(h[1]).all())
torch.cuda.synchronize()
print(f"{env_id}: {sim.kernel_config['kernel']} x{n}, {steps} steps, {ndone} episode ends (auto-reset), obs finite: "
      f"{bool(torch.isfinite(obs).all())}")
