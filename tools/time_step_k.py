#!/usr/bin/env python
"""Device-resident rollout throughput through mmz_step_k (K steps per host call, one cached CUDA graph; GPU only).

    python tools/time_step_k.py [ENV_ID] [N_ENVS] [K] [CALLS]
Unlike bench.py there is no L2 flush between steps and no host work between launches: this is what a rollout collector
that owns its actions in advance (or a policy captured in the same graph) sees.
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, "mujoco-maze_b200")
import mujoco_maze  # noqa: E402,F401
from mujoco_maze import gym  # noqa: E402
from mujoco_maze.backend import BatchedSim  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "PointUMaze-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 5
model = gym.make(env_id, num_envs=1).unwrapped.model
sim = BatchedSim(model, n, auto_reset=True)
r = np.asarray(model.meta["act_ctrlrange"], float)
lo, hi = (torch.tensor(x, device="cuda", dtype=torch.float32) for x in (r[:, 0], r[:, 1]))
g = torch.Generator(device="cuda").manual_seed(1)
acts = lo + (hi - lo) * torch.rand((K, n, sim.nu), device="cuda", generator=g)
sim.reset(seed=0)
out = sim.step_k(acts)  # records the graph
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(calls):
    out = sim.step_k(acts, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"env": env_id, "envs": n, "K": K, "calls": calls, "ms_per_step": ms / (K * calls),
                  "env_steps_per_sec": n * K * calls / (ms * 1e-3), "kernel": sim.kernel_config["kernel"]}))
