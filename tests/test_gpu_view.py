"""GPU parity of MazeEnv.get_top_down_view (reference maze_env.py:262-349).

`maze_view_kernel` (fp32, gather form) against (1) the raster the UNMODIFIED reference method produced for the same
torso / block positions (tests/golden/reference_top_down_view.json) and (2) the oracle after a physics step, where the
view must come from the same stale body positions as the rest of the observation. Tolerance 1e-5 absolute on the
static comparison (weights are sums of products of numbers in [0, 1]; fp32 round-off of the cell fractions), and
after a step the view may differ by what the step's position error moves it: 2e-3.
"""
import json
import os

import numpy as np
import pytest

from test_top_down_view import CASES, G, state_for, view_task

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA")
    return torch


def _agent(kind):
    from mujoco_maze.ant import AntEnv
    from mujoco_maze.point import PointEnv

    return PointEnv if kind == "point" else AntEnv


@pytest.mark.parametrize("key", CASES)
def test_view_matches_reference_raster(key, torch_cuda):
    from mujoco_maze.backend import BatchedSim
    from mujoco_maze.model_compiler import compile_maze_model

    case = G["cases"][CASES.index(key)]
    model = compile_maze_model(_agent(case["agent"]), view_task(case["task"], case["scaling"]), case["scaling"])
    n = len(case["samples"])
    q = np.stack([state_for(model, s["robot"], s["blocks"]) for s in case["samples"]])
    sim = BatchedSim(model, n)
    sim.set_state(q, np.zeros((n, int(model.nv))), np.full(n, 7, dtype=np.int32))
    launches0 = sim.launch_count
    obs = sim.observe().cpu().numpy()
    assert obs.shape == (n, int(model.obs_dim))
    want = np.array([s["view"] for s in case["samples"]])
    np.testing.assert_allclose(obs[:, -76:-1], want, rtol=0, atol=1e-5)
    np.testing.assert_allclose(obs[:, -1], 0.007, atol=1e-7)        # t after the view (maze_env.py:369)
    np.testing.assert_allclose(obs[:, :2], q[:, :2], atol=1e-6)     # state part before it
    assert sim.launch_count - launches0 == 2                      # the observe kernel, then the view kernel
    sim.close()


@pytest.mark.parametrize("key", ["GoalRewardPush-point", "GoalRewardFall-ant", "GoalRewardUMaze-ant", "GoalRewardPush-ant"])
def test_view_after_step_matches_oracle(key, torch_cuda, oracle_lib):
    from mujoco_maze.backend import BatchedSim
    from mujoco_maze.model_compiler import compile_maze_model

    case = G["cases"][CASES.index(key)]
    model = compile_maze_model(_agent(case["agent"]), view_task(case["task"], case["scaling"]), case["scaling"])
    rng = np.random.default_rng(5)
    n, nq, nv, nu = 48, int(model.nq), int(model.nv), int(model.nu)
    s = float(model.cell_size)
    q = np.tile(np.asarray(model.qpos0, float)[:nq], (n, 1))
    q[:, :2] += rng.uniform(-0.3 * s, 0.3 * s, size=(n, 2))
    v = rng.normal(scale=0.3, size=(n, nv))
    lo, hi = np.asarray(model.act_ctrlrange, float)[:nu].T
    a = rng.uniform(lo, hi, size=(n, nu))
    sim = BatchedSim(model, n, auto_reset=False)
    sim.set_state(q, v, np.zeros(n, dtype=np.int32))
    obs, rew, done, info = sim.step(a)
    obs = obs.cpu().numpy()
    o = oracle_lib.OracleEnv(model)
    o.L.ora_set_warmstart(o.h, 1)
    want = np.zeros_like(obs, dtype=float)
    for i in range(n):
        o.set_state(q[i], v[i], 0)
        want[i] = o.step(a[i])[0]
    err = np.abs(obs[:, -76:-1] - want[:, -76:-1]).max(1)
    os.makedirs("gpurun_out/parity", exist_ok=True)
    json.dump(dict(case=key, n=n, view_err_max=float(err.max()), view_err_median=float(np.median(err))),
              open(f"gpurun_out/parity/view_{key}.json", "w"))
    assert np.quantile(err, 0.95) <= 2e-3 and float(np.abs(want[:, -76:-1]).max()) > 0.5
    np.testing.assert_allclose(obs[:, -1], 0.001, atol=1e-7)
    sim.close()
