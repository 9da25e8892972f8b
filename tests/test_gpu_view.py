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


@pytest.mark.parametrize("key", ["GoalRewardPush-point", "GoalRewardUMaze-ant"])
def test_view_with_auto_reset_keeps_the_columns(key, torch_cuda, oracle_lib):
    """A TOP_DOWN_VIEW task under the in-kernel auto-reset (VectorMazeEnv): the environments that do NOT reset must
    get the state part, the view and t * 0.001 in the same columns as without auto-reset (the lanes-per-environment
    kernel once copied obs_dim entries of an obs_core-sized buffer there)."""
    from mujoco_maze.backend import BatchedSim
    from mujoco_maze.model_compiler import compile_maze_model

    case = G["cases"][CASES.index(key)]
    model = compile_maze_model(_agent(case["agent"]), view_task(case["task"], case["scaling"]), case["scaling"])
    rng = np.random.default_rng(6)
    n, nq, nv, nu = 40, int(model.nq), int(model.nv), int(model.nu)
    s = float(model.cell_size)
    q = np.tile(np.asarray(model.qpos0, float)[:nq], (n, 1))
    q[:, :2] += rng.uniform(-0.3 * s, 0.3 * s, size=(n, 2))
    v = rng.normal(scale=0.3, size=(n, nv))
    lo, hi = np.asarray(model.act_ctrlrange, float)[:nu].T
    a = rng.uniform(lo, hi, size=(n, nu))
    t0 = np.full(n, 41, dtype=np.int32)
    t0[n // 2:] = 999  # the second half hits the TimeLimit in this step and restarts inside the launch
    outs = []
    for auto in (False, True):
        sim = BatchedSim(model, n, auto_reset=auto)
        sim.set_state(q, v, t0)
        obs, rew, done, info = sim.step(a)
        outs.append((obs.cpu().numpy().copy(), done.cpu().numpy().copy()))
        sim.close()
    (plain, d0), (auto, d1) = outs
    h = n // 2
    assert (d0[:h] & 1).sum() == 0 and ((d1[h:] & 1) == 1).all()
    np.testing.assert_array_equal(auto[:h], plain[:h])                 # running episodes: identical, column for column
    np.testing.assert_allclose(auto[:h, -1], 0.042, atol=1e-7)
    np.testing.assert_allclose(auto[h:, -1], 0.0, atol=1e-7)           # fresh episodes start at t = 0 ...
    assert float(np.abs(auto[h:, -76:-1]).max()) > 0.5                 # ... with a view of their reset position
    o = oracle_lib.OracleEnv(model)
    o.L.ora_set_warmstart(o.h, 1)
    for i in range(0, h, 4):
        o.set_state(q[i], v[i], int(t0[i]))
        want = o.step(a[i])[0]
        assert np.abs(auto[i, -76:-1] - want[-76:-1]).max() <= 5e-3
        np.testing.assert_allclose(auto[i, :2], want[:2], atol=1e-3)
