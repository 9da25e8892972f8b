"""The checks of the reference's own tests/test_envs.py, through the same calls (`gym.make(id)`, `reset()`,
`step(action_space.sample())`, `unwrapped.has_extended_obs / _observe_balls / _task`), on the GPU-backed package.

Reference lines in brackets. One environment per call, numpy in / numpy out, exactly the reference's calling
convention; the step itself is the CUDA kernel (there is no CPU path)."""
import numpy as np
import pytest

import mujoco_maze
from mujoco_maze import gym

pytestmark = pytest.mark.gpu
MAZE_IDS = list(mujoco_maze.TaskRegistry.keys())


def _has(env_id):
    specs = getattr(getattr(gym, "registry", None), "env_specs", None)
    return specs is None or env_id in specs


def _reset_step(env_id, **kw):
    env = gym.make(env_id, **kw)
    s0 = env.reset()
    s, r, done, info = env.step(env.action_space.sample())
    assert isinstance(s0, np.ndarray) and isinstance(s, np.ndarray) and isinstance(r, float) and isinstance(done, bool)
    assert np.isfinite(s).all() and "position" in info
    return env, s0, s, r


@pytest.mark.parametrize("maze_id", MAZE_IDS)
def test_ant_maze(maze_id):  # [tests/test_envs.py:7-17] (the reference skips Billiard; here it runs too)
    for i in range(2):
        if not _has(f"Ant{maze_id}-v{i}"):
            continue
        env, s0, s, _ = _reset_step(f"Ant{maze_id}-v{i}")
        if not env.unwrapped.has_extended_obs:
            assert s0.shape == (30,) and s.shape == (30,)
        env.close()


@pytest.mark.parametrize("maze_id", MAZE_IDS)
def test_point_maze(maze_id):  # [tests/test_envs.py:20-36]
    for i in range(2):
        if not _has(f"Point{maze_id}-v{i}"):
            continue
        env, s0, s, r = _reset_step(f"Point{maze_id}-v{i}")
        if not env.unwrapped.has_extended_obs:
            assert s0.shape == (7,) and s.shape == (7,)
        if env.unwrapped._observe_balls:
            assert s0.shape == (10,) and s.shape == (10,)
        if i == 0:
            assert r != 0.0
        else:
            assert r == env.unwrapped._task.PENALTY
            assert r < 0.0
        env.close()


@pytest.mark.parametrize("maze_id", ["2Rooms", "4Rooms", "Billiard"])
def test_subgoal_envs(maze_id):  # [tests/test_envs.py:39-50]
    env, s0, s, _ = _reset_step(f"Point{maze_id}-v2")
    if not env.unwrapped.has_extended_obs:
        assert s0.shape == (7,) and s.shape == (7,)
    elif env.unwrapped._observe_balls:
        assert s0.shape == (10,) and s.shape == (10,)
    assert len(env.unwrapped._task.goals) > 1
    env.close()


@pytest.mark.parametrize("agent,dim", [("Reacher", 9), ("Swimmer", 11)])
@pytest.mark.parametrize("maze_id", MAZE_IDS)
def test_reacher_and_swimmer_maze(maze_id, agent, dim):  # [tests/test_envs.py:53-78]
    if any(x in maze_id for x in ("Fall", "Push", "Block", "Billiard")):
        return
    for i in range(2):
        if not _has(f"{agent}{maze_id}-v{i}"):
            continue
        env, s0, s, _ = _reset_step(f"{agent}{maze_id}-v{i}")
        if not env.unwrapped.has_extended_obs:
            assert s0.shape == (dim,) and s.shape == (dim,)
        env.close()


@pytest.mark.parametrize("v", [0, 1])
def test_maze_args(v):  # [tests/test_envs.py:81-86]
    env, s0, s, _ = _reset_step(f"PointTRoom-v{v}", task_kwargs={"goal": (-2.0, -3.0)})
    assert s0.shape == (7,) and s.shape == (7,)
    assert np.allclose(env.unwrapped._task.goals[0].pos, np.array([-2.0, -3.0]) * env.unwrapped._task.scale)
    env.close()
