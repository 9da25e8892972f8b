"""MazeEnv.get_top_down_view (reference maze_env.py:262-349): the oracle against the reference's own method.

tests/golden/reference_top_down_view.json was produced by tests/golden/gen_view_goldens.py, which calls the
UNMODIFIED reference method. Here the same torso / block positions are put into the oracle through set_state and the
75 view entries of its observation are compared with the reference's raster (fp64 both sides).
"""
import json
import os

import numpy as np
import pytest

from mujoco_maze import maze_task as T
from mujoco_maze.ant import AntEnv
from mujoco_maze.model_compiler import compile_maze_model
from mujoco_maze.point import PointEnv

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_top_down_view.json")))
CASES = [f'{c["task"]}-{c["agent"]}' for c in G["cases"]]


def view_task(name, scaling):
    cls = getattr(T, name)
    return type(name + "View", (cls,), dict(TOP_DOWN_VIEW=True))(scaling)


def state_for(model, robot, blocks):
    """qpos that puts the torso at `robot` and every movable block at its golden xy."""
    qpos = np.array(model.qpos0[: int(model.nq)], float)
    qpos[0], qpos[1] = robot
    for name, xy in blocks.items():
        b = model.names["body"].index(name)
        for j in range(int(model.njnt)):
            if int(model.jnt_body[j]) != b:
                continue
            axis = np.asarray(model.jnt_axis[j], float)
            k = int(np.argmax(np.abs(axis)))
            if k < 2:
                qpos[int(model.jnt_qadr[j])] = xy[k] - float(model.body_pos[b][k])
    return qpos


@pytest.mark.parametrize("key", CASES)
def test_oracle_view_matches_reference(key, oracle_lib):
    case = G["cases"][CASES.index(key)]
    agent = PointEnv if case["agent"] == "point" else AntEnv
    model = compile_maze_model(agent, view_task(case["task"], case["scaling"]), case["scaling"])
    assert int(model.view_dim) == 75 and int(model.nviewb) == 1 + len(case["samples"][0]["blocks"])
    plain = compile_maze_model(agent, getattr(T, case["task"])(case["scaling"]), case["scaling"])
    assert int(model.obs_dim) == int(plain.obs_dim) + 75
    o = oracle_lib.OracleEnv(model)
    core = int(model.obs_dim) - 76
    for s in case["samples"]:
        o.set_state(state_for(model, s["robot"], s["blocks"]), np.zeros(int(model.nv)), t=7)
        obs = o.observe()
        # the state part and the trailing t are where they are without the view (maze_env.py:368-369)
        assert obs[0] == pytest.approx(s["robot"][0]) and obs[-1] == pytest.approx(0.007)
        np.testing.assert_allclose(obs[core: core + 75], s["view"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("key", ["GoalRewardUMaze-point", "GoalRewardPush-point"])
def test_view_is_agent_independent(key, oracle_lib):
    """The reference method only reads the maze, the scaling and two body positions: the same goldens hold for a
    Swimmer in that maze (torso = its first body, x / y = its first two slide joints)."""
    from mujoco_maze.swimmer import SwimmerEnv

    case = G["cases"][CASES.index(key)]
    model = compile_maze_model(SwimmerEnv, view_task(case["task"], case["scaling"]), case["scaling"])
    assert model.names["body"][int(model.obj_body[int(model.nobj)])] == "torso"
    o = oracle_lib.OracleEnv(model)
    core = int(model.obs_dim) - 76
    for s in case["samples"][:8]:
        o.set_state(state_for(model, s["robot"], s["blocks"]), np.zeros(int(model.nv)), t=0)
        np.testing.assert_allclose(o.observe()[core: core + 75], s["view"], rtol=0, atol=1e-9)


def test_view_is_off_for_every_registered_task():
    # maze_task.py:68 - no upstream task overrides TOP_DOWN_VIEW, so registered ids never pay for the view
    for maze_id in T.TaskRegistry.keys():
        for cls in T.TaskRegistry.tasks(maze_id):
            assert cls.TOP_DOWN_VIEW is False


def test_env_with_view_has_reference_obs_shape_and_no_cpu_fallback():
    """MazeEnv of a TOP_DOWN_VIEW task: obs = state part + 75 view entries + t (maze_env.py:368-369); without CUDA the
    view accessor fails loudly like every other product path."""
    import torch

    from mujoco_maze import maze_env
    from mujoco_maze.backend import MmzError

    cls = type("PushView", (T.GoalRewardPush,), dict(TOP_DOWN_VIEW=True))
    env = maze_env.MazeEnv(PointEnv, cls, maze_size_scaling=4.0)
    plain = maze_env.MazeEnv(PointEnv, T.GoalRewardPush, maze_size_scaling=4.0)
    assert env.observation_space.shape == (plain.observation_space.shape[0] + 75,)
    assert env.has_extended_obs
    with pytest.raises(ValueError):
        plain.get_top_down_view()
    if not torch.cuda.is_available():
        with pytest.raises(MmzError):
            env.get_top_down_view()
        with pytest.raises(MmzError):
            env.render(mode="rgb_array")
    with pytest.raises(NotImplementedError):
        env.render(mode="human")
