import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mujoco-maze_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def compile_test_xml(name, rows="R", scale=4.0, kind="swimmer", **kw):
    """Compile tests/assets/<name> as a generic torque agent in an (almost) empty maze."""
    from mujoco_maze.agent_model import AgentModel
    from mujoco_maze.maze_task import MazeTask, _maze
    from mujoco_maze.model_compiler import compile_maze_model

    class _Agent(AgentModel):
        FILE = os.path.join(ASSETS, name)
        MANUAL_COLLISION = False
        FRAME_SKIP = 1
        KERNEL_KIND = kind

    class _Task(MazeTask):
        REWARD_THRESHOLD = 0.0
        PENALTY = 0.0
        create_maze = _maze(rows)

        def reward(self, obs):
            return 0.0

    return compile_maze_model(_Agent, _Task(scale), scale, **kw)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import mmz_oracle

    mmz_oracle.build()
    return mmz_oracle


def make_model(env_id, **kw):
    import mujoco_maze
    from mujoco_maze import gym

    return gym.make(env_id, **kw).unwrapped.model
