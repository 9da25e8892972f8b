"""mujoco_maze — B200-native batched maze-navigation environments.

Same user-facing surface as kngwyu/mujoco-maze (reference mujoco_maze/__init__.py:
importing the package registers `Point/Ant/Swimmer/Reacher{Maze}-v{i}` with entry point
`mujoco_maze.maze_env:MazeEnv`), but `MazeEnv.step` advances N lock-step
environments in one hand-written sm_100a CUDA kernel behind the C ABI of
include/mmz.h. `gym` is used when installed; otherwise `mujoco_maze.gym` is a
minimal in-repo shim of the slice of gym 0.20 the API needs.
"""

try:  # pragma: no cover - gym is absent in the build image
    import gym  # type: ignore

    if not hasattr(gym, "envs") or not hasattr(gym.envs, "register"):
        raise ImportError
except ImportError:
    from mujoco_maze import gym_shim as gym  # noqa: F401

from mujoco_maze.ant import AntEnv  # noqa: E402
from mujoco_maze.maze_task import TaskRegistry  # noqa: E402
from mujoco_maze.point import PointEnv  # noqa: E402
from mujoco_maze.reacher import ReacherEnv  # noqa: E402
from mujoco_maze.swimmer import SwimmerEnv  # noqa: E402

__version__ = "0.2.0+b200.r2"

MAX_EPISODE_STEPS = 1000  # reference __init__.py:31,47,76

_AGENTS = (  # (id prefix, descriptor class, which Scaling field selects the maze size)
    ("Point", PointEnv, "point"),
    ("Ant", AntEnv, "ant"),
    ("Swimmer", SwimmerEnv, "swimmer"),
    ("Reacher", ReacherEnv, "swimmer"),  # the reference keys Reacher on the swimmer scale too (__init__.py:51-64, quirk Q14)
)


def _register_all() -> None:
    for maze_id in TaskRegistry.keys():
        for version, task_cls in enumerate(TaskRegistry.tasks(maze_id)):
            for prefix, agent_cls, scale_field in _AGENTS:
                scale = getattr(task_cls.MAZE_SIZE_SCALING, scale_field)
                if scale is None:
                    continue
                env_id = f"{prefix}{maze_id}-v{version}"
                if env_id in getattr(gym.envs.registry, "env_specs", {}):
                    continue  # re-import
                gym.envs.register(
                    id=env_id,
                    entry_point="mujoco_maze.maze_env:MazeEnv",
                    kwargs=dict(
                        model_cls=agent_cls,
                        maze_task=task_cls,
                        maze_size_scaling=scale,
                        inner_reward_scaling=task_cls.INNER_REWARD_SCALING,
                    ),
                    max_episode_steps=MAX_EPISODE_STEPS,
                    reward_threshold=task_cls.REWARD_THRESHOLD,
                )


_register_all()


def make(env_id: str, **kwargs):
    """`gym.make` for this package's ids that is safe for BATCHED environments under the real gym.

    A batched environment (`num_envs=N`) counts its TimeLimit steps per environment inside the kernel and returns tensors;
    gym's own scalar `TimeLimit` wrapper (which `gym.make` adds from the registered `max_episode_steps`) evaluates
    `not done` on them and fails. The in-repo shim skips the wrapper for batched environments; with the real gym
    installed, build them through this function, which constructs the registered entry point directly."""
    if kwargs.get("num_envs") is None:
        return gym.make(env_id, **kwargs)
    spec = gym.spec(env_id) if hasattr(gym, "spec") else gym.envs.registry.spec(env_id)
    merged = dict(getattr(spec, "kwargs", None) or getattr(spec, "_kwargs", None) or {})
    merged.update(kwargs)
    from mujoco_maze.maze_env import MazeEnv

    env = MazeEnv(**merged)
    env.spec = spec
    return env


def install_gym_shim() -> None:
    """Expose the in-repo shim as `import gym` when the real package is missing."""
    import sys

    sys.modules.setdefault("gym", gym)
