"""Agent descriptors.

In the reference an `AgentModel` *is* a MuJoCo simulation (`MujocoEnv`,
agent_model.py:12-21). Here the simulation lives in the batched CUDA backend, so
an agent class is a descriptor: the class attributes the reference defines
(FILE, MANUAL_COLLISION, ORI_IND, RADIUS, OBJBALL_TYPE; agent_model.py:13-17)
plus the constants its `step` / `_get_obs` / `reset_model` imply, which the
model compiler bakes into the blob. Instances keep the reference's accessor
methods (`get_xy`, `set_xy`, `get_ori`, `_get_obs`) and forward them to the
backend the owning `MazeEnv` binds.
"""

from typing import Optional

import numpy as np

from mujoco_maze import gym


class AgentModel:
    FILE: str
    MANUAL_COLLISION: bool
    ORI_IND: Optional[int] = None
    RADIUS: Optional[float] = None
    OBJBALL_TYPE: Optional[str] = None
    # --- additions consumed by the model compiler -------------------------
    FRAME_SKIP: int = 1
    KERNEL_KIND: str = "generic"  # "point" | "ant" | "swimmer"

    def __init__(self, file_path: Optional[str] = None, **kwargs) -> None:
        self.file_path = file_path
        self._env = None  # bound by MazeEnv

    # -- wiring ------------------------------------------------------------
    def _bind(self, env) -> None:
        self._env = env

    @property
    def dt(self) -> float:
        return float(self._env.model.timestep) * self.FRAME_SKIP

    @property
    def action_space(self) -> gym.spaces.Box:
        rng = self._env.model.meta["act_ctrlrange"]
        return gym.spaces.Box(rng[:, 0].astype(np.float32), rng[:, 1].astype(np.float32))

    @property
    def observation_space(self) -> gym.spaces.Box:
        n = int(self._env.model.n_agent_q + self._env.model.n_agent_v)
        high = np.inf * np.ones(n, dtype=np.float32)
        return gym.spaces.Box(-high, high)

    # -- reference accessors -------------------------------------------------
    def _get_obs(self):
        """qpos[:n_agent_q] ++ qvel[:n_agent_v] (point.py:63-69, ant.py:75-82, swimmer.py:49-53)."""
        qpos, qvel, _ = self._env._state()
        m = self._env.model
        return self._env._out(self._env._cat(qpos[:, : int(m.n_agent_q)], qvel[:, : int(m.n_agent_v)]))

    def get_xy(self):
        qpos, _, _ = self._env._state()
        return self._env._out(qpos[:, :2].clone())

    def set_xy(self, xy) -> None:
        self._env._set_xy(xy)

    def get_ori(self):
        qpos, _, _ = self._env._state()
        return self._env._out(qpos[:, self.ORI_IND])

    def close(self) -> None:
        pass
