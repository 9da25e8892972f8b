"""Point robot descriptor (reference point.py:17-92).

Step semantics baked into the kernel (STEP_TELEPORT): heading += a[1] wrapped
once into [-pi, pi]; xy += a[0] * (cos, sin); every qvel clipped to
+-VELOCITY_LIMITS; one RK4 `mj_step` with zero control (point.py:44-61).
"""

import numpy as np

from mujoco_maze import gym
from mujoco_maze.agent_model import AgentModel


class PointEnv(AgentModel):
    FILE: str = "point.xml"
    ORI_IND: int = 2
    MANUAL_COLLISION: bool = True
    RADIUS: float = 0.4
    OBJBALL_TYPE: str = "hinge"
    VELOCITY_LIMITS: float = 10.0
    FRAME_SKIP: int = 1
    KERNEL_KIND: str = "point"

    @property
    def observation_space(self) -> gym.spaces.Box:
        high = np.inf * np.ones(6, dtype=np.float32)
        high[3:] = self.VELOCITY_LIMITS * 1.2
        high[self.ORI_IND] = np.pi
        return gym.spaces.Box(-high, high)
