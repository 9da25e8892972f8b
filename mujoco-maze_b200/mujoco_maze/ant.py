"""Ant robot descriptor (reference ant.py:38-111).

Step semantics baked into the kernel (STEP_TORQUE, frame_skip 5): ctrl = clamp(a),
5 RK4 `mj_step`s, inner reward = w * |dxy / dt| - c * sum(a^2) with the raw action
(ant.py:56-73). `forward_reward_fn`: the reference's two functions are fused into the kernel
(`forward_reward_vnorm` |v|, `forward_reward_vabs` |vx| + |vy|; ant.py:18-23); any other callable is evaluated
by the host wrapper on the step's xy velocity (MazeEnv.step) and added to the kernel's reward.
"""

from typing import Callable

import numpy as np

from mujoco_maze.agent_model import AgentModel

ForwardRewardFn = Callable[[float, float], float]


def forward_reward_vabs(xy_velocity) -> float:
    return np.sum(np.abs(xy_velocity))


def forward_reward_vnorm(xy_velocity) -> float:
    return np.linalg.norm(xy_velocity)


def forward_reward_kind(fn) -> int:
    """include/mmz_model.h MMZ_FWD_*: 0 vnorm, 1 vabs (both in-kernel), 2 = evaluated on the host."""
    return 0 if fn is forward_reward_vnorm else 1 if fn is forward_reward_vabs else 2


def q_inv(a):
    return [a[0], -a[1], -a[2], -a[3]]


def q_mult(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return [
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ]


class AntEnv(AgentModel):
    FILE: str = "ant.xml"
    ORI_IND: int = 3
    MANUAL_COLLISION: bool = False
    OBJBALL_TYPE: str = "freejoint"
    FRAME_SKIP: int = 5
    KERNEL_KIND: str = "ant"

    def __init__(
        self,
        file_path: str = None,
        forward_reward_weight: float = 1.0,
        ctrl_cost_weight: float = 1e-4,
        forward_reward_fn: ForwardRewardFn = forward_reward_vnorm,
    ) -> None:
        super().__init__(file_path)
        self._forward_reward_weight = forward_reward_weight
        self._ctrl_cost_weight = ctrl_cost_weight
        self._forward_reward_fn = forward_reward_fn

    def get_ori(self):
        """Heading of the torso x axis projected on the ground plane (ant.py:98-103)."""
        import torch

        qpos, _, _ = self._env._state()
        w, x, y, z = (qpos[:, self.ORI_IND + k] for k in range(4))
        # first column of the rotation matrix of (w, x, y, z), unnormalised like upstream
        ox = w * w + x * x - y * y - z * z
        oy = 2 * (x * y + w * z)
        return self._env._out(torch.atan2(oy, ox))
