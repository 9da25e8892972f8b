"""Swimmer robot descriptor (reference swimmer.py:16-76): STEP_TORQUE, frame_skip 4."""

from mujoco_maze.agent_model import AgentModel
from mujoco_maze.ant import ForwardRewardFn, forward_reward_vnorm


class SwimmerEnv(AgentModel):
    FILE: str = "swimmer.xml"
    MANUAL_COLLISION: bool = False
    FRAME_SKIP: int = 4
    KERNEL_KIND: str = "swimmer"

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
