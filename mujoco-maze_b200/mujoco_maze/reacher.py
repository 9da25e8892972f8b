"""Reacher robot descriptor (reference reacher.py:15-76): a two-link swimmer, STEP_TORQUE, frame_skip 4.

The reference marks it "not tested" (README.md:129-130) and registers it on the swimmer maze scale
(__init__.py:51-64). It runs through the same kernel path as the Swimmer: no contacts, fluid drag, one
limited hinge, one motor; observation = qpos ++ qvel (+ t) = 9 entries (reference tests/test_envs.py:63-64).
"""

from mujoco_maze.agent_model import AgentModel
from mujoco_maze.ant import ForwardRewardFn, forward_reward_vnorm


class ReacherEnv(AgentModel):
    FILE: str = "reacher.xml"
    MANUAL_COLLISION: bool = False
    FRAME_SKIP: int = 4
    KERNEL_KIND: str = "swimmer"  # same step / observation / reset rules as the Swimmer (reacher.py:36-68)

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
