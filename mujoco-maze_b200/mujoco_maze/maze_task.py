"""Task catalogue: maze maps, goals, reward / termination rules.

Host-side mirror of the reference task library (maze_task.py:26-807). Tasks are
plain Python objects evaluated ONCE at env construction: the map becomes the
wall grid, the goals and the *resolved* reward/termination rule become
constants of the model blob, and the per-step evaluation happens inside the
CUDA step kernel. `kernel_rule(task)` performs that resolution by looking at
which function object Python's MRO actually binds (reference quirk Q1: in 16 of
the 18 `DistReward*` classes the goal/penalty reward shadows the distance
mix-in, maze_task.py:93-99,125).

Maps are written as compact strings, one character per cell:
  B wall   . empty   R robot start   C chasm   O object ball
  M XY block   Y YZ block   Z XYZ block
"""

from abc import ABC, abstractmethod
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple, Type

import numpy as np

from mujoco_maze.maze_env_utils import MazeCell


class Rgb(NamedTuple):
    red: float
    green: float
    blue: float

    def rgba_str(self) -> str:
        return f"{self.red} {self.green} {self.blue} 1"


RED = Rgb(0.7, 0.1, 0.1)
GREEN = Rgb(0.1, 0.7, 0.1)
BLUE = Rgb(0.1, 0.1, 0.7)

_CELL_OF = {
    "B": MazeCell.BLOCK,
    ".": MazeCell.EMPTY,
    "R": MazeCell.ROBOT,
    "C": MazeCell.CHASM,
    "O": MazeCell.OBJECT_BALL,
    "M": MazeCell.XY_BLOCK,
    "Y": MazeCell.YZ_BLOCK,
    "Z": MazeCell.XYZ_BLOCK,
}


def parse_map(rows: str) -> List[List[MazeCell]]:
    """'BBB/BRB/BBB' -> list of lists of MazeCell."""
    return [[_CELL_OF[c] for c in row] for row in rows.split("/")]


def _maze(rows: str):
    """Build the `create_maze` staticmethod of a task class from a map string."""

    def create_maze() -> List[List[MazeCell]]:
        return parse_map(rows)

    return staticmethod(create_maze)


class MazeGoal:
    def __init__(
        self,
        pos: np.ndarray,
        reward_scale: float = 1.0,
        rgb: Rgb = RED,
        threshold: float = 0.6,
        custom_size: Optional[float] = None,
    ) -> None:
        assert 0.0 <= reward_scale <= 1.0
        self.pos = pos
        self.dim = pos.shape[0]
        self.reward_scale = reward_scale
        self.rgb = rgb
        self.threshold = threshold
        self.custom_size = custom_size

    def _offset(self, obs: np.ndarray) -> np.ndarray:
        return np.asarray(obs)[: self.dim] - self.pos

    def neighbor(self, obs: np.ndarray) -> bool:
        return bool(np.linalg.norm(self._offset(obs)) <= self.threshold)

    def euc_dist(self, obs: np.ndarray) -> float:
        return float(np.sum(np.square(self._offset(obs))) ** 0.5)


class Scaling(NamedTuple):
    ant: Optional[float]
    point: Optional[float]
    swimmer: Optional[float]


class MazeTask(ABC):
    REWARD_THRESHOLD: float
    PENALTY: Optional[float] = None
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=8.0, point=4.0, swimmer=4.0)
    INNER_REWARD_SCALING: float = 0.01
    OBSERVE_BLOCKS: bool = False  # Fall / Push / BlockMaze
    OBSERVE_BALLS: bool = False  # Billiard
    OBJECT_BALL_SIZE: float = 1.0
    PUT_SPIN_NEAR_AGENT: bool = False  # unused upstream
    TOP_DOWN_VIEW: bool = False  # unused upstream

    def __init__(self, scale: float) -> None:
        self.goals: List[MazeGoal] = []
        self.scale = scale

    def sample_goals(self) -> bool:
        return False

    def termination(self, obs: np.ndarray) -> bool:
        return any(g.neighbor(obs) for g in self.goals)

    @abstractmethod
    def reward(self, obs: np.ndarray) -> float:
        ...

    @staticmethod
    @abstractmethod
    def create_maze() -> List[List[MazeCell]]:
        ...


# ---------------------------------------------------------------------------
# The five reward rules and two termination rules the registry resolves to.
# They are module-level functions so that `kernel_rule` can identify them by
# identity after Python's MRO has picked one.
# ---------------------------------------------------------------------------
def _first_goal_hit(goals: Sequence[MazeGoal], where: np.ndarray) -> Optional[MazeGoal]:
    for g in goals:
        if g.neighbor(where):
            return g
    return None


def _reward_reach(self, obs):  # R0: 1 on termination, else PENALTY
    return 1.0 if self.termination(obs) else self.PENALTY


def _reward_scaled(self, obs):  # R1: reward_scale of the first goal reached
    g = _first_goal_hit(self.goals, obs)
    return self.PENALTY if g is None else g.reward_scale


def _reward_scaled_object(self, obs):  # R2: R1 evaluated on the object position
    g = _first_goal_hit(self.goals, obs[3:6])
    return self.PENALTY if g is None else g.reward_scale


def _reward_dist_object(self, obs):  # R3: -distance(object, goal0) / scale
    return -self.goals[0].euc_dist(obs[3:6]) / self.scale


def _reward_zero(self, _obs):  # R4
    return 0.0


def _reward_dist(self, obs):  # R5: -distance(agent, goal0) / scale
    return -self.goals[0].euc_dist(obs) / self.scale


def _termination_object(self, obs):  # T1: any goal reached by the object
    return _first_goal_hit(self.goals, obs[3:6]) is not None


class DistRewardMixIn:
    REWARD_THRESHOLD: float = -1000.0
    goals: List[MazeGoal]
    scale: float
    reward = _reward_dist


# Rule ids shared with the model blob (include/mmz_model.h: MMZ_REWARD_*, MMZ_TERM_*).
REWARD_REACH, REWARD_SCALED, REWARD_SCALED_OBJECT = 0, 1, 2
REWARD_DIST_OBJECT, REWARD_ZERO, REWARD_DIST, REWARD_HOST = 3, 4, 5, 6
TERM_AGENT, TERM_OBJECT, TERM_HOST = 0, 1, 2

_REWARD_IDS = {
    _reward_reach: REWARD_REACH,
    _reward_scaled: REWARD_SCALED,
    _reward_scaled_object: REWARD_SCALED_OBJECT,
    _reward_dist_object: REWARD_DIST_OBJECT,
    _reward_zero: REWARD_ZERO,
    _reward_dist: REWARD_DIST,
}


def kernel_rule(task: MazeTask) -> Tuple[int, int]:
    """(reward rule id, termination rule id) that Python resolves for `task`.

    A user subclass with its own `reward`/`termination` (README.md:91-113) maps
    to *_HOST: the kernel then leaves that output to the host wrapper, which
    calls the Python method per env on the returned observation.
    """
    cls = type(task)
    term_fn = cls.termination
    if term_fn is MazeTask.termination:
        term = TERM_AGENT
    elif term_fn is _termination_object:
        term = TERM_OBJECT
    else:
        term = TERM_HOST
    reward = _REWARD_IDS.get(cls.reward, REWARD_HOST)
    if reward == REWARD_REACH and term == TERM_HOST:
        reward = REWARD_HOST  # R0 is defined through termination()
    return reward, term


# ---------------------------------------------------------------------------
# Goal-reaching family: one goal, R0
# ---------------------------------------------------------------------------
class GoalRewardUMaze(MazeTask):
    REWARD_THRESHOLD: float = 0.9
    PENALTY: float = -0.0001
    reward = _reward_reach
    create_maze = _maze("BBBBB/BR..B/BBB.B/B...B/BBBBB")

    def __init__(self, scale: float) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array([0.0, 2.0 * scale]))]


class DistRewardUMaze(GoalRewardUMaze, DistRewardMixIn):
    pass


class GoalRewardSimpleRoom(GoalRewardUMaze):
    create_maze = _maze("BBBBB/BR..B/BBBBB")

    def __init__(self, scale: float) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array([2.0 * scale, 0.0]))]


class DistRewardSimpleRoom(GoalRewardSimpleRoom, DistRewardMixIn):
    pass


class GoalRewardSquareRoom(GoalRewardUMaze):
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=2.5, point=4.0, swimmer=2.0)
    create_maze = _maze("BBBBB/B...B/B.R.B/B...B/BBBBB")

    def __init__(self, scale: float, goal: Tuple[float, float] = (1.0, 0.0)) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array(goal) * scale)]


class NoRewardSquareRoom(GoalRewardSquareRoom):
    reward = _reward_zero

    def __init__(self, scale: float) -> None:
        super().__init__(scale)


class DistRewardSquareRoom(GoalRewardSquareRoom, DistRewardMixIn):
    pass


class GoalRewardPush(GoalRewardUMaze):
    OBSERVE_BLOCKS: bool = True
    create_maze = _maze("BBBBB/B.RBB/B.M.B/BB.BB/BBBBB")

    def __init__(self, scale: float) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array([0.0, 2.375]) * scale)]


class DistRewardPush(GoalRewardPush, DistRewardMixIn):
    pass


class GoalRewardMultiPush(GoalRewardUMaze):
    OBSERVE_BLOCKS: bool = True
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=2.0, point=6.0, swimmer=None)
    create_maze = _maze("BBBBBB/BBB.BB/B..M.B/B.R.BB/B..M.B/BBB.BB/BBBBBB")

    def __init__(self, scale: float, goal: Tuple[float, float] = (1.0, -2)) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array(goal) * scale)]


class DistRewardMultiPush(GoalRewardMultiPush, DistRewardMixIn):
    pass


class NoRewardMultiPush(GoalRewardMultiPush):
    reward = _reward_zero


class GoalRewardMultiPushSmall(GoalRewardMultiPush):
    create_maze = _maze("BBBBBB/BB.BBB/B.M.BB/BBRM.B/B.M.BB/BB.BBB/BBBBBB")

    def __init__(self, scale: float, goal: Tuple[float, float] = (1.0, -1.0)) -> None:
        super().__init__(scale, goal)


class DistRewardMultiPushSmall(GoalRewardMultiPushSmall, DistRewardMixIn):
    pass


class NoRewardMultiPushSmall(GoalRewardMultiPushSmall):
    reward = _reward_zero


class GoalRewardPushMaze(GoalRewardUMaze):
    OBSERVE_BLOCKS: bool = True
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=2.0, point=6.0, swimmer=None)
    create_maze = _maze("BBBBBBB/B..RM.B/BBBB.BB/B.M.MBB/BB.B.BB/BBBBBBB")

    def __init__(self, scale: float, goal: Tuple[float, float] = (3.0, 0.0)) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array(goal) * scale)]


class DistRewardPushMaze(GoalRewardPushMaze, DistRewardMixIn):
    pass


class NoRewardPushMaze(GoalRewardPushMaze):
    reward = _reward_zero


class GoalRewardFall(GoalRewardUMaze):
    OBSERVE_BLOCKS: bool = True
    create_maze = _maze("BBBB/BR.B/B.YB/BCCB/B..B/BBBB")

    def __init__(self, scale: float) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array([0.0, 3.375, 4.5]) * scale)]


class DistRewardFall(GoalRewardFall, DistRewardMixIn):
    pass


class GoalRewardMultiFall(GoalRewardUMaze):
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=2.0, point=None, swimmer=None)
    OBSERVE_BLOCKS: bool = True
    PENALTY: float = -0.0001
    create_maze = _maze("BBBBBB/BR.C.B/B.ZC.B/BCCBBB/B..BBB/BBBBBB")

    def __init__(self, scale: float, goal: Tuple[int, int] = (3.0, 1.0)) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array([*goal, 0.5]) * scale)]


class DistRewardMultiFall(GoalRewardMultiFall, DistRewardMixIn):
    pass


class NoRewardMultiFall(GoalRewardFall):  # sic: Fall's map and goal (quirk Q7)
    reward = _reward_zero


class GoalRewardLongCorridor(GoalRewardUMaze):
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=2.0, point=4.0, swimmer=2.0)
    create_maze = _maze(
        "BBBBBBBBB/BRB...B.B/B.B.B.B.B/B.B.B.B.B/B...B...B/BBBBBBBBB"
    )

    def __init__(self, scale: float, goal: Tuple[float, float] = (1.0, 3.0)) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array(goal) * scale)]


class DistRewardLongCorridor(GoalRewardLongCorridor, DistRewardMixIn):
    pass


class GoalRewardBlockMaze(GoalRewardUMaze):
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=8.0, point=4.0, swimmer=None)
    OBSERVE_BLOCKS: bool = True
    create_maze = _maze("BBBBB/BR..B/BBBMB/B...B/B...B/BBBBB")

    def __init__(self, scale: float) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array([0.0, 3.0 * scale]))]


class DistRewardBlockMaze(GoalRewardBlockMaze, DistRewardMixIn):
    pass


# ---------------------------------------------------------------------------
# Room family: possibly several goals, R1
# ---------------------------------------------------------------------------
class _RoomTask(MazeTask):
    REWARD_THRESHOLD: float = 0.9
    PENALTY: float = -0.0001
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=4.0, point=4.0, swimmer=4.0)
    reward = _reward_scaled


def _sub_goal(xy, scale, **kw) -> MazeGoal:
    return MazeGoal(np.array(xy) * scale, reward_scale=0.5, rgb=GREEN, **kw)


class GoalReward2Rooms(_RoomTask):
    create_maze = _maze(
        "BBBBBBBB/B...B..B/B...B..B/B.R.B..B/B...B..B/B......B/BBBBBBBB"
    )

    def __init__(self, scale: float, goal: Tuple[int, int] = (4.0, -2.0)) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array(goal) * scale)]


class DistReward2Rooms(GoalReward2Rooms, DistRewardMixIn):
    pass


class SubGoal2Rooms(GoalReward2Rooms):
    def __init__(
        self,
        scale: float,
        primary_goal: Tuple[float, float] = (4.0, -2.0),
        subgoals: List[Tuple[float, float]] = [(1.0, -2.0), (-1.0, 2.0)],
    ) -> None:
        super().__init__(scale, primary_goal)
        self.goals.extend(_sub_goal(xy, scale) for xy in subgoals)


class GoalReward4Rooms(_RoomTask):
    create_maze = _maze(
        "BBBBBBBBB/B...B...B/B.......B/B...B...B/BB.BBB.BB/"
        "B...B...B/B.......B/BR..B...B/BBBBBBBBB"
    )

    def __init__(self, scale: float) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array([6.0 * scale, -6.0 * scale]))]


class DistReward4Rooms(GoalReward4Rooms, DistRewardMixIn):
    pass


class SubGoal4Rooms(GoalReward4Rooms):
    def __init__(self, scale: float) -> None:
        super().__init__(scale)
        self.goals += [
            MazeGoal(np.array([0.0 * scale, -6.0 * scale]), 0.5, GREEN),
            MazeGoal(np.array([6.0 * scale, 0.0 * scale]), 0.5, GREEN),
        ]


class GoalRewardTRoom(_RoomTask):
    create_maze = _maze("BBBBBBB/B..B..B/B..B..B/B.BBB.B/B..R..B/BBBBBBB")

    def __init__(self, scale: float, goal: Tuple[float, float] = (2.0, -3.0)) -> None:
        super().__init__(scale)
        self.goals = [MazeGoal(np.array(goal) * scale)]


class DistRewardTRoom(GoalRewardTRoom, DistRewardMixIn):
    pass


class SubGoalTRoom(GoalRewardTRoom):
    def __init__(
        self,
        scale: float,
        primary_goal: Tuple[float, float] = (2.0, -3.0),
        subgoal: Tuple[float, float] = (-2.0, -3.0),
    ) -> None:
        super().__init__(scale, primary_goal)
        self.goals.append(_sub_goal(subgoal, scale))


class NoRewardCorridor(MazeTask):
    REWARD_THRESHOLD: float = 0.0
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=4.0, point=4.0, swimmer=1.0)
    reward = _reward_zero
    create_maze = _maze(
        "BBBBBBBBB/B..B....B/B..B....B/B.....BBB/B...R...B/"
        "BBB.....B/B....B..B/B....B..B/BBBBBBBBB"
    )


class GoalRewardCorridor(NoRewardCorridor):
    REWARD_THRESHOLD: float = 0.9
    PENALTY: float = -0.0001
    reward = _reward_scaled

    def __init__(self, scale: float, goal: Tuple[float, float] = (3.0, -3.0)) -> None:
        super().__init__(scale)
        self.goals.append(MazeGoal(np.array(goal) * scale))


class DistRewardCorridor(GoalRewardCorridor, DistRewardMixIn):
    pass


# ---------------------------------------------------------------------------
# Object families: goal reached by a block / ball (obs[3:6]); R2, R3, T1
# ---------------------------------------------------------------------------
class _ObjectTask(MazeTask):
    REWARD_THRESHOLD: float = 0.9
    PENALTY: float = -0.0001
    GOAL_SIZE: float = 0.3
    reward = _reward_scaled_object
    termination = _termination_object


class GoalRewardBlockCarry(_ObjectTask):
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=2.0, point=3.0, swimmer=None)
    OBSERVE_BLOCKS: bool = True
    create_maze = _maze("BBBBB/B...B/BRM.B/B...B/BBBBB")

    def __init__(self, scale: float, goal: Tuple[float, float] = (2.0, 0.0)) -> None:
        super().__init__(scale)
        self.goals.append(
            MazeGoal(
                np.array(goal) * scale,
                threshold=self.GOAL_SIZE + 0.5,
                custom_size=self.GOAL_SIZE,
            )
        )


class DistRewardBlockCarry(GoalRewardBlockCarry):
    reward = _reward_dist_object


class NoRewardBlockCarry(GoalRewardBlockCarry):
    reward = _reward_zero


class GoalRewardBilliard(_ObjectTask):
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=None, point=3.0, swimmer=None)
    OBSERVE_BALLS: bool = True
    create_maze = _maze(
        "BBBBBBB/B.....B/B.....B/B..O..B/B..R..B/B.....B/BBBBBBB"
    )

    def __init__(self, scale: float, goal: Tuple[float, float] = (2.0, -3.0)) -> None:
        super().__init__(scale)
        self.goals.append(self._ball_goal(np.array(goal) * scale))

    def _threshold(self) -> float:
        return self.OBJECT_BALL_SIZE + self.GOAL_SIZE

    def _ball_goal(self, pos: np.ndarray, **kw) -> MazeGoal:
        return MazeGoal(
            pos, threshold=self._threshold(), custom_size=self.GOAL_SIZE, **kw
        )


class DistRewardBilliard(GoalRewardBilliard):
    reward = _reward_dist_object


class NoRewardBilliard(GoalRewardBilliard):
    reward = _reward_zero

    def __init__(self, scale: float) -> None:
        MazeTask.__init__(self, scale)  # no goals at all: never terminates


class SubGoalBilliard(GoalRewardBilliard):
    def __init__(
        self,
        scale: float,
        primary_goal: Tuple[float, float] = (2.0, -3.0),
        subgoals: List[Tuple[float, float]] = [(-2.0, -3.0), (-2.0, 1.0), (2.0, 1.0)],
    ) -> None:
        super().__init__(scale, primary_goal)
        for xy in subgoals:
            self.goals.append(
                self._ball_goal(np.array(xy) * scale, reward_scale=0.5, rgb=GREEN)
            )


class BanditBilliard(SubGoalBilliard):
    create_maze = _maze(
        "BBBBBBB/B..BB.B/B.....B/BRO.BBB/B.....B/B.....B/BBBBBBB"
    )

    def __init__(
        self,
        scale: float,
        primary_goal: Tuple[float, float] = (4.0, -2.0),
        subgoals: List[Tuple[float, float]] = [(4.0, 2.0)],
    ) -> None:
        super().__init__(scale, primary_goal, subgoals)


class GoalRewardSmallBilliard(GoalRewardBilliard):
    MAZE_SIZE_SCALING: Scaling = Scaling(ant=2.0, point=4.0, swimmer=None)
    OBJECT_BALL_SIZE: float = 0.4
    GOAL_SIZE: float = 0.2
    create_maze = _maze("BBBBB/B...B/B.O.B/B.R.B/BBBBB")

    def __init__(self, scale: float, goal: Tuple[float, float] = (-1.0, -2.0)) -> None:
        super().__init__(scale, goal)


class DistRewardSmallBilliard(GoalRewardSmallBilliard, DistRewardMixIn):
    pass


class NoRewardSmallBilliard(GoalRewardSmallBilliard):
    reward = _reward_zero


class TaskRegistry:
    REGISTRY: Dict[str, List[Type[MazeTask]]] = {
        maze_id: [globals()[f"{prefix}{maze_id}"] for prefix in prefixes]
        for maze_id, prefixes in (
            ("SimpleRoom", ("DistReward", "GoalReward")),
            ("SquareRoom", ("DistReward", "GoalReward", "NoReward")),
            ("UMaze", ("DistReward", "GoalReward")),
            ("Push", ("DistReward", "GoalReward")),
            ("MultiPush", ("DistReward", "GoalReward", "NoReward")),
            ("MultiPushSmall", ("DistReward", "GoalReward", "NoReward")),
            ("PushMaze", ("DistReward", "GoalReward", "NoReward")),
            ("Fall", ("DistReward", "GoalReward")),
            ("MultiFall", ("DistReward", "GoalReward", "NoReward")),
            ("2Rooms", ("DistReward", "GoalReward", "SubGoal")),
            ("4Rooms", ("DistReward", "GoalReward", "SubGoal")),
            ("TRoom", ("DistReward", "GoalReward", "SubGoal")),
            ("BlockMaze", ("DistReward", "GoalReward")),
            ("Corridor", ("DistReward", "GoalReward", "NoReward")),
            ("LongCorridor", ("DistReward", "GoalReward")),
            ("BlockCarry", ("DistReward", "GoalReward", "NoReward")),
            ("Billiard", ("DistReward", "GoalReward", "SubGoal", "Bandit", "NoReward")),
            ("SmallBilliard", ("DistReward", "GoalReward", "NoReward")),
        )
    }

    @staticmethod
    def keys() -> List[str]:
        return list(TaskRegistry.REGISTRY.keys())

    @staticmethod
    def tasks(key: str) -> List[Type[MazeTask]]:
        return TaskRegistry.REGISTRY[key]
