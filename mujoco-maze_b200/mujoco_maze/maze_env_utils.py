"""Maze geometry: cell kinds, wall segments and the segment-intersection clamp.

Host-side (cold path) mirror of the reference's geometry helpers. The batched
hot path never calls into this module: `CollisionDetector.segments()` is
evaluated once at construction and baked into the model blob; the clamp itself
runs inside the CUDA step kernel (csrc/mmz_kernels.cu, `segment_clamp`).

Public surface follows the reference so custom tasks keep working:
  MazeCell                      maze_env_utils.py:19-81
  Line                          maze_env_utils.py:84-128
  Collision                     maze_env_utils.py:131-142
  CollisionDetector             maze_env_utils.py:145-206
"""

from enum import Enum
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

Point = complex


class MazeCell(Enum):
    ROBOT = -1
    EMPTY = 0
    BLOCK = 1
    CHASM = 2
    OBJECT_BALL = 3
    XY_BLOCK = 14
    XZ_BLOCK = 15
    YZ_BLOCK = 16
    XYZ_BLOCK = 17
    XY_HALF_BLOCK = 18
    SPIN = 19

    # --- static kinds -----------------------------------------------------
    def is_block(self) -> bool:
        return self is MazeCell.BLOCK

    def is_chasm(self) -> bool:
        return self is MazeCell.CHASM

    def is_object_ball(self) -> bool:
        return self is MazeCell.OBJECT_BALL

    def is_robot(self) -> bool:
        return self is MazeCell.ROBOT

    def is_empty(self) -> bool:
        return self.value <= 0  # ROBOT or EMPTY

    def is_wall_or_chasm(self) -> bool:
        return self.value in (1, 2)

    # --- movable kinds: (x, y, z, spin, half) capability table --------------
    def _caps(self) -> Tuple[bool, bool, bool, bool, bool]:
        return _MOVE_CAPS.get(self.value, (False,) * 5)

    def can_move_x(self) -> bool:
        return self._caps()[0]

    def can_move_y(self) -> bool:
        return self._caps()[1]

    def can_move_z(self) -> bool:
        return self._caps()[2]

    def can_spin(self) -> bool:
        return self._caps()[3]

    def is_half_block(self) -> bool:
        return self._caps()[4]

    def can_move(self) -> bool:
        return any(self._caps()[:3])


#                   x      y      z      spin   half
_MOVE_CAPS = {
    14: (True, True, False, False, False),  # XY_BLOCK
    15: (True, False, True, False, False),  # XZ_BLOCK
    16: (False, True, True, False, False),  # YZ_BLOCK
    17: (True, True, True, False, False),  # XYZ_BLOCK
    18: (True, True, False, False, True),  # XY_HALF_BLOCK
    19: (True, True, False, True, False),  # SPIN
}


def _as_point(p: Union[Sequence[float], Point]) -> Point:
    return p if isinstance(p, complex) else complex(float(p[0]), float(p[1]))


def _cross(a: Point, b: Point) -> float:
    """z component of a x b for 2-D vectors stored as complex numbers."""
    return a.real * b.imag - a.imag * b.real


def _dot(a: Point, b: Point) -> float:
    return a.real * b.real + a.imag * b.imag


class Line:
    """A 2-D segment p1 -> p2 (points kept as Python complex, as upstream)."""

    def __init__(self, p1, p2) -> None:
        self.p1 = _as_point(p1)
        self.p2 = _as_point(p2)
        self.v1 = self.p2 - self.p1
        self.conj_v1 = self.v1.conjugate()
        self.norm = abs(self.v1)

    def _intersect(self, other: "Line") -> bool:
        # other's two end points are on opposite sides of (or on) our carrier line
        side_a = _cross(self.v1, other.p1 - self.p1)
        side_b = _cross(self.v1, other.p2 - self.p1)
        return side_a * side_b <= 0.0

    def _projection(self, p: Point) -> Point:
        t = _dot(p - self.p1, self.v1) / _dot(self.v1, self.v1)
        return self.p1 + t * self.v1

    def reflection(self, p: Point) -> Point:
        foot = self._projection(p)
        return foot + (foot - p)

    def distance(self, p: Point) -> float:
        return abs(p - self._projection(p))

    def _cross_point(self, other: "Line") -> Point:
        d = other.p2 - other.p1
        denom = _cross(self.v1, d)
        numer = _cross(self.v1, self.p2 - other.p1)
        return other.p1 + (numer / denom) * d  # ZeroDivisionError when parallel (as upstream)

    def intersect(self, other: "Line") -> Optional[Point]:
        if self._intersect(other) and other._intersect(self):
            return self._cross_point(other)
        return None

    def __repr__(self) -> str:
        return (
            f"Line(({self.p1.real}, {self.p1.imag}) -> "
            f"({self.p2.real}, {self.p2.imag}))"
        )


class Collision:
    def __init__(self, point: Point, reflection: Point) -> None:
        self._point = point
        self._reflection = reflection

    @property
    def point(self) -> np.ndarray:
        return np.array([self._point.real, self._point.imag])

    def rest(self) -> np.ndarray:
        d = self._reflection - self._point
        return np.array([d.real, d.imag])


class CollisionDetector:
    """Wall faces pushed out by `radius`; `detect` finds the nearest crossing.

    Face emission order (cells row-major, neighbours W, N, E, S) matches
    maze_env_utils.py:151-184 because ties in `detect` resolve to the first
    segment and the kernel reproduces that order.
    """

    EPS: float = 0.05
    NEIGHBORS: List[Tuple[int, int]] = [[0, -1], [-1, 0], [0, 1], [1, 0]]

    def __init__(self, structure, size_scaling, torso_x, torso_y, radius) -> None:
        rows, cols = len(structure), len(structure[0])
        reach = 0.5 * size_scaling + radius
        self.lines: List[Line] = []

        def walkable(i: int, j: int) -> bool:
            return 0 <= i < rows and 0 <= j < cols and structure[i][j].is_empty()

        for i in range(rows):
            for j in range(cols):
                if not structure[i][j].is_block():
                    continue
                cx = j * size_scaling - torso_x
                cy = i * size_scaling - torso_y
                lo_x, hi_x, lo_y, hi_y = cx - reach, cx + reach, cy - reach, cy + reach
                for dx, dy in self.NEIGHBORS:
                    if not walkable(i + dy, j + dx):
                        continue
                    if dx != 0:  # vertical face on the west/east side
                        x = hi_x if dx > 0 else lo_x
                        seg = ((x, lo_y), (x, hi_y))
                    else:  # horizontal face on the north/south side
                        y = hi_y if dy > 0 else lo_y
                        seg = ((lo_x, y), (hi_x, y))
                    self.lines.append(Line(*seg))

    def segments(self) -> np.ndarray:
        """[L, 4] float64 array (x1, y1, x2, y2) — the constants the kernel stages."""
        out = np.zeros((len(self.lines), 4))
        for k, ln in enumerate(self.lines):
            out[k] = (ln.p1.real, ln.p1.imag, ln.p2.real, ln.p2.imag)
        return out

    def detect(self, old_pos, new_pos) -> Optional[Collision]:
        move = Line(old_pos, new_pos)
        if move.norm <= 1e-8:
            return None
        best, best_dist = None, None
        for wall in self.lines:
            hit = wall.intersect(move)
            if hit is None:
                continue
            d = abs(hit - move.p1)
            if best is None or d < best_dist:
                best, best_dist = Collision(hit, wall.reflection(move.p2)), d
        return best
