"""Generate reference-pinned golden vectors from the reference's importable Python half.

Runs ONLY in the build container (needs /root/reference). It imports the
UNMODIFIED reference modules `mujoco_maze.maze_env_utils` and
`mujoco_maze.maze_task` (numpy-only) through a stub package that bypasses
`mujoco_maze/__init__.py` (which needs gym/mujoco), and writes
tests/golden/reference_python_half.json. The committed JSON is what the tests
read; /root/reference is never touched at test time.

    python tests/golden/gen_reference_goldens.py
"""
import importlib
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/mujoco_maze"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_python_half.json")


def load_reference():
    for k in [k for k in sys.modules if k == "mujoco_maze" or k.startswith("mujoco_maze.")]:
        del sys.modules[k]
    pkg = types.ModuleType("mujoco_maze")
    pkg.__path__ = [REF]
    sys.modules["mujoco_maze"] = pkg
    utils = importlib.import_module("mujoco_maze.maze_env_utils")
    task = importlib.import_module("mujoco_maze.maze_task")
    return utils, task


def main():
    utils, task = load_reference()
    rng = np.random.default_rng(20261017)
    out = {"tasks": {}, "line": {}, "detect": {}}

    # --- Line known answers (tests/test_intersect.py:10-11, 23-26) plus random pairs
    cases = []
    for _ in range(64):
        p = rng.uniform(-5, 5, size=(4, 2))
        l1, l2 = utils.Line(p[0], p[1]), utils.Line(p[2], p[3])
        x = l1.intersect(l2)
        q = complex(*rng.uniform(-5, 5, size=2))
        refl = l1.reflection(q)
        cases.append(
            dict(
                pts=p.tolist(),
                hit=None if x is None else [x.real, x.imag],
                q=[q.real, q.imag],
                dist=l1.distance(q),
                refl=[refl.real, refl.imag],
            )
        )
    out["line"]["random"] = cases

    # --- every registered task class
    for maze_id in task.TaskRegistry.keys():
        for v, cls in enumerate(task.TaskRegistry.tasks(maze_id)):
            sc = cls.MAZE_SIZE_SCALING
            rec = dict(
                cls=cls.__name__,
                scaling=[sc.ant, sc.point, sc.swimmer],
                inner=cls.INNER_REWARD_SCALING,
                observe_blocks=cls.OBSERVE_BLOCKS,
                observe_balls=cls.OBSERVE_BALLS,
                ball_size=cls.OBJECT_BALL_SIZE,
                penalty=cls.PENALTY,
                reward_threshold=cls.REWARD_THRESHOLD,
                map=[[c.value for c in row] for row in cls.create_maze()],
                per_scale={},
            )
            for scale in sorted({s for s in sc if s is not None}):
                t = cls(scale)
                goals = [
                    dict(pos=np.asarray(g.pos, float).tolist(), dim=int(g.dim), w=float(g.reward_scale),
                         thr=float(g.threshold), custom_size=g.custom_size)
                    for g in t.goals
                ]
                # random obs of length 12; half of them sampled near a goal
                obs_list, rew, term = [], [], []
                for k in range(12):
                    obs = rng.uniform(-3 * scale, 3 * scale, size=12)
                    if t.goals and k % 2 == 0:
                        g = t.goals[k // 2 % len(t.goals)]
                        jitter = rng.normal(size=g.dim) * g.threshold * 0.6
                        where = rng.integers(0, 2)  # agent slot or object slot
                        base = 3 * int(where)
                        obs[base: base + g.dim] = np.asarray(g.pos) + jitter
                    obs_list.append(obs.tolist())
                    rew.append(float(t.reward(obs)))
                    term.append(bool(t.termination(obs)))
                rec["per_scale"][repr(float(scale))] = dict(goals=goals, obs=obs_list, reward=rew, term=term)
            out["tasks"][f"{maze_id}-v{v}"] = rec

    # --- wall segments + detect() for the manual-collision agent (Point, r=0.4) and ball radii
    E = utils.MazeCell
    for maze_id in task.TaskRegistry.keys():
        cls = task.TaskRegistry.tasks(maze_id)[0]
        scale = cls.MAZE_SIZE_SCALING.point
        if scale is None:
            continue
        structure = cls.create_maze()
        tx = ty = None
        for i, row in enumerate(structure):
            for j, c in enumerate(row):
                if c == E.ROBOT and tx is None:
                    tx, ty = j * scale, i * scale
        for radius in (0.4, cls.OBJECT_BALL_SIZE):
            det = utils.CollisionDetector(structure, scale, tx, ty, radius)
            segs = [[l.p1.real, l.p1.imag, l.p2.real, l.p2.imag] for l in det.lines]
            moves = []
            h, w = len(structure), len(structure[0])
            for k in range(80):
                old = np.array([rng.uniform(-tx - scale, (w - 1) * scale - tx + scale),
                                rng.uniform(-ty - scale, (h - 1) * scale - ty + scale)])
                step = rng.normal(size=2) * (0.1 if k % 4 else 0.6 * scale)
                if k % 40 == 39:
                    step = np.zeros(2)
                new = old + step
                try:
                    c = det.detect(old, new)
                except ZeroDivisionError:
                    continue  # exactly collinear: reference raises; kernel treats as no hit
                rec = dict(old=old.tolist(), new=new.tolist(), hit=c is not None)
                if c is not None:
                    pos = c.point + 0.8 * c.rest()
                    try:
                        second = det.detect(old, pos) is not None
                    except ZeroDivisionError:
                        continue
                    rec.update(point=c.point.tolist(), rest=c.rest().tolist(), pos=pos.tolist(), second=second)
                moves.append(rec)
            out["detect"][f"{maze_id}/r{radius}"] = dict(scale=scale, torso=[tx, ty], radius=radius,
                                                        segments=segs, moves=moves)
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
