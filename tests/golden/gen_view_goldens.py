"""Golden vectors for MazeEnv.get_top_down_view, produced by the reference's own method.

Runs ONLY in the build container (needs /root/reference). The reference's `maze_env.py` imports gym and its
MuJoCo binding at module level; neither is installed, so empty stand-in modules are put in `sys.modules` first.
`MazeEnv.get_top_down_view` (reference maze_env.py:262-349) itself is UNMODIFIED: it is called unbound on a plain
namespace that carries exactly the attributes the method reads (`_view`, `_xy_to_rowcol`, `_maze_structure`,
`_maze_size_scaling`, `_init_torso_x/y`, `movable_blocks`, `wrapped_env.get_body_com`), set up the way
`MazeEnv.__init__` sets them (maze_env.py:56-65, 90-95, 114, 155).

    python tests/golden/gen_view_goldens.py   ->  tests/golden/reference_top_down_view.json
"""
import importlib
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/mujoco_maze"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_top_down_view.json")


def load_reference_maze_env():
    for k in [k for k in sys.modules if k == "mujoco_maze" or k.startswith("mujoco_maze.") or k == "gym" or k.startswith("gym.")]:
        del sys.modules[k]

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod

    gym = stub("gym", Env=type("Env", (), {}))
    stub("gym.core", ObsType=object)
    gym.spaces = stub("gym.spaces", Space=object, Box=object)
    stub("gym.utils", EzPickle=type("EzPickle", (), {}))
    stub("gym.envs")
    stub("gym.envs.mujoco")
    stub("gym.envs.mujoco.mujoco_env", MujocoEnv=type("MujocoEnv", (), {}))
    pkg = types.ModuleType("mujoco_maze")
    pkg.__path__ = [REF]
    sys.modules["mujoco_maze"] = pkg
    return importlib.import_module("mujoco_maze.maze_env"), importlib.import_module("mujoco_maze.maze_task")


def main():
    maze_env, maze_task = load_reference_maze_env()
    rng = np.random.default_rng(20261018)
    out = {"cases": []}
    # (task class, agent kind, scaling): flat mazes, movable blocks, chasms, a 9x9 grid
    setups = [
        ("GoalRewardUMaze", "point", 4.0), ("GoalRewardUMaze", "ant", 8.0),
        ("GoalRewardPush", "point", 4.0), ("GoalRewardPush", "ant", 8.0),
        ("GoalRewardMultiPush", "point", 4.0),
        ("GoalRewardFall", "point", 4.0), ("GoalRewardFall", "ant", 8.0),
        ("GoalReward4Rooms", "point", 4.0), ("GoalReward4Rooms", "ant", 4.0),
    ]
    for cls_name, agent, scaling in setups:
        cls = getattr(maze_task, cls_name)
        structure = cls.create_maze()
        h, w = len(structure), len(structure[0])
        tx = ty = None
        for i in range(h):
            for j in range(w):
                if structure[i][j].is_robot() and tx is None:
                    tx, ty = j * scaling, i * scaling
        blocks0 = {f"movable_{i}_{j}": (j * scaling - tx, i * scaling - ty,
                                        structure[i][j].can_move_x(), structure[i][j].can_move_y())
                   for i in range(h) for j in range(w) if structure[i][j].can_move()}
        samples = []
        for k in range(24):
            # robot anywhere over the maze (and a little outside), blocks displaced from their cells
            robot = np.array([rng.uniform(-tx - 0.75 * scaling, (w - 1) * scaling - tx + 0.75 * scaling),
                              rng.uniform(-ty - 0.75 * scaling, (h - 1) * scaling - ty + 0.75 * scaling)])
            pos = {"torso": np.array([robot[0], robot[1], 0.5])}
            for name, (bx, by, move_x, move_y) in blocks0.items():
                d = rng.uniform(-0.8 * scaling, 0.8 * scaling, size=2)  # only along the block's slide joints
                pos[name] = np.array([bx + d[0] * move_x, by + d[1] * move_y, 0.0])
            env = types.SimpleNamespace(
                _view=np.zeros([5, 5, 3]),
                _xy_to_rowcol=lambda x, y, s=scaling: (2 + (y + s / 2) / s, 2 + (x + s / 2) / s),
                _maze_structure=structure, _maze_size_scaling=scaling, _init_torso_x=tx, _init_torso_y=ty,
                movable_blocks=list(blocks0.keys()),
                wrapped_env=types.SimpleNamespace(get_body_com=lambda name, pos=pos: pos[name]),
            )
            view = maze_env.MazeEnv.get_top_down_view(env)
            samples.append(dict(robot=robot.tolist(), blocks={n: pos[n][:2].tolist() for n in blocks0},
                                view=np.asarray(view, float).ravel().tolist()))
        out["cases"].append(dict(task=cls_name, agent=agent, scaling=scaling, samples=samples))
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
