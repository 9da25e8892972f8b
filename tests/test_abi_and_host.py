"""CPU-side checks of the drop-in boundary (no GPU, no compute calls into libmmz):
  * libmmz.so loads and exports every function include/mmz.h declares, with the ABI version the binding expects;
  * error behaviour that does not need a device: NULL / bad arguments return negative codes and set mmz_last_error;
  * the product path fails loudly without CUDA (no CPU fallback) and never imports oracle/;
  * the host-side mirror of the reference's interface: registered ids, spaces, custom MazeTask recipe (README.md:79-127).
"""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "mujoco-maze_b200")


def _lib_path():
    p = os.path.join(PKG, "libmmz.so")
    if not os.path.exists(p):
        sys.path.insert(0, PKG)
        import build_native

        build_native.build(verbose=False)
    return p


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mmz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmz_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 16 and "mmz_step" in names and "mmz_step_host" in names
    lib = ctypes.CDLL(_lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"include/mmz.h declares {missing} but libmmz.so does not export them"
    # and the Python binding declares a signature for every one of them
    from mujoco_maze import backend

    assert sorted(backend.EXPORTED_SYMBOLS) == names
    lib2 = backend.load_library(_lib_path())
    assert lib2.mmz_abi_version() == backend.ABI_VERSION


def test_model_header_matches_python_layout():
    """include/mmz_model.h is generated from model_layout.py: sizes must agree with the blob the compiler emits."""
    from conftest import make_model

    hdr = open(os.path.join(ROOT, "include", "mmz_model.h")).read()
    nint = int(re.search(r"#define MMZ_MODEL_NINT (\d+)", hdr).group(1))
    nreal = int(re.search(r"#define MMZ_MODEL_NREAL (\d+)", hdr).group(1))
    model = make_model("AntUMaze-v0")
    assert len(model.blob(4)) == 4 * nint + 4 * nreal
    assert len(model.blob(8)) == 4 * nint + 8 * nreal


def test_bad_arguments_return_error_codes_without_a_device():
    from mujoco_maze import backend

    lib = backend.load_library(_lib_path())
    h = ctypes.c_void_p()
    assert lib.mmz_create(None, 0, 4, 0, 0, ctypes.byref(h)) == -1  # MMZ_ERR_INVALID
    assert b"null" in lib.mmz_last_error()
    blob = b"\0" * 16
    assert lib.mmz_create(blob, len(blob), 4, 0, 0, ctypes.byref(h)) == -2  # MMZ_ERR_MODEL: wrong size
    assert b"bytes" in lib.mmz_last_error()
    assert lib.mmz_step(None, None, None, None, None, None, None) == -1
    assert lib.mmz_dims(None, None, None, None, None, None) == -1
    assert lib.mmz_launch_count(None) == 0
    lib.mmz_destroy(None)  # no-op


def test_no_cpu_fallback_and_no_oracle_import():
    """Without CUDA the step path must raise; importing the package must not pull in oracle/."""
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import mujoco_maze\n"
        "from mujoco_maze import gym\n"
        "from mujoco_maze.backend import MmzError\n"
        "env = gym.make('PointUMaze-v0')\n"
        "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'product imported oracle/'\n"
        "import torch\n"
        "if torch.cuda.is_available():\n"
        "    print('HAS_CUDA')\n"
        "else:\n"
        "    try:\n"
        "        env.reset()\n"
        "    except MmzError as e:\n"
        "        print('RAISED', e)\n"
        "    else:\n"
        "        raise SystemExit('reset() ran without CUDA: there must be no CPU fallback')\n"
    ) % (PKG, ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "RAISED" in out.stdout or "HAS_CUDA" in out.stdout
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle/"
                assert "mmz_oracle" not in text, f"{f} references the oracle"


def test_registered_ids_spaces_and_custom_task():
    import mujoco_maze  # noqa: F401
    from mujoco_maze import gym
    from mujoco_maze.maze_env_utils import MazeCell
    from mujoco_maze.maze_task import MazeGoal, MazeTask
    from mujoco_maze.point import PointEnv

    ids = [s for s in gym.registry_ids()] if hasattr(gym, "registry_ids") else None
    env = gym.make("AntUMaze-v0").unwrapped
    assert env.observation_space.shape == (30,) and env.action_space.shape == (8,)
    assert np.allclose(env.action_space.high, 30.0)
    p = gym.make("PointUMaze-v1").unwrapped
    assert p.observation_space.shape == (7,) and np.allclose(p.action_space.high, [1.0, 0.25])
    specs = gym.registry.env_specs if hasattr(gym, "registry") and hasattr(gym.registry, "env_specs") else None
    if specs is not None:  # the in-repo shim: the reference registers 48 Point + 45 Ant + 26 Swimmer + 26 Reacher ids
        assert len([k for k in specs if not k.endswith("-vtest")]) == 145
        assert sum(k.startswith("Reacher") for k in specs) == 26 and sum(k.startswith("Ant") for k in specs) == 45
    r = gym.make("ReacherUMaze-v0").unwrapped
    assert r.observation_space.shape == (9,) and r.action_space.shape == (1,)  # reference tests/test_envs.py:63-64

    # README.md:79-127: a user-defined task registers and compiles (host-side reward falls back to Python)
    class GoalRewardEMaze(MazeTask):
        REWARD_THRESHOLD = 0.9
        PENALTY = -0.0001

        def __init__(self, scale):
            super().__init__(scale)
            self.goals = [MazeGoal(np.array([0.0, 4.0]) * scale)]

        def reward(self, obs):
            return 1.0 if self.termination(obs) else self.PENALTY

        @staticmethod
        def create_maze():
            E, B, R = MazeCell.EMPTY, MazeCell.BLOCK, MazeCell.ROBOT
            return [[B, B, B, B, B], [B, R, E, E, B], [B, B, B, E, B], [B, E, E, E, B], [B, B, B, B, B]]

    gym.register(id="PointEMaze-vtest", entry_point="mujoco_maze.maze_env:MazeEnv",
                 kwargs=dict(model_cls=PointEnv, maze_task=GoalRewardEMaze, maze_size_scaling=4.0, inner_reward_scaling=0.0))
    e = gym.make("PointEMaze-vtest").unwrapped
    assert e.model.nseg > 0 and e.observation_space.shape == (7,)


def test_every_registered_id_compiles_with_the_reference_shapes():
    """reference tests/test_envs.py: every id makes; obs (30,) Ant / (7,) or (10,) Point / (11,) Swimmer / (9,) Reacher
    unless the task observes blocks or balls; Point*-v2 sub-goal tasks have several goals."""
    import mujoco_maze  # noqa: F401
    from mujoco_maze import gym

    specs = getattr(getattr(gym, "registry", None), "env_specs", None)
    if specs is None:
        pytest.skip("real gym installed: its registry also holds foreign ids")
    base = {"Ant": 30, "Point": 7, "Swimmer": 11, "Reacher": 9}
    n = 0
    for env_id in sorted(k for k in specs if not k.endswith("-vtest")):  # (another test registers a custom id)
        env = gym.make(env_id).unwrapped
        prefix = next(p for p in base if env_id.startswith(p))
        extra = 3 * (len(env.movable_blocks) if env._observe_blocks else 0) + 3 * (len(env.object_balls) if env._observe_balls else 0)
        want = base[prefix] + extra
        if prefix in ("Swimmer", "Reacher"):  # their _get_obs takes ALL of qpos / qvel, block joints included (swimmer.py:49-53)
            want = int(env.model.nq) + int(env.model.nv) + 1 + extra
        assert env.observation_space.shape == (want,), env_id
        assert env.has_extended_obs == (extra > 0 or env._top_down_view)
        blob = env.model.blob(4)
        assert len(blob) > 8000
        n += 1
    assert n == 145
    for env_id in ("Point2Rooms-v2", "Point4Rooms-v2", "PointBilliard-v2"):  # reference tests/test_envs.py:39-50
        assert len(gym.make(env_id).unwrapped._task.goals) > 1


def test_ant_body_masses_follow_the_mujoco_2_capsule_convention():
    """The reference runs on MuJoCo 2.0 (mujoco-py 2.0.2.13, poetry.lock:145-146), whose compiler takes a capsule's volume
    as pi r^2 L + pi r^3; the compiled model must reproduce the body masses that MuJoCo 2.0 reports for this geometry
    and density (gym Ant-v2 `model.body_mass`: torso 0.32725, leg links 0.036477, ankle links 0.064911 - SURVEY.md
    appendix A.1). `legacy_capsule_volume=True` is the default of compile_maze_model and what every parity run uses;
    with the exact 4/3 pi r^3 caps of later MuJoCo releases the links weigh 0.039158 / 0.067592."""
    from mujoco_maze import model_compiler as mc

    assets = os.path.join(ROOT, "mujoco-maze_b200", "mujoco_maze", "assets")
    sc = mc.parse_mjcf(os.path.join(assets, "ant.xml"))
    legacy = mc.flatten(sc, True)["body_mass"]
    np.testing.assert_allclose(legacy[0], 0.32725, rtol=2e-5)
    np.testing.assert_allclose(legacy[[1, 2, 4, 5, 7, 8, 10, 11]], 0.036477, rtol=2e-5)
    np.testing.assert_allclose(legacy[[3, 6, 9, 12]], 0.064911, rtol=2e-5)
    exact = mc.flatten(sc, False)["body_mass"]
    np.testing.assert_allclose(exact[[1, 3]], [0.039158, 0.067592], rtol=2e-5)
    # the model every env id compiles to: welded links folded into their parents, same total mass, legacy convention
    from conftest import make_model

    m = make_model("AntUMaze-v0")
    assert int(m.nbody) == 9
    np.testing.assert_allclose(np.asarray(m.body_mass)[:9].sum(), legacy.sum(), rtol=1e-6)
    np.testing.assert_allclose(np.asarray(m.body_mass)[0], legacy[0] + 4 * 0.036477, rtol=2e-5)
    assert bool(m.meta.get("legacy_capsule_volume", True)) is True


def test_package_make_builds_batched_envs_without_a_scalar_timelimit(monkeypatch):
    """With the real gym installed `gym.make` would wrap a batched environment in gym's scalar TimeLimit (whose `not done`
    fails on tensors). `mujoco_maze.make` builds batched environments from the registered spec directly; simulated here by a
    `gym.make` that always wraps."""
    import mujoco_maze
    from mujoco_maze import gym

    def wrapping_make(env_id, **kw):
        return gym.wrappers.TimeLimit(gym.spec(env_id).make(**kw), 1000)

    monkeypatch.setattr(mujoco_maze.gym, "make", wrapping_make)
    scalar = mujoco_maze.make("PointUMaze-v0")
    assert type(scalar).__name__ == "TimeLimit"                       # reference-shaped env: as upstream
    batched = mujoco_maze.make("PointUMaze-v0", num_envs=8)
    assert type(batched).__name__ == "MazeEnv" and batched.is_batched and batched.spec.id == "PointUMaze-v0"
    assert int(batched.model.max_episode_steps) == 1000              # counted inside the kernel instead
