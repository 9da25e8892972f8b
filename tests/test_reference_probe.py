"""bench.py's probe for the REAL reference (gym + MuJoCo + baseline/_ref) and the MuJoCo cross-check plumbing.

MuJoCo is not installable in the build image, so the probe must say "unavailable" here and the CPU legs fall back to the
fp64 restatement (kind "port"). That the probe would FIND and TIME a working reference is shown with a fake one: a
directory holding stand-in `gym`, `mujoco_py` and `mujoco_maze` modules, passed through MMZ_REF_PATH. The replay half of
tools/mujoco_crosscheck.py is exercised with a dump the oracle wrote itself (errors must be ~0: it tests the plumbing,
not MuJoCo).
"""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOOP = os.path.join(ROOT, "tools", "reference_loop.py")

FAKE_GYM = '''
import numpy as np
__version__ = "0.20-fake"
class _Space:
    low = np.array([-1.0, -0.25]); high = np.array([1.0, 0.25])
    def sample(self): return np.random.uniform(self.low, self.high)
    def seed(self, s): np.random.seed(s)
class _Data:
    def __init__(self): self.qpos = np.zeros(3); self.qvel = np.zeros(3); self.ncon = 0
class _Sim:
    def __init__(self): self.data = _Data()
class _Agent:
    def __init__(self): self.sim = _Sim()
class _Env:
    def __init__(self, id): self.id, self.action_space, self.wrapped_env, self.t = id, _Space(), _Agent(), 0
    @property
    def unwrapped(self): return self
    def seed(self, s): pass
    def reset(self): self.t = 0; self.wrapped_env.sim.data.qpos[:] = 0; return np.zeros(7)
    def step(self, a):
        self.t += 1
        d = self.wrapped_env.sim.data
        d.qpos[:2] += 0.01 * np.asarray(a); d.qvel[:2] = a
        return np.concatenate([d.qpos, d.qvel, [self.t * 0.001]]), -0.0001, self.t >= 1000, {}
def make(id): return _Env(id)
'''


def fake_reference(tmp_path):
    for pkg, body in (("gym", FAKE_GYM), ("mujoco_py", '__version__ = "2.0.2.13-fake"\n'), ("mujoco_maze", "")):
        d = tmp_path / pkg
        d.mkdir()
        (d / "__init__.py").write_text(textwrap.dedent(body))
    return str(tmp_path)


def run_loop(args, env=None):
    r = subprocess.run([sys.executable, LOOP, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env={**os.environ, **(env or {})})
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_probe_reports_unavailable_in_this_image():
    p = run_loop(["--probe"])
    assert p["available"] is False and p["why"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mujoco_crosscheck.py")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "unavailable" in json.loads(r.stdout.strip().splitlines()[-1])


def test_probe_finds_and_times_a_working_reference(tmp_path):
    env = {"MMZ_REF_PATH": fake_reference(tmp_path)}
    p = run_loop(["--probe"], env)
    assert p["available"] is True and p["mujoco"].startswith("mujoco_py 2.0.2.13") and p["gym"] == "0.20-fake"
    t = run_loop(["--time", "PointUMaze-v0", "--steps", "200", "--warmup", "5", "--procs", "2"], env)
    assert t["procs"] == 2 and len(t["per_proc"]) == 2 and t["env_steps_per_sec"] > 0
    out = tmp_path / "dump.json"
    d = run_loop(["--dump", "PointUMaze-v0", "--n", "2", "--steps", "3", "--out", str(out)], env)
    assert d["episodes"] == 2
    rec = json.load(open(out))
    assert len(rec["episodes"][0]["steps"]) == 3 and len(rec["episodes"][0]["steps"][0]["qpos"]) == 3


def test_bench_cpu_legs_prefer_the_real_reference(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    assert bench.probe_reference()["available"] is False  # this image: the port is what gets timed
    monkeypatch.setenv("MMZ_REF_PATH", fake_reference(tmp_path))
    assert bench.probe_reference()["available"] is True
    monkeypatch.setattr(bench, "host_cores", lambda: 2)
    b = bench.cpu_baseline(None, "PointUMaze-v0", 0.05)  # the model is only needed by the port
    assert b["kind"] == "reference" and b["cores"] == 2 and b["value"] > 0 and "gym.make('PointUMaze-v0')" in b["sample"]


def test_crosscheck_replay_of_an_oracle_written_dump(tmp_path, oracle_lib):
    """compare() must reproduce a dump exactly when the 'reference' that wrote it is the oracle itself."""
    import mujoco_compare as mc
    from conftest import make_model

    model = make_model("AntUMaze-v0")
    o = oracle_lib.OracleEnv(model)
    rng = np.random.default_rng(4)
    nq, nv, nu = int(model.nq), int(model.nv), int(model.nu)
    episodes = []
    for _ in range(2):
        q = np.asarray(model.qpos0, float)[:nq] + np.concatenate([rng.uniform(-0.1, 0.1, 3), np.zeros(4), rng.uniform(-0.1, 0.1, nq - 7)])
        v = 0.1 * rng.normal(size=nv)
        o.set_state(q, v, 0)
        ep = {"qpos0": q.tolist(), "qvel0": v.tolist(), "obs0": [], "steps": []}
        for k in range(3):
            a = rng.uniform(-30, 30, nu)
            obs, rew, bits, _ = o.step(a)
            qq, vv, _ = o.get_state()
            ep["steps"].append({"action": a.tolist(), "qpos": qq.tolist(), "qvel": vv.tolist(), "ncon": o.counts()["ncon"],
                                "obs": np.asarray(obs).tolist(), "reward": float(rew), "done": bool(bits & 1)})
        episodes.append(ep)
    path = tmp_path / "AntUMaze-v0.json"
    json.dump({"env_id": "AntUMaze-v0", "n": 2, "steps": 3, "seed": 0, "episodes": episodes}, open(path, "w"))
    st = mc.compare(str(path))
    assert st["steps"] == 6 and st["done_mismatch"] == 0 and st["ncon_mismatch"] == 0
    assert st["qpos"] < 1e-9 and st["qvel"] < 1e-7 and st["obs"] < 1e-7 and st["reward"] < 1e-9


def test_oracle_against_committed_mujoco_dumps(oracle_lib):
    """Consumes tests/golden/mujoco_crosscheck/*.json (written by tools/mujoco_crosscheck.py on a box with MuJoCo).
    None is committed - MuJoCo cannot be installed in the build image - so this SKIPS and the physics half of the oracle
    stays "parity unpinned" (DESIGN.md section 4). With dumps present the restatement has to match real MuJoCo on
    teacher-forced steps: positions 1e-4, velocities 1e-3 (relative to 1 + |x|), same `done`, same contact counts in
    at least 98 % of the steps (the capsule-box / box-box narrow phases are NOT MuJoCo's algorithms)."""
    import glob

    import pytest

    import mujoco_compare as mc

    dumps = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "mujoco_crosscheck", "*.json")))
    if not dumps:
        pytest.skip("no MuJoCo dump committed (tools/mujoco_crosscheck.py needs gym + mujoco-py): physics parity unpinned")
    for d in dumps:
        st = mc.compare(d)
        print(json.dumps(st))
        assert st["done_mismatch"] == 0, st
        assert st["qpos"] <= 1e-4 and st["qvel"] <= 1e-3, st
        assert st["ncon_mismatch"] <= 0.02 * st["steps"], st
