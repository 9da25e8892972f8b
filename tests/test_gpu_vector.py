"""The gym.vector-style wrapper on the GPU: TimeLimit truncation and in-kernel auto-reset through the public API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vector_env_truncates_and_auto_resets():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA")
    import mujoco_maze  # noqa: F401
    from mujoco_maze.vector import VectorMazeEnv

    n = 96
    venv = VectorMazeEnv("PointUMaze-v1", n, seed=3)
    assert venv.observation_space.shape == (n, 7) and venv.action_space.shape == (n, 2)
    obs = venv.reset()
    assert obs.shape == (n, 7) and float(obs[:, -1].abs().max()) == 0.0  # t * 0.001 == 0 after reset
    # jump to the end of the episode instead of stepping 1000 times
    q, v, t = venv.get_state()
    venv.env.sim.set_state(q, v, torch.full((n,), 997, dtype=torch.int32, device=q.device))
    zeros = torch.zeros((n, 2), device=q.device)
    for k in range(2):
        obs, rew, done, info = venv.step(zeros)
        assert not bool(done.any()) and not bool(info["TimeLimit.truncated"].any())
        assert abs(float(obs[0, -1]) - (998 + k) * 0.001) < 1e-6
    obs, rew, done, info = venv.step(zeros)  # step 1000: truncated, re-initialised inside the launch
    assert bool(done.all()) and bool(info["TimeLimit.truncated"].all())
    assert float(obs[:, -1].abs().max()) == 0.0                      # first observation of the next episode
    assert float(obs[:, :3].abs().max()) <= 0.1 + 1e-6               # reset_model noise (point.py:72-75)
    assert bool((rew < 0).all())                                     # v1: the penalty of the step that ended (tests/test_envs.py:35-36)
    _, _, t = venv.get_state()
    assert int(t.max()) == 0
    obs, rew, done, info = venv.step(venv.sample_actions())
    assert not bool(done.any()) and abs(float(obs[0, -1]) - 0.001) < 1e-7
    # numpy flavour
    venv.close()
    venv = VectorMazeEnv("AntUMaze-v0", 40, seed=1, as_numpy=True)
    obs = venv.reset()
    obs, rew, done, info = venv.step(np.zeros((40, 8), dtype=np.float32))
    assert isinstance(obs, np.ndarray) and obs.shape == (40, 30) and rew.shape == (40,) and done.dtype == bool
    assert set(info) >= {"position", "reward_forward", "reward_ctrl", "TimeLimit.truncated"}
    venv.close()
