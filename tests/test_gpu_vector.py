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


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["PointUMaze-v0", "AntUMaze-v0", "SwimmerUMaze-v0"])
def test_step_k_equals_k_single_steps_bit_for_bit(env_id):
    """mmz_step_k (K launches replayed from one CUDA graph) against K mmz_step calls: same observations, rewards, done bits
    and final state, with the TimeLimit (and the in-kernel auto-reset) firing in the MIDDLE of the K-block."""
    import torch

    from conftest import make_model
    from mujoco_maze.backend import BatchedSim

    n, K = 96, 6
    model = make_model(env_id, num_envs=n)
    sims = [BatchedSim(model, n, auto_reset=True) for _ in range(2)]
    lo, hi = (torch.as_tensor(np.asarray(model.act_ctrlrange, np.float32)[: sims[0].nu, k], device="cuda") for k in (0, 1))
    g = torch.Generator(device="cuda").manual_seed(3)
    acts = lo + (hi - lo) * torch.rand((2 * K, n, sims[0].nu), device="cuda", generator=g)
    for sim in sims:
        sim.reset(seed=8)
        q, v, t = sim.get_state()
        t[:] = 1000 - 3                      # every episode ends at the third step of the first block
        sim.set_state(q, v, t)
    single = [tuple(x.clone() for x in sims[0].step(acts[k])) for k in range(2 * K)]
    out = None
    blocks = []
    for b in range(2):                        # the second call replays the recorded graph with the same buffers
        a = acts[b * K:(b + 1) * K].contiguous() if b == 0 else a.copy_(acts[b * K:(b + 1) * K])
        out = sims[1].step_k(a, out)
        blocks.append(tuple(x.clone() for x in out))
    for b in range(2):
        for k in range(K):
            o, r, d, i = single[b * K + k]
            assert torch.equal(blocks[b][0][k], o) and torch.equal(blocks[b][1][k], r)
            assert torch.equal(blocks[b][2][k], d) and torch.equal(blocks[b][3][k], i)
    assert int((blocks[0][2][2] & 1).sum()) == n and int((blocks[0][2][2] & 2).sum()) > 0   # truncated at step 3 ...
    assert int(blocks[0][2][3].sum()) == 0                                                  # ... and running again at step 4
    for x, y in zip(sims[0].get_state(), sims[1].get_state()):
        assert torch.equal(x, y)
    assert sims[1].launch_count - 2 >= 2 * K                                               # (reset + set_state refresh, then K per block)
    for sim in sims:
        sim.close()
