"""The Python side of MazeEnv.step on the GPU: forward_reward_fn variants (reference ant.py:18-23, 44-53), user-defined
task rules (README.md:79-127), the aliasing contract of step(), and argument validation in front of the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA")
    return torch


def _rollout(env, torch, steps=3, seed=0):
    n = env.unwrapped.num_envs
    g = torch.Generator(device="cuda").manual_seed(seed)
    outs = []
    for _ in range(steps):
        a = 60 * torch.rand((n, 8), device="cuda", generator=g) - 30
        outs.append((a, *env.step(a)))
    return outs


def test_forward_reward_vabs_is_fused_and_matches_the_oracle(torch_cuda, oracle_lib):
    from mujoco_maze import gym
    from mujoco_maze.ant import forward_reward_vabs

    torch = torch_cuda
    n = 64
    env = gym.make("AntUMaze-v0", num_envs=n, device="cuda:0", forward_reward_fn=forward_reward_vabs)
    assert int(env.unwrapped.model.forward_reward_kind) == 1
    env.reset(seed=4)
    sim = env.unwrapped.sim
    q, v, _ = (x.cpu().numpy().astype(np.float64) for x in sim.get_state())
    (a, obs, rew, done, info), = _rollout(env, torch, steps=1)
    dt = 0.02 * 5
    vel = (info["position"].cpu().numpy() - q[:, :2]) / dt
    np.testing.assert_allclose(info["reward_forward"].cpu().numpy(), np.abs(vel).sum(1), rtol=2e-4, atol=2e-5)
    o = oracle_lib.OracleEnv(env.unwrapped.model)
    o.L.ora_set_warmstart(o.h, 1)
    for i in range(0, n, 4):
        o.set_state(q[i], v[i], 0)
        _, rr, _, ri = o.step(a[i].cpu().numpy().astype(np.float64))
        assert abs(float(rew[i]) - rr) <= 1e-5 + 2e-4 * abs(rr), (float(rew[i]), rr)
        assert abs(float(info["reward_forward"][i]) - ri[2]) <= 1e-5 + 2e-4 * abs(ri[2])


def test_any_forward_reward_callable_runs_on_the_host(torch_cuda):
    """reference ant.py:44-53 accepts any callable: here 'velocity along +x'. Same physics as the default env; the reward
    differs by exactly weight * (fn(v) - |v|) and info['reward_forward'] is fn(v)."""
    from mujoco_maze import gym

    torch = torch_cuda
    n = 48
    plain = gym.make("AntUMaze-v0", num_envs=n, device="cuda:0")
    custom = gym.make("AntUMaze-v0", num_envs=n, device="cuda:0", forward_reward_fn=lambda v: v[0], forward_reward_weight=2.0)
    assert int(custom.unwrapped.model.forward_reward_kind) == 2
    o0, o1 = plain.reset(seed=9), custom.reset(seed=9)
    assert torch.equal(o0, o1)
    xy = o0[:, :2].clone()
    for (a, obs, rew, done, info), (_, cobs, crew, cdone, cinfo) in zip(_rollout(plain, torch), _rollout(custom, torch)):
        assert torch.equal(obs, cobs)
        vel = (info["position"] - xy) / 0.1
        want_fwd = vel[:, 0]
        torch.testing.assert_close(cinfo["reward_forward"], want_fwd, rtol=1e-5, atol=1e-6)
        scale = float(plain.unwrapped._inner_reward_scaling)
        want = rew - scale * info["reward_forward"] + scale * 2.0 * want_fwd
        torch.testing.assert_close(crew, want, rtol=1e-4, atol=1e-5)
        xy = obs[:, :2].clone()
    with pytest.raises(ValueError):
        gym.make("AntUMaze-v0", num_envs=n, device="cuda:0", auto_reset=True, forward_reward_fn=lambda v: v[0])


def test_custom_task_rules_scalar_loop_and_batch_protocol(torch_cuda):
    """README custom-task recipe: a MazeTask subclass with its own reward / termination. The scalar methods are called per
    environment; `reward_batch` / `termination_batch` keep a large batch on the device. Both must agree."""
    from mujoco_maze import maze_task as mt
    from mujoco_maze.maze_env import MazeEnv
    from mujoco_maze.point import PointEnv

    torch = torch_cuda

    class Scalar(mt.GoalRewardUMaze):
        def reward(self, obs):
            return -float(np.abs(obs[:2] - self.goals[0].pos).sum())

        def termination(self, obs):
            return bool(obs[0] > 0.05)

    class Batch(Scalar):
        def reward_batch(self, obs):
            goal = torch.as_tensor(self.goals[0].pos, dtype=obs.dtype, device=obs.device)
            return -(obs[:, :2] - goal).abs().sum(1)

        def termination_batch(self, obs):
            return obs[:, 0] > 0.05

    n = 80
    envs = [MazeEnv(PointEnv, task, num_envs=n, device="cuda:0", maze_size_scaling=4.0) for task in (Scalar, Batch)]
    outs = []
    for env in envs:
        assert env._host_reward and env._host_term
        env.reset(seed=11)
        g = torch.Generator(device="cuda").manual_seed(2)
        for _ in range(3):
            a = torch.rand((n, 2), device="cuda", generator=g) * torch.tensor([2.0, 0.5], device="cuda") - torch.tensor([1.0, 0.25], device="cuda")
            obs, rew, done, info = env.step(a)
        outs.append((obs, rew, done))
    (o0, r0, d0), (o1, r1, d1) = outs
    assert torch.equal(o0, o1) and torch.equal(d0, d1) and bool(d0.any()) and not bool(d0.all())
    torch.testing.assert_close(r0, r1, rtol=1e-5, atol=1e-6)
    want = -(o0[:, :2].double().cpu().numpy() - envs[0]._task.goals[0].pos).__abs__().sum(1)
    np.testing.assert_allclose(r0.cpu().numpy(), want, rtol=1e-5, atol=1e-5)   # Point has no inner reward
    with pytest.raises(ValueError):
        MazeEnv(PointEnv, Scalar, num_envs=n, device="cuda:0", maze_size_scaling=4.0, auto_reset=True)


def test_step_returns_copies_unless_asked_not_to(torch_cuda):
    from mujoco_maze import gym

    torch = torch_cuda
    env = gym.make("PointUMaze-v0", num_envs=32, device="cuda:0")
    env.reset(seed=1)
    a = torch.zeros((32, 2), device="cuda")
    a[:, 0] = 0.5
    obs1, r1, d1, i1 = env.step(a)
    keep = obs1.clone()
    obs2, *_ = env.step(a)
    assert torch.equal(obs1, keep) and not torch.equal(obs1, obs2)          # the caller's tensor survives the next step
    b1, *_ = env.step(a, copy=False)
    snap = b1.clone()
    b2, *_ = env.step(a, copy=False)
    assert b1.data_ptr() == b2.data_ptr() and not torch.equal(b2, snap)      # the engine's buffer: overwritten in place


def test_buffers_are_validated_before_their_addresses_reach_the_c_abi(torch_cuda):
    from conftest import make_model
    from mujoco_maze.backend import BatchedSim

    torch = torch_cuda
    n = 32
    sim = BatchedSim(make_model("PointUMaze-v0"), n)
    sim.reset(seed=0)
    ok = dict(action=torch.zeros((n, 2), device="cuda"), obs=torch.empty((n, sim.obs_dim), device="cuda"),
              reward=torch.empty((n,), device="cuda"), done=torch.empty((n,), device="cuda", dtype=torch.uint8))
    sim.step_into(**ok)
    for key, bad in (("action", torch.zeros((n, 2), device="cuda", dtype=torch.float64)),     # dtype
                     ("action", torch.zeros((2, n), device="cuda").t()),                        # not contiguous
                     ("obs", torch.empty((n, sim.obs_dim + 1), device="cuda")),                 # shape
                     ("reward", torch.empty((n,))),                                             # host tensor on the device path
                     ("done", torch.empty((n,), device="cuda", dtype=torch.bool))):
        with pytest.raises(ValueError):
            sim.step_into(**{**ok, key: bad})
    host = {k: v.cpu() for k, v in ok.items()}
    with pytest.raises(ValueError):                                                             # pageable, not pinned
        sim.step_host(host["action"], host["obs"], host["reward"], host["done"])
    pinned = {k: v.pin_memory() for k, v in host.items()}
    sim.step_host(pinned["action"], pinned["obs"], pinned["reward"], pinned["done"])
    assert bool(torch.isfinite(pinned["obs"]).all())
    sim.close()
