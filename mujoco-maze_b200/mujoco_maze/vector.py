"""gym.vector-style wrapper over the batched environment (the caller side of the hot path).

The reference leaves running many environments to the caller (one `MazeEnv` per process under a vector-env
wrapper). Here one `MazeEnv(num_envs=N, auto_reset=True)` already IS the vector of environments: every `step`
is one kernel launch, an environment that terminates or hits the 1000-step TimeLimit (reference __init__.py:31)
is re-initialised inside the same launch, and - as in `gym.vector` - the observation returned for it is the first
observation of its next episode while `dones` / `infos["TimeLimit.truncated"]` describe the episode that ended.
"""

from typing import Optional

import numpy as np

from mujoco_maze import gym


class VectorMazeEnv:
    def __init__(self, env_id: str, num_envs: int, device: str = "cuda:0", seed: int = 0, as_numpy: bool = False,
                 env_offset: int = 0, **kwargs) -> None:
        self.env = gym.make(env_id, num_envs=int(num_envs), device=device, auto_reset=True, seed=seed,
                            env_offset=env_offset, **kwargs).unwrapped
        self.num_envs = int(num_envs)
        self.as_numpy = bool(as_numpy)
        self.single_observation_space = self.env.observation_space
        self.single_action_space = self.env.action_space
        so, sa = self.single_observation_space, self.single_action_space
        self.observation_space = gym.spaces.Box(np.tile(so.low, (self.num_envs, 1)), np.tile(so.high, (self.num_envs, 1)))
        self.action_space = gym.spaces.Box(np.tile(sa.low, (self.num_envs, 1)), np.tile(sa.high, (self.num_envs, 1)))
        self._pending = None

    # ------------------------------------------------------------------ helpers
    def _out(self, x):
        return x.detach().cpu().numpy() if self.as_numpy else x

    def _actions(self, actions):
        import torch

        a = torch.as_tensor(np.asarray(actions) if not torch.is_tensor(actions) else actions, dtype=torch.float32,
                            device=self.env.device)
        return a.reshape(self.num_envs, -1)

    # ------------------------------------------------------------------ gym.vector API
    def reset(self, seed: Optional[int] = None):
        return self._out(self.env.reset(seed=seed))

    def step(self, actions):
        obs, reward, done, info = self.env.step(self._actions(actions))
        return self._out(obs), self._out(reward), self._out(done), {k: self._out(v) for k, v in info.items()}

    def step_async(self, actions) -> None:
        self._pending = self._actions(actions)

    def step_wait(self):
        if self._pending is None:
            raise RuntimeError("step_wait() without step_async()")
        a, self._pending = self._pending, None
        return self.step(a)

    def sample_actions(self):
        """Uniform actions over the control ranges (what `action_space.sample()` draws), on the device."""
        import torch

        lo = torch.as_tensor(self.single_action_space.low, device=self.env.device)
        hi = torch.as_tensor(self.single_action_space.high, device=self.env.device)
        return lo + (hi - lo) * torch.rand((self.num_envs, lo.numel()), device=self.env.device)

    def get_state(self):
        return self.env.sim.get_state()

    def close(self) -> None:
        self.env.close()
