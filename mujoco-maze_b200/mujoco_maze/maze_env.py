"""MazeEnv: the reference's environment class over the batched CUDA step engine.

Mirrors the public surface of reference maze_env.py:27-486 — constructor
keywords (:28-44), `step` (:448-481) -> (obs, reward, done, info), `reset`
(:371-382), `_get_obs` (:351-369), `action_space` (:422-424),
`observation_space` (:235-246), `has_extended_obs`, `get_ori` — but holds N
environments. Everything the reference does per step in Python/MuJoCo happens
in one kernel launch (`mmz_step`, include/mmz.h); this class only marshals
tensors.

Two calling conventions:
  * `num_envs=None` (what plain `gym.make(id)` gives): one environment, numpy in /
    numpy out, scalars for reward/done — shaped exactly like the reference so its
    tests read unchanged. Still runs on the GPU; there is no CPU path.
  * `num_envs=N`: torch CUDA tensors `[N, ...]` in and out, per-env episode
    counters and TimeLimit truncation inside the kernel, optional auto-reset.

Decision on the reference's half-migrated API: `reset()` returns
the observation (gym 0.20 / the reference's own tests); `reset(return_info=True)`
returns `(obs, info)`.
"""

import itertools as it
from typing import Any, List, Optional, Tuple, Type

import numpy as np

from mujoco_maze import gym, maze_env_utils, maze_task
from mujoco_maze import maze_task as _tasks  # the ctor keyword `maze_task` shadows the module
from mujoco_maze.agent_model import AgentModel
from mujoco_maze.model_compiler import MazeModel, compile_maze_model


class MazeEnv(gym.Env):
    def __init__(
        self,
        model_cls: Type[AgentModel],
        maze_task: Type[maze_task.MazeTask] = maze_task.MazeTask,
        include_position: bool = True,  # accepted and ignored, as upstream (quirk Q13)
        maze_height: float = 0.5,
        maze_size_scaling: float = 4.0,
        inner_reward_scaling: float = 1.0,
        restitution_coef: float = 0.8,
        task_kwargs: dict = {},
        websock_port: Optional[int] = None,
        camera_move_x: Optional[float] = None,
        camera_move_y: Optional[float] = None,
        camera_zoom: Optional[float] = None,
        image_shape: Tuple[int, int] = (600, 480),
        num_envs: Optional[int] = None,
        device: str = "cuda:0",
        auto_reset: bool = False,
        env_offset: int = 0,
        seed: int = 0,
        max_episode_steps: int = 1000,
        **kwargs,
    ) -> None:
        self.is_batched = num_envs is not None
        self.num_envs = int(num_envs) if self.is_batched else 1
        self.device = device
        self._auto_reset = bool(auto_reset)
        self._env_offset = int(env_offset)  # global index of env 0 when this batch is one shard of a larger one
        self._seed = int(seed)
        self._episode = 0

        self._task = maze_task(maze_size_scaling, **task_kwargs)
        self._maze_height = maze_height
        self._maze_size_scaling = maze_size_scaling
        self._inner_reward_scaling = inner_reward_scaling
        self._restitution_coef = restitution_coef
        self._observe_blocks = self._task.OBSERVE_BLOCKS
        self._observe_balls = self._task.OBSERVE_BALLS
        self._top_down_view = self._task.TOP_DOWN_VIEW
        self._put_spin_near_agent = self._task.PUT_SPIN_NEAR_AGENT

        self.wrapped_env = model_cls(file_path=None, **kwargs)
        from mujoco_maze.ant import forward_reward_kind

        fwd_fn = getattr(self.wrapped_env, "_forward_reward_fn", None)
        self._forward_kind = forward_reward_kind(fwd_fn) if fwd_fn is not None else 0
        self._prev_xy = None  # host-evaluated forward_reward_fn: xy before the step
        self.model: MazeModel = compile_maze_model(
            model_cls,
            self._task,
            maze_size_scaling,
            maze_height=maze_height,
            inner_reward_scaling=inner_reward_scaling,
            restitution_coef=restitution_coef,
            forward_reward_weight=getattr(self.wrapped_env, "_forward_reward_weight", 1.0),
            ctrl_cost_weight=getattr(self.wrapped_env, "_ctrl_cost_weight", 1e-4),
            forward_reward_kind=self._forward_kind,
            # the scalar env is wrapped by gym's TimeLimit on the host, exactly like upstream
            max_episode_steps=max_episode_steps if self.is_batched else 0,
        )
        self.wrapped_env._bind(self)

        meta = self.model.meta
        self._maze_structure = meta["structure"]
        self.elevated = meta["elevated"]
        self.blocks = meta["blocks"]
        self._init_torso_x, self._init_torso_y = meta["torso_xy"]
        self._init_positions = [(x - self._init_torso_x, y - self._init_torso_y) for x, y in self._find_all_robots()]
        self.movable_blocks: List[str] = list(meta["movable_blocks"])
        self.object_balls: List[str] = list(meta["object_balls"])
        if model_cls.MANUAL_COLLISION:
            # host copy, for introspection only: the kernel holds the same segments
            self._collision = maze_env_utils.CollisionDetector(
                self._maze_structure, maze_size_scaling, self._init_torso_x, self._init_torso_y, model_cls.RADIUS
            )
        else:
            self._collision = None
        self._host_reward = int(self.model.reward_rule) == _tasks.REWARD_HOST
        self._host_term = int(self.model.term_rule) == _tasks.TERM_HOST
        if self._auto_reset and (self._host_reward or self._host_term or self._forward_kind == 2):
            # The kernel restarts an episode inside the launch that ends it: a termination only the host can see would
            # never reset its environment, and after a TimeLimit reset the returned observation already belongs to the
            # next episode, so host rules would be evaluated on the wrong state.
            raise ValueError("auto_reset=True needs the task's reward / termination (and forward_reward_fn) to run inside "
                             "the kernel; this task or agent defines its own in Python. Use auto_reset=False and reset(mask=done).")

        self.observation_space = self._get_obs_space()
        self._websock_port = websock_port
        self._image_shape = image_shape
        self._sim = None
        self.t = 0

    # ------------------------------------------------------------------ backend
    @property
    def sim(self):
        """The device engine, created on first use (so the class is importable without a GPU)."""
        if self._sim is None:
            from mujoco_maze.backend import BatchedSim

            self._sim = BatchedSim(self.model, self.num_envs, self.device, auto_reset=self._auto_reset,
                                   env_offset=self._env_offset)
        return self._sim

    def _state(self):
        return self.sim.get_state()

    def _cat(self, *xs):
        import torch

        return torch.cat(xs, dim=1)

    def _out(self, x):
        """Batched: device tensor as is. Scalar env: numpy float64 without the batch axis."""
        if self.is_batched:
            return x
        return x[0].detach().cpu().numpy().astype(np.float64)

    def _set_xy(self, xy) -> None:
        import torch

        qpos, qvel, t = self.sim.get_state()
        xy = torch.as_tensor(np.asarray(xy) if not torch.is_tensor(xy) else xy, dtype=torch.float32, device=qpos.device)
        qpos[:, :2] = xy.reshape(-1, 2)
        self.sim.set_state(qpos, qvel, t)

    # ------------------------------------------------------------------ spaces
    @property
    def has_extended_obs(self) -> bool:
        return self._top_down_view or self._observe_blocks or self._observe_balls

    @property
    def action_space(self):
        return self.wrapped_env.action_space

    def get_ori(self):
        return self.wrapped_env.get_ori()

    def _get_obs_space(self):
        n = int(self.model.obs_dim)
        high = np.inf * np.ones(n, dtype=np.float32)
        low = -high
        inner = self.wrapped_env.observation_space
        k = min(inner.shape[0], n)
        high[:k], low[:k] = inner.high[:k], inner.low[:k]
        low[0], high[0], low[1], high[1] = self._xy_limits()
        return gym.spaces.Box(low, high)

    def _xy_limits(self) -> Tuple[float, float, float, float]:
        open_cells = [
            (i, j)
            for i, row in enumerate(self._maze_structure)
            for j, c in enumerate(row)
            if not c.is_block()
        ]
        s = self._maze_size_scaling
        js = [j for _, j in open_cells]
        is_ = [i for i, _ in open_cells]
        xmin, xmax = (min(js) - 0.5) * s - self._init_torso_x, (max(js) + 0.5) * s - self._init_torso_x
        ymin, ymax = (min(is_) - 0.5) * s - self._init_torso_y, (max(is_) + 0.5) * s - self._init_torso_y
        return xmin, xmax, ymin, ymax

    def _find_robot(self) -> Tuple[float, float]:
        robots = self._find_all_robots()
        if not robots:
            raise ValueError("No robot in maze specification.")
        return robots[0]

    def _find_all_robots(self) -> List[Tuple[float, float]]:
        s = self._maze_size_scaling
        st = self._maze_structure
        return [(j * s, i * s) for i, j in it.product(range(len(st)), range(len(st[0]))) if st[i][j].is_robot()]

    # ------------------------------------------------------------------ episode API
    def _get_obs(self):
        return self._out(self.sim.observe())

    def get_top_down_view(self):
        """The 5x5x3 egocentric raster (walls, chasms, movable blocks) of reference maze_env.py:262-349.

        Computed on the GPU by `maze_view_kernel` for tasks with TOP_DOWN_VIEW and delivered as the 75 entries
        before the trailing `t * 0.001` of the observation (maze_env.py:353-354, 369); this accessor reshapes them
        to `[5, 5, 3]` (`[N, 5, 5, 3]` for a batch).
        """
        if not self._top_down_view:
            raise ValueError("the task does not set TOP_DOWN_VIEW")
        obs = self._get_obs()
        return obs[..., -76:-1].reshape(*obs.shape[:-1], 5, 5, 3)

    def reset(self, seed: Optional[int] = None, return_info: bool = False, mask=None, **kwargs):
        if seed is not None:
            self._seed = int(seed)
        self.t = 0
        self._episode += 1
        obs = self.sim.reset(seed=(self._seed << 20) + self._episode, mask=mask)
        if self._forward_kind == 2:
            self._prev_xy = obs[:, :2].clone()
        obs = self._out(obs if not self.is_batched else obs.clone())
        return (obs, {}) if return_info else obs

    def step(self, action, copy: bool = True):
        """One MazeEnv.step (maze_env.py:448-481) of every environment in ONE kernel launch.

        Batched envs return device tensors. With `copy=True` (default) they are fresh tensors the caller may keep; with
        `copy=False` they are the engine's own output buffers, overwritten by the next `step` / `reset` (saves four
        device-to-device copies per step for callers that consume them right away)."""
        self.t += 1
        sim = self.sim
        if not self.is_batched:
            action = np.asarray(action, dtype=np.float32).reshape(1, -1)
        obs, reward, done, info_t = sim.step(action)
        if self.is_batched and copy:
            obs, reward, done, info_t = obs.clone(), reward.clone(), done.clone(), info_t.clone()
        if self._forward_kind == 2:
            reward, info_t = self._host_forward_reward(obs, reward, info_t)
        if self._host_reward or self._host_term:
            reward, done = self._host_rules(obs, reward, done, info_t)
        has_inner = self.model.step_kind == 0  # torque agents report reward_forward / reward_ctrl
        if self.is_batched:
            info = {"position": info_t[:, :2], "TimeLimit.truncated": (done & 2) != 0, "unstable": (done & 4) != 0}
            if has_inner:
                info["reward_forward"], info["reward_ctrl"] = info_t[:, 2], info_t[:, 3]
            return obs, reward, (done & 1) != 0, info
        o = obs[0].cpu().numpy().astype(np.float64)
        i_np = info_t[0].cpu().numpy().astype(np.float64)
        info = {"position": i_np[:2].copy()}
        if has_inner:
            info["reward_forward"], info["reward_ctrl"] = float(i_np[2]), float(i_np[3])
        r = float(reward[0].item())
        if not (self._host_reward or self._host_term):
            # The scalar (reference-shaped) call returns Python floats like the reference: the outer reward is one of a
            # few exact constants (PENALTY, 1.0, a goal's reward_scale; maze_task.py) that fp32 cannot represent, and the
            # reference's tests compare it with `==` (tests/test_envs.py:32-36). Snap the kernel's value to the constant
            # it is the fp32 image of; distances and the inner reward stay as computed on the device.
            inner = 0.0
            if has_inner:
                inner = float(self._inner_reward_scaling) * (
                    float(getattr(self.wrapped_env, "_forward_reward_weight", 1.0)) * float(i_np[2]) + float(i_np[3]))
            outer = r - inner
            for c in self._exact_outer_rewards():
                if abs(outer - c) <= 2e-7 * max(1.0, abs(c)):  # within fp32 rounding of the constant
                    r = inner + c
                    break
        return o, r, bool(int(done[0].item()) & 1), info

    def _exact_outer_rewards(self):
        consts = [float(self._task.PENALTY), 1.0, 0.0]
        consts += [float(g.reward_scale) for g in self._task.goals]
        return consts

    def _host_forward_reward(self, obs, reward, info_t):
        """A `forward_reward_fn` that is neither of the reference's two (ant.py:18-23, 44-53): the kernel left the
        forward term out (MMZ_FWD_HOST); evaluate the callable on the step's xy velocity and add it. The callable is
        the reference's scalar signature fn(xy_velocity[2]); a result of shape [N] for the [N, 2] batch is used as is,
        anything else falls back to one call per environment."""
        import torch

        xy = info_t[:, :2]
        prev = self._prev_xy if self._prev_xy is not None else xy
        dt = float(self.model.timestep) * int(self.model.frame_skip)
        vel = ((xy - prev) / dt).detach().cpu().numpy().astype(np.float64)
        fn = self.wrapped_env._forward_reward_fn
        fwd = None
        try:
            out = np.asarray(fn(vel), dtype=np.float64)
            if out.shape == (vel.shape[0],) and vel.shape[0] > 1:
                fwd = out
        except Exception:  # noqa: BLE001  (not written for batches)
            fwd = None
        if fwd is None:
            fwd = np.array([float(fn(v)) for v in vel])
        fwd_t = torch.as_tensor(fwd, dtype=torch.float32, device=reward.device)
        w = float(self._inner_reward_scaling) * float(getattr(self.wrapped_env, "_forward_reward_weight", 1.0))
        info_t = info_t.clone()
        info_t[:, 2] = fwd_t
        self._prev_xy = obs[:, :2].clone()
        return reward + w * fwd_t, info_t

    def _host_rules(self, obs, reward, done, info_t):
        """User-defined MazeTask.reward / termination (README custom-task recipe, reference README.md:79-127).

        The kernel returned the scaled inner reward (outer rule REWARD_HOST adds nothing) and done without the goal
        test (TERM_HOST). A task may offer batch versions - `reward_batch(obs)` / `termination_batch(obs)`, `[N, D]`
        device tensor in, `[N]` tensor out - which run without leaving the device; otherwise the reference's scalar
        methods are called per environment on ONE host copy of the observations (fine for small batches; a 65 536-env
        batch needs the batch versions)."""
        import torch

        task = self._task
        rb, tb = getattr(task, "reward_batch", None), getattr(task, "termination_batch", None)
        need_loop = (self._host_reward and rb is None) or (self._host_term and tb is None)
        o = obs.detach().cpu().numpy().astype(np.float64) if need_loop else None
        if self._host_reward:
            if rb is not None:
                reward = reward + torch.as_tensor(rb(obs), dtype=torch.float32, device=reward.device).reshape(-1)
            else:
                add = np.fromiter((float(task.reward(row)) for row in o), dtype=np.float64, count=o.shape[0])
                reward = reward + torch.as_tensor(add, dtype=torch.float32, device=reward.device)
        if self._host_term:
            if tb is not None:
                hit = torch.as_tensor(tb(obs), device=done.device).reshape(-1).to(torch.bool)
            else:
                hit = torch.as_tensor(np.fromiter((bool(task.termination(row)) for row in o), dtype=bool, count=o.shape[0]),
                                      device=done.device)
            done = done | hit.to(torch.uint8)
        return reward, done

    def set_marker(self) -> None:
        """Reference maze_env.py:384-387 moves the goal sites of the MuJoCo scene before rendering. The rasteriser draws
        the goal discs straight from the compiled task goals, so there is nothing to move; kept for API parity."""

    def render(self, mode="rgb_array", width: int = 256, height: int = 256, env_ids: Any = None, **kwargs) -> Any:
        """Top-down RGB image(s) from the batched CUDA rasteriser (`mmz_render`).

        The reference renders through MuJoCo's OpenGL context (maze_env.py:389-420); here every requested
        environment is drawn by one kernel launch. `mode="rgb_array"`: uint8 `[height, width, 3]` numpy array for a
        single environment, `[count, height, width, 3]` CUDA tensor for a batch (`env_ids`: None = all, an int, or a
        `(first, count)` range). `mode="human"` (an on-screen / websocket viewer) is not rebuilt.
        """
        if mode != "rgb_array":
            raise NotImplementedError("only mode='rgb_array' is rebuilt; the on-screen and websocket viewers are not")
        if env_ids is None:
            first, count = 0, self.num_envs
        elif isinstance(env_ids, int):
            first, count = env_ids, 1
        else:
            first, count = env_ids
        rgb = self.sim.render(width, height, first, count)
        return rgb if self.is_batched else rgb[0].cpu().numpy()

    def close(self) -> None:
        if self._sim is not None:
            self._sim.close()
            self._sim = None
