"""Multi-GPU sharding of a batch of environments: one process per GPU, no data-path collective.

Environments never exchange data (the reference holds one independent `mjData` per `MazeEnv`,
reference maze_env.py:218), so a batch of `total` environments is split into contiguous
env-index ranges, one per rank. The reset noise is keyed by the GLOBAL env index
(`env_offset`, include/mmz.h: mmz_set_env_offset), which makes 1 GPU x N and G GPUs x N/G
produce the same trajectories. The only collective is optional: all-gathering the observations
when the caller wants one tensor. Two ways: `ObsGatherer` (NCCL on GPUs; the same code runs over gloo on CPU
tensors, which is how tests/test_sharding_gloo.py covers it) and `PeerObsGatherer`, where the step kernel itself
stores the observations into every rank's gathered tensor over NVLink (symmetric memory; no collective at all).
"""

from typing import Optional, Tuple


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """(first global env index, number of envs) of `rank`; the remainder goes to the first ranks."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if total < world:
        raise ValueError(f"{total} environments cannot be split over {world} ranks")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def make_sharded(env_id: str, total_envs: int, rank: Optional[int] = None, world: Optional[int] = None,
                 device: Optional[str] = None, **kwargs):
    """gym.make(env_id) for this rank's shard of `total_envs` environments (rank/world default to torch.distributed)."""
    import torch.distributed as dist

    from mujoco_maze import gym

    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    start, count = shard_bounds(total_envs, world, rank)
    if device is None:
        import os

        device = f"cuda:{int(os.environ.get('LOCAL_RANK', rank))}"
    return gym.make(env_id, num_envs=count, device=device, env_offset=start, **kwargs)


class ObsGatherer:
    """All-gathers per-rank `[n_r, obs_dim]` observation shards into one `[total, obs_dim]` tensor, in env order."""

    def __init__(self, total_envs: int, obs_dim: int, device, dtype=None, group=None):
        import torch
        import torch.distributed as dist

        self.dist, self.torch, self.group = dist, torch, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total, self.obs_dim = int(total_envs), int(obs_dim)
        self.bounds = [shard_bounds(self.total, self.world, r) for r in range(self.world)]
        self.equal = len({c for _, c in self.bounds}) == 1
        dtype = dtype or torch.float32
        self.out = torch.empty((self.total, self.obs_dim), dtype=dtype, device=device)
        if not self.equal:  # ragged shards: gather padded blocks, then compact
            self.maxc = max(c for _, c in self.bounds)
            self.pad = torch.zeros((self.maxc, self.obs_dim), dtype=dtype, device=device)
            self.blocks = torch.empty((self.world * self.maxc, self.obs_dim), dtype=dtype, device=device)

    def __call__(self, local_obs):
        start, count = self.bounds[self.rank]
        if tuple(local_obs.shape) != (count, self.obs_dim):
            raise ValueError(f"rank {self.rank} expected a shard of shape {(count, self.obs_dim)}, got {tuple(local_obs.shape)}")
        if self.world == 1:
            self.out.copy_(local_obs)
        elif self.equal:
            self.dist.all_gather_into_tensor(self.out, local_obs.contiguous(), group=self.group)
        else:
            self.pad[:count].copy_(local_obs)
            self.dist.all_gather_into_tensor(self.blocks, self.pad, group=self.group)
            for r, (s, c) in enumerate(self.bounds):
                self.out[s:s + c].copy_(self.blocks[r * self.maxc:r * self.maxc + c])
        return self.out


class PeerObsGatherer:
    """Fused gather: `mmz_step` stores each observation row a second time into the `[total, obs_dim]` tensor of EVERY rank
    (peer-mapped symmetric memory, or one NVLS multicast address), from inside the step kernel - the NVLink transfers
    overlap the physics of the blocks still running, and nothing runs after the kernel but `sync()`, a device-side
    barrier on the current stream (signal pads of the same symmetric allocation).

        g = PeerObsGatherer(sim, total_envs, first_env)      # collective: every rank of the group calls it
        sim.step(actions); g.sync(); policy(g.out)          # g.out == ObsGatherer's result, bit for bit
    """

    def __init__(self, sim, total_envs: int, first_env: int, group=None, multicast: bool = False):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.sim, self.torch = sim, torch
        group = group or dist.group.WORLD
        self.world = dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("the fused gather holds at most 8 peer addresses (one NVSwitch node)")
        with torch.cuda.device(sim.device):
            self.out = symm_mem.empty((int(total_envs), sim.obs_dim), dtype=torch.float32, device=sim.device)
            self.out.zero_()
            self.hdl = symm_mem.rendezvous(self.out, group)
        self.multicast = bool(multicast and self.hdl.has_multicast_support and self.hdl.multicast_ptr)
        if self.multicast:
            ptrs = [self.hdl.multicast_ptr + (self.out.data_ptr() - self.hdl.buffer_ptrs[self.hdl.rank])]
        else:  # one view of every rank's buffer (this rank's own included), same offset inside the allocation
            self._views = [self.hdl.get_buffer(r, tuple(self.out.shape), torch.float32) for r in range(self.world)]
            ptrs = [v.data_ptr() for v in self._views]
        self._ptrs = ptrs
        sim.set_obs_peers(ptrs, int(first_env), self.multicast)
        self.sync()

    def sync(self):
        """All ranks' stores of the step just launched are complete and visible after this (stream-ordered)."""
        self.hdl.barrier(channel=0)
        return self.out

    def close(self):
        self.sim.set_obs_peers([], 0)
