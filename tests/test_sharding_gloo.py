"""Host-side sharding logic over 2 gloo ranks on CPU: shard bounds tile the env index range,
the observation all-gather reassembles shards in env order (equal and ragged), and make_sharded passes the
global env offset through. No GPU and no compute calls into libmmz."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mujoco-maze_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

from mujoco_maze.sharding import ObsGatherer, shard_bounds  # noqa: E402


def test_shard_bounds_tile_the_range():
    for total, world in [(8, 1), (8, 2), (10, 4), (524288, 8), (7, 7), (65537, 8)]:
        covered = []
        for r in range(world):
            s, c = shard_bounds(total, world, r)
            assert c >= total // world and c <= total // world + 1
            covered.extend(range(s, s + c))
        assert covered == list(range(total))
    with pytest.raises(ValueError):
        shard_bounds(3, 4, 0)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, obs_dim, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = shard_bounds(total, world, rank)
        # an "observation" that encodes the global env index, as the kernel's env_offset keying does
        idx = torch.arange(start, start + count, dtype=torch.float32)
        local = idx[:, None] * 10 + torch.arange(obs_dim, dtype=torch.float32)[None, :]
        g = ObsGatherer(total, obs_dim, device="cpu")
        full = g(local).clone()
        again = g(local + 1.0).clone()  # buffers are reused across steps
        q.put((rank, full.numpy(), again.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 65])  # equal shards -> all_gather_into_tensor; ragged -> padded path
def test_obs_all_gather_two_gloo_ranks(total):
    world, obs_dim = 2, 30
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, obs_dim, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(total, dtype=np.float32)[:, None] * 10 + np.arange(obs_dim, dtype=np.float32)[None, :]
    for rank, full, again in results:
        np.testing.assert_array_equal(full, expect)
        np.testing.assert_array_equal(again, expect + 1.0)


def test_make_sharded_passes_global_offset(monkeypatch):
    from mujoco_maze import gym, sharding

    seen = {}

    def fake_make(env_id, **kw):
        seen.update(kw, env_id=env_id)
        return "env"

    monkeypatch.setattr(gym, "make", fake_make)
    assert sharding.make_sharded("Ant4Rooms-v0", 524288, rank=3, world=8, device="cuda:3", auto_reset=True) == "env"
    assert seen == {"env_id": "Ant4Rooms-v0", "num_envs": 65536, "device": "cuda:3", "env_offset": 3 * 65536, "auto_reset": True}
