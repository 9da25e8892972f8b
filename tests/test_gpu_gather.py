"""The fused observation gather (include/mmz.h: mmz_set_obs_peers; BASELINE configs[3]).

One GPU: the step kernel must store every observation row a second time at `row_offset + env` of each buffer it is given,
bit-identical to what it returns, for both kernel families. Two GPUs (skipped on a one-GPU box): two ranks over symmetric
memory - the tensor every rank ends up with must equal dist.all_gather_into_tensor of the per-rank observations."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import make_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env_id", ["Ant4Rooms-v0", "PointUMaze-v0", "AntPush-v0"])
def test_step_stores_observations_into_peer_buffers(env_id):
    import torch

    from mujoco_maze.backend import BatchedSim, MmzError

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA")
    n, total, off = 96, 300, 131
    model = make_model(env_id, num_envs=n)
    sim = BatchedSim(model, n, auto_reset=True)
    sim.reset(seed=2)
    q, v, t = sim.get_state()
    t[: n // 3] = 999                     # a third of the batch restarts inside the launch: the gathered row is the NEW episode's
    sim.set_state(q, v, t)
    bufs = [torch.full((total, sim.obs_dim), -7.0, device="cuda") for _ in range(3)]
    sim.set_obs_peers([b.data_ptr() for b in bufs], off)
    lo, hi = (torch.as_tensor(np.asarray(model.act_ctrlrange, np.float32)[: sim.nu, k], device="cuda") for k in (0, 1))
    a = lo + (hi - lo) * torch.rand((n, sim.nu), device="cuda")
    obs, _, done, _ = sim.step(a)
    torch.cuda.synchronize()
    assert int((done & 1).sum()) == n // 3
    for b in bufs:
        assert torch.equal(b[off:off + n], obs)
        assert bool((b[:off] == -7.0).all()) and bool((b[off + n:] == -7.0).all())
    sim.set_obs_peers([], 0)              # off again: the buffers stay as they are
    bufs[0].fill_(-7.0)
    sim.step(a)
    torch.cuda.synchronize()
    assert bool((bufs[0] == -7.0).all())
    with pytest.raises(MmzError):
        sim.set_obs_peers([b.data_ptr() for b in bufs] * 3, 0)   # more than 8 buffers
    sim.close()


def test_two_rank_fused_gather_equals_all_gather():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "gather_worker.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-3000:]
