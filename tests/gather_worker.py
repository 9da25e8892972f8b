"""torchrun worker of tests/test_gpu_gather.py: PeerObsGatherer against dist.all_gather_into_tensor on two ranks."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mujoco-maze_b200"))
import mujoco_maze  # noqa: E402,F401
from mujoco_maze.sharding import PeerObsGatherer, make_sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 4096
env = make_sharded("Ant4Rooms-v0", world * n, auto_reset=True)
sim = env.unwrapped.sim
env.reset(seed=0)
for multicast in (False, True):
    g = PeerObsGatherer(sim, world * n, rank * n, multicast=multicast)
    gen = torch.Generator(device="cuda").manual_seed(5 + rank)
    ref = torch.empty((world * n, sim.obs_dim), device="cuda")
    for k in range(4):
        a = 60 * torch.rand((n, sim.nu), device="cuda", generator=gen) - 30
        obs, *_ = sim.step(a)
        out = g.sync()
        dist.all_gather_into_tensor(ref, obs.contiguous())
        torch.cuda.synchronize()
        assert torch.equal(out, ref), f"rank {rank} step {k} multicast={g.multicast}: fused gather != all_gather"
        g.sync()  # nobody starts the next step (which overwrites the rows) before everyone has compared
    if rank == 0:
        print(f"fused gather ok (multicast requested {multicast}, used {g.multicast})", flush=True)
    g.close()
dist.barrier()
if rank == 0:
    print("GATHER_OK", flush=True)
dist.destroy_process_group()
