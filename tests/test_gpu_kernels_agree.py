"""The two CUDA mappings of the same step must agree: the hybrid kernel (mmz_hkernel.cuh, default for the Ant
family) against the lanes-per-environment kernel (mmz_dyn.cuh, MMZ_KERNEL=groups), same states and actions,
through the C ABI. fp32 on both sides, different summation orders: tolerance 2e-4 relative on qpos / qvel."""
import os

import numpy as np
import pytest

from conftest import make_model

pytestmark = pytest.mark.gpu


def _run(model, env_id, kernel, q, v, acts):
    import torch

    from mujoco_maze.backend import BatchedSim

    old = os.environ.get("MMZ_KERNEL")
    if kernel:
        os.environ["MMZ_KERNEL"] = kernel
    else:
        os.environ.pop("MMZ_KERNEL", None)
    try:
        sim = BatchedSim(model, q.shape[0])
    finally:
        if old is None:
            os.environ.pop("MMZ_KERNEL", None)
        else:
            os.environ["MMZ_KERNEL"] = old
    sim.set_state(q, v, np.zeros(q.shape[0], dtype=np.int32))
    outs = []
    for a in acts:
        obs, rew, done, info = sim.step(torch.as_tensor(a, device="cuda"))
        outs.append((obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy()))
    qq, vv, _ = sim.get_state()
    cfg = dict(sim.kernel_config)
    sim.close()
    return outs, qq.cpu().numpy(), vv.cpu().numpy(), cfg


@pytest.mark.parametrize("env_id", ["AntUMaze-v0", "Ant4Rooms-v0", "AntPush-v0", "AntFall-v0", "PointUMaze-v0", "Point4Rooms-v0",
                                    "PointPush-v0", "PointFall-v0"])
def test_hybrid_and_groups_kernels_agree(env_id):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from test_gpu_parity import sample_actions, sample_states

    rng = np.random.default_rng(5)
    model = make_model(env_id)
    n = 200  # not a multiple of 32: exercises the padding environments of the last block
    q, v = sample_states(model, env_id, n, rng)
    if env_id.startswith("Point"):
        v = np.clip(v, -9, 9)
    acts = [sample_actions(model, n, rng).astype(np.float32) for _ in range(3)]
    oh, qh, vh, cfg_h = _run(model, env_id, None, q, v, acts)
    og, qg, vg, cfg_g = _run(model, env_id, "groups", q, v, acts)
    assert cfg_h != cfg_g, "both runs used the same kernel"
    # single step: tight; after 3 free-running steps contact-rich environments may have diverged chaotically
    o1 = np.abs(oh[0][0] - og[0][0]) / (1 + np.abs(og[0][0]))
    assert np.quantile(o1.max(1), 0.98) < 2e-4, float(o1.max())
    assert (oh[0][2] == og[0][2]).all()
    np.testing.assert_allclose(oh[0][1], og[0][1], atol=1e-4)
    e3 = np.abs(qh - qg) / (1 + np.abs(qg))
    assert np.median(e3.max(1)) < 1e-4


def test_kernel_selection_per_model():
    """Which kernel serves which model (mmz_kernel_name): the hybrid for the Ant family with and without movable blocks and
    for the Point (small batches, or with movable blocks), the lanes-per-environment kernel (8 / 16 / 32 lanes) with only the features the model needs for everything else."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mujoco_maze.backend import BatchedSim

    want = {
        "AntUMaze-v0": "maze_hkernel<14,0>", "Ant4Rooms-v0": "maze_hkernel<14,0>", "AntPush-v0": "maze_hkernel<16,1>",
        "AntFall-v0": "maze_hkernel<16,1>", "PointUMaze-v0": "maze_hkernel<4,1>", "PointPush-v0": "maze_hkernel<16,1>",
        "SwimmerUMaze-v0": "maze_kernel<8,8,2>",
        "ReacherUMaze-v0": "maze_kernel<8,4,7>", "AntMultiPush-v0": "maze_kernel<32,20,7>", "PointBilliard-v0": "maze_kernel<8,8,7>",
        "AntSmallBilliard-v0": "maze_kernel<32,20,7>",
    }
    got = {}
    for env_id in want:
        sim = BatchedSim(make_model(env_id), 32)
        got[env_id] = sim.kernel_config["kernel"]
        sim.close()
    assert got == want
    # a single-body Point in a large batch stays on the lanes kernel (64 environments in flight per SM)
    sim = BatchedSim(make_model("PointUMaze-v0"), 32768)
    assert sim.kernel_config["kernel"] == "maze_kernel<8,4,1>"
    sim.close()
