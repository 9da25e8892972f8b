"""The batched top-down rasteriser behind MazeEnv.render(mode="rgb_array") (`mmz_render`).

The reference renders with MuJoCo's OpenGL context (maze_env.py:389-420); there is no pixel-parity target, so these
are geometric checks: the colour under known world points (wall cell, free floor, goal site, agent, movable block),
that the agent's pixels follow set_state, and that a batch renders every environment independently.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WALL, BLOCK, AGENT = (102, 102, 102), (0.9, 0.1, 0.1), (0.8, 0.6, 0.4)


def pixel(img, model, x, y):
    """Colour at world (x, y) of an image from mmz_render (window = maze bounding box + half a cell)."""
    h, w = img.shape[:2]
    s, (ox, oy) = float(model.cell_size), np.asarray(model.origin, float)
    x0, x1 = -ox - s, (int(model.grid_w) - 1) * s - ox + s
    y0, y1 = -oy - s, (int(model.grid_h) - 1) * s - oy + s
    px, py = int((x - x0) / (x1 - x0) * w), int((y1 - y) / (y1 - y0) * h)
    return tuple(int(v) for v in img[py, px])


def ratio(c, ref):
    """A shaded colour keeps the channel ratios of its base colour."""
    c, ref = np.asarray(c, float), np.asarray(ref, float)
    k = c.sum() / (255 * ref.sum())
    return 0.75 <= k <= 1.01 and np.abs(c / 255 - k * ref).max() < 0.02


def test_render_point_umaze_and_ant_push():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA")
    import mujoco_maze
    from mujoco_maze import gym

    env = gym.make("PointUMaze-v0")
    env.reset()
    img = env.render(mode="rgb_array", width=200, height=200)
    m = env.unwrapped.model
    assert isinstance(img, np.ndarray) and img.shape == (200, 200, 3) and img.dtype == np.uint8
    s = float(m.cell_size)
    assert pixel(img, m, -s, 0.0) == WALL and pixel(img, m, s, s) == WALL  # outer wall, the U's middle bar
    assert ratio(pixel(img, m, 0.0, 0.0), AGENT)                          # the point robot starts at the origin
    floor = pixel(img, m, s, 0.25 * s)
    assert floor[1] > floor[0] and floor[1] > 150                          # greenish floor
    goal = np.asarray(env.unwrapped._task.goals[0].pos, float)
    g = pixel(img, m, goal[0], goal[1])
    assert g[0] > 200 and g[1] < 80                                        # red goal site
    with pytest.raises(NotImplementedError):
        env.render(mode="human")
    env.close()

    n = 8
    env = gym.make("AntPush-v0", num_envs=n, device="cuda:0")
    env.reset()
    sim, m = env.unwrapped.sim, env.unwrapped.model
    q, v, t = sim.get_state()
    s = float(m.cell_size)
    shift = torch.linspace(-0.3 * s, 0.3 * s, n, device=q.device)
    q[:, 0] = shift                      # every environment puts its ant somewhere else
    sim.set_state(q, v, t)
    imgs = env.render(mode="rgb_array", width=160, height=160)
    assert imgs.shape == (n, 160, 160, 3) and imgs.dtype == torch.uint8 and imgs.is_cuda
    imgs = imgs.cpu().numpy()
    b = m.names["body"].index(env.unwrapped.model.meta["movable_blocks"][0])
    bx, by = float(m.body_pos[b][0]), float(m.body_pos[b][1])
    for i in range(n):
        assert ratio(pixel(imgs[i], m, float(shift[i]), 0.0), AGENT), i   # torso sphere at its own x
        assert ratio(pixel(imgs[i], m, bx, by), BLOCK), i                 # the movable block at its cell
    assert not np.array_equal(imgs[0], imgs[-1])
    sub = env.render(mode="rgb_array", width=160, height=160, env_ids=(3, 2)).cpu().numpy()
    assert sub.shape[0] == 2 and np.array_equal(sub[0], imgs[3]) and np.array_equal(sub[1], imgs[4])
    try:
        from PIL import Image

        os.makedirs("gpurun_out", exist_ok=True)
        Image.fromarray(np.concatenate([imgs[0], imgs[-1]], axis=1)).save("gpurun_out/render_AntPush-v0.png")
        Image.fromarray(img).save("gpurun_out/render_PointUMaze-v0.png")
    except ImportError:
        pass
    env.close()


@pytest.mark.parametrize("env_id", ["PointUMaze-v0", "AntPush-v0", "AntFall-v0", "PointBilliard-v0", "Ant4Rooms-v0"])
def test_render_matches_the_cpu_restatement_pixel_for_pixel(env_id, oracle_lib):
    """The comparator of the rasteriser: oracle/render_oracle.py restates its specification in numpy (fp64), with body poses
    from the fp64 oracle's kinematics. Whole images of randomly posed environments must agree: at least 99.5 % of the pixels
    within 2 grey levels (silhouette edges may fall on the other side of a pixel centre in fp32), mean error < 0.5 level."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA")
    from conftest import make_model
    from mujoco_maze.backend import BatchedSim
    from oracle import render_oracle
    from test_gpu_parity import sample_states

    rng = np.random.default_rng(21)
    model = make_model(env_id)
    n, W, H = 6, 192, 160
    q, v = sample_states(model, env_id, n, rng)
    sim = BatchedSim(model, n)
    sim.set_state(q, v, np.zeros(n, dtype=np.int32))
    got = sim.render(W, H).cpu().numpy()
    o = oracle_lib.OracleEnv(model)
    worst_frac, worst_mean, moving_px = 1.0, 0.0, 0
    for i in range(n):
        o.set_state(q[i], v[i], 0)
        o.forward()
        xpos, xquat = o.xpos()
        want = render_oracle.render(model, xpos, xquat, W, H)
        diff = np.abs(got[i].astype(int) - want.astype(int)).max(-1)
        worst_frac = min(worst_frac, float((diff <= 2).mean()))
        worst_mean = max(worst_mean, float(diff.mean()))
        moving_px += int((want != render_oracle.render(model, xpos + np.array([1e3, 0, 0]), xquat, W, H)).any(-1).sum())
    assert moving_px > 60, "the restatement draws (almost) no moving geom: nothing would be compared"
    assert worst_frac >= 0.995 and worst_mean < 0.5, (env_id, worst_frac, worst_mean)
    sim.close()
