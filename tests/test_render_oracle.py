"""The numpy restatement of the rasteriser (oracle/render_oracle.py) on CPU: known colours under known world points.
(Its comparison with the CUDA rasteriser is tests/test_gpu_render.py::test_render_matches_the_cpu_restatement_pixel_for_pixel.)"""
import numpy as np

from conftest import make_model


def _pixel(img, model, x, y):
    from oracle import render_oracle

    h, w = img.shape[:2]
    x0, y0, x1, y1 = render_oracle.window(model)
    return tuple(int(v) for v in img[int((y1 - y) / (y1 - y0) * h), int((x - x0) / (x1 - x0) * w)])


def test_restatement_draws_walls_floor_goal_agent_and_block(oracle_lib):
    from oracle import render_oracle

    for env_id in ("PointUMaze-v0", "AntPush-v0"):
        model = make_model(env_id)
        o = oracle_lib.OracleEnv(model)
        nq, nv = int(model.nq), int(model.nv)
        o.set_state(np.asarray(model.qpos0, float)[:nq], np.zeros(nv), 0)
        o.forward()
        xpos, xquat = o.xpos()
        img = render_oracle.render(model, xpos, xquat, 200, 200)
        s = float(model.cell_size)
        assert img.shape == (200, 200, 3) and img.dtype == np.uint8
        ox, oy = np.asarray(model.origin, float)
        assert _pixel(img, model, -ox, -oy) == (102, 102, 102)                  # cell (0, 0): outer wall
        agent = np.asarray(_pixel(img, model, 0.0, 0.0), float) / 255            # the robot starts at the origin
        assert abs(agent[0] / agent[2] - 2.0) < 0.1 and abs(agent[1] / agent[2] - 1.5) < 0.1   # (0.8, 0.6, 0.4) shaded
        goal = np.asarray(model.goal_pos, float)[0]
        if env_id == "PointUMaze-v0":
            g = _pixel(img, model, goal[0], goal[1])
            assert g[0] > 200 and g[1] < 80                                      # red goal disc
            f = _pixel(img, model, s, 0.25 * s)
            assert f[1] > f[0] and f[1] > 150                                    # greenish floor
        else:
            b = model.names["body"].index(model.meta["movable_blocks"][0])
            blk = np.asarray(_pixel(img, model, float(model.body_pos[b][0]), float(model.body_pos[b][1])), float) / 255
            assert blk[0] > 0.6 and blk[1] < 0.15 and blk[2] < 0.15              # (0.9, 0.1, 0.1) movable block
