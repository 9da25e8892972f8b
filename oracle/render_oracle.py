"""CPU restatement (numpy, fp64) of the batched top-down rasteriser `maze_render_kernel` - TEST INFRASTRUCTURE ONLY.

The reference's `MazeEnv.render(mode="rgb_array")` (maze_env.py:389-420) reads pixels back from MuJoCo's OpenGL context;
there is no such context here and no pixel-parity target against it. What CAN be checked is that the CUDA rasteriser draws
what its specification says, for every pixel: this module restates that specification independently - body poses come
from the fp64 oracle's kinematics (oracle/mmz_oracle.c), the image from vectorised numpy - and
tests/test_gpu_render.py compares whole images.

Specification (colours are the reference's rgba values): orthographic view from +z of the window [maze bounding box + one
cell]; row 0 = largest y. Static layer from the maze grid: wall boxes (0.4, 0.4, 0.4; maze_env.py:133,148), floor
(0.8, 0.9, 0.8; ant.xml:20 / point.xml:17) or platforms (0.9 grey) with a unit checker darkened to 85 %, chasms nearly black;
goal sites as discs of radius 0.1 * scaling (maze_env.py:199-210) in (0.9, 0.15, 0.15). Moving layer: the highest hit of the
ray down the z axis against every moving geom above the static layer (spheres and capsules in closed form, boxes by the
slab test); agent geoms (0.8, 0.6, 0.4), movable blocks (0.9, 0.1, 0.1; maze_env.py:600), object balls (0.1, 0.1, 0.7;
maze_env.py:500), shaded 0.8 + 0.2 * clamp(z / (2 * wall half height)).
"""
import numpy as np

GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 2, 3, 6
CELL_WALL, CELL_PLATFORM = 1, 2


def _quat2mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def window(model):
    s, (ox, oy) = float(model.cell_size), np.asarray(model.origin, float)
    return -ox - s, -oy - s, (int(model.grid_w) - 1) * s - ox + s, (int(model.grid_h) - 1) * s - oy + s


def render(model, xpos, xquat, width, height):
    """uint8 [height, width, 3] for one environment whose body poses are (xpos [nb, 3], xquat [nb, 4])."""
    m = model
    s = float(m.cell_size)
    x0, y0, x1, y1 = window(m)
    px, py = np.meshgrid(np.arange(width), np.arange(height))
    x = x0 + (px + 0.5) * (x1 - x0) / width
    y = y1 - (py + 0.5) * (y1 - y0) / height
    ox, oy = np.asarray(m.origin, float)
    gw, gh = int(m.grid_w), int(m.grid_h)
    j = np.floor((x + ox) / s + 0.5).astype(int)
    i = np.floor((y + oy) / s + 0.5).astype(int)
    inside = (i >= 0) & (i < gh) & (j >= 0) & (j < gw)
    grid = np.asarray(m.grid).reshape(-1)[: gw * gh].reshape(gh, gw)
    cell = np.where(inside, grid[np.clip(i, 0, gh - 1), np.clip(j, 0, gw - 1)], 0)
    dark = ((np.floor(x).astype(int) + np.floor(y).astype(int)) & 1) != 0
    wall_h = float(np.asarray(m.wall_half)[2])
    elevated = bool(int(m.elevated))
    c = np.zeros((height, width, 3))
    top = np.zeros((height, width))
    wall = (cell & CELL_WALL) != 0
    chasm = ~wall & elevated & ((cell & CELL_PLATFORM) == 0)
    ground = ~wall & ~chasm
    c[wall] = (0.4, 0.4, 0.4)
    top[wall] = float(m.wall_z) + wall_h
    c[chasm] = (0.05, 0.05, 0.08)
    top[chasm] = float(m.floor_z)
    base = np.array((0.9, 0.9, 0.9) if elevated else (0.8, 0.9, 0.8))
    c[ground] = base[None, :] * np.where(dark[ground], 0.85, 1.0)[:, None]
    top[ground] = float(m.plat_z) + wall_h if elevated else float(m.floor_z)
    for g in range(int(m.ngoal)):
        gp = np.asarray(m.goal_pos, float)[g]
        disc = ((x - gp[0]) ** 2 + (y - gp[1]) ** 2 <= (0.1 * s) ** 2) & ~wall
        c[disc] = (0.9, 0.15, 0.15)
    agent_root = int(np.asarray(m.body_root)[0])
    for g in range(int(m.ngeom)):
        b = int(np.asarray(m.geom_body)[g])
        Rb = _quat2mat(xquat[b])
        gpos = xpos[b] + Rb @ np.asarray(m.geom_pos, float)[g]
        R = _quat2mat(_quat_mul(xquat[b] / np.linalg.norm(xquat[b]), np.asarray(m.geom_quat, float)[g]))
        gtype = int(np.asarray(m.geom_type)[g])
        size = np.asarray(m.geom_size, float)[g]
        hit = np.full((height, width), -1e30)
        if gtype in (GEOM_SPHERE, GEOM_CAPSULE):
            r = size[0]
            cx, cy, cz = (np.full((height, width), v) for v in gpos)
            if gtype == GEOM_CAPSULE:
                h, ax, ay, az = size[1], R[0, 2], R[1, 2], R[2, 2]
                den = ax * ax + ay * ay
                t = ((x - gpos[0]) * ax + (y - gpos[1]) * ay) / den if den > 1e-12 else np.full((height, width), h if az > 0 else -h)
                t = np.clip(t, -h, h)
                cx, cy, cz = gpos[0] + t * ax, gpos[1] + t * ay, gpos[2] + t * az
            d2 = (x - cx) ** 2 + (y - cy) ** 2
            ok = d2 <= r * r
            hit[ok] = (cz + np.sqrt(np.maximum(r * r - d2, 0.0)))[ok]
        elif gtype == GEOM_BOX:
            rel = np.stack([x - gpos[0], y - gpos[1], np.full_like(x, 1e3 - gpos[2])], -1)
            o = rel @ R                      # R^T rel
            d = R.T @ np.array([0.0, 0.0, -1.0])
            t0 = np.full((height, width), -1e30)
            t1 = np.full((height, width), 1e30)
            for k in range(3):
                if abs(d[k]) < 1e-9:
                    t0 = np.where(np.abs(o[..., k]) > size[k], 1e30, t0)
                else:
                    a, b2 = (-size[k] - o[..., k]) / d[k], (size[k] - o[..., k]) / d[k]
                    t0 = np.maximum(t0, np.minimum(a, b2))
                    t1 = np.minimum(t1, np.maximum(a, b2))
            ok = t0 <= t1
            hit[ok] = (1e3 - t0)[ok]
        else:
            continue
        over = hit > top
        top = np.where(over, hit, top)
        k = 0.8 + 0.2 * np.clip(hit / (2 * wall_h + 1e-6), 0.0, 1.0)
        is_agent = int(np.asarray(m.body_root)[b]) == agent_root
        col = np.array((0.8, 0.6, 0.4) if is_agent else (0.9, 0.1, 0.1) if gtype == GEOM_BOX else (0.1, 0.1, 0.7))
        c[over] = col[None, :] * k[over][:, None]
    return np.minimum(255.0, c * 255.0 + 0.5).astype(np.uint8)
