"""Binary layout of the model blob shared by the CUDA kernel and the CPU oracle.

This table is the single source of truth; `include/mmz_model.h` is generated
from it (`python -m mujoco_maze.model_layout > include/mmz_model.h`, checked by
tests/test_model_layout.py). All integer fields come first (int32), then all
real fields (`mmz_real`: float for the kernel, double for the oracle), so the
struct has no padding in either precision.
"""

from typing import Dict, List, Tuple

import numpy as np

MAGIC = 0x4D4D5A31  # 'MMZ1'
VERSION = 5

CAPS = dict(
    MAXBODY=16,  # moving bodies (world excluded)
    MAXJNT=20,
    MAXDOF=20,
    MAXQ=24,
    MAXGEOM=20,  # geoms attached to moving bodies
    MAXACT=8,
    MAXGOAL=4,
    MAXSEG=64,  # wall segments of the manual clamp (registry max 48)
    MAXCELL=144,  # maze grid cells (registry max 9x9)
    MAXOBJ=8,  # latched bodies: observed ones spliced into obs (blocks / balls, <= 4) + the top-down view's
)

# MuJoCo's enum values, kept so that model dumps read familiar.
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 0, 2, 3, 6

STEP_TORQUE, STEP_TELEPORT = 0, 1  # AntEnv/SwimmerEnv.step vs PointEnv.step
RESET_POINT, RESET_ANT, RESET_SWIMMER = 0, 1, 2
CELL_WALL, CELL_PLATFORM, CELL_CHASM = 1, 2, 4  # bit flags of grid[]
VIEW_DIM = 75  # 5 x 5 x 3 egocentric raster (maze_env.py:95)

B, J, D, Q, G, A, GO, S, C, O = (
    CAPS["MAXBODY"], CAPS["MAXJNT"], CAPS["MAXDOF"], CAPS["MAXQ"], CAPS["MAXGEOM"],
    CAPS["MAXACT"], CAPS["MAXGOAL"], CAPS["MAXSEG"], CAPS["MAXCELL"], CAPS["MAXOBJ"],
)

# (name, shape, comment)
INT_FIELDS: List[Tuple[str, Tuple[int, ...], str]] = [
    ("magic", (), ""),
    ("version", (), ""),
    ("real_bytes", (), "4 or 8"),
    ("total_bytes", (), "sizeof(mmz_model)"),
    ("nbody", (), "moving bodies"),
    ("njnt", (), ""),
    ("nv", (), ""),
    ("nq", (), ""),
    ("ngeom", (), "geoms on moving bodies"),
    ("nu", (), ""),
    ("ngoal", (), ""),
    ("nseg", (), ""),
    ("grid_h", (), ""),
    ("grid_w", (), ""),
    ("step_kind", (), "MMZ_STEP_*"),
    ("frame_skip", (), ""),
    ("manual_collision", (), "segment clamp on the agent xy (point.py:30)"),
    ("collision_on", (), "0: option collision=predefined with no pairs (swimmer.xml:3)"),
    ("has_floor", (), ""),
    ("elevated", (), ""),
    ("reward_rule", (), "MMZ_REWARD_* (resolved, survey A9)"),
    ("term_rule", (), "MMZ_TERM_*"),
    ("max_episode_steps", (), "TimeLimit (__init__.py:31)"),
    ("obs_dim", (), ""),
    ("n_agent_q", (), "agent qpos entries copied to obs"),
    ("n_agent_v", (), "agent qvel entries copied to obs"),
    ("nobj", (), "observed bodies spliced after obs[:3]"),
    ("reset_kind", (), "MMZ_RESET_*"),
    ("forward_reward_kind", (), "MMZ_FWD_*: forward_reward_fn of the torque agents (ant.py:18-23,44-53)"),
    ("nviewb", (), "bodies latched for the top-down view after the observed ones: torso, then movable blocks"),
    ("view_dim", (), "0, or 75 = 5x5x3 top-down view between the state part of obs and t (maze_env.py:353-369)"),
    ("obj_body", (O,), "nobj observed bodies, then nviewb view bodies"),
    ("body_parent", (B,), "-1 = world"),
    ("body_jntadr", (B,), ""),
    ("body_jntnum", (B,), ""),
    ("body_dofadr", (B,), ""),
    ("body_dofnum", (B,), ""),
    ("body_level", (B,), "depth in its tree, roots = 0"),
    ("body_root", (B,), "root body of its tree"),
    ("body_dofmask", (B,), "bit d set: dof d moves this body"),
    ("jnt_type", (J,), ""),
    ("jnt_body", (J,), ""),
    ("jnt_qadr", (J,), ""),
    ("jnt_dadr", (J,), ""),
    ("jnt_limited", (J,), ""),
    ("dof_body", (D,), ""),
    ("dof_jnt", (D,), ""),
    ("dof_parent", (D,), "-1 = none"),
    ("geom_type", (G,), ""),
    ("geom_body", (G,), ""),
    ("geom_contype", (G,), ""),
    ("geom_conaffinity", (G,), ""),
    ("geom_condim", (G,), ""),
    ("act_dof", (A,), ""),
    ("act_limited", (A,), ""),
    ("goal_dim", (GO,), ""),
    ("grid", (C,), "row-major, bit0 wall box (BLOCK cell), bit1 platform box, bit2 CHASM cell"),
]

REAL_FIELDS: List[Tuple[str, Tuple[int, ...], str]] = [
    ("timestep", (), ""),
    ("gravity", (3,), ""),
    ("density", (), "fluid"),
    ("viscosity", (), "fluid"),
    ("inner_reward_scale", (), "maze_env.py:477"),
    ("forward_reward_weight", (), "ant.py:47"),
    ("ctrl_cost_weight", (), "ant.py:48"),
    ("restitution", (), "maze_env.py:36"),
    ("penalty", (), ""),
    ("task_scale", (), "MazeTask.scale"),
    ("vel_limit", (), "point.py:33"),
    ("reset_noise", (), "0.1"),
    ("cell_size", (), ""),
    ("origin", (2,), "robot cell centre (torso_x, torso_y)"),
    ("wall_half", (3,), "half extents of a wall box"),
    ("wall_z", (), "centre z of wall boxes"),
    ("plat_z", (), "centre z of platform boxes"),
    ("wall_margin", (), ""),
    ("wall_friction", (3,), ""),
    ("wall_solref", (2,), ""),
    ("wall_solimp", (5,), ""),
    ("floor_z", (), ""),
    ("floor_margin", (), ""),
    ("floor_friction", (3,), ""),
    ("floor_solref", (2,), ""),
    ("floor_solimp", (5,), ""),
    ("body_pos", (B, 3), ""),
    ("body_quat", (B, 4), ""),
    ("body_ipos", (B, 3), ""),
    ("body_iquat", (B, 4), ""),
    ("body_mass", (B,), ""),
    ("body_inertia", (B, 3), ""),
    ("jnt_pos", (J, 3), ""),
    ("jnt_axis", (J, 3), ""),
    ("jnt_range", (J, 2), ""),
    ("jnt_margin", (J,), ""),
    ("jnt_solref", (J, 2), ""),
    ("jnt_solimp", (J, 5), ""),
    ("qpos0", (Q,), ""),
    ("dof_armature", (D,), ""),
    ("dof_damping", (D,), ""),
    ("dof_invweight0", (D,), ""),
    ("geom_size", (G, 3), ""),
    ("geom_pos", (G, 3), ""),
    ("geom_quat", (G, 4), ""),
    ("geom_margin", (G,), ""),
    ("geom_friction", (G, 3), ""),
    ("geom_solref", (G, 2), ""),
    ("geom_solimp", (G, 5), ""),
    ("geom_invweight", (G,), "translational body_invweight0 of the geom's (unmerged) body"),
    ("act_gear", (A,), ""),
    ("act_ctrlrange", (A, 2), ""),
    ("goal_pos", (GO, 3), ""),
    ("goal_thr", (GO,), ""),
    ("goal_scale", (GO,), ""),
    ("seg", (S, 4), "x1 y1 x2 y2"),
]


def _count(shape: Tuple[int, ...]) -> int:
    n = 1
    for s in shape:
        n *= s
    return n


N_INT = sum(_count(s) for _, s, _ in INT_FIELDS)
if N_INT % 2:  # keep the real section 8-byte aligned in the double layout
    INT_FIELDS.append(("pad_", (), ""))
    N_INT += 1
N_REAL = sum(_count(s) for _, s, _ in REAL_FIELDS)


def blob_bytes(real_bytes: int) -> int:
    return 4 * N_INT + real_bytes * N_REAL


def pack(fields: Dict[str, np.ndarray], real_bytes: int) -> bytes:
    """Serialise `fields` (name -> array-like) to the blob layout."""
    rdt = np.float32 if real_bytes == 4 else np.float64
    out = bytearray()
    vals = dict(fields)
    vals.update(magic=MAGIC, version=VERSION, real_bytes=real_bytes,
                total_bytes=blob_bytes(real_bytes), pad_=0)
    for table, dt in ((INT_FIELDS, np.int32), (REAL_FIELDS, rdt)):
        for name, shape, _ in table:
            buf = np.zeros(shape, dtype=dt)
            if name in vals:
                src = np.asarray(vals[name], dtype=dt)
                if shape == ():
                    buf[...] = src
                else:
                    if any(a > b for a, b in zip(src.shape, shape)) or src.ndim != len(shape):
                        raise ValueError(f"model field {name}: shape {src.shape} exceeds capacity {shape}")
                    buf[tuple(slice(0, n) for n in src.shape)] = src
            out += buf.tobytes()
    assert len(out) == blob_bytes(real_bytes)
    return bytes(out)


def unpack(blob: bytes) -> Dict[str, np.ndarray]:
    """Inverse of `pack` (used by tests and debugging)."""
    head = np.frombuffer(blob[:16], dtype=np.int32)
    assert head[0] == MAGIC, "bad magic"
    real_bytes = int(head[2])
    rdt = np.float32 if real_bytes == 4 else np.float64
    off, res = 0, {}
    for table, dt in ((INT_FIELDS, np.int32), (REAL_FIELDS, rdt)):
        for name, shape, _ in table:
            n = _count(shape)
            arr = np.frombuffer(blob, dtype=dt, count=n, offset=off).reshape(shape)
            res[name] = arr.copy()
            off += n * np.dtype(dt).itemsize
    return res


def header_text() -> str:
    lines = [
        "/* mmz_model.h - binary layout of the maze model blob.",
        " *",
        " * GENERATED from mujoco-maze_b200/mujoco_maze/model_layout.py - do not edit.",
        " * The blob is produced on the host by the model compiler (the stand-in for",
        " * MuJoCo's MJCF compiler + MazeEnv.__init__ geometry injection,",
        " * reference maze_env.py:97-218) and consumed by mmz_create() (include/mmz.h)",
        " * and by the CPU oracle. Define MMZ_REAL_IS_DOUBLE for the oracle's layout.",
        " */",
        "#ifndef MMZ_MODEL_H",
        "#define MMZ_MODEL_H",
        "#include <stdint.h>",
        "",
        "#ifdef MMZ_REAL_IS_DOUBLE",
        "typedef double mmz_real;",
        "#else",
        "typedef float mmz_real;",
        "#endif",
        "",
        f"#define MMZ_MAGIC 0x{MAGIC:08X}",
        f"#define MMZ_VERSION {VERSION}",
    ]
    for k, v in CAPS.items():
        lines.append(f"#define MMZ_{k} {v}")
    lines += [
        "",
        f"#define MMZ_JNT_FREE {JNT_FREE}",
        f"#define MMZ_JNT_BALL {JNT_BALL}",
        f"#define MMZ_JNT_SLIDE {JNT_SLIDE}",
        f"#define MMZ_JNT_HINGE {JNT_HINGE}",
        f"#define MMZ_GEOM_PLANE {GEOM_PLANE}",
        f"#define MMZ_GEOM_SPHERE {GEOM_SPHERE}",
        f"#define MMZ_GEOM_CAPSULE {GEOM_CAPSULE}",
        f"#define MMZ_GEOM_BOX {GEOM_BOX}",
        f"#define MMZ_STEP_TORQUE {STEP_TORQUE}",
        f"#define MMZ_STEP_TELEPORT {STEP_TELEPORT}",
        f"#define MMZ_RESET_POINT {RESET_POINT}",
        f"#define MMZ_RESET_ANT {RESET_ANT}",
        f"#define MMZ_RESET_SWIMMER {RESET_SWIMMER}",
        f"#define MMZ_CELL_WALL {CELL_WALL}",
        f"#define MMZ_CELL_PLATFORM {CELL_PLATFORM}",
        f"#define MMZ_CELL_CHASM {CELL_CHASM}",
        f"#define MMZ_VIEW_DIM {VIEW_DIM}",
        "/* resolved reward / termination rules (SURVEY.md section 8(a) row A9) */",
        "#define MMZ_REWARD_REACH 0         /* 1.0 if terminated else penalty        maze_task.py:110-111 */",
        "#define MMZ_REWARD_SCALED 1        /* first reached goal's reward_scale     maze_task.py:356-360 */",
        "#define MMZ_REWARD_SCALED_OBJECT 2 /* same on obs[3:6]                      maze_task.py:592-597 */",
        "#define MMZ_REWARD_DIST_OBJECT 3   /* -|obs[3:6]-goal0|/scale               maze_task.py:619-621 */",
        "#define MMZ_REWARD_ZERO 4          /* NoReward*                                                  */",
        "#define MMZ_REWARD_DIST 5          /* -|obs[:dim]-goal0|/scale              maze_task.py:98-99   */",
        "#define MMZ_REWARD_HOST 6          /* user-defined: outer reward left to the host wrapper       */",
        "#define MMZ_TERM_AGENT 0           /* any goal within threshold of obs[:dim] maze_task.py:77-81  */",
        "#define MMZ_TERM_OBJECT 1          /* ... of obs[3:6]                       maze_task.py:599-604 */",
        "#define MMZ_TERM_HOST 2",
        "/* forward_reward_fn of AntEnv / SwimmerEnv (reference ant.py:18-23) on the xy velocity of the step */",
        "#define MMZ_FWD_VNORM 0            /* forward_reward_vnorm: |v|, the default                     */",
        "#define MMZ_FWD_VABS 1             /* forward_reward_vabs: |vx| + |vy|                           */",
        "#define MMZ_FWD_HOST 2             /* any other callable: the host wrapper adds weight * fn(v)   */",
        "",
        "typedef struct mmz_model {",
    ]

    def decl(ctype, name, shape, comment):
        dims = "".join(f"[{s}]" for s in shape)
        c = f" /* {comment} */" if comment else ""
        return f"  {ctype} {name}{dims};{c}"

    for name, shape, comment in INT_FIELDS:
        lines.append(decl("int32_t", name, shape, comment))
    for name, shape, comment in REAL_FIELDS:
        lines.append(decl("mmz_real", name, shape, comment))
    lines += [
        "} mmz_model;",
        "",
        f"#define MMZ_MODEL_NINT {N_INT}",
        f"#define MMZ_MODEL_NREAL {N_REAL}",
        "",
        "#endif /* MMZ_MODEL_H */",
        "",
    ]
    return "\n".join(lines)


if __name__ == "__main__":
    print(header_text(), end="")
