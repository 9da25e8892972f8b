"""Host-side model compiler: MJCF asset + maze task -> constant model blob.

Cold path, runs once per env id. It replaces two things the reference does at
construction time:
  * `MazeEnv.__init__` rewriting the agent MJCF with one box geom per wall cell,
    movable-block bodies and goal sites (reference maze_env.py:97-218, 563-660);
  * MuJoCo's own MJCF compiler [EXT, not in the reference tree]: default
    classes, `fromto` capsules, `inertiafromgeom`, `angle="degree"`, and the
    `invweight0` tables used by the soft-constraint regulariser.
The output is the flat `mmz_model` struct (include/mmz_model.h) that
`mmz_create` uploads and the step kernel stages into shared memory.

Only the MJCF subset used by assets/{point,ant,swimmer}.xml is understood:
one top-level <default> with <joint>/<geom>, nested <body> trees with
free/slide/hinge joints, plane/sphere/capsule/box geoms, <motor> actuators.
"""

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from mujoco_maze import model_layout as L
from mujoco_maze.maze_env_utils import CollisionDetector, MazeCell

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
MJ_MINVAL = 1e-15

# MuJoCo 2.0 built-in defaults [EXT] for the attributes the assets leave unset.
_GEOM_BUILTIN = dict(
    type="sphere", contype="1", conaffinity="1", condim="3", margin="0",
    friction="1 0.005 0.0001", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2",
    density="1000", pos="0 0 0",
)
_JOINT_BUILTIN = dict(
    type="hinge", limited="false", armature="0", damping="0", margin="0",
    pos="0 0 0", axis="0 0 1", range="0 0",
    solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2",
)


# ---------------------------------------------------------------------------
# small quaternion / rotation helpers (w, x, y, z), float64
# ---------------------------------------------------------------------------
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def mat_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(max(1.0 + R[i, i] - R[j, j] - R[k, k], 0.0)) * 2
        q = [0.0] * 4
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q = np.array(q)
    return q / np.linalg.norm(q)


def quat_z_to_vec(v):
    """Rotation taking +z onto direction v (how MJCF `fromto` orients a capsule)."""
    v = np.asarray(v, float)
    v = v / np.linalg.norm(v)
    axis = np.cross([0.0, 0.0, 1.0], v)
    s = np.linalg.norm(axis)
    ang = math.atan2(s, v[2])
    axis = np.array([1.0, 0.0, 0.0]) if s < 1e-12 else axis / s
    return np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * axis])


def _floats(text: str, n: Optional[int] = None) -> np.ndarray:
    cleaned = "".join(ch if (ch.isdigit() or ch in "+-.eE ") else " " for ch in text)  # quirk Q12: "0 0s 1.3"
    vals = np.array([float(t) for t in cleaned.split()])
    if n is not None and len(vals) < n:
        vals = np.concatenate([vals, np.zeros(n - len(vals))])
    return vals


# ---------------------------------------------------------------------------
# scene description
# ---------------------------------------------------------------------------
@dataclass
class Geom:
    name: str
    type: int
    size: np.ndarray  # 3
    pos: np.ndarray
    quat: np.ndarray
    contype: int
    conaffinity: int
    condim: int
    margin: float
    friction: np.ndarray  # 3
    solref: np.ndarray  # 2
    solimp: np.ndarray  # 5
    density: float
    mass: Optional[float] = None


@dataclass
class Joint:
    name: str
    type: int
    pos: np.ndarray
    axis: np.ndarray
    limited: bool
    range: np.ndarray
    armature: float
    damping: float
    margin: float
    solref: np.ndarray
    solimp: np.ndarray


@dataclass
class Body:
    name: str
    pos: np.ndarray
    quat: np.ndarray
    parent: int  # index into Scene.bodies, -1 = world
    joints: List[Joint] = field(default_factory=list)
    geoms: List[Geom] = field(default_factory=list)


@dataclass
class Actuator:
    joint: str
    gear: float
    ctrlrange: np.ndarray
    limited: bool


@dataclass
class Scene:
    timestep: float = 0.002
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    density: float = 0.0
    viscosity: float = 0.0
    collision_on: bool = True
    world_geoms: List[Geom] = field(default_factory=list)
    bodies: List[Body] = field(default_factory=list)
    actuators: List[Actuator] = field(default_factory=list)
    geom_default: Dict[str, str] = field(default_factory=dict)
    joint_default: Dict[str, str] = field(default_factory=dict)
    degrees: bool = True

    def body_index(self, name: str) -> int:
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        raise KeyError(name)


_GEOM_TYPES = {"plane": L.GEOM_PLANE, "sphere": L.GEOM_SPHERE, "capsule": L.GEOM_CAPSULE, "box": L.GEOM_BOX}
_JNT_TYPES = {"free": L.JNT_FREE, "ball": L.JNT_BALL, "slide": L.JNT_SLIDE, "hinge": L.JNT_HINGE}


def make_geom(attrs: Dict[str, str], defaults: Dict[str, str]) -> Geom:
    a = dict(_GEOM_BUILTIN)
    a.update(defaults)
    a.update(attrs)
    gtype = _GEOM_TYPES[a["type"]]
    size = _floats(a.get("size", "0"), 3)[:3]
    pos = _floats(a["pos"], 3)[:3]
    quat = _floats(a["quat"], 4)[:4] if "quat" in a else np.array([1.0, 0, 0, 0])
    if "fromto" in a:
        ft = _floats(a["fromto"], 6)
        p0, p1 = ft[:3], ft[3:6]
        pos = 0.5 * (p0 + p1)
        quat = quat_z_to_vec(p1 - p0)
        size = np.array([size[0], 0.5 * np.linalg.norm(p1 - p0), 0.0])
    solimp = _floats(a["solimp"])
    solimp = np.concatenate([solimp, [0.9, 0.95, 0.001, 0.5, 2.0][len(solimp):]])
    return Geom(
        name=a.get("name", ""), type=gtype, size=size, pos=pos, quat=quat,
        contype=int(a["contype"]), conaffinity=int(a["conaffinity"]), condim=int(a["condim"]),
        margin=float(a["margin"]), friction=_floats(a["friction"], 3)[:3],
        solref=_floats(a["solref"], 2)[:2], solimp=solimp, density=float(a["density"]),
        mass=float(a["mass"]) if "mass" in a else None,
    )


def make_joint(attrs: Dict[str, str], defaults: Dict[str, str], degrees: bool) -> Joint:
    a = dict(_JOINT_BUILTIN)
    a.update(defaults)
    a.update(attrs)
    jtype = _JNT_TYPES[a["type"]]
    rng = _floats(a["range"], 2)[:2]
    if degrees and jtype == L.JNT_HINGE:
        rng = np.deg2rad(rng)
    axis = _floats(a["axis"], 3)[:3]
    n = np.linalg.norm(axis)
    axis = axis / n if n > 0 else np.array([0.0, 0.0, 1.0])
    solimp = _floats(a["solimplimit"])
    solimp = np.concatenate([solimp, [0.9, 0.95, 0.001, 0.5, 2.0][len(solimp):]])
    limited = a["limited"].lower() == "true" and jtype in (L.JNT_SLIDE, L.JNT_HINGE)
    return Joint(
        name=a.get("name", ""), type=jtype, pos=_floats(a["pos"], 3)[:3], axis=axis,
        limited=limited, range=rng, armature=float(a["armature"]), damping=float(a["damping"]),
        margin=float(a["margin"]), solref=_floats(a["solreflimit"], 2)[:2], solimp=solimp,
    )


def parse_mjcf(path: str) -> Scene:
    root = ET.parse(path).getroot()
    sc = Scene()
    comp = root.find("compiler")
    if comp is not None:
        sc.degrees = comp.get("angle", "degree") == "degree"
    opt = root.find("option")
    if opt is not None:
        sc.timestep = float(opt.get("timestep", sc.timestep))
        sc.density = float(opt.get("density", 0.0))
        sc.viscosity = float(opt.get("viscosity", 0.0))
        if "gravity" in opt.attrib:
            sc.gravity = _floats(opt.get("gravity"), 3)[:3]
        # collision="predefined" with no <contact><pair>: nothing ever collides (quirk Q8)
        sc.collision_on = opt.get("collision", "all") != "predefined" or root.find("contact/pair") is not None
        if opt.get("integrator", "Euler") != "RK4":
            raise NotImplementedError("only integrator=RK4 is supported (all reference assets use it)")
    dflt = root.find("default")
    if dflt is not None:
        g, j = dflt.find("geom"), dflt.find("joint")
        sc.geom_default = dict(g.attrib) if g is not None else {}
        sc.joint_default = dict(j.attrib) if j is not None else {}
    world = root.find("worldbody")

    def walk(elem, parent: int):
        for child in elem:
            if child.tag == "geom":
                g = make_geom(child.attrib, sc.geom_default)
                (sc.world_geoms if parent < 0 else sc.bodies[parent].geoms).append(g)
            elif child.tag in ("joint", "freejoint"):
                attrs = dict(child.attrib)
                dflts = sc.joint_default
                if child.tag == "freejoint":
                    attrs["type"], dflts = "free", {}
                sc.bodies[parent].joints.append(make_joint(attrs, dflts, sc.degrees))
            elif child.tag == "body":
                b = Body(
                    name=child.get("name", f"body{len(sc.bodies)}"),
                    pos=_floats(child.get("pos", "0 0 0"), 3)[:3],
                    quat=_floats(child.get("quat", "1 0 0 0"), 4)[:4],
                    parent=parent,
                )
                sc.bodies.append(b)
                walk(child, len(sc.bodies) - 1)

    walk(world, -1)
    act = root.find("actuator")
    if act is not None:
        for m in act:
            if m.tag != "motor":
                raise NotImplementedError(f"actuator <{m.tag}>")
            sc.actuators.append(Actuator(
                joint=m.get("joint"), gear=_floats(m.get("gear", "1"))[0],
                ctrlrange=_floats(m.get("ctrlrange", "0 0"), 2)[:2],
                limited=m.get("ctrllimited", "false").lower() == "true",
            ))
    return sc


# ---------------------------------------------------------------------------
# inertia from geoms
# ---------------------------------------------------------------------------
def geom_mass_inertia(g: Geom, legacy_capsule_volume: bool) -> Tuple[float, np.ndarray]:
    """Mass and diagonal inertia of a geom in its own frame (MuJoCo `inertiafromgeom`)."""
    if g.type == L.GEOM_SPHERE:
        r = g.size[0]
        vol = 4.0 / 3.0 * math.pi * r ** 3
        m = g.mass if g.mass is not None else g.density * vol
        return m, np.full(3, 0.4 * m * r * r)
    if g.type == L.GEOM_BOX:
        sx, sy, sz = g.size
        vol = 8 * sx * sy * sz
        m = g.mass if g.mass is not None else g.density * vol
        return m, m / 3.0 * np.array([sy * sy + sz * sz, sx * sx + sz * sz, sx * sx + sy * sy])
    if g.type == L.GEOM_CAPSULE:
        r, h = g.size[0], g.size[1]
        height = 2 * h
        # MuJoCo <= 2.0 used pi r^2 L + pi r^3 (reproduces gym Ant-v2 body_mass 0.036477 / 0.064911);
        # later releases use the exact pi r^2 L + 4/3 pi r^3. SURVEY.md appendix A.1.
        cap = math.pi * r ** 3 if legacy_capsule_volume else 4.0 / 3.0 * math.pi * r ** 3
        vol = math.pi * r * r * height + cap
        m = g.mass if g.mass is not None else g.density * vol
        ms = m * 4 * r / (4 * r + 3 * height)  # two half spheres
        mc = m - ms
        it = mc * (3 * r * r + height * height) / 12 + 0.4 * ms * r * r + ms * height * (3 * r + 2 * height) / 8
        ia = mc * r * r / 2 + 0.4 * ms * r * r
        return m, np.array([it, it, ia])
    return 0.0, np.zeros(3)


def _body_inertial(elements) -> Tuple[float, np.ndarray, np.ndarray, np.ndarray]:
    """Combine (mass, com, 3x3 inertia about com) elements -> mass, ipos, iquat, principal moments."""
    mass = sum(m for m, _, _ in elements)
    if mass <= 0:
        return 0.0, np.zeros(3), np.array([1.0, 0, 0, 0]), np.zeros(3)
    com = sum(m * c for m, c, _ in elements) / mass
    I = np.zeros((3, 3))
    for m, c, Ic in elements:
        d = c - com
        I += Ic + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    if np.allclose(I, np.diag(np.diag(I)), atol=1e-14 * max(1.0, np.abs(I).max())):
        return mass, com, np.array([1.0, 0, 0, 0]), np.diag(I).copy()
    w, V = np.linalg.eigh(I)
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    return mass, com, mat_to_quat(V), w


# ---------------------------------------------------------------------------
# flattened arrays + numpy kinematics / mass matrix (used for invweight0 and tests)
# ---------------------------------------------------------------------------
def flatten(sc: Scene, legacy_capsule_volume: bool = True) -> Dict[str, np.ndarray]:
    nb = len(sc.bodies)
    f: Dict[str, list] = {k: [] for k in (
        "body_parent", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum", "body_level", "body_root",
        "body_dofmask", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia",
        "jnt_type", "jnt_body", "jnt_qadr", "jnt_dadr", "jnt_limited", "jnt_pos", "jnt_axis", "jnt_range",
        "jnt_margin", "jnt_solref", "jnt_solimp", "qpos0", "dof_body", "dof_jnt", "dof_parent", "dof_armature",
        "dof_damping", "geom_type", "geom_body", "geom_contype", "geom_conaffinity", "geom_condim", "geom_size",
        "geom_pos", "geom_quat", "geom_margin", "geom_friction", "geom_solref", "geom_solimp",
    )}
    names = dict(body=[], jnt=[], geom=[])
    last_dof_of_body: List[int] = []
    nq = nv = 0
    for bi, b in enumerate(sc.bodies):
        p = b.parent
        f["body_parent"].append(p)
        f["body_level"].append(0 if p < 0 else f["body_level"][p] + 1)
        f["body_root"].append(bi if p < 0 else f["body_root"][p])
        f["body_pos"].append(b.pos)
        f["body_quat"].append(b.quat / np.linalg.norm(b.quat))
        names["body"].append(b.name)
        elems = []
        for g in b.geoms:
            m, Id = geom_mass_inertia(g, legacy_capsule_volume)
            R = quat_to_mat(g.quat)
            elems.append((m, g.pos, R @ np.diag(Id) @ R.T))
        mass, ipos, iquat, inertia = _body_inertial(elems)
        f["body_mass"].append(mass)
        f["body_ipos"].append(ipos)
        f["body_iquat"].append(iquat)
        f["body_inertia"].append(inertia)
        f["body_jntadr"].append(len(f["jnt_type"]))
        f["body_jntnum"].append(len(b.joints))
        f["body_dofadr"].append(nv)
        prev_dof = -1 if p < 0 else last_dof_of_body[p]
        mask = 0 if p < 0 else f["body_dofmask"][p]
        for j in b.joints:
            names["jnt"].append(j.name)
            f["jnt_type"].append(j.type)
            f["jnt_body"].append(bi)
            f["jnt_qadr"].append(nq)
            f["jnt_dadr"].append(nv)
            f["jnt_limited"].append(int(j.limited))
            f["jnt_pos"].append(j.pos)
            f["jnt_axis"].append(j.axis)
            f["jnt_range"].append(j.range)
            f["jnt_margin"].append(j.margin)
            f["jnt_solref"].append(j.solref)
            f["jnt_solimp"].append(j.solimp)
            if j.type == L.JNT_FREE:
                if p >= 0:
                    raise ValueError("free joint below another moving body")
                f["qpos0"] += list(b.pos) + list(b.quat / np.linalg.norm(b.quat))
                ndof, nqj = 6, 7
            elif j.type == L.JNT_BALL:
                raise NotImplementedError("ball joints (SPIN blocks) are not supported")
            else:
                f["qpos0"].append(0.0)
                ndof, nqj = 1, 1
            for _ in range(ndof):
                f["dof_body"].append(bi)
                f["dof_jnt"].append(len(f["jnt_type"]) - 1)
                f["dof_parent"].append(prev_dof)
                f["dof_armature"].append(j.armature)
                f["dof_damping"].append(j.damping)
                mask |= 1 << nv
                prev_dof = nv
                nv += 1
            nq += nqj
        f["body_dofnum"].append(nv - f["body_dofadr"][-1])
        f["body_dofmask"].append(mask)
        last_dof_of_body.append(prev_dof)
        for g in b.geoms:
            names["geom"].append(g.name)
            f["geom_type"].append(g.type)
            f["geom_body"].append(bi)
            f["geom_contype"].append(g.contype)
            f["geom_conaffinity"].append(g.conaffinity)
            f["geom_condim"].append(g.condim)
            f["geom_size"].append(g.size)
            f["geom_pos"].append(g.pos)
            f["geom_quat"].append(g.quat / np.linalg.norm(g.quat))
            f["geom_margin"].append(g.margin)
            f["geom_friction"].append(g.friction)
            f["geom_solref"].append(g.solref)
            f["geom_solimp"].append(g.solimp)
    arr = {}
    for k, v in f.items():
        is_int = k.split("_")[-1] in ("parent", "jntadr", "jntnum", "dofadr", "dofnum", "level", "root", "dofmask",
                                      "type", "body", "qadr", "dadr", "limited", "jnt", "contype", "conaffinity",
                                      "condim")
        arr[k] = np.array(v, dtype=np.int64 if is_int else np.float64)
        if len(v) == 0:
            arr[k] = np.zeros((0,), dtype=arr[k].dtype)
    arr.update(nbody=nb, njnt=len(f["jnt_type"]), nv=nv, nq=nq, ngeom=len(f["geom_type"]))
    arr["_names"] = names
    return arr


def np_kinematics(m: Dict[str, np.ndarray], qpos: np.ndarray):
    """World pose of every body: xpos [nb,3], xquat [nb,4] (numpy, float64)."""
    nb = m["nbody"]
    xpos, xquat = np.zeros((nb, 3)), np.zeros((nb, 4))
    for b in range(nb):
        p = m["body_parent"][b]
        if p < 0:
            ppos, pquat = np.zeros(3), np.array([1.0, 0, 0, 0])
        else:
            ppos, pquat = xpos[p], xquat[p]
        pos = ppos + quat_to_mat(pquat) @ m["body_pos"][b]
        quat = quat_mul(pquat, m["body_quat"][b])
        for j in range(m["body_jntadr"][b], m["body_jntadr"][b] + m["body_jntnum"][b]):
            qa, t = m["jnt_qadr"][j], m["jnt_type"][j]
            if t == L.JNT_FREE:
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
            elif t == L.JNT_SLIDE:
                pos = pos + quat_to_mat(quat) @ m["jnt_axis"][j] * (qpos[qa] - m["qpos0"][qa])
            elif t == L.JNT_HINGE:
                R = quat_to_mat(quat)
                anchor = pos + R @ m["jnt_pos"][j]
                ang = qpos[qa] - m["qpos0"][qa]
                dq = np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * m["jnt_axis"][j]])
                quat = quat_mul(quat, dq)
                pos = anchor - quat_to_mat(quat) @ m["jnt_pos"][j]
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    return xpos, xquat


def np_jacobian(m, qpos, xpos, xquat, body: int, point: np.ndarray):
    """Translational and rotational Jacobians (3 x nv each) of `point` fixed to `body`."""
    nv = m["nv"]
    jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
    b = body
    while b >= 0:
        R = quat_to_mat(xquat[b])
        # joints of body b act in the frame *before* the later joints of the same body; for the
        # single-joint and free-joint bodies used here the body frame after all joints is equivalent.
        for j in range(m["body_jntadr"][b], m["body_jntadr"][b] + m["body_jntnum"][b]):
            d, t = m["jnt_dadr"][j], m["jnt_type"][j]
            if t == L.JNT_FREE:
                jp[:, d:d + 3] = np.eye(3)
                for k in range(3):
                    ax = R[:, k]
                    jr[:, d + 3 + k] = ax
                    jp[:, d + 3 + k] = np.cross(ax, point - xpos[b])
            else:
                # axis direction: slide axes of earlier joints are unaffected by later hinges only if
                # they precede them; the body frame is exact for hinges, and slides here always come
                # first on z-rotating bodies whose slide axes are x / y: use the pre-hinge frame.
                ax = _joint_axis_world(m, qpos, xquat, b, j)
                if t == L.JNT_SLIDE:
                    jp[:, d] = ax
                else:
                    anchor = xpos[b] + R @ m["jnt_pos"][j]
                    jr[:, d] = ax
                    jp[:, d] = np.cross(ax, point - anchor)
        b = m["body_parent"][b]
    return jp, jr


def _joint_axis_world(m, qpos, xquat, b: int, j: int) -> np.ndarray:
    """World direction of joint j's axis: body frame with the *later* hinges of the same body undone."""
    quat = xquat[b]
    for jj in range(m["body_jntadr"][b] + m["body_jntnum"][b] - 1, j, -1):
        if m["jnt_type"][jj] == L.JNT_HINGE:
            qa = m["jnt_qadr"][jj]
            ang = -(qpos[qa] - m["qpos0"][qa])
            dq = np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * m["jnt_axis"][jj]])
            quat = quat_mul(quat, dq)
    return quat_to_mat(quat) @ m["jnt_axis"][j]


def np_mass_matrix(m, qpos) -> np.ndarray:
    """M(q) = sum_b m Jp^T Jp + Jr^T I_world Jr + diag(armature): an independent check of CRB."""
    xpos, xquat = np_kinematics(m, qpos)
    nv = m["nv"]
    M = np.diag(m["dof_armature"].astype(float)) if nv else np.zeros((0, 0))
    for b in range(m["nbody"]):
        if m["body_mass"][b] <= 0:
            continue
        R = quat_to_mat(xquat[b])
        com = xpos[b] + R @ m["body_ipos"][b]
        Ri = R @ quat_to_mat(m["body_iquat"][b])
        Iw = Ri @ np.diag(m["body_inertia"][b]) @ Ri.T
        jp, jr = np_jacobian(m, qpos, xpos, xquat, b, com)
        M = M + m["body_mass"][b] * jp.T @ jp + jr.T @ Iw @ jr
    return M


def compute_invweight0(m) -> Tuple[np.ndarray, np.ndarray]:
    """(body_invweight0 [nb,2], dof_invweight0 [nv]) at qpos0 — MuJoCo's `mj_setConst` tables [EXT]."""
    qpos0 = m["qpos0"]
    nv = m["nv"]
    M = np_mass_matrix(m, qpos0)
    Minv = np.linalg.inv(M) if nv else M
    xpos, xquat = np_kinematics(m, qpos0)
    bw = np.zeros((m["nbody"], 2))
    for b in range(m["nbody"]):
        com = xpos[b] + quat_to_mat(xquat[b]) @ m["body_ipos"][b]
        jp, jr = np_jacobian(m, qpos0, xpos, xquat, b, com)
        bw[b, 0] = max(MJ_MINVAL, np.trace(jp @ Minv @ jp.T) / 3)
        bw[b, 1] = max(MJ_MINVAL, np.trace(jr @ Minv @ jr.T) / 3)
    dw = np.diag(Minv).copy() if nv else np.zeros(0)
    for j in range(m["njnt"]):
        if m["jnt_type"][j] == L.JNT_FREE:
            d = m["jnt_dadr"][j]
            dw[d:d + 3] = dw[d:d + 3].mean()
            dw[d + 3:d + 6] = dw[d + 3:d + 6].mean()
    return bw, dw


def merge_welded_bodies(sc: Scene) -> Tuple[Scene, List[int]]:
    """Fold joint-less child bodies into their parents (same dynamics, fewer bodies).

    Returns the new scene and, for every geom of the new scene in order, the index
    of the body it belonged to in the old scene (needed for per-geom invweight).
    """
    keep = [i for i, b in enumerate(sc.bodies) if b.joints or b.parent < 0]
    remap, xform = {}, {}  # old body -> (new body index, pos, quat) of old frame in new body's frame
    new_bodies: List[Body] = []
    geom_src: List[List[int]] = []
    for i, b in enumerate(sc.bodies):
        if i in keep:
            if b.parent < 0:
                npar, pos, quat = -1, b.pos, b.quat
            else:
                npar, ppos, pquat = xform[b.parent]
                pos = ppos + quat_to_mat(pquat) @ b.pos
                quat = quat_mul(pquat, b.quat)
            nb = Body(name=b.name, pos=pos, quat=quat, parent=npar, joints=list(b.joints), geoms=list(b.geoms))
            new_bodies.append(nb)
            geom_src.append([i] * len(b.geoms))
            xform[i] = (len(new_bodies) - 1, np.zeros(3), np.array([1.0, 0, 0, 0]))
        else:
            host, ppos, pquat = xform[b.parent]
            pos = ppos + quat_to_mat(pquat) @ b.pos
            quat = quat_mul(pquat, b.quat)
            xform[i] = (host, pos, quat)
            R = quat_to_mat(quat)
            for g in b.geoms:
                g2 = Geom(**{**g.__dict__})
                g2.pos = pos + R @ g.pos
                g2.quat = quat_mul(quat, g.quat)
                new_bodies[host].geoms.append(g2)
                geom_src[host].append(i)
    out = Scene(**{**sc.__dict__})
    out.bodies = new_bodies
    return out, [s for lst in geom_src for s in lst]


# ---------------------------------------------------------------------------
# maze injection + final model
# ---------------------------------------------------------------------------
class MazeModel:
    """Compiled model: `fields` (numpy, float64/int64) + blob serialisation."""

    def __init__(self, fields: Dict[str, np.ndarray], names: Dict[str, List[str]], meta: Dict):
        self.fields = fields
        self.names = names
        self.meta = meta

    def __getattr__(self, k):
        try:
            return self.__dict__["fields"][k]
        except KeyError:
            raise AttributeError(k)

    def blob(self, real_bytes: int = 4) -> bytes:
        return L.pack(self.fields, real_bytes)


def add_movable_block(sc: Scene, cell: MazeCell, i: int, j: int, s: float, x: float, y: float, h: float,
                      height_offset: float) -> str:
    """Movable block body (reference maze_env.py:563-660)."""
    if cell.can_spin():
        raise NotImplementedError("SPIN blocks (ball joint) are unused upstream and not supported")
    falling = cell.can_move_z()
    shrink = 0.99 if falling else (0.5 if cell.is_half_block() else 1.0)
    half = 0.5 * s * shrink
    name = f"movable_{i}_{j}"
    body = Body(name=name, pos=np.array([x, y, h]), quat=np.array([1.0, 0, 0, 0]), parent=-1)
    body.geoms.append(make_geom(
        dict(name=f"block_{i}_{j}", type="box", pos="0 0 0", size=f"{half} {half} {h}",
             mass="0.001" if falling else "0.0002", contype="1", conaffinity="1"),
        sc.geom_default))
    common = dict(type="slide", armature="0", damping="0.0", margin="0.01", pos="0 0 0",
                  limited="true" if falling else "false", range=f"{-s} {s}")
    if cell.can_move_x():
        body.joints.append(make_joint(dict(common, name=f"movable_x_{i}_{j}", axis="1 0 0"), sc.joint_default, False))
    if cell.can_move_y():
        body.joints.append(make_joint(dict(common, name=f"movable_y_{i}_{j}", axis="0 1 0"), sc.joint_default, False))
    if cell.can_move_z():
        body.joints.append(make_joint(
            dict(common, name=f"movable_z_{i}_{j}", axis="0 0 1", limited="true", range=f"{-height_offset} 0"),
            sc.joint_default, False))
    sc.bodies.append(body)
    return name


def add_object_ball(sc: Scene, kind: Optional[str], i: int, j: int, x: float, y: float, size: float) -> str:
    """Object ball body (reference maze_env.py:489-560): a sphere of radius `size` resting on the floor.

    "hinge" (Point, point.py:32): slide x, slide y and an unlimited hinge about z, mass 1e-4 * size^3;
    "freejoint" (Ant, ant.py:42): a free joint, mass from the default density. solimp 0.9 0.99 0.001 in both."""
    name = f"objball_{i}_{j}"
    body = Body(name=name, pos=np.array([x, y, 0.0]), quat=np.array([1.0, 0, 0, 0]), parent=-1)
    attrs = dict(name=f"{name}_geom", type="sphere", size=f"{size}", pos=f"0.0 0.0 {size}", contype="1", conaffinity="1",
                 solimp="0.9 0.99 0.001")
    if kind == "hinge":
        attrs["mass"] = f"{0.0001 * size ** 3}"
        body.geoms.append(make_geom(attrs, sc.geom_default))
        for jn, ax in (("x", "1 0 0"), ("y", "0 1 0")):
            body.joints.append(make_joint(dict(name=f"{name}_{jn}", axis=ax, pos="0 0 0", type="slide"), sc.joint_default, False))
        body.joints.append(make_joint(dict(name=f"{name}_rot", axis="0 0 1", pos="0 0 0", type="hinge", limited="false"),
                                      sc.joint_default, False))
    elif kind == "freejoint":
        body.geoms.append(make_geom(attrs, sc.geom_default))
        # <freejoint> takes no defaults (armature 0, damping 0) [EXT: MJCF reference]
        body.joints.append(make_joint(dict(name=f"{name}_root", type="free"), {}, True))
    else:
        raise ValueError(f"OBJBALL_TYPE is not registered for {kind}")  # reference maze_env.py:188-191
    sc.bodies.append(body)
    return name


def compile_maze_model(
    agent,  # AgentModel subclass (class attributes only are read)
    task,  # MazeTask instance
    maze_size_scaling: float,
    maze_height: float = 0.5,
    inner_reward_scaling: float = 1.0,
    restitution_coef: float = 0.8,
    forward_reward_weight: float = 1.0,
    ctrl_cost_weight: float = 1e-4,
    forward_reward_kind: int = 0,
    max_episode_steps: int = 1000,
    merge_welded: bool = True,
    legacy_capsule_volume: bool = True,
) -> MazeModel:
    from mujoco_maze import maze_task as mt

    s = float(maze_size_scaling)
    structure = task.create_maze()
    rows, cols = len(structure), len(structure[0])
    if rows * cols > L.CAPS["MAXCELL"]:
        raise ValueError(f"maze of {rows}x{cols} cells exceeds MAXCELL={L.CAPS['MAXCELL']}")
    robots = [(j * s, i * s) for i in range(rows) for j in range(cols) if structure[i][j].is_robot()]
    if not robots:
        raise ValueError("No robot in maze specification.")
    torso_x, torso_y = robots[0]
    elevated = any(c.is_chasm() for row in structure for c in row)
    has_blocks = any(c.can_move() for row in structure for c in row)

    sc = parse_mjcf(os.path.join(ASSET_DIR, agent.FILE))
    height_offset = 0.0
    if elevated:
        height_offset = maze_height * s
        torso = sc.bodies[sc.body_index("torso")]
        torso.pos = np.array([0.0, 0.0, float(f"{0.75 + height_offset:.2f}")])
    explicit_solimp = set()
    if has_blocks:
        # reference maze_env.py:108-112: the default class gets solimp .995 .995 .01, so every geom that
        # did not spell out its own solimp changes. Re-parse with the new default to honour that rule.
        sc_new = dict(sc.geom_default, solimp=".995 .995 .01")
        root = ET.parse(os.path.join(ASSET_DIR, agent.FILE)).getroot()
        explicit_solimp = {g.get("name") for g in root.iter("geom") if "solimp" in g.attrib}
        new_solimp = make_geom({}, sc_new).solimp
        for g in sc.world_geoms + [g for b in sc.bodies for g in b.geoms]:
            if g.name not in explicit_solimp:
                g.solimp = new_solimp.copy()
        sc.geom_default = sc_new

    h = maze_height / 2 * s
    grid = np.zeros(rows * cols, dtype=np.int64)
    obj_names: List[str] = []
    ball_names: List[str] = []
    for i in range(rows):
        for j in range(cols):
            cell = structure[i][j]
            if cell.is_robot() and task.PUT_SPIN_NEAR_AGENT:
                cell = MazeCell.SPIN
            x, y = j * s - torso_x, i * s - torso_y
            if elevated and not cell.is_chasm():
                grid[i * cols + j] |= L.CELL_PLATFORM
            if cell.is_chasm():
                grid[i * cols + j] |= L.CELL_CHASM
            if cell.is_block():
                grid[i * cols + j] |= L.CELL_WALL
            elif cell.can_move():
                obj_names.append(add_movable_block(sc, cell, i, j, s, x, y, h, height_offset))
            elif cell.is_object_ball():
                ball_names.append(add_object_ball(sc, agent.OBJBALL_TYPE, i, j, x, y, float(task.OBJECT_BALL_SIZE)))

    wall = make_geom(dict(type="box", contype="1", conaffinity="1"), sc.geom_default)
    floor = next((g for g in sc.world_geoms if g.type == L.GEOM_PLANE), None)

    # --- flatten (unmerged for invweight0), then optionally merge welded bodies
    flat_full = flatten(sc, legacy_capsule_volume)
    bw_full, dw = compute_invweight0(flat_full)
    if merge_welded:
        sc_m, geom_src = merge_welded_bodies(sc)
        flat = flatten(sc_m, legacy_capsule_volume)
    else:
        sc_m, flat = sc, flat_full
        geom_src = list(flat_full["geom_body"])
    names = flat.pop("_names")
    flat_full.pop("_names", None)
    assert flat["nv"] == flat_full["nv"] and np.allclose(flat["qpos0"], flat_full["qpos0"])
    for cap, key in (("MAXBODY", "nbody"), ("MAXJNT", "njnt"), ("MAXDOF", "nv"), ("MAXQ", "nq"), ("MAXGEOM", "ngeom")):
        if flat[key] > L.CAPS[cap]:
            raise ValueError(f"model has {key}={flat[key]} > {cap}={L.CAPS[cap]}")

    fields: Dict[str, np.ndarray] = dict(flat)
    fields["dof_invweight0"] = dw
    fields["geom_invweight"] = np.array([bw_full[b, 0] for b in geom_src])

    # --- actuators
    jnt_names = names["jnt"]
    act_dof, act_gear, act_range, act_lim = [], [], [], []
    for a in sc.actuators:
        jid = jnt_names.index(a.joint)
        act_dof.append(int(flat["jnt_dadr"][jid]))
        act_gear.append(a.gear)
        act_range.append(a.ctrlrange)
        act_lim.append(int(a.limited))
    fields.update(nu=len(act_dof), act_dof=np.array(act_dof, dtype=np.int64), act_gear=np.array(act_gear),
                  act_ctrlrange=np.array(act_range).reshape(-1, 2), act_limited=np.array(act_lim, dtype=np.int64))

    # --- task constants
    reward_rule, term_rule = mt.kernel_rule(task)
    goals = list(task.goals)
    if len(goals) > L.CAPS["MAXGOAL"]:
        raise ValueError("too many goals")
    gpos = np.zeros((len(goals), 3))
    for k, g in enumerate(goals):
        gpos[k, : g.dim] = np.asarray(g.pos, float)
    # observed bodies, in the reference's order: balls, then blocks (maze_env.py:360-366)
    obs_bodies = [names["body"].index(n) for n in ball_names] if task.OBSERVE_BALLS else []
    obs_bodies += [names["body"].index(n) for n in obj_names] if task.OBSERVE_BLOCKS else []
    if len(obs_bodies) > 4:
        raise ValueError("too many observed bodies")
    # get_top_down_view (maze_env.py:262-349) reads the torso and every movable block from data.xpos
    view_bodies = [names["body"].index("torso")] + [names["body"].index(n) for n in obj_names] if task.TOP_DOWN_VIEW else []
    view_dim = L.VIEW_DIM if task.TOP_DOWN_VIEW else 0
    if len(obs_bodies) + len(view_bodies) > L.CAPS["MAXOBJ"]:
        raise ValueError("too many movable blocks for the top-down view")

    kind = agent.KERNEL_KIND
    if kind == "point":
        naq, nav, step_kind, reset_kind = 3, 3, L.STEP_TELEPORT, L.RESET_POINT
    elif kind == "ant":
        naq, nav, step_kind, reset_kind = 15, 14, L.STEP_TORQUE, L.RESET_ANT
    elif kind == "swimmer":
        naq, nav, step_kind, reset_kind = flat["nq"], flat["nv"], L.STEP_TORQUE, L.RESET_SWIMMER
    else:
        raise ValueError(f"unknown agent kind {kind}")
    obs_dim = naq + nav + 3 * len(obs_bodies) + view_dim + 1

    segs = np.zeros((0, 4))
    if agent.MANUAL_COLLISION:
        if agent.RADIUS is None:
            raise ValueError("Manual collision needs radius of the model")
        segs = CollisionDetector(structure, s, torso_x, torso_y, agent.RADIUS).segments()
        if len(segs) > L.CAPS["MAXSEG"]:
            raise ValueError("too many wall segments")

    fields.update(
        ngoal=len(goals), nseg=len(segs), grid_h=rows, grid_w=cols, step_kind=step_kind,
        frame_skip=agent.FRAME_SKIP, manual_collision=int(agent.MANUAL_COLLISION),
        collision_on=int(sc.collision_on), has_floor=int(floor is not None), elevated=int(elevated),
        reward_rule=reward_rule, term_rule=term_rule, max_episode_steps=max_episode_steps,
        obs_dim=obs_dim, n_agent_q=naq, n_agent_v=nav, nobj=len(obs_bodies), reset_kind=reset_kind, forward_reward_kind=int(forward_reward_kind),
        obj_body=np.array(obs_bodies + view_bodies, dtype=np.int64), nviewb=len(view_bodies), view_dim=view_dim, goal_dim=np.array([g.dim for g in goals], dtype=np.int64),
        grid=grid,
        timestep=sc.timestep, gravity=sc.gravity, density=sc.density, viscosity=sc.viscosity,
        inner_reward_scale=inner_reward_scaling, forward_reward_weight=forward_reward_weight,
        ctrl_cost_weight=ctrl_cost_weight, restitution=restitution_coef,
        penalty=0.0 if task.PENALTY is None else task.PENALTY, task_scale=task.scale,
        vel_limit=getattr(agent, "VELOCITY_LIMITS", 0.0), reset_noise=0.1,
        cell_size=s, origin=np.array([torso_x, torso_y]), wall_half=np.array([0.5 * s, 0.5 * s, h]),
        wall_z=h + height_offset, plat_z=h,
        wall_margin=wall.margin, wall_friction=wall.friction, wall_solref=wall.solref, wall_solimp=wall.solimp,
        goal_pos=gpos, goal_thr=np.array([g.threshold for g in goals]),
        goal_scale=np.array([g.reward_scale for g in goals]), seg=segs,
    )
    if floor is not None:
        fields.update(floor_z=floor.pos[2], floor_margin=floor.margin, floor_friction=floor.friction,
                      floor_solref=floor.solref, floor_solimp=floor.solimp)
    meta = dict(torso_xy=(torso_x, torso_y), height_offset=height_offset, elevated=elevated, blocks=has_blocks,
                movable_blocks=obj_names, object_balls=ball_names, structure=structure,
                body_invweight0_full=bw_full, legacy_capsule_volume=legacy_capsule_volume,
                merge_welded=merge_welded, act_ctrlrange=np.array(act_range).reshape(-1, 2))
    return MazeModel(fields, names, meta)
