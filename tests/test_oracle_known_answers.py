"""Known answers for the CPU restatement: its narrow phase (oracle/mmz_oracle.c:400-647) in configurations whose contacts
follow from elementary geometry - the cases in which any correct narrow phase, MuJoCo's mjc_PlaneSphere / mjc_SphereBox /
mjc_BoxBox included, must produce the same points: distance negative by the penetration, position midway between the two
surfaces, normal from geom 1 to geom 2 (the conventions of mmz_narrow.cuh:1-8). Geometry of the scenes: the Point of
assets/point.xml (sphere r = 0.5 at z = 0.5, arrow box 0.5 x 0.1 x 0.1 at 0.6 ahead) in the U maze of maze_task.py (cells of
4, walls 2 high: the cell west of the start cell is a wall whose east face is the plane x = -2), the Ant's torso sphere
(r = 0.25) over the floor. The last tests evaluate MuJoCo's documented formulas for the reference acceleration and the
regularisation of contact and joint-limit rows independently and compare."""
import numpy as np
import pytest

from conftest import make_model


@pytest.fixture(scope="module")
def point(oracle_lib):
    model = make_model("PointUMaze-v0")
    assert float(model.cell_size) == 4.0 and np.allclose(np.asarray(model.wall_half), [2, 2, 1])
    return oracle_lib.OracleEnv(model)


def test_sphere_against_wall_face(point):
    """The Point's sphere 0.1 deep in the wall west of the start cell, the arrow pointing away from it."""
    point.set_state([-1.6, 0.0, 0.0], [0, 0, 0])
    point.forward([0, 0])
    cons = point.contacts()
    assert len(cons) == 1
    c = cons[0]
    assert c["dist"] == pytest.approx(-0.1, abs=1e-12)
    assert np.allclose(c["frame"][0], [-1, 0, 0], atol=1e-12)           # from the sphere into the wall
    assert np.allclose(c["pos"], [-2.05, 0.0, 0.5], atol=1e-12)         # midway between x = -2.1 (sphere) and x = -2 (wall)
    assert (c["body1"], c["body2"]) == (0, -1)


def test_sphere_just_outside_the_wall_is_no_contact(point):
    point.set_state([-1.5 + 1e-9, 0.0, 0.0], [0, 0, 0])                 # the sphere's west pole 1e-9 short of the face, margin 0
    point.forward([0, 0])
    assert point.counts()["ncon"] == 0


def test_box_face_against_wall_face(point):
    """The arrow box head-on into the same wall, 0.05 deep: the four corners of its front face, all at the same depth."""
    point.set_state([-0.95, 0.0, np.pi], [0, 0, 0])
    point.forward([0, 0])
    cons = point.contacts()
    assert len(cons) == 4                                                # the sphere is 0.45 short of the wall
    pts = sorted((round(c["pos"][1], 9), round(c["pos"][2], 9)) for c in cons)
    assert pts == [(-0.1, 0.4), (-0.1, 0.6), (0.1, 0.4), (0.1, 0.6)]
    for c in cons:
        assert c["dist"] == pytest.approx(-0.05, abs=1e-9)
        assert c["pos"][0] == pytest.approx(-2.025, abs=1e-9)            # midway between x = -2.05 (arrow) and x = -2 (wall)
        assert np.allclose(c["frame"][0], [1, 0, 0], atol=1e-9)          # geom 1 = the wall (lower geom id), geom 2 = the arrow
        assert (c["body1"], c["body2"]) == (-1, 0)


def test_box_face_tilted_against_wall_face(point):
    """The arrow yawed by 0.2 rad: its leading vertical edge enters first - two contacts, depth from the edge's x."""
    th = np.pi - 0.2
    x0 = -0.922
    point.set_state([x0, 0.0, th], [0, 0, 0])
    point.forward([0, 0])
    cons = point.contacts()
    # corners of the front face in the world: centre + R (0.5, +-0.1, .)
    c, s = np.cos(th), np.sin(th)
    cx, cy = x0 + 0.6 * c, 0.6 * s
    corners = [(cx + 0.5 * c - sy * 0.1 * s, cy + 0.5 * s + sy * 0.1 * c) for sy in (-1, 1)]
    deep = [p for p in corners if p[0] < -2.0]
    assert len(deep) == 1 and len(cons) == 2                             # one vertical edge (two corners, z = 0.4 and 0.6) is inside
    for k in cons:
        assert k["dist"] == pytest.approx(deep[0][0] + 2.0, abs=1e-9)
        assert k["pos"][1] == pytest.approx(deep[0][1], abs=1e-9)
        assert np.allclose(k["frame"][0], [1, 0, 0], atol=1e-9)
    assert sorted(round(k["pos"][2], 9) for k in cons) == [0.4, 0.6]


def test_sphere_against_floor(oracle_lib):
    """The Ant's torso sphere (r = 0.25) 0.05 deep in the floor: plane contact under its centre (geom 1 = the plane)."""
    model = make_model("AntUMaze-v0")
    o = oracle_lib.OracleEnv(model)
    q = np.asarray(model.qpos0, float)[: int(model.nq)].copy()
    q[0:3] = [0.3, -0.2, 0.2]
    o.set_state(q, np.zeros(int(model.nv)))
    o.forward(np.zeros(int(model.nu)))
    under = [c for c in o.contacts() if np.allclose(c["pos"][:2], [0.3, -0.2], atol=1e-12)]
    assert len(under) == 1
    c = under[0]
    assert c["dist"] == pytest.approx(-0.05, abs=1e-12)
    assert np.allclose(c["frame"][0], [0, 0, 1], atol=1e-12)
    assert c["pos"][2] == pytest.approx(-0.025, abs=1e-12)               # midway between the floor and the sphere's lowest point
    assert (c["body1"], c["body2"]) == (-1, 0)


def test_capsule_end_against_wall_face(oracle_lib):
    """The Ant at its reference pose (legs stretched out horizontally at z = 0.75, ankle capsules r = 0.08 ending 0.8 out on
    both diagonals, ant.xml:22-66) with the tips of its two west ankles 0.03 deep in the wall west of the start cell (east
    face: x = -4): one contact per ankle at the end sphere - the other end of each capsule is 0.4 further from the wall."""
    model = make_model("AntUMaze-v0")
    assert float(model.cell_size) == 8.0
    o = oracle_lib.OracleEnv(model)
    q = np.asarray(model.qpos0, float)[: int(model.nq)].copy()
    q[0] = -3.15
    o.set_state(q, np.zeros(int(model.nv)))
    o.forward(np.zeros(int(model.nu)))
    cons = o.contacts()
    assert len(cons) == 2
    assert sorted(round(c["pos"][1], 9) for c in cons) == [-0.8, 0.8]
    assert sorted(c["body1"] for c in cons) == [4, 6] and all(c["body2"] == -1 for c in cons)
    for c in cons:
        assert c["dist"] == pytest.approx(-0.03, abs=1e-9)
        assert np.allclose(c["frame"][0], [-1, 0, 0], atol=1e-9)         # from the capsule into the wall
        assert c["pos"][0] == pytest.approx(-4.015, abs=1e-9)            # midway between x = -4.03 (capsule) and x = -4 (wall)
        assert c["pos"][2] == pytest.approx(0.75, abs=1e-9)


def test_capsules_flat_on_the_floor(oracle_lib):
    """The Ant at its reference pose lowered to z = 0.05: every leg capsule (r = 0.08) lies 0.03 deep in the floor and touches it
    with BOTH end spheres, the torso sphere (r = 0.25) is 0.2 deep. Capsule ends along each diagonal: 0 and 0.2 (the hip stub
    on the torso), 0.2 and 0.4 (the leg), 0.4 and 0.8 (the ankle)."""
    model = make_model("AntUMaze-v0")
    o = oracle_lib.OracleEnv(model)
    q = np.asarray(model.qpos0, float)[: int(model.nq)].copy()
    q[2] = 0.05
    o.set_state(q, np.zeros(int(model.nv)))
    o.forward(np.zeros(int(model.nu)))
    cons = o.contacts()
    assert o.counts()["overflow"] == 0
    for c in cons:
        assert np.allclose(c["frame"][0], [0, 0, 1], atol=1e-12) and c["body1"] == -1
        assert c["pos"][2] == pytest.approx(0.5 * c["dist"], abs=1e-12)  # midway between the floor and the lowest point
    sphere = [c for c in cons if abs(c["dist"] + 0.2) < 1e-9]
    caps = [c for c in cons if abs(c["dist"] + 0.03) < 1e-9]
    assert len(sphere) == 1 and len(caps) == len(cons) - 1 == 24
    want = sorted((round(sx * r, 9), round(sy * r, 9)) for sx in (-1, 1) for sy in (-1, 1) for r in (0.0, 0.2, 0.2, 0.4, 0.4, 0.8))
    got = sorted((round(c["pos"][0], 9) + 0.0, round(c["pos"][1], 9) + 0.0) for c in caps)
    assert got == want


def test_contact_reference_acceleration_follows_the_documented_formula(point):
    """MuJoCo's soft-constraint reference acceleration (documentation, "Solver parameters"): a_ref = -b v - k d(r) r with
    b = 2 / (dmax tau), k = 1 / (dmax^2 tau^2 zeta^2), solref = (tau, zeta) = (0.02, 1) made safe as tau >= 2 timestep = 0.04,
    and solimp mixed with equal weights between the two geoms: dmax = (0.99 [point.xml] + 0.95 [maze box default]) / 2; 0.1 deep
    is far beyond the impedance width (0.001), so d(r) = dmax. All four pyramid rows of the contact share r and, for a
    velocity along the normal, v."""
    tau, zeta, dmax, r = 0.04, 1.0, 0.5 * (0.99 + 0.95), -0.1
    k, b = 1.0 / (dmax ** 2 * tau ** 2 * zeta ** 2), 2.0 / (dmax * tau)
    for vx in (0.0, -0.3, 0.2):
        point.set_state([-1.6, 0.0, 0.0], [vx, 0, 0])
        point.forward([0, 0])
        J, D, aref, f = point.efc()
        assert J.shape == (4, 3) and np.allclose(J[:, 0], 1.0)          # d(dist)/dt = +vx: the wall is to the west
        assert np.allclose(aref, -b * vx - k * dmax * r, rtol=1e-12)
        assert np.allclose(D, D[0]) and D[0] > 0                          # the four edges of the pyramid share their regularisation


def test_contact_regularisation_follows_the_documented_formula(point):
    """R of a pyramidal contact row as MuJoCo builds it (mj_makeImpedance, documentation "Solver parameters" / "Contact"):
    R = 2 mu^2 R_first, R_first = (1 - d) / d * A with the approximate inverse inertia A = (1 + mu^2) (w_1 + w_2), w = the
    translational body_invweight0 = trace(J M^-1 J^T) / 3 in the reference pose. The Point slides along x and y with its whole
    mass m and cannot move along z: w = (1/m + 1/m + 0) / 3; the wall is the world (w = 0). D = 1 / R."""
    model = make_model("PointUMaze-v0")
    mass = float(np.asarray(model.body_mass)[0])
    d, mu = 0.5 * (0.99 + 0.95), 1.0
    w = 2.0 / (3.0 * mass)
    R = 2.0 * mu * mu * (1.0 - d) / d * (1.0 + mu * mu) * w
    point.set_state([-1.6, 0.0, 0.0], [0, 0, 0])
    point.forward([0, 0])
    assert point.contacts()[0]["mu"] == mu
    _, D, _, _ = point.efc()
    assert np.allclose(D, 1.0 / R, rtol=1e-9)


def test_joint_limit_rows_follow_the_documented_formulas(oracle_lib):
    """The Ant in its reference pose, high above the floor: its four ankle hinges sit at 0, outside their ranges (ant.xml:
    [30, 70] or [-70, -30] degrees), so exactly four limit rows are active, 0.5236 rad deep. Row: J = +1 at a violated lower
    limit, -1 at an upper one; a_ref = k d |r| with tau = 2 timestep (solref made safe), d = dmax = 0.95 (far beyond the
    impedance width); R = (1 - d) / d * dof_invweight0 with dof_invweight0 = (M^-1)_dd in the reference pose (hinges)."""
    model = make_model("AntUMaze-v0")
    nq, nv = int(model.nq), int(model.nv)
    o = oracle_lib.OracleEnv(model)
    q = np.asarray(model.qpos0, float)[:nq].copy()
    q[2] = 3.0
    o.set_state(q, np.zeros(nv))
    o.forward(np.zeros(int(model.nu)))
    assert o.counts()["ncon"] == 0
    J, D, aref, _ = o.efc()
    assert J.shape == (4, nv)
    rng_ = np.asarray(model.jnt_range, float)
    Minv = np.linalg.inv(o.mass_matrix())
    tau, d, depth = 2 * float(model.timestep), 0.95, np.deg2rad(30.0)
    k = 1.0 / (d * d * tau * tau)
    for row, dof in enumerate((7, 9, 11, 13)):
        jnt = dof - 5                                                    # joint 0 is the free joint (dofs 0..5)
        lower = rng_[jnt][0] > 0                                         # range above 0: the LOWER limit is violated
        want = np.zeros(nv)
        want[dof] = 1.0 if lower else -1.0
        assert np.allclose(J[row], want)
        assert aref[row] == pytest.approx(k * d * depth, rel=1e-6)
        assert D[row] == pytest.approx(1.0 / ((1 - d) / d * Minv[dof, dof]), rel=1e-6)


def test_actuation_passive_and_bias_forces_of_the_ant_in_known_states(oracle_lib):
    """Smooth forces from first principles. ant.xml:8,23,71-78: motors (gear 1) on the eight hinges in the order hip_4, ankle_4,
    hip_1, ankle_1, ... with ctrlrange [-30, 30] (clamped), joint damping 1 and armature 1 on every hinge, none on the free
    joint. (1) actuation = clip(ctrl) on the motor's dof; (2) passive = -damping * qvel on the hinges, nothing on the free
    joint; (3) at rest the bias force is gravity alone: total mass * 9.81 along the free joint's z, no net torque on the
    torso in the symmetric reference pose; (4) the free joint's translations carry the whole mass, a hinge its armature plus
    the inertia of what hangs on it."""
    model = make_model("AntUMaze-v0")
    nq, nv, nu = int(model.nq), int(model.nv), int(model.nu)
    o = oracle_lib.OracleEnv(model)
    q = np.asarray(model.qpos0, float)[:nq].copy()
    q[2] = 3.0
    ctrl = np.array([5.0, -2.5, 30.0, -30.0, 45.0, -70.0, 0.0, 1.0])  # two of them outside the control range
    v = np.zeros(nv)
    v[6:] = np.linspace(-1.0, 1.0, 8)
    o.set_state(q, v)
    o.forward(ctrl)
    act = o.vec("qfrc_act")
    want = np.zeros(nv)
    # dofs 6..13 = hip_1, ankle_1, hip_2, ankle_2, hip_3, ankle_3, hip_4, ankle_4 (body order); motors start at hip_4
    for k, dof in enumerate((12, 13, 6, 7, 8, 9, 10, 11)):
        want[dof] = np.clip(ctrl[k], -30.0, 30.0)
    assert np.allclose(act, want)
    pas = o.vec("qfrc_passive")
    assert np.allclose(pas[:6], 0) and np.allclose(pas[6:], -1.0 * v[6:])
    o.set_state(q, np.zeros(nv))
    o.forward(np.zeros(nu))
    bias = o.vec("qfrc_bias")
    mass = float(np.asarray(model.body_mass)[: int(model.nbody)].sum())
    assert bias[2] == pytest.approx(mass * 9.81, rel=1e-12)
    assert np.allclose(bias[:2], 0, atol=1e-12) and np.allclose(bias[3:6], 0, atol=1e-9)
    M = o.mass_matrix()
    assert np.allclose(np.diag(M)[:3], mass)
    assert np.all(np.diag(M)[6:] > 1.0)
