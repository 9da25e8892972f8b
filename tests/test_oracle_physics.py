"""Closed-form and conservation checks of the fp64 oracle's physics (no MuJoCo needed).

The physics half of the oracle is parity-UNPINNED (no MuJoCo binary exists in
this image), so these tests validate the restatement itself: SURVEY.md section 4
"oracle-free physics checks" (1)-(7).
"""
import numpy as np
import pytest

from conftest import compile_test_xml, make_model
from mujoco_maze.model_compiler import np_kinematics, np_mass_matrix, quat_to_mat


def _energy(model, o, gravity):
    q, v, _ = o.get_state()
    o.forward(None)
    M = o.mass_matrix()
    xpos, xquat = o.xpos()
    ke = 0.5 * v @ M @ v
    pe = 0.0
    for b in range(int(model.nbody)):
        com = xpos[b] + quat_to_mat(xquat[b]) @ model.body_ipos[b]
        pe -= model.body_mass[b] * gravity @ com
    return ke, pe


def test_mass_matrix_matches_jacobian_form(oracle_lib):
    rng = np.random.default_rng(1)
    for env_id in ("AntUMaze-v0", "PointUMaze-v0", "SwimmerUMaze-v0", "AntPush-v0"):
        model = make_model(env_id)
        o = oracle_lib.OracleEnv(model)
        for _ in range(5):
            q = model.qpos0 + rng.uniform(-0.5, 0.5, size=int(model.nq))
            if env_id.startswith("Ant"):
                q[3:7] /= np.linalg.norm(q[3:7])
            o.set_state(q, rng.normal(size=int(model.nv)))
            o.forward(None)
            M = o.mass_matrix()
            Mp = np_mass_matrix(model.fields, q)
            assert np.abs(M - Mp).max() < 1e-12
            assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0


def test_merged_and_unmerged_models_agree(oracle_lib):
    """Folding welded bodies into their parents must not change M, bias or qacc."""
    import mujoco_maze
    from mujoco_maze.ant import AntEnv
    from mujoco_maze.maze_task import GoalRewardUMaze
    from mujoco_maze.model_compiler import compile_maze_model

    a = compile_maze_model(AntEnv, GoalRewardUMaze(8.0), 8.0, merge_welded=True)
    b = compile_maze_model(AntEnv, GoalRewardUMaze(8.0), 8.0, merge_welded=False)
    assert int(a.nbody) == 9 and int(b.nbody) == 13
    oa, ob = oracle_lib.OracleEnv(a), oracle_lib.OracleEnv(b)
    rng = np.random.default_rng(2)
    for k in range(6):
        q = a.qpos0 + rng.uniform(-0.3, 0.3, size=15)
        q[2] = 0.45 + 0.1 * k  # some in contact, some not
        q[3:7] /= np.linalg.norm(q[3:7])
        v = rng.normal(size=14)
        act = rng.uniform(-30, 30, size=8)
        oa.set_state(q, v)
        ob.set_state(q, v)
        qa, qb = oa.forward(act), ob.forward(act)
        assert oa.counts()["ncon"] == ob.counts()["ncon"]
        assert np.abs(oa.mass_matrix() - ob.mass_matrix()).max() < 1e-12
        assert np.abs(oa.vec("qfrc_bias") - ob.vec("qfrc_bias")).max() < 1e-10
        assert np.abs(qa - qb).max() < 1e-7 * (1 + np.abs(qa).max())


def test_free_fall_is_exact(oracle_lib):
    """Ant torso before first contact: RK4 integrates z = z0 - g t^2 / 2 exactly."""
    model = make_model("AntUMaze-v0")
    o = oracle_lib.OracleEnv(model)
    q = model.qpos0.copy()
    q[2] = 5.0
    q[7:] = [0.0, 0.9, 0.0, -0.9, 0.0, -0.9, 0.0, 0.9]  # ankles inside their ranges: no limit forces
    o.set_state(q, np.zeros(14))
    h, n = float(model.timestep), 20
    for _ in range(n):
        assert o.mj_step(None) == 0
        assert o.counts()["nefc"] == 0
    qn, vn, _ = o.get_state()
    t = n * h
    assert abs(qn[2] - (5.0 - 0.5 * 9.81 * t * t)) < 1e-10
    assert abs(vn[2] + 9.81 * t) < 1e-10
    assert np.abs(vn[6:]).max() < 1e-9 and np.abs(qn[3:7] - [1, 0, 0, 0]).max() < 1e-12


def test_damped_hinge_decay(oracle_lib):
    model = compile_test_xml("hinge_damped.xml")
    o = oracle_lib.OracleEnv(model)
    o.set_state(np.zeros(1), np.array([2.0]))
    o.forward(None)
    inertia = o.mass_matrix()[0, 0]  # includes armature 1
    h, n = float(model.timestep), 50
    for _ in range(n):
        o.mj_step(None)
    _, v, _ = o.get_state()
    exact = 2.0 * np.exp(-n * h * 1.0 / inertia)
    assert abs(v[0] - exact) < 1e-7  # RK4 global error O(h^4)


def test_pendulum_energy_conservation(oracle_lib):
    model = compile_test_xml("pendulum2.xml")
    o = oracle_lib.OracleEnv(model)
    g = np.array([0.0, 0.0, -9.81])
    o.set_state(np.array([0.7, -0.4, 0.05]), np.array([0.5, -1.0, 0.3]))
    ke0, pe0 = _energy(model, o, g)
    for _ in range(1000):
        o.mj_step(None)
    ke1, pe1 = _energy(model, o, g)
    assert abs(ke0 - ke1) > 1e-3  # energy actually moved between kinetic and potential
    assert abs((ke0 + pe0) - (ke1 + pe1)) < 1e-7 * max(1.0, abs(ke0 + pe0))


def _momenta(model, o):
    """Linear and angular momentum (about the world origin) from body twists via finite differences."""
    q, v, _ = o.get_state()
    f = model.fields
    eps = 1e-7
    # integrate positions a tiny step to differentiate body frames numerically
    def frames(qq):
        xpos, xquat = np_kinematics(f, qq)
        return xpos, np.array([quat_to_mat(x) for x in xquat])

    q2 = q.copy()
    # manifold step identical to integrate_pos
    q2[:3] += eps * v[:3]
    w = v[3:6]
    ang = np.linalg.norm(w) * eps
    if ang > 0:
        ax = w / np.linalg.norm(w)
        dq = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
        from mujoco_maze.model_compiler import quat_mul
        q2[3:7] = quat_mul(q[3:7], dq)
    q2[7:] += eps * v[6:]
    x0, R0 = frames(q)
    x1, R1 = frames(q2)
    P, Lm = np.zeros(3), np.zeros(3)
    for b in range(int(model.nbody)):
        m = model.body_mass[b]
        c0 = x0[b] + R0[b] @ model.body_ipos[b]
        c1 = x1[b] + R1[b] @ model.body_ipos[b]
        vc = (c1 - c0) / eps
        dR = (R1[b] - R0[b]) / eps @ R0[b].T
        omega = np.array([dR[2, 1], dR[0, 2], dR[1, 0]])
        Ri = R0[b] @ quat_to_mat(model.body_iquat[b])
        Iw = Ri @ np.diag(model.body_inertia[b]) @ Ri.T
        P += m * vc
        Lm += Iw @ omega + m * np.cross(c0, vc)
    return P, Lm


def test_floating_body_conserves_momentum_and_energy(oracle_lib):
    model = compile_test_xml("floating.xml")
    assert int(model.nv) == 9
    o = oracle_lib.OracleEnv(model)
    rng = np.random.default_rng(3)
    q = model.qpos0.copy()
    q[3:7] = [0.8, 0.2, -0.5, 0.1]
    q[3:7] /= np.linalg.norm(q[3:7])
    q[7:] = [0.3, -0.6, 0.4]
    v = rng.normal(size=9)
    o.set_state(q, v)
    ke0, _ = _energy(model, o, np.zeros(3))
    P0, L0 = _momenta(model, o)
    for _ in range(500):
        o.mj_step(None)
    ke1, _ = _energy(model, o, np.zeros(3))
    P1, L1 = _momenta(model, o)
    qn, _, _ = o.get_state()
    assert abs(np.linalg.norm(qn[3:7]) - 1) < 1e-12
    assert abs(ke0 - ke1) < 1e-8 * ke0
    assert np.abs(P0 - P1).max() < 1e-5 * (1 + np.abs(P0).max())  # finite-difference momenta: looser
    assert np.abs(L0 - L1).max() < 1e-5 * (1 + np.abs(L0).max())


def test_solver_kkt_and_warmstart_independence(oracle_lib):
    model = make_model("AntUMaze-v0")
    o = oracle_lib.OracleEnv(model)
    rng = np.random.default_rng(4)
    seen_contacts = 0
    for k in range(12):
        q = model.qpos0 + rng.uniform(-0.2, 0.2, size=15)
        q[7:] += [0.0, 1.0, 0.0, -1.0, 0.0, -1.0, 0.0, 1.0]  # feet down
        q[2] = 0.40 + 0.02 * k
        q[3:7] /= np.linalg.norm(q[3:7])
        v = rng.normal(size=14)
        act = rng.uniform(-30, 30, size=8)
        o.set_state(q, v)
        o.L.ora_set_warmstart(o.h, 0)
        a = o.forward(act)
        J, D, aref, f = o.efc()
        M = o.mass_matrix()
        seen_contacts += o.counts()["ncon"]
        # stationarity: M (a - a_smooth) = J^T f, with f = -D min(0, J a - aref) >= 0
        jar = J @ a - aref
        assert np.all(f >= 0)
        assert np.allclose(f, -D * np.minimum(0, jar), rtol=1e-9, atol=1e-9)
        res = M @ a - o.vec("qfrc_smooth") - J.T @ f
        assert np.abs(res).max() < 1e-8 * (1 + np.abs(o.vec("qfrc_smooth")).max())
        # unique minimiser: a different starting point converges to the same answer
        o.L.ora_set_warmstart(o.h, 1)  # starts from the previous qacc (a) perturbed by the new forward
        o.set_state(q, v)
        a2 = o.forward(act)
        assert np.abs(a - a2).max() < 1e-7 * (1 + np.abs(a).max())
    assert seen_contacts > 10


def test_swimmer_dissipates_energy(oracle_lib):
    model = make_model("SwimmerUMaze-v0")
    o = oracle_lib.OracleEnv(model)
    rng = np.random.default_rng(5)
    o.set_state(rng.uniform(-0.1, 0.1, size=5), rng.uniform(-1, 1, size=5))
    last = None
    for _ in range(40):
        ke, _ = _energy(model, o, np.zeros(3))
        if last is not None:
            assert ke <= last * (1 + 1e-12)
        last = ke
        for _ in range(5):
            o.mj_step(None)
    assert o.counts()["ncon"] == 0  # collision="predefined": walls are decorative (quirk Q8)


def test_ant_settles_on_four_ankles(oracle_lib):
    model = make_model("AntUMaze-v0")
    o = oracle_lib.OracleEnv(model)
    o.set_state(model.qpos0, np.zeros(14))
    for _ in range(800):  # armature 1 on feather-weight links: the collapse onto the ankle stops is slow
        assert o.mj_step(None) == 0
    q, v, _ = o.get_state()
    assert np.abs(v).max() < 1e-3
    assert 0.3 < q[2] < 0.75
    ank = np.abs(q[[8, 10, 12, 14]])
    assert np.all(ank > np.deg2rad(30) - 0.05) and np.all(ank < np.deg2rad(70) + 0.05)
    assert o.counts()["ncon"] >= 4
