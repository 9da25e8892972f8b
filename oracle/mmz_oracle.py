"""ctypes wrapper of the CPU float64 oracle (oracle/mmz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs. The product package never imports it.
PARITY STATUS: Python half pinned by tests/golden (generated from the real
reference); physics parity UNPINNED (no MuJoCo available) — see the C header.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libmmz_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mmz_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "mmz_model.h")
    stale = (not os.path.exists(LIB)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIB) for p in (src, hdr))
    if force or stale:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-fopenmp", "-shared", "-o", LIB, src, "-lm"])
    return LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        dp, ip, vp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p
        L.ora_create.argtypes, L.ora_create.restype = [ctypes.c_char_p, ctypes.c_size_t], vp
        L.ora_destroy.argtypes = [vp]
        L.ora_model_bytes.restype = ctypes.c_size_t
        L.ora_set_warmstart.argtypes = [vp, ctypes.c_int]
        L.ora_set_state.argtypes = [vp, dp, dp, ctypes.c_int]
        L.ora_get_state.argtypes = [vp, dp, dp, ip]
        L.ora_observe.argtypes = [vp, dp]
        L.ora_forward.argtypes = [vp, dp]
        L.ora_get_vec.argtypes = [vp, ctypes.c_int, dp]
        L.ora_get_M.argtypes = [vp, dp]
        L.ora_get_counts.argtypes = [vp, ip]
        L.ora_get_contact.argtypes = [vp, ctypes.c_int, dp]
        L.ora_get_efc.argtypes = [vp, dp, dp, dp, dp]
        L.ora_get_xpos.argtypes = [vp, dp, dp]
        L.ora_mj_step.argtypes, L.ora_mj_step.restype = [vp, dp], ctypes.c_int
        L.ora_detect.argtypes, L.ora_detect.restype = [vp, dp, dp, dp, dp], ctypes.c_int
        L.ora_task_rules.argtypes = [vp, dp, dp, ip]
        L.ora_step.argtypes, L.ora_step.restype = [vp, dp, dp, dp, dp], ctypes.c_int
        L.ora_rollout.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, dp, dp, dp, ctypes.c_int,
                                  ctypes.c_int, dp, dp]
        L.ora_rollout.restype = ctypes.c_long
        L.ora_batch_create.argtypes, L.ora_batch_create.restype = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int], vp
        L.ora_batch_destroy.argtypes = [vp]
        L.ora_batch_set_state.argtypes = [vp, dp, dp]
        L.ora_batch_step.argtypes, L.ora_batch_step.restype = [vp, dp, dp, dp, ip, ctypes.c_int], ctypes.c_long
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _arr(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


VEC = dict(qacc=0, qacc_smooth=1, qfrc_bias=2, qfrc_passive=3, qfrc_smooth=4, qfrc_act=5)


class OracleEnv:
    """One environment of the fp64 restatement. `model` is a mujoco_maze MazeModel."""

    def __init__(self, model):
        self.L = lib()
        self.model = model
        self.blob = model.blob(8)
        assert self.L.ora_model_bytes() == len(self.blob), "oracle built against a different mmz_model.h"
        self.h = self.L.ora_create(self.blob, len(self.blob))
        if not self.h:
            raise RuntimeError("ora_create rejected the model blob")
        self.nq, self.nv, self.nu = int(model.nq), int(model.nv), int(model.nu)
        self.obs_dim, self.nbody = int(model.obs_dim), int(model.nbody)

    def __del__(self):
        try:
            if self.h:
                self.L.ora_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_state(self, qpos, qvel, t=0):
        self.L.ora_set_state(self.h, _d(_arr(qpos, self.nq)), _d(_arr(qvel, self.nv)), int(t))

    def get_state(self):
        q, v, t = np.zeros(self.nq), np.zeros(self.nv), ctypes.c_int()
        self.L.ora_get_state(self.h, _d(q), _d(v), ctypes.byref(t))
        return q, v, t.value

    def observe(self):
        o = np.zeros(self.obs_dim)
        self.L.ora_observe(self.h, _d(o))
        return o

    def forward(self, action=None):
        a = _arr(np.zeros(max(self.nu, 1)) if action is None else action)
        self.L.ora_forward(self.h, _d(a))
        return self.vec("qacc")

    def vec(self, name):
        out = np.zeros(self.nv)
        self.L.ora_get_vec(self.h, VEC[name], _d(out))
        return out

    def mass_matrix(self):
        out = np.zeros((self.nv, self.nv))
        self.L.ora_get_M(self.h, _d(out))
        return out

    def counts(self):
        c = (ctypes.c_int * 4)()
        self.L.ora_get_counts(self.h, c)
        return dict(ncon=c[0], nefc=c[1], niter=c[2], overflow=c[3])

    def contacts(self):
        out = []
        for i in range(self.counts()["ncon"]):
            b = np.zeros(17)
            self.L.ora_get_contact(self.h, i, _d(b))
            out.append(dict(dist=b[0], pos=b[1:4].copy(), frame=b[4:13].reshape(3, 3).copy(), body1=int(b[13]),
                            body2=int(b[14]), mu=b[15], margin=b[16]))
        return out

    def efc(self):
        n = self.counts()["nefc"]
        J, D, aref, f = np.zeros((max(n, 1), self.nv)), np.zeros(max(n, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
        self.L.ora_get_efc(self.h, _d(J), _d(D), _d(aref), _d(f))
        return J[:n], D[:n], aref[:n], f[:n]

    def xpos(self):
        p, q = np.zeros((self.nbody, 3)), np.zeros((self.nbody, 4))
        self.L.ora_get_xpos(self.h, _d(p), _d(q))
        return p, q

    def mj_step(self, ctrl=None):
        c = _arr(np.zeros(max(self.nu, 1)) if ctrl is None else ctrl)
        return self.L.ora_mj_step(self.h, _d(c))

    def detect(self, old_xy, new_xy):
        pt, rf = np.zeros(2), np.zeros(2)
        hit = self.L.ora_detect(self.h, _d(_arr(old_xy, 2)), _d(_arr(new_xy, 2)), _d(pt), _d(rf))
        return (pt, rf) if hit else None

    def task_rules(self, obs):
        o = _arr(obs)
        r, d = ctypes.c_double(), ctypes.c_int()
        pad = np.zeros(max(o.size, 16))
        pad[: o.size] = o
        self.L.ora_task_rules(self.h, _d(pad), ctypes.byref(r), ctypes.byref(d))
        return r.value, bool(d.value)

    def step(self, action):
        """MazeEnv.step -> (obs, reward, done_bits, info[4])."""
        obs, info, r = np.zeros(self.obs_dim), np.zeros(4), ctypes.c_double()
        bits = self.L.ora_step(self.h, _d(_arr(action, self.nu)), _d(obs), ctypes.byref(r), _d(info))
        return obs, r.value, bits, info


def rollout(model, qpos, qvel, actions, nthreads=1):
    """Advance n envs for `steps` steps on `nthreads` host threads; returns (env_steps, obs, reward)."""
    L = lib()
    blob = model.blob(8)
    qpos, qvel, actions = _arr(qpos), _arr(qvel), _arr(actions)
    n, steps = qpos.shape[0], actions.shape[0]
    obs, rew = np.zeros((n, int(model.obs_dim))), np.zeros(n)
    done = L.ora_rollout(blob, len(blob), n, _d(qpos), _d(qvel), _d(actions), steps, int(nthreads), _d(obs), _d(rew))
    return done, obs, rew


class OracleBatch:
    """n persistent environments of the fp64 restatement stepped by OpenMP threads (CPU baseline)."""

    def __init__(self, model, n):
        self.L = lib()
        blob = model.blob(8)
        self.n, self.nq, self.nv, self.nu = int(n), int(model.nq), int(model.nv), int(model.nu)
        self.obs_dim = int(model.obs_dim)
        self.h = self.L.ora_batch_create(blob, len(blob), self.n)
        if not self.h:
            raise RuntimeError("ora_batch_create rejected the model blob")
        self.obs = np.zeros((self.n, self.obs_dim))
        self.reward = np.zeros(self.n)
        self.done = np.zeros(self.n, dtype=np.int32)

    def set_state(self, qpos, qvel):
        self.L.ora_batch_set_state(self.h, _d(_arr(qpos, self.n * self.nq)), _d(_arr(qvel, self.n * self.nv)))

    def step(self, actions, nthreads=1):
        a = _arr(actions, self.n * self.nu)
        self.L.ora_batch_step(self.h, _d(a), _d(self.obs), _d(self.reward),
                              self.done.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(nthreads))
        return self.obs, self.reward, self.done

    def close(self):
        if self.h:
            self.L.ora_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
