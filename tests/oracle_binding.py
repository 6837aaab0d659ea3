"""ctypes binding of oracle/liblocosim_ref.so -- TEST INFRASTRUCTURE ONLY (never imported by the package)."""
import ctypes

import numpy as np


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        vp, ip = ctypes.c_void_p, ctypes.c_int
        lib.ref_create.restype = vp
        lib.ref_create.argtypes = [vp, ip, vp, ip]
        lib.refenv_create.restype = vp
        lib.refenv_create.argtypes = [vp, ip, vp, ip, vp, ip, vp, ip]
        lib.refenv_sim.restype = vp
        lib.refenv_sim.argtypes = [vp]
        lib.refenv_set_user.restype = None
        lib.refenv_set_user.argtypes = [vp, vp]
        lib.refenv_set_rotation.restype = None
        lib.refenv_set_rotation.argtypes = [vp, ctypes.c_double]
        lib.refenv_obs_dim.restype = ip
        lib.refenv_obs_dim.argtypes = [vp]
        for f, args in [("ref_destroy", [vp]), ("refenv_destroy", [vp]), ("ref_reset", [vp, vp, vp]),
                        ("ref_get_state", [vp, vp, vp]), ("ref_step", [vp, vp, ip]),
                        ("refenv_reset_to", [vp, ip, ip, vp]), ("refenv_step", [vp, vp, vp, vp, vp]),
                        ("ref_get_warmstart", [vp, vp]), ("ref_set_warmstart", [vp, vp])]:
            getattr(lib, f).restype = None
            getattr(lib, f).argtypes = args
        lib.ref_rollout.restype = ctypes.c_long
        lib.ref_rollout.argtypes = [vp, ip, vp, ip, vp, ip, vp, ip, ip, ip, ip, ctypes.c_ulonglong, vp, vp]
        lib.ref_nv.restype = ip
        lib.ref_nv.argtypes = [vp]
        lib.ref_debug_convex.restype = ctypes.c_long
        lib.ref_debug_convex.argtypes = [ip]

    def convex_hits(self):
        """Contacts reported so far by the oracle's general convex routine (mjc_Convex / MPR), process-wide counter."""
        return int(self.lib.ref_debug_convex(1))

    def env(self, model_blobs, task_blobs):
        return OracleEnv(self, model_blobs, task_blobs)

    def rollout(self, model_blobs, task_blobs, n_envs, n_steps, nthreads, seed=0):
        mi, mr = [np.ascontiguousarray(x) for x in model_blobs]
        ti, tr = [np.ascontiguousarray(x) for x in task_blobs]
        resets = ctypes.c_long(0)
        n = self.lib.ref_rollout(_p(mi), len(mi), _p(mr), len(mr), _p(ti), len(ti), _p(tr), len(tr), n_envs, n_steps,
                                 nthreads, seed, None, ctypes.byref(resets))
        return n, resets.value


class OracleEnv:
    def __init__(self, o, model_blobs, task_blobs):
        self.lib = o.lib
        mi, mr = model_blobs
        ti, tr = task_blobs
        self._keep = [np.ascontiguousarray(mi, dtype=np.int32), np.ascontiguousarray(mr, dtype=np.float64),
                      np.ascontiguousarray(ti, dtype=np.int32), np.ascontiguousarray(tr, dtype=np.float64)]
        a, b, c, d = self._keep
        self.h = ctypes.c_void_p(self.lib.refenv_create(_p(a), len(a), _p(b), len(b), _p(c), len(c), _p(d), len(d)))
        assert self.h.value, "oracle rejected the blobs"
        self.sim = ctypes.c_void_p(self.lib.refenv_sim(self.h))
        self.obs_dim = self.lib.refenv_obs_dim(self.h)
        self.nv = self.lib.ref_nv(self.sim)

    def set_user(self, user):
        u = np.zeros(4)
        u[:len(user)] = user
        self.lib.refenv_set_user(self.h, _p(u))

    def set_rotation(self, angle):
        self.lib.refenv_set_rotation(self.h, float(angle))

    def reset_to(self, traj_no, step_no):
        obs = np.zeros(self.obs_dim)
        self.lib.refenv_reset_to(self.h, int(traj_no), int(step_no), _p(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs = np.zeros(self.obs_dim)
        r = ctypes.c_double(0)
        ab = ctypes.c_int(0)
        self.lib.refenv_step(self.h, _p(a), _p(obs), ctypes.byref(r), ctypes.byref(ab))
        return obs, r.value, bool(ab.value)

    def get_state(self):
        q, v = np.zeros(self.nv), np.zeros(self.nv)
        self.lib.ref_get_state(self.sim, _p(q), _p(v))
        return q, v

    def set_state(self, qpos, qvel, warmstart=None):
        q = np.ascontiguousarray(qpos, dtype=np.float64)
        v = np.ascontiguousarray(qvel, dtype=np.float64)
        self.lib.ref_reset(self.sim, _p(q), _p(v))
        if warmstart is not None:
            w = np.ascontiguousarray(warmstart, dtype=np.float64)
            self.lib.ref_set_warmstart(self.sim, _p(w))

    def close(self):
        if self.h:
            self.lib.refenv_destroy(self.h)
            self.h = None


def load(path):
    return Oracle(ctypes.CDLL(path))
