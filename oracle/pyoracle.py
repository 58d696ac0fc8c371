"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY; used by tests/, smoke() and bench.py cpu_baseline)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        L.orc_last_error.restype = C.c_char_p
        L.orc_sample_height.restype = C.c_double
        for name in ("orc_destroy", "orc_update", "orc_env_step", "orc_reset", "orc_get_state", "orc_set_state",
                     "orc_get_last_tau", "orc_get_poli_state", "orc_get_net_out", "orc_rbd", "orc_forward_dynamics",
                     "orc_net_eval", "orc_com", "orc_reset_tuples", "orc_eval_stats", "orc_set_phys",
                     "orc_set_explore"):
            getattr(L, name).restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, pack, num_envs=1, mode=0, terrain_seeds=None, rng_seed=1234):
        self.L = lib()
        seeds = None
        if terrain_seeds is not None:
            seeds = np.ascontiguousarray(terrain_seeds, dtype=np.uint64)
        self.h = self.L.orc_create(pack.encode(), num_envs, mode, _p(seeds), C.c_uint64(rng_seed))
        if not self.h:
            raise RuntimeError(self.L.orc_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.n = num_envs
        self.ndof = self.L.orc_num_dof(self.h)
        self.nj = self.L.orc_num_joints(self.h)
        self.S = self.L.orc_state_size(self.h)
        self.A = self.L.orc_action_size(self.h)
        self.num_params = self.L.orc_num_params(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def update(self, dt=1.0 / 30.0, threads=1):
        self.L.orc_update(self.h, C.c_double(dt), threads)

    def env_step(self, env=0, h=1.0 / 600.0):
        self.L.orc_env_step(self.h, env, C.c_double(h))

    def reset(self, env=0):
        self.L.orc_reset(self.h, env)

    def set_phys(self, p7):
        a = np.ascontiguousarray(p7, dtype=np.float64)
        self.L.orc_set_phys(self.h, _p(a))

    def set_explore(self, enable, rate, temp, base_rate):
        self.L.orc_set_explore(self.h, int(enable), C.c_double(rate), C.c_double(temp), C.c_double(base_rate))

    def get_state(self, env=0):
        q = np.zeros(self.ndof); qd = np.zeros(self.ndof); tau = np.zeros(self.ndof)
        c = np.zeros(self.nj, np.uint8)
        self.L.orc_get_state(self.h, env, _p(q), _p(qd), _p(tau), _p(c))
        return q, qd, tau, c

    def set_state(self, env=0, q=None, qd=None, tau=None, contact=None):
        f = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
        q, qd, tau, contact = f(q, np.float64), f(qd, np.float64), f(tau, np.float64), f(contact, np.uint8)
        self.L.orc_set_state(self.h, env, _p(q), _p(qd), _p(tau), _p(contact))

    def get_ctrl(self, env=0):
        out = np.zeros(160)
        n = self.L.orc_get_ctrl(self.h, env, _p(out))
        return out[:n]

    def last_tau(self, env=0):
        t = np.zeros(self.ndof)
        self.L.orc_get_last_tau(self.h, env, _p(t))
        return t

    def poli_state(self, env=0):
        s = np.zeros(self.S)
        self.L.orc_get_poli_state(self.h, env, _p(s))
        return s

    def net_out(self, env=0, n=90):
        y = np.zeros(n)
        self.L.orc_get_net_out(self.h, env, _p(y))
        return y

    def terrain(self, env=0, seg=0, cap=1024):
        d = np.zeros(cap, np.float32)
        mx = C.c_double(0)
        fl = C.c_int(0)
        n = self.L.orc_get_terrain(self.h, env, seg, _p(d), cap, C.byref(mx), C.byref(fl))
        return d[:n].copy(), mx.value, fl.value

    def sample_height(self, x, env=0):
        return self.L.orc_sample_height(self.h, env, C.c_double(x))

    def rbd(self, env=0):
        M = np.zeros((self.ndof, self.ndof)); Cb = np.zeros(self.ndof)
        self.L.orc_rbd(self.h, env, _p(M), _p(Cb))
        return M, Cb

    def forward_dynamics(self, tau, dt=1.0 / 3000.0, env=0):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        qdd = np.zeros(self.ndof)
        self.L.orc_forward_dynamics(self.h, env, _p(tau), C.c_double(dt), _p(qdd))
        return qdd

    def net_eval(self, x, n_out=90):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(n_out)
        self.L.orc_net_eval(self.h, _p(x), _p(y))
        return y

    def com(self, env=0):
        c = np.zeros(2); v = np.zeros(2)
        self.L.orc_com(self.h, env, _p(c), _p(v))
        return c, v

    def tuples(self):
        n = self.L.orc_num_tuples(self.h)
        W = 1 + self.S + self.A + self.S
        rows = np.zeros((max(n, 1), W)); flags = np.zeros(max(n, 1), np.uint32); ids = np.zeros(max(n, 1), np.int32)
        m = self.L.orc_get_tuples(self.h, _p(rows), _p(flags), _p(ids), n)
        return rows[:m], flags[:m], ids[:m]

    def reset_tuples(self):
        self.L.orc_reset_tuples(self.h)

    def eval_stats(self):
        c = C.c_int64(0); e = C.c_int64(0); a = C.c_double(0); s = C.c_int64(0)
        self.L.orc_eval_stats(self.h, C.byref(c), C.byref(e), C.byref(a), C.byref(s))
        return dict(cycles=c.value, episodes=e.value, avg_dist=a.value, steps=s.value)

    def dist_log(self, env=0, cap=4096):
        d = np.zeros(cap)
        n = self.L.orc_dist_log(self.h, env, _p(d), cap)
        return d[:n].copy()
