"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY; used by tests/, smoke() and bench.py cpu_baseline)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def host_has_avx2_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":", 1)[1].split())
                    return "avx2" in fl and "fma" in fl
    except OSError:
        pass
    return False


_FAST = None


def lib(fast=False):
    """fast=True: the -march=x86-64-v3 build (bench.py's CPU baseline); falls back to the portable build on a host without AVX2/FMA"""
    global _LIB, _FAST
    if fast and host_has_avx2_fma():
        if _FAST is None:
            path = os.path.join(_HERE, "_build", "liboracle_fast.so")
            if not os.path.exists(path):
                build()
            _FAST = _bind(C.CDLL(path))
        return _FAST
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = _bind(C.CDLL(path))
    return _LIB


def _bind(L):
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    L.orc_last_error.restype = C.c_char_p
    L.orc_sample_height.restype = C.c_double
    for name in ("orc_destroy", "orc_update", "orc_env_step", "orc_reset", "orc_get_state", "orc_set_state",
                 "orc_get_last_tau", "orc_get_poli_state", "orc_get_net_out", "orc_rbd", "orc_forward_dynamics",
                 "orc_net_eval", "orc_com", "orc_reset_tuples", "orc_eval_stats", "orc_set_phys",
                 "orc_set_explore"):
        getattr(L, name).restype = None
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, pack, num_envs=1, mode=0, terrain_seeds=None, rng_seed=1234, fast=False):
        self.L = lib(fast)
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

    def flags(self, env=0):
        """(has_fallen, has_stumbled, cycle_count)"""
        out = np.zeros(3, np.int32)
        self.L.orc_flags.restype = None
        self.L.orc_flags(self.h, env, _p(out))
        return out.tolist()

    def calc_reward(self, env=0):
        self.L.orc_calc_reward.restype = C.c_double
        return self.L.orc_calc_reward(self.h, env)

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

    def rbd_extra(self, what, env=0):
        """what: 'gravity' [ndof], 'jacobian' [6, ndof], 'joint_pos' [nj, 3] of the controller's RBD model at the env's state."""
        idx = {"gravity": 0, "jacobian": 1, "joint_pos": 2}[what]
        out = np.zeros({0: self.ndof, 1: 6 * self.ndof, 2: 3 * self.nj}[idx])
        self.L.orc_rbd_extra.restype = None
        self.L.orc_rbd_extra(self.h, env, idx, _p(out))
        return out if idx == 0 else out.reshape((6, self.ndof) if idx == 1 else (self.nj, 3))

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


class OracleTrainer:
    """CPU restatement of cMACETrainer + the Caffe SGD solver step (oracle/trainer.h)."""

    DEFAULTS = dict(replay_cap=500000, num_init_samples=200, num_steps_per_iter=1, freeze_target_iters=0,
                    init_input_offset_scale=1, discount=0.9, base_lr=1e-3, momentum=0.9, weight_decay=5e-4, seed=1)

    def __init__(self, pack, **kw):
        self.L = lib()
        L = self.L
        L.orc_trainer_create.restype = C.c_void_p
        L.orc_trainer_create.argtypes = [C.c_char_p, C.c_void_p]
        L.orc_trainer_loss_grad.restype = C.c_double
        for name in ("orc_trainer_destroy", "orc_trainer_add_tuples", "orc_trainer_train", "orc_trainer_get",
                     "orc_trainer_set_theta", "orc_trainer_counters", "orc_trainer_losses", "orc_trainer_eval_batch"):
            getattr(L, name).restype = None
        p = dict(self.DEFAULTS)
        p.update(kw)
        self.params = p
        arr = np.array([p[k] for k in ("replay_cap", "num_init_samples", "num_steps_per_iter", "freeze_target_iters",
                                       "init_input_offset_scale", "discount", "base_lr", "momentum", "weight_decay", "seed")], float)
        h = L.orc_trainer_create(pack.encode(), _p(arr))
        if not h:
            raise RuntimeError(L.orc_last_error().decode())
        self.h = C.c_void_p(h)
        self.num_params = L.orc_trainer_num_params(self.h)
        self.W = L.orc_trainer_tuple_width(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_trainer_destroy(self.h)
            self.h = None

    def add_tuples(self, rows, flags):
        rows = np.ascontiguousarray(rows, np.float64)
        flags = np.ascontiguousarray(flags, np.uint32)
        assert rows.shape[1] == self.W
        self.L.orc_trainer_add_tuples(self.h, _p(rows), _p(flags), rows.shape[0])

    def train(self):
        self.L.orc_trainer_train(self.h)

    def get(self, what, n=None):
        idx = {"theta": 0, "target": 1, "history": 2, "in_off": 3, "in_scale": 4, "out_off": 5, "out_scale": 6}[what]
        if n is None:
            n = self.num_params if idx < 3 else (self.n_in if idx < 5 else self.n_out)
        out = np.zeros(n)
        self.L.orc_trainer_get(self.h, idx, _p(out))
        return out

    def set_theta(self, theta):
        theta = np.ascontiguousarray(theta, np.float64)
        self.L.orc_trainer_set_theta(self.h, _p(theta))

    def counters(self):
        c = np.zeros(9, np.int64)
        self.L.orc_trainer_counters(self.h, _p(c))
        return dict(zip(("iter", "actor_iter", "stage", "num", "head", "total", "critic", "actor", "actor_batch"), c.tolist()))

    def losses(self):
        l = np.zeros(2)
        self.L.orc_trainer_losses(self.h, _p(l))
        return l

    def lists(self, which, cap=1 << 20):
        out = np.zeros(cap, np.int32)
        n = self.L.orc_trainer_lists(self.h, {"critic": 0, "actor": 1, "actor_batch": 2, "last_critic": 3, "last_actor": 4}[which], _p(out), cap)
        return out[:n].copy()

    def rows(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        rows = np.zeros((ids.size, self.W), np.float32)
        flags = np.zeros(ids.size, np.int32)
        self.L.orc_trainer_get_rows.restype = None
        self.L.orc_trainer_get_rows(self.h, _p(ids), ids.size, _p(rows), _p(flags))
        return rows, flags

    def loss_grad(self, X, Y, want_grad=True):
        X = np.ascontiguousarray(X, np.float64); Y = np.ascontiguousarray(Y, np.float64)
        g = np.zeros(self.num_params) if want_grad else None
        loss = self.L.orc_trainer_loss_grad(self.h, _p(X), _p(Y), X.shape[0], _p(g))
        return loss, g

    def eval_batch(self, X, target=False):
        X = np.ascontiguousarray(X, np.float64)
        Y = np.zeros((X.shape[0], self.n_out))
        self.L.orc_trainer_eval_batch(self.h, int(target), _p(X), X.shape[0], _p(Y))
        return Y

    @property
    def n_in(self):
        return (self.W - 1 - self._a()) // 2

    def _a(self):
        # tuple row = 1 + S + A + S with A = 1 + frag; frag is 29 (dog) / 28 (raptor): S = (W - 1 - A) / 2 must be integral
        for a in (30, 29):
            if (self.W - 1 - a) % 2 == 0:
                return a
        raise RuntimeError("unexpected tuple width")

    @property
    def n_out(self):
        return 3 + 3 * (self._a() - 1)
