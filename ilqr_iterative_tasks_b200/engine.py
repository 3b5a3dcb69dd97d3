"""ctypes binding of libi2lqr.so (include/i2lqr.h): the only door from Python into the CUDA path.

There is no CPU fallback.  If the shared library is missing, or no sm_100-class GPU is usable, every entry
point raises RuntimeError -- the exception the reference's caller already handles
(iterative_ilqr/utils/base.py:146-155).
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import OUTER_ITERS, Obstacle, Params  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("I2LQR_LIB") or os.path.join(_HERE, "libi2lqr.so")  # I2LQR_LIB: A/B builds of the same library
_lib = None

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)

# every symbol include/i2lqr.h declares
EXPORTS = [
    "i2lqr_abi_version", "i2lqr_last_error", "i2lqr_default_params", "i2lqr_create", "i2lqr_destroy",
    "i2lqr_set_params", "i2lqr_solve_batch", "i2lqr_solve_batch_device", "i2lqr_add_lap", "i2lqr_num_laps",
    "i2lqr_set_obstacle", "i2lqr_select_knn", "i2lqr_solve_step", "i2lqr_get_step_trace",
    "i2lqr_episodes_begin", "i2lqr_episodes_run_lap", "i2lqr_episodes_status", "i2lqr_episodes_set_noise", "i2lqr_last_stats", "i2lqr_sync_stats",
    "i2lqr_measure_fp64_peak", "i2lqr_split_row_doubles", "i2lqr_split_begin", "i2lqr_split_solve", "i2lqr_split_merge",
    "i2lqr_split_end", "i2lqr_dynamics_batch", "i2lqr_linearize_batch", "i2lqr_cost_derivatives_batch",
    "i2lqr_cost_final_batch", "i2lqr_backward_pass_batch", "i2lqr_forward_pass_batch",
    "i2lqr_add_lap_all", "i2lqr_set_obstacles", "i2lqr_spec_stats", "i2lqr_spec_timeline", "i2lqr_set_obstacle_list",
]  # fmt: skip


def lib():
    """Loads libi2lqr.so; raises RuntimeError (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH
            )
        L = C.CDLL(LIB_PATH)
        L.i2lqr_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("i2lqr: " + lib().i2lqr_last_error().decode())


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def measure_fp64_peak(device=0, reps=5):
    t = C.c_double()
    _check(lib().i2lqr_measure_fp64_peak(C.c_int(device), C.c_int(reps), C.byref(t)))
    return t.value


class Engine:
    """One GPU, one stream, `n_episodes` independent safe sets (see i2lqr_create)."""

    spec_set_size = int(os.environ.get("I2LQR_SPEC_S", "32"))  # mirrors the library's tuning switch (developer tools only)

    def __init__(self, params=None, device=0, n_episodes=0, lap_capacity=0):
        self.params = params if params is not None else _abi.default_params()
        self.device = device
        self.n_episodes = n_episodes
        self._h = C.c_void_p()
        _check(lib().i2lqr_create(C.byref(self.params), C.c_int(device), C.c_int(n_episodes), C.c_int(lap_capacity), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().i2lqr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters ------------------------------------------------------------------------------------
    def set_params(self, params):
        _check(lib().i2lqr_set_params(self._h, C.byref(params)))
        self.params = params

    @property
    def N(self):
        return self.params.num_horizon

    # -- flat batch ------------------------------------------------------------------------------------
    @staticmethod
    def _obs_arg(obs, B):
        if obs is None:
            return None, 0, None
        if isinstance(obs, Obstacle):
            if not obs.present:
                return None, 0, None
            return C.cast(C.byref(obs), C.c_void_p), 1, obs
        arr = np.ascontiguousarray(obs, dtype=_abi.OBSTACLE_DTYPE)
        if arr.shape[0] not in (1, B):
            raise ValueError("obstacle array needs 1 or B entries")
        return C.c_void_p(arr.ctypes.data), arr.shape[0], arr

    def solve_batch(self, x0, x_term, lamb_in=None, obs=None, want=("uvar", "xvar", "lamb_out", "iters"), out=None, uvar_init=None):
        """B independent `ilqr()` solves with host arrays (H2D + solve + D2H inside the call).

        Returns a dict with uvar (B,2,N), xvar (B,4,N+1), lamb_out (B), iters (B).  `out` may carry
        preallocated (e.g. pinned) arrays of those names.
        """
        x0 = _f64(x0).reshape(-1, 4)
        B, N = x0.shape[0], self.N
        x_term = _f64(x_term, (B, 4))
        lamb = None if lamb_in is None else _f64(lamb_in, (B,))
        uinit = None if uvar_init is None else _f64(uvar_init, (B, 2, N))
        optr, n_obs, keep = self._obs_arg(obs, B)
        res = dict(out) if out else {}
        shapes = {"uvar": ((B, 2, N), np.float64), "xvar": ((B, 4, N + 1), np.float64), "lamb_out": ((B,), np.float64), "iters": ((B,), np.int32)}
        for k in want:
            if k not in res:
                res[k] = np.empty(shapes[k][0], dtype=shapes[k][1])
        it = res.get("iters")
        _check(lib().i2lqr_solve_batch(
            self._h, C.c_int64(B), _d(x0), _d(x_term), _d(lamb), optr, C.c_int64(n_obs), _d(uinit),
            _d(res.get("uvar")), _d(res.get("xvar")), _d(res.get("lamb_out")), None if it is None else it.ctypes.data_as(_i32p),
        ))  # fmt: skip
        del keep
        return res

    def solve_batch_device(self, B, x0, x_term, lamb_in, obs, n_obs, uvar, xvar, lamb_out, iters, stream=None, uvar_init=0):
        """Raw device-pointer form (ints from e.g. torch.Tensor.data_ptr()); asynchronous."""
        vp = lambda p: C.c_void_p(p) if p else None  # noqa: E731
        _check(lib().i2lqr_solve_batch_device(
            self._h, C.c_int64(B), vp(x0), vp(x_term), vp(lamb_in), vp(obs), C.c_int64(n_obs), vp(uvar_init),
            vp(uvar), vp(xvar), vp(lamb_out), vp(iters), vp(stream),
        ))  # fmt: skip

    def sync_stats(self, stream=None):
        _check(lib().i2lqr_sync_stats(self._h, C.c_void_p(stream) if stream else None))
        return self.last_stats()

    # -- safe set --------------------------------------------------------------------------------------
    def add_lap(self, episode, x):
        x = _f64(x).reshape(-1, 4)
        _check(lib().i2lqr_add_lap(self._h, C.c_int(episode), _d(x), C.c_int(x.shape[0])))

    def add_lap_all(self, x):
        """The same lap appended to every episode's safe set (one upload)."""
        x = _f64(x).reshape(-1, 4)
        _check(lib().i2lqr_add_lap_all(self._h, _d(x), C.c_int(x.shape[0])))

    def set_obstacles(self, obstacles, first=0):
        """Obstacles of episodes [first, first + len(obstacles)) from a structured array (None clears all)."""
        if obstacles is None:
            _check(lib().i2lqr_set_obstacles(self._h, C.c_int(0), C.c_int(self.n_episodes), None))
            return
        arr = np.ascontiguousarray(obstacles, dtype=_abi.OBSTACLE_DTYPE)
        _check(lib().i2lqr_set_obstacles(self._h, C.c_int(first), C.c_int(arr.shape[0]), C.c_void_p(arr.ctypes.data)))

    def num_laps(self, episode=0):
        n = C.c_int()
        _check(lib().i2lqr_num_laps(self._h, C.c_int(episode), C.byref(n)))
        return n.value

    def set_obstacle(self, episode, obstacle):
        ptr = None if obstacle is None else C.byref(obstacle)
        _check(lib().i2lqr_set_obstacle(self._h, C.c_int(episode), ptr))

    def set_obstacle_list(self, episode, obstacles):
        """Extension: up to four obstacles for one episode (their barrier terms add); a sequence of `Obstacle` structs or a
        structured array."""
        if obstacles is None or len(obstacles) == 0:
            _check(lib().i2lqr_set_obstacle_list(self._h, C.c_int(episode), None, C.c_int(0)))
            return
        if isinstance(obstacles[0], Obstacle):
            arr = (Obstacle * len(obstacles))(*obstacles)
            _check(lib().i2lqr_set_obstacle_list(self._h, C.c_int(episode), arr, C.c_int(len(obstacles))))
        else:
            arr = np.ascontiguousarray(obstacles, dtype=_abi.OBSTACLE_DTYPE)
            _check(lib().i2lqr_set_obstacle_list(self._h, C.c_int(episode), C.c_void_p(arr.ctypes.data), C.c_int(arr.shape[0])))

    def select_knn(self, episode, lap, x0, K):
        x0 = _f64(x0, (4,))
        idx = np.empty(K, dtype=np.int64)
        n = C.c_int32()
        _check(lib().i2lqr_select_knn(self._h, C.c_int(episode), C.c_int(lap), _d(x0), C.c_int(K), idx.ctypes.data_as(_i64p), C.byref(n)))
        return idx[: n.value]

    # -- controller step -------------------------------------------------------------------------------
    def _step_outputs(self, E):
        N = self.N
        return {
            "u": np.empty((E, 2)), "x_pred": np.empty((E, 4, N + 1)), "u_pred": np.empty((E, 2, N)),
            "winner_lap": np.empty(E, dtype=np.int32), "winner_rank": np.empty(E, dtype=np.int32),
            "winner_idx": np.empty(E, dtype=np.int64), "winner_cost": np.empty(E), "shrink": np.empty(E, dtype=np.int32),
        }  # fmt: skip

    def solve_step(self, x, n_episodes=None):
        x = _f64(x).reshape(-1, 4)
        E = x.shape[0] if n_episodes is None else n_episodes
        out = self._step_outputs(E)
        _check(lib().i2lqr_solve_step(
            self._h, C.c_int(E), _d(x), _d(out["u"]), _d(out["x_pred"]), _d(out["u_pred"]),
            out["winner_lap"].ctypes.data_as(_i32p), out["winner_rank"].ctypes.data_as(_i32p),
            out["winner_idx"].ctypes.data_as(_i64p), _d(out["winner_cost"]), out["shrink"].ctypes.data_as(_i32p),
        ))  # fmt: skip
        return out

    def step_trace(self, n_episodes=1):
        K, L = self.params.num_ss_points, self.params.num_ss_iter
        shp = (n_episodes, OUTER_ITERS, L, K)
        t = {"sel_idx": np.empty(shp, dtype=np.int64), "cand_cost": np.empty(shp), "cand_lamb": np.empty(shp), "cand_iters": np.empty(shp, dtype=np.int32)}
        _check(lib().i2lqr_get_step_trace(
            self._h, C.c_int(n_episodes), t["sel_idx"].ctypes.data_as(_i64p), _d(t["cand_cost"]), _d(t["cand_lamb"]),
            t["cand_iters"].ctypes.data_as(_i32p),
        ))  # fmt: skip
        return t

    # -- controller step split over handles (one per GPU) --------------------------------------------------
    def split_row_doubles(self):
        n = C.c_int64()
        _check(lib().i2lqr_split_row_doubles(self._h, C.byref(n)))
        return n.value

    def split_begin(self, x, stream=None):
        x = _f64(x).reshape(-1, 4)
        _check(lib().i2lqr_split_begin(self._h, C.c_int(x.shape[0]), _d(x), C.c_void_p(stream) if stream else None))
        return x.shape[0]

    def split_solve(self, E, it, part, n_parts, rows_ptr, stream=None):
        _check(lib().i2lqr_split_solve(self._h, C.c_int(E), C.c_int(it), C.c_int(part), C.c_int(n_parts), C.c_void_p(rows_ptr), C.c_void_p(stream) if stream else None))

    def split_merge(self, E, it, part, n_parts, all_rows_ptr, stream=None):
        _check(lib().i2lqr_split_merge(self._h, C.c_int(E), C.c_int(it), C.c_int(part), C.c_int(n_parts), C.c_void_p(all_rows_ptr), C.c_void_p(stream) if stream else None))

    def split_end(self, E, stream=None):
        out = self._step_outputs(E)
        _check(lib().i2lqr_split_end(
            self._h, C.c_int(E), _d(out["u"]), _d(out["x_pred"]), _d(out["u_pred"]),
            out["winner_lap"].ctypes.data_as(_i32p), out["winner_rank"].ctypes.data_as(_i32p),
            out["winner_idx"].ctypes.data_as(_i64p), _d(out["winner_cost"]), out["shrink"].ctypes.data_as(_i32p),
            C.c_void_p(stream) if stream else None,
        ))  # fmt: skip
        return out

    # -- the functions an iteration is made of (include/i2lqr.h, "one by one") ------------------------------
    def dynamics(self, x, u):
        x, u = _f64(x).reshape(-1, 4), _f64(u).reshape(-1, 2)
        out = np.empty_like(x)
        _check(lib().i2lqr_dynamics_batch(self._h, C.c_int64(x.shape[0]), _d(x), _d(u), _d(out)))
        return out

    def linearize(self, v, theta, accel, want_A=True, want_B=True):
        N = self.N
        theta = _f64(theta).reshape(-1, N)
        B = theta.shape[0]
        v = None if v is None else _f64(v, (B, N))
        accel = None if accel is None else _f64(accel, (B, N))
        A = np.empty((B, 4, 4, N)) if want_A else None
        Bm = np.empty((B, 4, 2, N)) if want_B else None
        _check(lib().i2lqr_linearize_batch(self._h, C.c_int64(B), _d(v), _d(theta), _d(accel), _d(A), _d(Bm)))
        return A, Bm

    def cost_derivatives(self, uvar, dX, xvar, obs=None):
        N = self.N
        xvar = _f64(xvar).reshape(-1, 4, N + 1)
        B = xvar.shape[0]
        uvar, dX = _f64(uvar, (B, 2, N)), _f64(dX, (B, 4, N + 1))
        optr, n_obs, keep = self._obs_arg(obs, B)
        l_u, l_uu, l_x, l_xx = np.empty((B, 2, N)), np.empty((B, 2, 2, N)), np.empty((B, 4, N)), np.empty((B, 4, 4, N))
        _check(lib().i2lqr_cost_derivatives_batch(self._h, C.c_int64(B), _d(uvar), _d(dX), _d(xvar), optr, C.c_int64(n_obs), _d(l_u), _d(l_uu), _d(l_x), _d(l_xx)))
        del keep
        return l_u, l_uu, l_x, l_xx

    def cost_final(self, xvar, x_term, obs=None):
        N = self.N
        xvar = _f64(xvar).reshape(-1, 4, N + 1)
        B = xvar.shape[0]
        x_term = _f64(x_term, (B, 4))
        optr, n_obs, keep = self._obs_arg(obs, B)
        l_x, l_xx = np.empty((B, 4)), np.empty((B, 4, 4))
        _check(lib().i2lqr_cost_final_batch(self._h, C.c_int64(B), _d(xvar), _d(x_term), optr, C.c_int64(n_obs), _d(l_x), _d(l_xx)))
        del keep
        return l_x, l_xx

    def backward_pass(self, xvar, uvar, x_term, lamb=None, obs=None):
        N = self.N
        xvar = _f64(xvar).reshape(-1, 4, N + 1)
        B = xvar.shape[0]
        uvar, x_term = _f64(uvar, (B, 2, N)), _f64(x_term, (B, 4))
        lamb = None if lamb is None else _f64(lamb, (B,))
        optr, n_obs, keep = self._obs_arg(obs, B)
        k, K = np.empty((B, 2, N)), np.empty((B, 2, 4, N))
        _check(lib().i2lqr_backward_pass_batch(self._h, C.c_int64(B), _d(xvar), _d(uvar), _d(x_term), _d(lamb), optr, C.c_int64(n_obs), _d(k), _d(K)))
        del keep
        return k, K

    def forward_pass(self, xvar, uvar, x_term, k, K):
        N = self.N
        xvar = _f64(xvar).reshape(-1, 4, N + 1)
        B = xvar.shape[0]
        uvar, x_term, k, K = _f64(uvar, (B, 2, N)), _f64(x_term, (B, 4)), _f64(k, (B, 2, N)), _f64(K, (B, 2, 4, N))
        xn, un, cost = np.empty((B, 4, N + 1)), np.empty((B, 2, N)), np.empty(B)
        _check(lib().i2lqr_forward_pass_batch(self._h, C.c_int64(B), _d(xvar), _d(uvar), _d(x_term), _d(k), _d(K), _d(xn), _d(un), _d(cost)))
        return xn, un, cost

    # -- batched closed-loop episodes ------------------------------------------------------------------
    def episodes_begin(self, x_init, goal):
        x_init = _f64(x_init, (self.n_episodes, 4))
        goal = _f64(goal, (4,))
        _check(lib().i2lqr_episodes_begin(self._h, _d(x_init), _d(goal)))

    def episodes_run_lap(self, want_traj=False):
        E, S = self.n_episodes, self.params.max_sim_steps
        lap_len = np.empty(E, dtype=np.int32)
        lx = np.empty((E, S + 1, 4)) if want_traj else None
        lu = np.empty((E, S, 2)) if want_traj else None
        _check(lib().i2lqr_episodes_run_lap(self._h, lap_len.ctypes.data_as(_i32p), _d(lx), _d(lu)))
        return (lap_len, lx, lu) if want_traj else lap_len

    def episodes_set_noise(self, seed):
        _check(lib().i2lqr_episodes_set_noise(self._h, C.c_uint64(seed)))

    def episodes_status(self):
        st = np.empty(self.n_episodes, dtype=np.int32)
        _check(lib().i2lqr_episodes_status(self._h, st.ctypes.data_as(_i32p)))
        return st

    def spec_stats(self):
        """Cumulative counters of the speculative one-launch controller step (include/i2lqr.h, i2lqr_spec_stats)."""
        v = (C.c_int64 * 8)()
        _check(lib().i2lqr_spec_stats(self._h, v))
        names = ("hits", "waited", "self_solved", "misses", "spec_finished", "spec_abandoned", "launches")
        out = {k: int(v[i]) for i, k in enumerate(names)}
        served = out["hits"] + out["waited"] + out["self_solved"] + out["misses"]
        out["hit_rate"] = (out["hits"] + out["waited"]) / served if served else None
        return out

    def spec_timeline(self, max_chains=64):
        """(3, c_max, 8) uint64 table of the last one-launch step (needs I2LQR_SPEC_TIMELINE=1), or None."""
        K, S = self.params.num_ss_points, 32
        buf = np.zeros(3 * max_chains * 8 + 4 * 3 * max_chains * K * S + 9216, dtype=np.uint64)
        cm = C.c_int32()
        _check(lib().i2lqr_spec_timeline(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(buf.size), C.byref(cm)))
        if not cm.value:
            return None
        n_items = 4 * 3 * cm.value * K * self.spec_set_size
        self.spec_item_debug = buf[3 * cm.value * 8 : 3 * cm.value * 8 + n_items]  # (entries, 4): start ns, end ns, rounds, iterations
        self.spec_worker_debug = buf[3 * cm.value * 8 + n_items : 3 * cm.value * 8 + n_items + 4096].reshape(1024, 4)
        self.spec_trip_profile = buf[3 * cm.value * 8 + n_items + 4096 : 3 * cm.value * 8 + n_items + 9216].reshape(1024, 5)  # I2LQR_TRIP_PROFILE builds
        return buf[: 3 * cm.value * 8].reshape(3, cm.value, 8)

    # -- instrumentation -------------------------------------------------------------------------------
    def last_stats(self):
        ms, n, it, sv = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().i2lqr_last_stats(self._h, C.byref(ms), C.byref(n), C.byref(it), C.byref(sv)))
        return {"kernel_ms": ms.value, "launches": n.value, "ilqr_iterations": it.value, "solves": sv.value}
