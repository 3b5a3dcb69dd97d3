"""ctypes front-end of oracle/_build/libi2lqr_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ilqr_iterative_tasks_b200._abi import OUTER_ITERS, Obstacle, Params  # struct layouts only

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libi2lqr_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "i2lqr_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_relax_cost.restype = C.c_double
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _obs_ptr(obs):
    if obs is None:
        return None, 0
    if isinstance(obs, Obstacle):
        return C.byref(obs), 1
    arr = np.ascontiguousarray(obs)
    return arr.ctypes.data_as(C.c_void_p), arr.shape[0]


def ilqr(params, x0, x_term, lamb_in, obstacle=None):
    """One candidate solve; returns (uvar (2,N), xvar (4,N+1), lamb_out, iters)."""
    N = params.num_horizon
    uvar = np.zeros((2, N))
    xvar = np.zeros((4, N + 1))
    lamb_out = C.c_double()
    iters = C.c_int32()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    x_term = np.ascontiguousarray(x_term, dtype=np.float64)
    optr = C.byref(obstacle) if obstacle is not None else None
    rc = lib().orc_ilqr(C.byref(params), optr, _d(x0), _d(x_term), C.c_double(lamb_in), _d(uvar), _d(xvar), C.byref(lamb_out), C.byref(iters))
    assert rc == 0
    return uvar, xvar, lamb_out.value, iters.value


def ilqr_batch(params, x0, x_term, lamb_in, obs=None, nthreads=1, uvar_init=None):
    """B independent solves on `nthreads` host threads; obs: None, one Obstacle, or a structured array (B)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    x_term = np.ascontiguousarray(x_term, dtype=np.float64)
    lamb_in = np.ascontiguousarray(lamb_in, dtype=np.float64)
    B, N = x0.shape[0], params.num_horizon
    uvar = np.zeros((B, 2, N))
    xvar = np.zeros((B, 4, N + 1))
    lamb_out = np.zeros(B)
    iters = np.zeros(B, dtype=np.int32)
    optr, n_obs = _obs_ptr(obs)
    uinit = None if uvar_init is None else np.ascontiguousarray(uvar_init, dtype=np.float64).reshape(B, 2, N)
    rc = lib().orc_ilqr_batch(
        C.byref(params), C.c_int64(B), _d(x0), _d(x_term), _d(lamb_in), optr, C.c_int64(n_obs), _d(uinit),
        _d(uvar), _d(xvar), _d(lamb_out), iters.ctypes.data_as(_ip), C.c_int(nthreads),
    )  # fmt: skip
    assert rc == 0
    return uvar, xvar, lamb_out, iters


def backward_pass(params, xvar, uvar, dX, x_term, lamb, obstacle=None):
    """control/iterative_ilqr.py:88-130 for one problem, reference layout: -> k (2,N), K (2,4,N)."""
    N = params.num_horizon
    xvar, uvar, dX = (np.ascontiguousarray(a, dtype=np.float64) for a in (xvar, uvar, dX))
    x_term = np.ascontiguousarray(x_term, dtype=np.float64)
    k, K = np.zeros((2, N)), np.zeros((2, 4, N))
    optr = C.byref(obstacle) if obstacle is not None else None
    rc = lib().orc_backward_pass(C.byref(params), optr, _d(xvar), _d(uvar), _d(dX), _d(x_term), C.c_double(lamb), _d(k), _d(K))
    assert rc == 0
    return k, K


def forward_pass(params, xvar, uvar, x_term, k, K):
    """control/iterative_ilqr.py:133-160 for one problem: -> xvar_new (4,N+1), uvar_new (2,N), cost_new."""
    N = params.num_horizon
    xvar, uvar, k, K = (np.ascontiguousarray(a, dtype=np.float64) for a in (xvar, uvar, k, K))
    x_term = np.ascontiguousarray(x_term, dtype=np.float64)
    xn, un, cost = np.zeros((4, N + 1)), np.zeros((2, N)), C.c_double()
    rc = lib().orc_forward_pass(C.byref(params), _d(xvar), _d(uvar), _d(x_term), _d(k), _d(K), _d(xn), _d(un), C.byref(cost))
    assert rc == 0
    return xn, un, cost.value


def np_argsort(v):
    """numpy.argsort(v) as the scalar introsort of NumPy does it (tie order included)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty(v.shape[0], dtype=np.int64)
    lib().orc_np_argsort(_d(v), C.c_int64(v.shape[0]), out.ctypes.data_as(_lp))
    return out


def select_knn(lap, x0, K):
    lap = np.ascontiguousarray(lap, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    idx = np.zeros(K, dtype=np.int64)
    n = lib().orc_select_knn(_d(lap), C.c_int(lap.shape[0]), _d(x0), C.c_int(K), idx.ctypes.data_as(_lp))
    return idx[:n]


def relax_cost(params, xN, x_term, qfun, it):
    xN = np.ascontiguousarray(xN, dtype=np.float64)
    x_term = np.ascontiguousarray(x_term, dtype=np.float64)
    return lib().orc_relax_cost(C.byref(params), _d(xN), _d(x_term), C.c_double(qfun), C.c_int(it))


def kinetic_bicycle(x, u, dt):
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros(4)
    lib().orc_kinetic_bicycle(_d(x), _d(u), C.c_double(dt), _d(out))
    return out


def set_extra_obstacles(obstacles):
    """EXTENSION (obstacle lists): further obstacles whose barriers are added to every following `ilqr` / `calc_input` /
    `backward_pass` call of this thread; None or [] switches back to the reference's single obstacle.  Returns the ctypes
    array, which the caller must keep alive while it is set."""
    from ilqr_iterative_tasks_b200 import _abi

    if not obstacles:
        lib().orc_set_extra_obstacles(None, C.c_int(0))
        return None
    arr = (_abi.Obstacle * len(obstacles))(*obstacles)
    lib().orc_set_extra_obstacles(arr, C.c_int(len(obstacles)))
    return arr


def calc_input(params, laps, x, obstacle=None):
    """Solve branch of calc_input for one controller; laps = list of (T,4) arrays (all stored laps)."""
    N, K, L = params.num_horizon, params.num_ss_points, params.num_ss_iter
    laps = [np.ascontiguousarray(l, dtype=np.float64) for l in laps]
    ptrs = (_dp * len(laps))(*[_d(l) for l in laps])
    lens = np.array([l.shape[0] for l in laps], dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = {
        "u": np.zeros(2), "x_pred": np.zeros((4, N + 1)), "u_pred": np.zeros((2, N)),
        "sel_idx": np.zeros((OUTER_ITERS, L, K), dtype=np.int64), "cand_cost": np.zeros((OUTER_ITERS, L, K)),
        "cand_lamb": np.zeros((OUTER_ITERS, L, K)), "cand_iters": np.zeros((OUTER_ITERS, L, K), dtype=np.int32),
    }  # fmt: skip
    wl, wr, sh = C.c_int32(), C.c_int32(), C.c_int32()
    wi, wc = C.c_int64(), C.c_double()
    optr = C.byref(obstacle) if obstacle is not None else None
    rc = lib().orc_calc_input(
        C.byref(params), optr, C.c_int(len(laps)), ptrs, lens.ctypes.data_as(_ip), _d(x),
        _d(out["u"]), _d(out["x_pred"]), _d(out["u_pred"]), C.byref(wl), C.byref(wr), C.byref(wi), C.byref(wc), C.byref(sh),
        out["sel_idx"].ctypes.data_as(_lp), _d(out["cand_cost"]), _d(out["cand_lamb"]), out["cand_iters"].ctypes.data_as(_ip),
    )  # fmt: skip
    assert rc == 0
    out.update(winner_lap=wl.value, winner_rank=wr.value, winner_idx=wi.value, winner_cost=wc.value, shrink=sh.value)
    return out
