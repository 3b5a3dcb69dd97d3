"""Developer screen for chain_core.cuh: the host emulation of one lane group (tools/host_core_check.cpp,
host_chain_solve) against the generic host-compiled core on lambda chains.  Not part of the product."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqr_iterative_tasks_b200 import _abi, synth  # noqa: E402

SO = "/tmp/libhostcore.so"
subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=fast", "-std=c++17", "-shared", "-fPIC", "-o", SO, os.path.join(ROOT, "tools", "host_core_check.cpp")])
lib = C.CDLL(SO)
lib.host_fill_kparams.restype = C.c_char_p
dp = C.POINTER(C.c_double)


def kparams(p):
    kp = C.create_string_buffer(lib.host_kparams_size())
    err = lib.host_fill_kparams(C.byref(p), kp)
    assert not err, err
    return kp


def generic_chain(p, x0, xts, obs):
    kp = kparams(p)
    N = p.num_horizon
    n = len(xts)
    U = np.zeros((n, 2, N)); X = np.zeros((n, 4, N + 1)); LO = np.zeros(n); IT = np.zeros(n, dtype=np.int32)
    lam = np.array([p.lamb])
    for j in range(n):
        u = np.zeros((1, 2, N)); x = np.zeros((1, 4, N + 1)); lo = np.zeros(1); it = np.zeros(1, dtype=np.int32)
        xx = np.ascontiguousarray(x0.reshape(1, 4)); xt = np.ascontiguousarray(xts[j].reshape(1, 4))
        lib.host_solve(kp, C.c_int64(1), xx.ctypes.data_as(dp), xt.ctypes.data_as(dp), lam.ctypes.data_as(dp),
                       C.byref(obs) if obs is not None else None, C.c_int64(1 if obs is not None else 0), None,
                       u.ctypes.data_as(dp), x.ctypes.data_as(dp), lo.ctypes.data_as(dp), it.ctypes.data_as(C.POINTER(C.c_int32)))
        U[j], X[j], LO[j], IT[j] = u[0], x[0], lo[0], it[0]
        lam = lo.copy()
    return U, X, LO, IT


def group_chain(p, x0, xts, obs, rungs):
    kp = kparams(p)
    N = p.num_horizon
    n = len(xts)
    U = np.zeros((n, 2, N)); X = np.zeros((n, 4, N + 1)); LO = np.zeros(n); IT = np.zeros(n, dtype=np.int32)
    xx = np.ascontiguousarray(x0); xt = np.ascontiguousarray(xts)
    lib.host_chain_solve(kp, C.c_int(rungs), C.c_int(n), xx.ctypes.data_as(dp), xt.ctypes.data_as(dp),
                         C.byref(obs) if obs is not None else None, U.ctypes.data_as(dp), X.ctypes.data_as(dp),
                         LO.ctypes.data_as(dp), IT.ctypes.data_as(C.POINTER(C.c_int32)))
    return U, X, LO, IT


rng = np.random.default_rng(7)
bad = 0
worst = 0.0
for trial in range(60):
    N = int(rng.choice([2, 3, 6, 6, 6, 7, 10, 13]))
    qr = trial % 5 == 4
    kw = {}
    if qr:
        kw = dict(matrix_Q=np.diag([0.1, 0.2, 0.05, 0.01]), matrix_R=np.array([[0.1, 0.02], [0.02, 0.3]]))
    p = _abi.default_params(num_horizon=N, **kw)
    d = synth.flat_candidates(16, N, seed=100 + trial)
    x0 = d["x0"][0]
    xts = d["x_term"] - d["x0"] + x0  # targets around this x0
    mode = int(d["mode"][0])
    obs = _abi.Obstacle(*[float(v) for v in d["obs"][0]], mode, 1) if trial % 7 != 6 else None
    ref = generic_chain(p, x0, xts, obs)
    for rungs in (1, 2, 4):
        got = group_chain(p, x0, xts, obs, rungs)
        same = np.array_equal(ref[3], got[3]) and np.array_equal(ref[2], got[2])
        ex = (np.abs(ref[1] - got[1]) / np.maximum(1.0, np.abs(ref[1]))).max(); eu = np.abs(ref[0] - got[0]).max()
        worst = max(worst, ex, eu)
        if not same or ex > 1e-9 or eu > 1e-9:  # host FMA contraction differs between the two code shapes by ulps
            bad += 1
            print("trial %d N %d qr %d rungs %d: discrete same %s, |dx| %.2e |du| %.2e iters %s vs %s" % (trial, N, qr, rungs, same, ex, eu, ref[3][:6], got[3][:6]))
print("host chain check: %d mismatching (trial, rungs) cases out of %d, worst difference %.2e" % (bad, 60 * 3, worst))
