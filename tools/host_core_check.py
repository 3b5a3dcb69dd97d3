"""Developer screen: run the host-compiled device core (tools/host_core_check.cpp) over the flat goldens and a
large oracle sample, report discrete-path mismatches and worst relative error.  Not part of the product."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqr_iterative_tasks_b200 import _abi, synth  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import helpers as H  # noqa: E402

SO = "/tmp/libhostcore.so"
subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=fast", "-std=c++17", "-shared", "-fPIC", "-o", SO, os.path.join(ROOT, "tools", "host_core_check.cpp")])
lib = C.CDLL(SO)
lib.host_fill_kparams.restype = C.c_char_p
dp = C.POINTER(C.c_double)


def solve(p, x0, xt, lamb, obs, uinit=None):
    kp = C.create_string_buffer(lib.host_kparams_size())
    err = lib.host_fill_kparams(C.byref(p), kp)
    assert not err, err
    B, N = x0.shape[0], p.num_horizon
    u = np.zeros((B, 2, N)); x = np.zeros((B, 4, N + 1)); lo = np.zeros(B); it = np.zeros(B, dtype=np.int32)
    x0 = np.ascontiguousarray(x0); xt = np.ascontiguousarray(xt); lamb = np.ascontiguousarray(lamb)
    optr = obs.ctypes.data_as(C.c_void_p) if obs is not None else None
    lib.host_solve(kp, C.c_int64(B), x0.ctypes.data_as(dp), xt.ctypes.data_as(dp), lamb.ctypes.data_as(dp), optr,
                   C.c_int64(0 if obs is None else B), None if uinit is None else np.ascontiguousarray(uinit).ctypes.data_as(dp),
                   u.ctypes.data_as(dp), x.ctypes.data_as(dp), lo.ctypes.data_as(dp), it.ctypes.data_as(C.POINTER(C.c_int32)))
    return u, x, lo, it


def ulps(a, ref):
    return np.abs(a - ref) / np.spacing(np.abs(ref))


rng = np.random.default_rng(0)
for lo, hi in ((-12, 3), (-700, -600), (-50, 50), (700, 709)):
    xs = rng.uniform(lo, hi, 200000)
    ys = np.empty_like(xs)
    lib.host_exp_fast(xs.ctypes.data_as(dp), ys.ctypes.data_as(dp), C.c_int64(xs.size))
    ref = np.exp(xs)
    ok = ref > 1e-300
    print("exp_fast [%g,%g]: max ulp vs numpy %.2f" % (lo, hi, ulps(ys[ok], ref[ok]).max()))
for span in (1.0, 10.0, 1000.0, 9.9e4):
    xs = rng.uniform(-span, span, 200000)
    s = np.empty_like(xs); c = np.empty_like(xs)
    lib.host_sincos_fast(xs.ctypes.data_as(dp), s.ctypes.data_as(dp), c.ctypes.data_as(dp), C.c_int64(xs.size))
    print("sincos_fast |x|<%g: max ulp sin %.2f cos %.2f" % (span, ulps(s, np.sin(xs)).max(), ulps(c, np.cos(xs)).max()))

g = H.load("ilqr_flat.npz")
for grp in H.FLAT_GROUPS:
    p, inp, obs, exp = H.flat_group(g, grp)
    u, x, lo, it = solve(p, inp["x0"], inp["x_term"], inp["lamb_in"], obs, inp["uvar_init"])
    ok = (it == exp["iters"]) & (lo == exp["lamb_out"])
    ex = H.rel_err(x[ok], exp["xvar"][ok]).max(); eu = H.rel_err(u[ok], exp["uvar"][ok]).max()
    print("%-9s discrete mismatches %d / %d   rel err x %.2e u %.2e" % (grp, (~ok).sum(), len(ok), ex, eu))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = synth.flat_candidates(n, 6, seed=99)
p = _abi.default_params(); obs = _abi.obstacle_array(d["obs"], d["mode"])
u, x, lo, it = solve(p, d["x0"], d["x_term"], d["lamb_in"], obs)
uo, xo, loo, ito = oracle.ilqr_batch(p, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=8)
ok = (it == ito) & (lo == loo)
print("oracle sample %d: discrete mismatches %d (%.4f%%), rel err x %.2e u %.2e, mean iters %.3f" % (n, (~ok).sum(), 100 * (~ok).mean(), H.rel_err(x[ok], xo[ok]).max(), H.rel_err(u[ok], uo[ok]).max(), it.mean()))

# ---- FP32 build vs FP64 (BASELINE config 4 gate: <= 1e-4 where the discrete path matches)
def solve32(p, x0, xt, lamb, obs):
    kp = C.create_string_buffer(lib.host_kparams_size())
    lib.host_fill_kparams(C.byref(p), kp)
    B, N = x0.shape[0], p.num_horizon
    u = np.zeros((B, 2, N)); x = np.zeros((B, 4, N + 1)); lo = np.zeros(B); it = np.zeros(B, dtype=np.int32)
    x0 = np.ascontiguousarray(x0); xt = np.ascontiguousarray(xt); lamb = np.ascontiguousarray(lamb)
    lib.host_solve_f32(kp, C.c_int64(B), x0.ctypes.data_as(dp), xt.ctypes.data_as(dp), lamb.ctypes.data_as(dp),
                       obs.ctypes.data_as(C.c_void_p), C.c_int64(B), u.ctypes.data_as(dp), x.ctypes.data_as(dp),
                       lo.ctypes.data_as(dp), it.ctypes.data_as(C.POINTER(C.c_int32)))
    return u, x, lo, it


for N in (6, 10, 25, 50):
    d = synth.flat_candidates(4096, N, seed=5 + N)
    p = _abi.default_params(num_horizon=N); obs = _abi.obstacle_array(d["obs"], d["mode"])
    u64, x64, l64, i64 = solve(p, d["x0"], d["x_term"], d["lamb_in"], obs)
    u32, x32, l32, i32 = solve32(p, d["x0"], d["x_term"], d["lamb_in"], obs)
    same = i64 == i32
    ex = H.rel_err(x32[same], x64[same]); eu = H.rel_err(u32[same], u64[same])
    print("FP32 N=%d: same iteration count %.1f%%; of those rel err x p50 %.1e p99 %.1e max %.1e | u max %.1e; within 1e-4: %.1f%%"
          % (N, 100 * same.mean(), np.median(ex), np.percentile(ex, 99), ex.max(), eu.max(), 100 * (np.maximum(ex, eu) <= 1e-4).mean()))
