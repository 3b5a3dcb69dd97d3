"""Shared helpers for the parity tests."""
import os

import numpy as np

from ilqr_iterative_tasks_b200 import _abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLAT_GROUPS = ["n6", "n6_lamb", "n6_noobs", "n6_qr", "n6_warm", "n2", "n10", "n25", "n50", "n100", "n200"]
STEP_SCENARIOS = ["stock", "moving_up", "moving_left", "pytest", "static_big", "k64"]
# FP64 parity bar of BASELINE.json north_star: states / inputs / costs within 1e-9 relative
RTOL = 1e-9


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def flat_group(g, grp):
    """-> (params, inputs dict, obstacle array or None, expected dict) for one group of ilqr_flat.npz."""
    N = int(g[grp + "/N"])
    p = _abi.default_params(num_horizon=N, matrix_Q=g[grp + "/Q"], matrix_R=g[grp + "/R"])
    obs = _abi.obstacle_array(g[grp + "/obs"], g[grp + "/mode"]) if int(g[grp + "/has_obstacle"]) else None
    inp = {k: g[grp + "/" + k] for k in ("x0", "x_term", "lamb_in", "uvar_init")}
    if not inp["uvar_init"].any():
        inp["uvar_init"] = None
    exp = {k: g[grp + "/" + k] for k in ("uvar", "xvar", "lamb_out", "iters")}
    return p, inp, obs, exp


def rel_err(a, ref):
    """max |a-ref| over a candidate's array, relative to max(1, max|ref|) of that candidate."""
    a = np.asarray(a)
    ref = np.asarray(ref)
    ax = tuple(range(1, ref.ndim))
    return np.abs(a - ref).max(axis=ax) / np.maximum(1.0, np.abs(ref).max(axis=ax))


def step_obstacle(g, s):
    mode = int(g["step/mode"][s])
    if mode < 0:
        return None
    return _abi.Obstacle(*[float(v) for v in g["step/obs"][s]], mode, 1)


def step_params(g):
    return _abi.default_params(num_horizon=int(g["N"]), num_ss_points=int(g["K"]), num_ss_iter=int(g["L"]))
