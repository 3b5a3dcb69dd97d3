"""Shared helpers for the parity tests."""
import os

import numpy as np

from ilqr_iterative_tasks_b200 import _abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLAT_GROUPS = ["n6", "n6_lamb", "n6_noobs", "n6_qr", "n6_warm", "n2", "n10", "n25", "n50", "n100", "n200"]
STEP_SCENARIOS = ["stock", "moving_up", "moving_left", "pytest", "static_big", "k64", "add_static", "add_up", "add_left"]
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


def same_discrete_path(out_iters, out_lamb, ref_iters, ref_lamb, what):
    """Mask of candidates whose discrete path (executed iterations, lambda_out bit for bit) equals the oracle's /
    reference's.  north_star asks for exact iteration counts and lambda, so the divergence count is printed (pytest -s /
    the captured-output section) and must be ZERO: a flipped accept / reject would need |J_new - J| within an ulp."""
    same = (np.asarray(out_iters) == np.asarray(ref_iters)) & (np.asarray(out_lamb) == np.asarray(ref_lamb))
    n_div = int((~same).sum())
    print("[parity] %s: discrete path diverged for %d of %d candidates" % (what, n_div, same.size))
    assert n_div == 0, "%s: discrete path diverged for %d of %d candidates (first: %s)" % (what, n_div, same.size, np.flatnonzero(~same)[:8])
    return same


def step_obstacle(g, s):
    mode = int(g["step/mode"][s])
    if mode < 0:
        return None
    return _abi.Obstacle(*[float(v) for v in g["step/obs"][s]], mode, 1)


def step_params(g):
    return _abi.default_params(num_horizon=int(g["N"]), num_ss_points=int(g["K"]), num_ss_iter=int(g["L"]))


def oracle_closed_loop(params, x0, obs_row, mode, initial_lap, n_laps, extra=None):
    """Closed loop of one episode driven by the CPU oracle: the reference's Simulator.sim
    (iterative_ilqr/utils/base.py:708-726), forward_one_step (:157-213), obstacle motion (:50-64), the
    controller's playback branch (:377-382) and the lap post-processing of
    iterative_ilqr/tests/ilqr_test.py:56-60.  Returns the list of laps (each (T,4), last row = goal)."""
    from oracle import oracle

    N, S, dt = params.num_horizon, params.max_sim_steps, params.timestep
    laps = [np.asarray(initial_lap, dtype=np.float64)]
    goal = laps[0][-1]
    x = np.asarray(x0, dtype=np.float64).copy()
    for _ in range(n_laps):
        obs = None if obs_row is None else _abi.Obstacle(*[float(v) for v in obs_row], int(mode), 1)
        # EXTENSION (obstacle lists): further obstacles as (row, mode) pairs; they move and reset like the first one
        more = [_abi.Obstacle(*[float(v) for v in r], int(m), 1) for r, m in (extra or [])]
        keep = oracle.set_extra_obstacles(more)
        horizon, u_old = N, None
        states = [x.copy()]
        for i in range(S):
            if horizon < N:  # playback
                u = u_old[:, 0].copy()
                u_old = u_old[:, 1:]
                horizon -= 1
            else:
                out = oracle.calc_input(params, laps, x, obs)
                u = out["u"].copy()
                u_old = out["u_pred"][:, 1:].copy()
                if out["shrink"]:
                    horizon -= 1
            x = oracle.kinetic_bicycle(x, u, dt)
            states.append(x.copy())
            for o in ([obs] if obs is not None else []) + more:
                if o.moving_option == 1:
                    o.y += o.spd * dt
                if o.moving_option == 2:
                    o.x -= o.spd * dt
            if keep is not None:
                for i, o in enumerate(more):  # the oracle reads the ctypes copy
                    keep[i].x, keep[i].y = o.x, o.y
            if np.linalg.norm(x - goal) <= params.goal_tol:
                break
        lap = np.array(states)
        lap[-1] = goal
        laps.append(lap)
        x = np.zeros(4)
        oracle.set_extra_obstacles(None)
    return laps
