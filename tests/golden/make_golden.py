"""Mint golden vectors by running the unmodified Python reference (HybridRobotics/ilqr-iterative-tasks).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py all          # every fixture, in parallel processes
    python tests/golden/make_golden.py flat|knn|steps_stock|...   # a single one

The reference's own tests hold no assertions and no golden vectors for this path, so
the fixtures below are the pin: outputs of the reference itself, executed here on NumPy 2.3 / SciPy 1.18,
committed as small .npz files together with this script.

Tie policy: `np.argsort` with the default kind (iterative_ilqr/utils/base.py:339) leaves equal
distances in an order that depends on which sort NumPy runs.  The reference pins numpy==1.19.5, whose float64
argsort is the scalar introsort; NumPy >= 1.25 dispatches to x86-simd-sort on AVX2 / AVX-512 hosts, with a
hardware-specific tie order.  Every fixture is therefore minted from the UNPATCHED reference with NumPy's SIMD
dispatch switched off through NumPy's own environment switch (this script re-executes itself with
NPY_DISABLE_CPU_FEATURES set), i.e. with the sort the pinned version uses.  No monkeypatching of argsort.
"""
import os
import subprocess
import sys

SCALAR_SORT = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX512_KNL AVX512_KNM AVX2"
if os.environ.get("NPY_DISABLE_CPU_FEATURES") is None:  # must be in the environment before numpy is imported
    os.environ["NPY_DISABLE_CPU_FEATURES"] = SCALAR_SORT
    os.execv(sys.executable, [sys.executable] + sys.argv)

import warnings

warnings.filterwarnings("ignore", message=".*NPY_DISABLE_CPU_FEATURES.*")

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ilqr_iterative_tasks_b200 import synth  # noqa: E402
sys.path.insert(0, HERE)
from refharness import Reference  # noqa: E402


def _param(ref, N, K=8, L=1, Q=None, R=None):
    kw = dict(num_ss_points=K, num_ss_iter=L, timestep=1, num_horizon=N)
    if Q is not None:
        kw["matrix_Q"] = Q
    if R is not None:
        kw["matrix_R"] = R
    return ref.base.iLqrParam(**kw)


def _run_flat_group(ref, N, B, seed, lamb_choices=None, with_obstacle=True, Q=None, R=None, warm=False):
    data = synth.flat_candidates(B, num_horizon=N, seed=seed)
    rng = np.random.default_rng(seed + 1)
    if lamb_choices is not None:
        data["lamb_in"] = rng.choice(np.array(lamb_choices, dtype=np.float64), size=B)
    param = _param(ref, N, Q=Q, R=R)
    # warm start of a direct ilqr() call; some inputs lie outside the clip bounds on purpose
    data["uvar_init"] = rng.standard_normal((B, 2, N)) * np.array([1.5, 0.9])[None, :, None] if warm else np.zeros((B, 2, N))
    uvar = np.zeros((B, 2, N))
    xvar = np.zeros((B, 4, N + 1))
    lamb_out = np.zeros(B)
    iters = np.zeros(B, dtype=np.int32)
    for b in range(B):
        obstacle = ref.make_obstacle(data["obs"][b], data["mode"][b]) if with_obstacle else None
        u, x, lo, it = ref.call_ilqr(
            param, N, data["x0"][b], data["x_term"][b], obstacle, float(data["lamb_in"][b]), uvar_init=data["uvar_init"][b]
        )
        uvar[b], xvar[b], lamb_out[b], iters[b] = u, x, lo, it
    out = dict(data)
    out.update(uvar=uvar, xvar=xvar, lamb_out=lamb_out, iters=iters, N=np.int32(N), has_obstacle=np.int32(with_obstacle))
    out["Q"] = np.zeros((4, 4)) if Q is None else np.array(Q, dtype=np.float64)
    out["R"] = np.zeros((2, 2)) if R is None else np.array(R, dtype=np.float64)
    return out


def make_flat():
    """Single `ilqr()` calls (iterative_ilqr/control/iterative_ilqr.py:7-85) on the config-4a/5 generator."""
    ref = Reference(stable_argsort=False)
    groups = {}
    groups["n6"] = _run_flat_group(ref, 6, 256, synth.SEED)
    groups["n6_lamb"] = _run_flat_group(ref, 6, 96, synth.SEED + 11, lamb_choices=[1e-3, 1e-1, 10.0, 1e2, 1e3, 1e4])
    groups["n6_noobs"] = _run_flat_group(ref, 6, 32, synth.SEED + 12, with_obstacle=False)
    groups["n6_qr"] = _run_flat_group(
        ref, 6, 48, synth.SEED + 13, Q=np.diag([0.1, 0.1, 0.05, 0.01]), R=np.diag([0.05, 0.05])
    )
    groups["n6_warm"] = _run_flat_group(ref, 6, 48, synth.SEED + 14, warm=True)
    for N, B in ((2, 16), (10, 24), (25, 16), (50, 8), (100, 4), (200, 4)):
        groups["n%d" % N] = _run_flat_group(ref, N, B, synth.SEED + N)
    flat = {}
    for g, d in groups.items():
        for k, v in d.items():
            flat["%s/%s" % (g, k)] = v
    np.savez_compressed(os.path.join(HERE, "ilqr_flat.npz"), **flat)
    ref.close()
    for g, d in groups.items():
        print(g, "iters mean %.2f max %d" % (d["iters"].mean(), d["iters"].max()))


def make_knn():
    """`iLqr.select_close_ss` (iterative_ilqr/utils/base.py:332-341)."""
    out = {}
    rng = np.random.default_rng(synth.SEED + 100)
    ref = Reference(stable_argsort=False)
    base = ref.base
    case = 0
    for T in (3, 5, 8, 22, 33, 64, 121, 500, 3000):
        for K in (8, 64):
            ctrl = base.iLqr(_param(ref, 2, K=K), system_param=base.KineticBicycleParam())
            lap = np.cumsum(rng.standard_normal((T, 4)) * np.array([2.0, 1.0, 0.3, 0.1]), axis=0)
            ctrl.ss.append(lap.T.copy())
            x0 = lap[rng.integers(0, T)] + rng.standard_normal(4) * 0.5
            with ref.quiet():
                idx = np.array(ctrl.select_close_ss(0, x0))
            # this group must be uniquely determined: the K+1 smallest distances pairwise distinct
            d = np.sort(np.abs(lap - x0).sum(axis=1))[: K + 1]
            assert np.all(np.diff(d) > 0)
            out["c%d/lap" % case] = lap
            out["c%d/x0" % case] = x0
            out["c%d/K" % case] = np.int32(K)
            out["c%d/idx" % case] = idx.astype(np.int64)
            out["c%d/stable" % case] = np.int32(0)
            case += 1
    ref.close()
    out["n_unpatched"] = np.int32(case)
    # tie cases: the stock first lap ends with 3 identical rows (flag name kept: 1 = ties reach the selection)
    ref = Reference(stable_argsort=False)
    base = ref.base
    xcl, _ = synth.initial_lap()
    for K, x0 in ((8, xcl[117] + 0.01), (8, xcl[-1].copy()), (64, xcl[100] - 0.2), (8, xcl[115] + np.array([0.3, 0, 0.1, 0]))):
        ctrl = base.iLqr(_param(ref, 2, K=K), system_param=base.KineticBicycleParam())
        ctrl.ss.append(xcl.T.copy())
        with ref.quiet():
            idx = np.array(ctrl.select_close_ss(0, x0))
        out["c%d/lap" % case] = xcl
        out["c%d/x0" % case] = x0
        out["c%d/K" % case] = np.int32(K)
        out["c%d/idx" % case] = idx.astype(np.int64)
        out["c%d/stable" % case] = np.int32(1)
        case += 1
    # many-way ties: integer grid lap
    grid = np.array([[i % 5, i // 5 % 5, 0.0, 0.0] for i in range(60)], dtype=np.float64)
    for K in (8, 64):
        ctrl = base.iLqr(_param(ref, 2, K=K), system_param=base.KineticBicycleParam())
        ctrl.ss.append(grid.T.copy())
        x0 = np.array([2.0, 2.0, 0.0, 0.0])
        with ref.quiet():
            idx = np.array(ctrl.select_close_ss(0, x0))
        out["c%d/lap" % case] = grid
        out["c%d/x0" % case] = x0
        out["c%d/K" % case] = np.int32(K)
        out["c%d/idx" % case] = idx.astype(np.int64)
        out["c%d/stable" % case] = np.int32(1)
        case += 1
    out["n_cases"] = np.int32(case)
    np.savez_compressed(os.path.join(HERE, "knn.npz"), **out)
    ref.close()
    print("knn cases", case)


def _obs_row(obstacle):
    if obstacle is None:
        return np.full(5, np.nan), -1
    spd = 0.0 if obstacle.spd is None else float(obstacle.spd)
    mode = 0 if obstacle.moving_option is None else int(obstacle.moving_option)
    return np.array([obstacle.x, obstacle.y, obstacle.width, obstacle.height, spd], dtype=np.float64), mode


def record_closed_loop(name, K, L, laps, obstacle_plan, append_goal_row=False, N=6):
    """Closed loop as the CLI script runs it (iterative_ilqr/tests/ilqr_test.py:8-78), every
    `calc_input` (iterative_ilqr/utils/base.py:371-479) recorded with its inputs, outputs and candidates.

    obstacle_plan(lap_index, base) -> Obstacle or None, evaluated before each lap
    (as iterative_ilqr/result/ilqr_test_add_moving_obstacle.py:63-75 does).
    """
    ref = Reference(stable_argsort=False)
    base = ref.base
    dt = 1
    with ref.quiet():
        ego = base.KineticBicycle(False, system_param=base.KineticBicycleParam())
        ego.set_state([0, 0, 0, 0])
        ego.set_timestep(dt)
        ego.get_traj()
        ego.set_zero_noise()
    param = _param(ref, N, K=K, L=L)
    ctrl = base.iLqr(param, obstacle=obstacle_plan(0, base), system_param=base.KineticBicycleParam())
    ctrl.add_trajectory(ego.xcl, ego.ucl)
    ctrl.set_timestep(dt)
    ego.set_ctrl_policy(ctrl)
    sim = base.Simulator()
    sim.set_robotic(ego)
    sim.set_timestep(dt)
    sim.set_traj()

    steps = []
    sel_log = []
    orig_sel = base.iLqr.select_close_ss

    def sel(self, it, x0):
        idx = orig_sel(self, it, x0)
        sel_log.append(np.array(idx, dtype=np.int64))
        return idx

    base.iLqr.select_close_ss = sel
    orig_calc = base.iLqr.calc_input

    def calc(self):
        rec = {}
        rec["x"] = np.array(self.x, dtype=np.float64).copy()
        rec["n_laps"] = self.iter
        rec["num_horizon_before"] = self.num_horizon
        rec["obs"], rec["mode"] = _obs_row(self.obstacle)
        rec["playback"] = int(self.num_horizon < self.ilqr_param.num_horizon)
        ref.trace = []
        del sel_log[:]
        orig_calc(self)
        rec["u"] = np.array(self.u, dtype=np.float64).copy()
        rec["num_horizon_after"] = self.num_horizon
        n_act = min(self.iter, L)
        if not rec["playback"]:
            rec["x_pred"] = np.array(self.x_pred).copy()
            rec["u_pred"] = np.array(self.u_pred).copy()
            sel_idx = np.full((3, L, K), -1, dtype=np.int64)
            c_lamb = np.full((3, L, K), np.nan)
            c_iters = np.full((3, L, K), -1, dtype=np.int32)
            c_xN = np.full((3, L, K, 4), np.nan)
            c_u0 = np.full((3, L, K, 2), np.nan)
            t = 0
            s = 0
            for it in range(3):
                for l in range(n_act):
                    idx = sel_log[s]
                    s += 1
                    sel_idx[it, l, : len(idx)] = idx
                    for r in range(len(idx)):
                        c = ref.trace[t]
                        t += 1
                        c_lamb[it, l, r] = c["lamb_out"]
                        c_iters[it, l, r] = c["iters"]
                        c_xN[it, l, r] = c["xvar"][:, -1]
                        c_u0[it, l, r] = c["uvar"][:, 0]
            assert t == len(ref.trace) and s == len(sel_log)
            rec.update(sel_idx=sel_idx, c_lamb=c_lamb, c_iters=c_iters, c_xN=c_xN, c_u0=c_u0)
        ref.trace = None
        steps.append(rec)

    base.iLqr.calc_input = calc
    lap_list = [np.array(ego.xcl).copy()]
    lap_first_step = []
    with ref.quiet():
        for lap in range(laps):
            ctrl.obstacle = obstacle_plan(lap, base)
            lap_first_step.append(len(steps))
            sim.sim(lap, sim_time=50)
            if append_goal_row:  # top-level tests/ilqr_test.py:51
                ego.data["state"][-1] = np.vstack((ego.data["state"][-1], ego.xcl[-1, :]))
            else:  # iterative_ilqr/tests/ilqr_test.py:59
                ego.data["state"][-1][-1, :] = ego.xcl[-1, :]
            ctrl.add_trajectory(ego.data["state"][-1], ego.data["input"][-1])
            lap_list.append(np.array(ego.data["state"][-1]).copy())
    base.iLqr.select_close_ss = orig_sel
    base.iLqr.calc_input = orig_calc

    out = {"K": np.int32(K), "L": np.int32(L), "N": np.int32(N), "n_steps": np.int32(len(steps)), "n_laps": np.int32(len(lap_list))}
    out["append_goal_row"] = np.int32(append_goal_row)
    out["lap_first_step"] = np.array(lap_first_step, dtype=np.int32)
    for i, lp in enumerate(lap_list):
        out["lap%d" % i] = lp
    for i, inp in enumerate(ego.data["input"]):
        out["lap_inputs%d" % (i + 1)] = np.array(inp)
    keys_all = ["x", "n_laps", "num_horizon_before", "obs", "mode", "playback", "u", "num_horizon_after"]
    for k in keys_all:
        out["step/" + k] = np.array([s[k] for s in steps])
    solve_ids = [i for i, s in enumerate(steps) if not s["playback"]]
    out["solve_ids"] = np.array(solve_ids, dtype=np.int32)
    for k in ["x_pred", "u_pred", "sel_idx", "c_lamb", "c_iters", "c_xN", "c_u0"]:
        out["solve/" + k] = np.array([steps[i][k] for i in solve_ids])
    np.savez_compressed(os.path.join(HERE, "steps_%s.npz" % name), **out)
    ref.close()
    print(name, "laps", [len(l) for l in lap_list], "steps", len(steps), "solve steps", len(solve_ids))


def make_steps_stock():
    record_closed_loop("stock", 8, 2, 3, lambda lap, base: base.Obstacle(31, -3, 8, 6))


def make_steps_moving_up():
    def plan(lap, base):
        return base.Obstacle(35, -16, 34, 34, spd=1, timestep=1, moving_option=1) if lap == 1 else None

    record_closed_loop("moving_up", 8, 2, 3, plan)


def make_steps_moving_left():
    def plan(lap, base):
        return base.Obstacle(50, -1, 35, 35, spd=0.2, timestep=1, moving_option=2) if lap == 1 else None

    record_closed_loop("moving_left", 8, 2, 2, plan)


def make_steps_pytest():
    record_closed_loop("pytest", 8, 1, 2, lambda lap, base: base.Obstacle(31, -2, 8, 6), append_goal_row=True)


def make_steps_k64():
    record_closed_loop("k64", 64, 3, 2, lambda lap, base: base.Obstacle(31, -3, 8, 6))


def make_steps_static_big():
    # result/ilqr_test_static_obstacle.py:36-39 obstacle constants
    record_closed_loop("static_big", 8, 2, 2, lambda lap, base: base.Obstacle(100, -5, 20, 40))


def make_steps_add_static():
    # result/ilqr_test_add_static_obstacle.py:49-58: a static obstacle appears on lap 5 of a run that has converged without one
    record_closed_loop("add_static", 8, 2, 7, lambda lap, base: base.Obstacle(35, 0, 30, 30) if lap == 5 else None)


def make_steps_add_up():
    # result/ilqr_test_add_moving_obstacle.py:17-23,63-75 (--moving-option up)
    def plan(lap, base):
        return base.Obstacle(35, -16, 34, 34, spd=1, timestep=1, moving_option=1) if lap == 5 else None

    record_closed_loop("add_up", 8, 2, 7, plan)


def make_steps_add_left():
    # result/ilqr_test_add_moving_obstacle.py:24-30,63-75 (--moving-option left)
    def plan(lap, base):
        return base.Obstacle(50, -1, 35, 35, spd=0.2, timestep=1, moving_option=2) if lap == 5 else None

    record_closed_loop("add_left", 8, 2, 7, plan)


def make_step_synth():
    """Config 4b: one full `calc_input` on a synthetic one-lap safe set for several horizons."""
    out = {}
    horizons = (6, 10, 25, 50)
    for N in horizons:
        ref = Reference(stable_argsort=False)
        base = ref.base
        lap, u, x = synth.synthetic_lap(N)
        obs = np.array([lap[N // 2, 0] + 1.0, lap[N // 2, 1] - 2.5, 4.0, 3.0, 0.3])
        obstacle = base.Obstacle(obs[0], obs[1], obs[2], obs[3], spd=obs[4], timestep=1, moving_option=1)
        ctrl = base.iLqr(_param(ref, N, K=8, L=1), obstacle=obstacle, system_param=base.KineticBicycleParam())
        ctrl.add_trajectory(lap, u)
        ctrl.set_timestep(1)
        ctrl.set_state(x)
        ref.trace = []
        with ref.quiet():
            ctrl.calc_input()
        g = "n%d/" % N
        out[g + "lap"], out[g + "x"], out[g + "obs"], out[g + "mode"] = lap, x, obs, np.int32(1)
        out[g + "u"], out[g + "x_pred"], out[g + "u_pred"] = np.array(ctrl.u), np.array(ctrl.x_pred), np.array(ctrl.u_pred)
        out[g + "num_horizon_after"] = np.int32(ctrl.num_horizon)
        out[g + "c_lamb"] = np.array([c["lamb_out"] for c in ref.trace])
        out[g + "c_iters"] = np.array([c["iters"] for c in ref.trace], dtype=np.int32)
        ref.trace = None
        ref.close()
        # fresh interpreter state is not needed: Reference() re-imports cached modules
    out["horizons"] = np.array(horizons, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "step_synth.npz"), **out)
    print("step_synth done")


def make_components():
    """The functions an iLQR iteration is made of, called one by one on mid-solve trajectories
    (iterative_ilqr/systems/kinetic_bicycle.py:10-52, control/ilqr_helper.py:9-150, control/iterative_ilqr.py:88-160)."""
    ref = Reference(stable_argsort=False)
    base, mod = ref.base, ref.ilqr_mod
    import control.ilqr_helper as helper  # the reference's module (on sys.path through Reference)
    import systems.kinetic_bicycle as sysmod

    sysp = base.KineticBicycleParam()
    out = {}
    groups = [("n6", 6, 64, None, None), ("n6_qr", 6, 48, np.diag([0.1, 0.1, 0.05, 0.01]) + 0.01, np.array([[0.05, 0.01], [0.01, 0.08]])),
              ("n25", 25, 16, None, None)]  # fmt: skip
    for name, N, B, Q, R in groups:
        data = synth.flat_candidates(B, num_horizon=N, seed=synth.SEED + 900 + N + (0 if Q is None else 1))
        rng = np.random.default_rng(synth.SEED + 77 + N)
        param = _param(ref, N, Q=Q, R=R)
        lamb = rng.choice(np.array([1e-2, 1.0, 100.0]), size=B)
        xvar, uvar = np.zeros((B, 4, N + 1)), np.zeros((B, 2, N))
        A, Bm = np.zeros((B, 4, 4, N)), np.zeros((B, 4, 2, N))
        l_u, l_uu, l_x, l_xx = np.zeros((B, 2, N)), np.zeros((B, 2, 2, N)), np.zeros((B, 4, N)), np.zeros((B, 4, 4, N))
        Vx, Vxx = np.zeros((B, 4)), np.zeros((B, 4, 4))
        gk, gK = np.zeros((B, 2, N)), np.zeros((B, 2, 4, N))
        xn, un, cn = np.zeros((B, 4, N + 1)), np.zeros((B, 2, N)), np.zeros(B)
        cc_b, cc_bdd = np.zeros((B, N, 2)), np.zeros((B, N, 2, 2))  # add_control_constraint per stage (ilqr_helper.py:83-103)
        for b in range(B):
            # a trajectory as ilqr() holds it mid-solve: clipped inputs, states rolled out with the model
            u = rng.standard_normal((2, N)) * np.array([[0.8], [0.25]])
            u[0] = np.clip(u[0], -sysp.a_max, sysp.a_max)
            u[1] = np.clip(u[1], -round(sysp.delta_max, 2), round(sysp.delta_max, 2))
            x = np.zeros((4, N + 1))
            x[:, 0] = data["x0"][b]
            for k in range(N):
                x[:, k + 1] = sysmod.kinetic_bicycle(x[:, k], u[:, k], 1)
            xvar[b], uvar[b] = x, u
            dX = x - np.zeros((4, 1))
            obstacle = ref.make_obstacle(data["obs"][b], data["mode"][b]) if b % 5 else None
            xt = np.array(data["x_term"][b])
            A[b] = sysmod.get_A_matrix(x[2, 1:], x[3, 1:], u[0, :], N, 1)
            Bm[b] = sysmod.get_B_matrix(x[3, 1:], N, 1)
            l_u[b], l_uu[b], l_x[b], l_xx[b] = helper.get_cost_derivation(u, dX, param, N, x, obstacle, sysp)
            Vx[b], Vxx[b] = helper.get_cost_final(x, xt, param.matrix_Qterminal, obstacle, param)
            for k in range(N):
                bd, bdd = helper.add_control_constraint(u[:, k], param, sysp)
                cc_b[b, k], cc_bdd[b, k] = np.asarray(bd)[:, 0], np.asarray(bdd)
            gk[b], gK[b] = mod._pristine[0](x, u, xt, dX, float(lamb[b]), N, 1, param, obstacle, sysp)
            xn[b], un[b], cn[b] = mod.forward_pass(x, u, xt, param, 1, N, gk[b], gK[b], sysp)
        has_obs = np.array([1 if b % 5 else 0 for b in range(B)], dtype=np.int32)
        g = name + "/"
        for k, v in dict(N=np.int32(N), x_term=data["x_term"], obs=data["obs"], mode=data["mode"], has_obs=has_obs, lamb=lamb,
                         Q=np.zeros((4, 4)) if Q is None else Q, R=np.zeros((2, 2)) if R is None else R, xvar=xvar, uvar=uvar, A=A, B=Bm,
                         l_u=l_u, l_uu=l_uu, l_x=l_x, l_xx=l_xx, cc_b=cc_b, cc_bdd=cc_bdd, Vx=Vx, Vxx=Vxx, k=gk, K=gK, xvar_new=xn, uvar_new=un, cost_new=cn).items():  # fmt: skip
            out[g + k] = v
    out["groups"] = np.array([g[0] for g in groups])
    # the scripted first lap as KineticBicycle.get_traj logs it (utils/base.py:103-138): states and the input log,
    # whose first row aliases the array the loop mutates on step 1
    ego = base.KineticBicycle(False, system_param=sysp)
    ego.set_state(np.zeros(4))
    ego.set_timestep(1)
    with ref.quiet():
        ego.get_traj()
    out["traj/xcl"], out["traj/ucl"] = np.array(ego.xcl), np.array(ego.ucl)
    np.savez_compressed(os.path.join(HERE, "components.npz"), **out)
    ref.close()
    print("components done")


def make_multi_obstacle():
    """EXTENSION goldens (obstacle lists, include/i2lqr.h i2lqr_set_obstacle_list).  The reference's cost functions take one
    Obstacle; a list is priced as the sum of its members' barrier terms.  That composition is done HERE, outside the
    reference, by wrapping the two cost functions `backward_pass` calls (control/iterative_ilqr.py:101-107): each
    member's contribution is the reference's own output for that obstacle minus its output for `None`.  The wrapped
    functions then drive the reference's unmodified `ilqr()`.  matrix_Q = matrix_R = 0 (the default), so the stage terms
    add exactly; the terminal gradient carries the rounding of one subtraction per member (< 1 ulp)."""
    ref = Reference(stable_argsort=False)
    mod = ref.ilqr_mod
    orig_cd, orig_cf = mod.get_cost_derivation, mod.get_cost_final

    def cost_derivation(ctrl_U, dX, ilqr_param, num_horizon, xvar, obstacle, sys_param):
        if not isinstance(obstacle, list):
            return orig_cd(ctrl_U, dX, ilqr_param, num_horizon, xvar, obstacle, sys_param)
        base = [np.array(a, dtype=np.float64) for a in orig_cd(ctrl_U, dX, ilqr_param, num_horizon, xvar, None, sys_param)]
        l_x, l_xx = base[2].copy(), base[3].copy()
        for o in obstacle:
            r = orig_cd(ctrl_U, dX, ilqr_param, num_horizon, xvar, o, sys_param)
            l_x = l_x + (np.asarray(r[2]) - base[2])
            l_xx = l_xx + (np.asarray(r[3]) - base[3])
        return base[0], base[1], l_x, l_xx

    def cost_final(x, x_terminal, matrix_Qterminal, obstacle, ilqr_param):
        if not isinstance(obstacle, list):
            return orig_cf(x, x_terminal, matrix_Qterminal, obstacle, ilqr_param)
        b0, b1 = [np.array(a, dtype=np.float64) for a in orig_cf(x, x_terminal, matrix_Qterminal, None, ilqr_param)]
        Vx, Vxx = b0.copy(), b1.copy()
        for o in obstacle:
            r = orig_cf(x, x_terminal, matrix_Qterminal, o, ilqr_param)
            Vx = Vx + (np.asarray(r[0]) - b0)
            Vxx = Vxx + (np.asarray(r[1]) - b1)
        return Vx, Vxx

    mod.get_cost_derivation, mod.get_cost_final = cost_derivation, cost_final
    try:
        N, B, M = 6, 64, 3
        data = synth.flat_candidates(B, num_horizon=N, seed=synth.SEED + 41)
        rng = np.random.default_rng(synth.SEED + 42)
        data["lamb_in"] = rng.choice(np.array([1.0, 1.0, 1.0, 1e-1, 10.0, 1e2]), size=B)
        param = _param(ref, N)
        members = np.stack([(np.arange(B) + s) % B for s in (0, 7, 13)], axis=1)  # candidate b faces obstacles b, b+7, b+13
        n_list = 1 + (np.arange(B) % M)  # lists of one, two and three obstacles
        uvar, xvar = np.zeros((B, 2, N)), np.zeros((B, 4, N + 1))
        lamb_out, iters = np.zeros(B), np.zeros(B, dtype=np.int32)
        for b in range(B):
            obstacles = [ref.make_obstacle(data["obs"][m], data["mode"][m]) for m in members[b, : n_list[b]]]
            u, x, lo, it = ref.call_ilqr(param, N, data["x0"][b], data["x_term"][b], obstacles, float(data["lamb_in"][b]))
            uvar[b], xvar[b], lamb_out[b], iters[b] = u, x, lo, it
    finally:
        mod.get_cost_derivation, mod.get_cost_final = orig_cd, orig_cf
    np.savez_compressed(os.path.join(HERE, "multi_obstacle.npz"), N=np.int32(N), x0=data["x0"], x_term=data["x_term"], lamb_in=data["lamb_in"],
                        obs=data["obs"], mode=data["mode"], members=members, n_list=n_list, uvar=uvar, xvar=xvar, lamb_out=lamb_out, iters=iters)  # fmt: skip
    ref.close()
    print("multi_obstacle done: iters mean %.2f max %d" % (iters.mean(), iters.max()))


TARGETS = {
    "multi_obstacle": make_multi_obstacle,
    "flat": make_flat,
    "knn": make_knn,
    "steps_stock": make_steps_stock,
    "steps_moving_up": make_steps_moving_up,
    "steps_moving_left": make_steps_moving_left,
    "steps_pytest": make_steps_pytest,
    "steps_k64": make_steps_k64,
    "steps_static_big": make_steps_static_big,
    "steps_add_static": make_steps_add_static,
    "steps_add_up": make_steps_add_up,
    "steps_add_left": make_steps_add_left,
    "step_synth": make_step_synth,
    "components": make_components,
}

if __name__ == "__main__":
    _probe = np.array([1, 1, 2, 2, 0, 0, 2, 2, 0, 0, 2, 1, 0, 2, 0, 1, 1], dtype=np.float64)
    if list(np.argsort(_probe)) != [8, 14, 12, 4, 5, 9, 11, 15, 0, 1, 16, 6, 10, 3, 2, 13, 7]:
        sys.exit("numpy.argsort is not running the scalar introsort (NPY_DISABLE_CPU_FEATURES not honoured?)")
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        procs = [(t, subprocess.Popen([sys.executable, os.path.abspath(__file__), t], env=env)) for t in TARGETS]
        bad = [t for t, p in procs if p.wait() != 0]
        print("failed:", bad if bad else "none")
        sys.exit(1 if bad else 0)
    TARGETS[what]()
