"""Pins the CPU oracle (oracle/i2lqr_oracle.c) to outputs of the unmodified Python reference.

tests/golden/*.npz were minted by tests/golden/make_golden.py from /root/reference in the build container.
Discrete results (selected indices, iteration counts, the lambda chain, winner, horizon-shrink flag) must
be exact; states and inputs within 1e-9 relative (the oracle uses libm and a Jacobi 2x2 eigen form where
the reference uses NumPy SIMD transcendentals and LAPACK dgeev).
"""
import os

import numpy as np
import pytest

from ilqr_iterative_tasks_b200 import _abi
from oracle import oracle
from tests import helpers as H


@pytest.mark.parametrize("grp", H.FLAT_GROUPS)
def test_ilqr_matches_reference(grp):
    g = H.load("ilqr_flat.npz")
    p, inp, obs, exp = H.flat_group(g, grp)
    u, x, lamb, iters = oracle.ilqr_batch(p, inp["x0"], inp["x_term"], inp["lamb_in"], obs, nthreads=4, uvar_init=inp["uvar_init"])
    assert np.array_equal(iters, exp["iters"])
    assert np.array_equal(lamb, exp["lamb_out"])  # bit-exact: same /10, *10 sequence on doubles
    assert H.rel_err(x, exp["xvar"]).max() <= H.RTOL
    assert H.rel_err(u, exp["uvar"]).max() <= H.RTOL


def test_single_call_equals_batch():
    g = H.load("ilqr_flat.npz")
    p, inp, obs, exp = H.flat_group(g, "n6")
    from ilqr_iterative_tasks_b200 import _abi

    for b in (0, 17, 255):
        o = _abi.Obstacle(*[float(v) for v in g["n6/obs"][b]], int(g["n6/mode"][b]), 1)
        u, x, lamb, it = oracle.ilqr(p, inp["x0"][b], inp["x_term"][b], float(inp["lamb_in"][b]), o)
        assert it == exp["iters"][b] and lamb == exp["lamb_out"][b]
        assert np.abs(x - exp["xvar"][b]).max() <= 1e-9 * max(1, np.abs(exp["xvar"][b]).max())


def test_knn_matches_reference():
    g = H.load("knn.npz")
    for c in range(int(g["n_cases"])):
        idx = oracle.select_knn(g["c%d/lap" % c], g["c%d/x0" % c], int(g["c%d/K" % c]))
        assert np.array_equal(idx, g["c%d/idx" % c]), "case %d" % c


_ARGSORT = r"""
import sys, numpy as np
sys.path.insert(0, r"{root}")
from oracle import oracle
probe = np.array([1, 1, 2, 2, 0, 0, 2, 2, 0, 0, 2, 1, 0, 2, 0, 1, 1], dtype=np.float64)
if list(np.argsort(probe)) != [8, 14, 12, 4, 5, 9, 11, 15, 0, 1, 16, 6, 10, 3, 2, 13, 7]:
    print("SKIP"); sys.exit(0)
rng = np.random.default_rng(7)
for trial in range(3000):
    T = int(rng.integers(1, 600)); mode = trial % 5
    if mode == 0: v = rng.standard_normal(T)
    elif mode == 1: v = rng.integers(0, max(2, T // 8), T).astype(float)
    elif mode == 2:
        v = rng.standard_normal(T); k = int(rng.integers(1, 6)); v[-k:] = v[-k - 1] if T > k else v[-1]
    elif mode == 3: v = np.sort(rng.integers(0, 5, T)).astype(float)[::-1].copy()
    else:
        v = np.full(T, 1.5); v[rng.integers(0, T, 3)] = 0.5
    assert np.array_equal(np.argsort(v), oracle.np_argsort(v)), (trial, T, mode)
k = 2048  # median-of-3 killer: drives the introsort into its heapsort fallback
a = np.zeros(2 * k + 1)
for i in range(1, k + 1):
    if i % 2 == 1:
        a[i], a[i + 1] = i, k + i
    a[k + i] = 2 * i
for v in (a[1:], np.floor(a[1:] / 7)):
    assert np.array_equal(np.argsort(v), oracle.np_argsort(v))
print("OK")
"""


def test_argsort_restatement_is_numpys_scalar_introsort():
    """The oracle's argsort against numpy.argsort itself, tie order included.  NumPy must run its scalar sort (what the
    reference's pinned numpy==1.19.5 has): a child interpreter is started with NumPy's SIMD dispatch disabled."""
    import subprocess
    import sys

    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES="AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", _ARGSORT.format(root=root)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    if "SKIP" in r.stdout:
        pytest.skip("this NumPy build does not fall back to the scalar argsort")
    assert "OK" in r.stdout


def test_knn_edge_cases():
    lap = np.array([[0.0, 0, 0, 0], [1.0, 0, 0, 0], [1.0, 0, 0, 0], [-1.0, 0, 0, 0]])
    # K > T returns T indices; short arrays go through the insertion sort, which keeps ties in index order
    assert list(oracle.select_knn(lap, np.zeros(4), 8)) == [0, 1, 2, 3]
    assert list(oracle.select_knn(lap, np.array([1.0, 0, 0, 0]), 2)) == [1, 2]
    assert list(oracle.select_knn(lap[:1], np.ones(4), 3)) == [0]


@pytest.mark.parametrize("name", H.STEP_SCENARIOS)
def test_calc_input_matches_reference(name):
    g = H.load("steps_%s.npz" % name)
    p = H.step_params(g)
    N = p.num_horizon
    laps = [g["lap%d" % i] for i in range(int(g["n_laps"]))]
    for si, s in enumerate(g["solve_ids"]):
        out = oracle.calc_input(p, laps[: int(g["step/n_laps"][s])], g["step/x"][s], H.step_obstacle(g, s))
        assert np.array_equal(out["sel_idx"], g["solve/sel_idx"][si]), "step %d" % s
        assert np.array_equal(out["cand_iters"], g["solve/c_iters"][si]), "step %d" % s
        assert np.array_equal(np.nan_to_num(out["cand_lamb"], nan=-1.0), np.nan_to_num(g["solve/c_lamb"][si], nan=-1.0))
        assert bool(out["shrink"]) == (int(g["step/num_horizon_after"][s]) < N)
        xr, ur = g["solve/x_pred"][si], g["solve/u_pred"][si]
        assert np.abs(out["x_pred"] - xr).max() <= H.RTOL * max(1.0, np.abs(xr).max())
        assert np.abs(out["u_pred"] - ur).max() <= H.RTOL * max(1.0, np.abs(ur).max())
        assert np.abs(out["u"] - g["step/u"][s]).max() <= H.RTOL * 2


def test_calc_input_synthetic_horizons():
    g = H.load("step_synth.npz")
    from ilqr_iterative_tasks_b200 import _abi

    for N in g["horizons"]:
        k = "n%d/" % N
        p = _abi.default_params(num_horizon=int(N), num_ss_points=8, num_ss_iter=1)
        o = _abi.Obstacle(*[float(v) for v in g[k + "obs"]], int(g[k + "mode"]), 1)
        out = oracle.calc_input(p, [g[k + "lap"]], g[k + "x"], o)
        assert np.array_equal(out["cand_iters"].ravel(), g[k + "c_iters"])
        assert np.array_equal(out["cand_lamb"].ravel(), g[k + "c_lamb"])
        assert np.abs(out["x_pred"] - g[k + "x_pred"]).max() <= H.RTOL * max(1.0, np.abs(g[k + "x_pred"]).max())
        assert np.abs(out["u_pred"] - g[k + "u_pred"]).max() <= H.RTOL * 2


def test_plant_step_pinned_by_reference_fixture():
    """data/closed_loop_feasible.txt (the reference's only known-answer file, 6 decimals): the scripted
    first lap regenerated through the oracle's kinetic_bicycle must reproduce it (stored as lap0)."""
    g = H.load("steps_stock.npz")
    lap0 = g["lap0"]
    from ilqr_iterative_tasks_b200 import synth

    xcl, ucl = synth.initial_lap()
    assert xcl.shape == (121, 4) and np.abs(xcl - lap0).max() < 1e-12
    x = np.zeros(4)
    for i in range(120):
        x = oracle.kinetic_bicycle(x, ucl[i], 1.0)
        assert np.abs(x - lap0[i + 1]).max() < 1e-9
    # the three duplicated final rows
    assert np.array_equal(lap0[-1], lap0[-2]) and np.array_equal(lap0[-1], lap0[-3])


def test_relax_cost_ladder():
    from ilqr_iterative_tasks_b200 import _abi

    p = _abi.default_params()
    xt = np.zeros(4)
    assert oracle.relax_cost(p, np.array([3.0, 4.0, 0, 0]), xt, 10, 0) == 10 + 6 + 100  # r=5 <= 80
    assert oracle.relax_cost(p, np.array([3.0, 4.0, 0, 0]), xt, 10, 1) == 10 + 6 + 100  # r=5 <= 8
    assert oracle.relax_cost(p, np.array([3.0, 4.0, 0, 0]), xt, 10, 2) == 10 + 6 + 700  # 5 <= 0.8*i -> i=7
    assert oracle.relax_cost(p, np.array([300.0, 400.0, 0, 0]), xt, 10, 1) == np.inf  # 500 > 440


def _component_case(g, grp, b):
    N = int(g[grp + "/N"])
    p = _abi.default_params(num_horizon=N, matrix_Q=g[grp + "/Q"], matrix_R=g[grp + "/R"])
    obs = _abi.Obstacle(*[float(v) for v in g[grp + "/obs"][b]], int(g[grp + "/mode"][b]), 1) if g[grp + "/has_obs"][b] else None
    return p, obs


@pytest.mark.parametrize("grp", ["n6", "n6_qr", "n25"])
def test_backward_forward_pass_match_reference(grp):
    """`backward_pass` / `forward_pass` (iterative_ilqr/control/iterative_ilqr.py:88-160) called on their own, on
    mid-solve trajectories: gains, trial trajectory and trial cost against the reference's outputs."""
    g = H.load("components.npz")
    B = g[grp + "/xvar"].shape[0]
    for b in range(B):
        p, obs = _component_case(g, grp, b)
        x, u, xt = g[grp + "/xvar"][b], g[grp + "/uvar"][b], g[grp + "/x_term"][b]
        k, K = oracle.backward_pass(p, x, u, x.copy(), xt, float(g[grp + "/lamb"][b]), obs)
        assert H.rel_err(k[None], g[grp + "/k"][b][None])[0] <= 1e-10, (grp, b)
        assert H.rel_err(K[None], g[grp + "/K"][b][None])[0] <= 1e-10, (grp, b)
        xn, un, c = oracle.forward_pass(p, x, u, xt, g[grp + "/k"][b], g[grp + "/K"][b])
        assert H.rel_err(xn[None], g[grp + "/xvar_new"][b][None])[0] <= 1e-12, (grp, b)
        assert H.rel_err(un[None], g[grp + "/uvar_new"][b][None])[0] <= 1e-12, (grp, b)
        assert abs(c - g[grp + "/cost_new"][b]) <= 1e-12 * max(1.0, abs(g[grp + "/cost_new"][b])), (grp, b)


def test_obstacle_list_matches_composed_reference():
    """EXTENSION (obstacle lists): the oracle's sum of member barriers against `multi_obstacle.npz` -- the reference's own
    `ilqr()` driven by cost functions that add up the reference's single-obstacle outputs (make_golden.py multi_obstacle)."""
    g = H.load("multi_obstacle.npz")
    p = _abi.default_params(num_horizon=int(g["N"]))
    for b in range(len(g["x0"])):
        members = g["members"][b, : g["n_list"][b]]
        obs = [_abi.Obstacle(*[float(v) for v in g["obs"][m]], int(g["mode"][m]), 1) for m in members]
        keep = oracle.set_extra_obstacles(obs[1:])
        try:
            u, x, lamb, iters = oracle.ilqr(p, g["x0"][b], g["x_term"][b], float(g["lamb_in"][b]), obs[0])
        finally:
            oracle.set_extra_obstacles(None)
        del keep
        assert iters == g["iters"][b] and lamb == g["lamb_out"][b], b
        assert np.abs(x - g["xvar"][b]).max() <= 4e-12 * max(1.0, np.abs(g["xvar"][b]).max()), b
        assert np.abs(u - g["uvar"][b]).max() <= 4e-12 * max(1.0, np.abs(g["uvar"][b]).max()), b
