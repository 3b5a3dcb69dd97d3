"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the oracle and the
golden vectors minted from the Python reference.  Discrete results exact; floats within 1e-9 relative."""
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from ilqr_iterative_tasks_b200 import engine

    return engine


def _check_flat(out, exp, tag=""):
    bad_it = int((out["iters"] != exp["iters"]).sum())
    bad_l = int((out["lamb_out"] != exp["lamb_out"]).sum())
    ex = H.rel_err(out["xvar"], exp["xvar"]).max()
    eu = H.rel_err(out["uvar"], exp["uvar"]).max()
    assert bad_it == 0, "%s: %d iteration-count mismatches" % (tag, bad_it)
    assert bad_l == 0, "%s: %d lamb_out mismatches" % (tag, bad_l)
    assert ex <= H.RTOL and eu <= H.RTOL, "%s: rel err x %.3e u %.3e" % (tag, ex, eu)


@pytest.mark.parametrize("grp", H.FLAT_GROUPS)
def test_flat_batch_vs_reference_golden(eng_mod, grp):
    g = H.load("ilqr_flat.npz")
    p, inp, obs, exp = H.flat_group(g, grp)
    e = eng_mod.Engine(p)
    out = e.solve_batch(inp["x0"], inp["x_term"], inp["lamb_in"], obs, uvar_init=inp["uvar_init"])
    _check_flat(out, exp, grp)
    st = e.last_stats()
    assert st["solves"] == len(exp["iters"]) and st["ilqr_iterations"] == int(exp["iters"].sum())
    assert st["launches"] == 1
    e.close()


def test_flat_batch_vs_oracle_sample(eng_mod):
    """Config 5 generator, a 4096-candidate sample against the oracle."""
    from ilqr_iterative_tasks_b200 import _abi, synth
    from oracle import oracle

    d = synth.flat_candidates(4096, 6)
    p = _abi.default_params()
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    e = eng_mod.Engine(p)
    out = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=8)
    same = H.same_discrete_path(out["iters"], out["lamb_out"], iters, lamb, "config-5 sample, 4096 candidates")
    ex = H.rel_err(out["xvar"][same], x[same]).max()
    eu = H.rel_err(out["uvar"][same], u[same]).max()
    assert ex <= H.RTOL and eu <= H.RTOL
    e.close()


@pytest.mark.parametrize("N", [10, 12, 13, 25])
def test_flat_batch_mid_horizons_vs_oracle(eng_mod, N):
    """Batches of >= 8192 candidates at horizons 7...25 run the tensor-memory variant with 256 / 512 columns per CTA
    (the reference goldens for these horizons are small batches, which take the shared-memory variants)."""
    from ilqr_iterative_tasks_b200 import _abi, synth
    from oracle import oracle

    B = 8192
    d = synth.flat_candidates(B, N, seed=synth.SEED + 300 + N)
    p = _abi.default_params(num_horizon=N)
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    e = eng_mod.Engine(p)
    out = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    e.close()
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=8)
    same = H.same_discrete_path(out["iters"], out["lamb_out"], iters, lamb, "tensor-memory variant, N=%d" % N)
    assert H.rel_err(out["xvar"][same], x[same]).max() <= H.RTOL
    assert H.rel_err(out["uvar"][same], u[same]).max() <= H.RTOL


@pytest.mark.parametrize("N,B,variant", [(26, 4096, "plain"), (50, 4096, "plain"), (50, 2048, "qr"), (50, 2048, "warm"), (200, 2048, "plain")])
def test_streaming_placement_vs_oracle(eng_mod, N, B, variant):
    """Long horizons (N >= 26, batches >= 2048) run the streaming placement: the working set in global memory, moved
    through a shared-memory ring by bulk asynchronous copies (StreamScratch).  Against the oracle on the same inputs,
    including non-zero Q / R and warm-started inputs; plus bit-identical reruns (the ring and its barriers carry no
    state from one call to the next)."""
    from ilqr_iterative_tasks_b200 import _abi, synth
    from oracle import oracle

    d = synth.flat_candidates(B, N, seed=synth.SEED + 500 + N)
    kw = {}
    if variant == "qr":
        kw = dict(matrix_Q=np.diag([0.1, 0.2, 0.05, 0.01]), matrix_R=np.array([[0.1, 0.02], [0.02, 0.3]]))
    p = _abi.default_params(num_horizon=N, **kw)
    uinit = None
    if variant == "warm":
        # mild perturbation: with O(1) random controls over 50 stages the solve is ill-conditioned enough that ulp-level
        # differences (FMA contraction) grow to 1e-3 with the discrete path still identical (tools/host_core_check.py
        # shows the same figure for the host-compiled core), which would test conditioning, not the placement
        uinit = np.random.default_rng(N).standard_normal((B, 2, N)) * np.array([0.15, 0.05])[None, :, None]
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    e = eng_mod.Engine(p)
    out = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs, uvar_init=uinit)
    again = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs, uvar_init=uinit)
    assert e.last_stats()["solves"] == B
    e.close()
    for k in out:
        assert np.array_equal(out[k], again[k]), k
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=8, uvar_init=uinit)
    same = H.same_discrete_path(out["iters"], out["lamb_out"], iters, lamb, "streaming placement, N=%d %s" % (N, variant))
    tol = 1e-8 if variant == "warm" else H.RTOL  # warm: up to 150 iterations of 50 stages amplify the last ulp
    assert H.rel_err(out["xvar"][same], x[same]).max() <= tol
    assert H.rel_err(out["uvar"][same], u[same]).max() <= tol


@pytest.mark.parametrize("N,variant", [(2, "plain"), (3, "plain"), (5, "plain"), (6, "shared_obstacle"), (6, "no_obstacle"), (6, "warm"), (4, "null_outputs")])
def test_lean_tensor_memory_placement_vs_oracle(eng_mod, N, variant):
    """Batches of >= 8192 candidates at horizons <= 6 run the LEAN tensor-memory variant (solve_kernel<double,0,6,0,128>: the
    candidate's start state parked in shared memory, its terminal target in the tensor-memory columns behind the gains, inputs
    loaded when the lane takes the candidate).  Horizons below 6 move the parked target's column offset; a shared obstacle, no
    obstacle, lambda_in = NULL and warm starts exercise the other branches of the candidate switch; declined outputs the store."""
    from ilqr_iterative_tasks_b200 import _abi, synth
    from oracle import oracle

    B = 8192 + 57
    d = synth.flat_candidates(B, N, seed=synth.SEED + 700 + N)
    p = _abi.default_params(num_horizon=N)
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    lamb_in, uinit = d["lamb_in"], None
    if variant == "shared_obstacle":
        obs, lamb_in = _abi.Obstacle(3.0, 0.5, 3.0, 2.0, 0.4, 1, 1), None
    elif variant == "no_obstacle":
        obs = None
    elif variant == "warm":
        uinit = np.random.default_rng(N).standard_normal((B, 2, N)) * np.array([0.5, 0.2])[None, :, None]
    e = eng_mod.Engine(p)
    out = e.solve_batch(d["x0"], d["x_term"], lamb_in, obs, uvar_init=uinit)
    assert e.last_stats()["solves"] == B
    if variant == "null_outputs":
        only = e.solve_batch(d["x0"], d["x_term"], lamb_in, obs, want=("lamb_out", "iters"))
        assert np.array_equal(only["iters"], out["iters"]) and np.array_equal(only["lamb_out"], out["lamb_out"])
    e.close()
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"], d["x_term"], np.ones(B) * p.lamb if lamb_in is None else lamb_in, obs, nthreads=8, uvar_init=uinit)
    same = H.same_discrete_path(out["iters"], out["lamb_out"], iters, lamb, "lean tensor-memory variant, N=%d %s" % (N, variant))
    assert H.rel_err(out["xvar"][same], x[same]).max() <= H.RTOL
    assert H.rel_err(out["uvar"][same], u[same]).max() <= H.RTOL


def test_flat_batch_size_boundaries(eng_mod):
    """Batch sizes around every launch-shape threshold (one lane, one warp, one CTA, the tensor-memory variant from
    8192 candidates, iteration-sliced passes from 131072): each equals the first B candidates of one big oracle run."""
    from ilqr_iterative_tasks_b200 import _abi, synth
    from oracle import oracle

    Bmax = 131072 + 33
    d = synth.flat_candidates(Bmax, 6, seed=synth.SEED + 77)
    p = _abi.default_params()
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=8)
    e = eng_mod.Engine(p)
    for B in (1, 31, 33, 127, 129, 8191, 8192, 8193, 131071, 131072, Bmax):
        out = e.solve_batch(d["x0"][:B], d["x_term"][:B], d["lamb_in"][:B], obs[:B])
        same = H.same_discrete_path(out["iters"], out["lamb_out"], iters[:B], lamb[:B], "batch size %d" % B)
        assert H.rel_err(out["xvar"][same], x[:B][same]).max() <= H.RTOL, B
        assert H.rel_err(out["uvar"][same], u[:B][same]).max() <= H.RTOL, B
        assert e.last_stats()["solves"] == B
    e.close()


def test_shared_and_absent_obstacle(eng_mod):
    from ilqr_iterative_tasks_b200 import _abi, synth
    from oracle import oracle

    d = synth.flat_candidates(512, 6, seed=5)
    p = _abi.default_params()
    e = eng_mod.Engine(p)
    one = _abi.Obstacle(3.0, 0.5, 3.0, 2.0, 0.4, 1, 1)
    out = e.solve_batch(d["x0"], d["x_term"], None, one)
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"], d["x_term"], np.ones(512), one, nthreads=8)
    _check_flat(out, {"iters": iters, "lamb_out": lamb, "xvar": x, "uvar": u}, "shared obstacle")
    out = e.solve_batch(d["x0"], d["x_term"], None, None)
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"], d["x_term"], np.ones(512), None, nthreads=8)
    _check_flat(out, {"iters": iters, "lamb_out": lamb, "xvar": x, "uvar": u}, "no obstacle")
    # empty batch is a no-op
    out = e.solve_batch(np.zeros((0, 4)), np.zeros((0, 4)))
    assert out["xvar"].shape == (0, 4, 7)
    e.close()


def test_host_pipeline_long_horizon(eng_mod):
    """The chunked three-stream host path with the hybrid (shared + global slab) scratch of N=16: several
    chunks in flight at once must not share working storage."""
    from ilqr_iterative_tasks_b200 import _abi, synth
    from oracle import oracle

    B, N = 3 * 32768 + 5, 16
    d = synth.flat_candidates(B, N, seed=77)
    p = _abi.default_params(num_horizon=N)
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    e = eng_mod.Engine(p)
    out = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    again = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    for k in out:
        assert np.array_equal(out[k], again[k]), k
    sel = np.random.default_rng(3).choice(B, 4096, replace=False)
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"][sel], d["x_term"][sel], d["lamb_in"][sel], obs[sel], nthreads=8)
    same = H.same_discrete_path(out["iters"][sel], out["lamb_out"][sel], iters, lamb, "host pipeline, N=16")
    assert H.rel_err(out["xvar"][sel][same], x[same]).max() <= H.RTOL
    assert e.last_stats()["solves"] == B
    e.close()


def test_knn_vs_reference_golden(eng_mod):
    from ilqr_iterative_tasks_b200 import _abi

    g = H.load("knn.npz")
    for c in range(int(g["n_cases"])):
        lap, K = g["c%d/lap" % c], int(g["c%d/K" % c])
        p = _abi.default_params(num_horizon=2, num_ss_points=K)
        e = eng_mod.Engine(p, n_episodes=1, lap_capacity=max(8, lap.shape[0]))
        e.add_lap(0, lap)
        idx = e.select_knn(0, 0, g["c%d/x0" % c], K)
        assert np.array_equal(idx, g["c%d/idx" % c]), "case %d" % c
        e.close()


def test_knn_tie_order_is_numpys(eng_mod):
    """Laps full of exactly equal distances (points on an integer grid, duplicated rows): the K-nearest kernel must
    emit NumPy's scalar-introsort order, which the oracle restates and is itself checked against NumPy
    (tests/test_oracle.py::test_argsort_restatement_is_numpys_scalar_introsort)."""
    from ilqr_iterative_tasks_b200 import _abi
    from oracle import oracle

    rng = np.random.default_rng(11)
    for case in range(40):
        T = int(rng.choice([5, 16, 17, 18, 33, 64, 121, 257, 500, 513, 700]))  # <= 512: ranked by counting, above: arg-min rounds
        K = int(rng.choice([8, 64]))
        if case % 3 == 0:
            lap = rng.integers(0, 4, (T, 4)).astype(np.float64)
        elif case % 3 == 1:
            lap = np.cumsum(rng.standard_normal((T, 4)), axis=0)
            lap[-int(rng.integers(2, 6)) :] = lap[-1]  # duplicated final rows, like the stock initial lap
        else:
            lap = np.repeat(rng.standard_normal((max(1, T // 4), 4)), 4, axis=0)[:T]
            T = lap.shape[0]
        x0 = lap[int(rng.integers(0, T))] + (0.0 if case % 2 else 0.5)
        p = _abi.default_params(num_horizon=2, num_ss_points=K)
        e = eng_mod.Engine(p, n_episodes=1, lap_capacity=max(8, T))
        e.add_lap(0, lap)
        assert np.array_equal(e.select_knn(0, 0, x0, K), oracle.select_knn(lap, x0, K)), (case, T, K)
        e.close()


@pytest.mark.parametrize("name", H.STEP_SCENARIOS)
def test_controller_step_vs_reference_golden(eng_mod, name):
    """Every recorded `calc_input` of a closed-loop reference run, replayed from its recorded inputs."""
    g = H.load("steps_%s.npz" % name)
    p = H.step_params(g)
    N = p.num_horizon
    e = eng_mod.Engine(p, n_episodes=1, lap_capacity=128)
    added = 0
    for si, s in enumerate(g["solve_ids"]):
        n_laps = int(g["step/n_laps"][s])
        while added < n_laps:
            e.add_lap(0, g["lap%d" % added])
            added += 1
        e.set_obstacle(0, H.step_obstacle(g, s))
        out = e.solve_step(g["step/x"][s].reshape(1, 4))
        tr = e.step_trace(1)
        tag = "%s step %d" % (name, s)
        assert np.array_equal(tr["sel_idx"][0], g["solve/sel_idx"][si]), tag
        assert np.array_equal(tr["cand_iters"][0], g["solve/c_iters"][si]), tag
        assert np.array_equal(np.nan_to_num(tr["cand_lamb"][0], nan=-1.0), np.nan_to_num(g["solve/c_lamb"][si], nan=-1.0)), tag
        assert bool(out["shrink"][0]) == (int(g["step/num_horizon_after"][s]) < N), tag
        xr, ur = g["solve/x_pred"][si], g["solve/u_pred"][si]
        assert np.abs(out["x_pred"][0] - xr).max() <= H.RTOL * max(1.0, np.abs(xr).max()), tag
        assert np.abs(out["u_pred"][0] - ur).max() <= H.RTOL * max(1.0, np.abs(ur).max()), tag
        assert np.abs(out["u"][0] - g["step/u"][s]).max() <= H.RTOL * 2, tag
    e.close()


def test_controller_step_synthetic_horizons(eng_mod):
    """Config 4b: one full calc_input per horizon (shared-memory and global-scratch kernels)."""
    from ilqr_iterative_tasks_b200 import _abi

    g = H.load("step_synth.npz")
    for N in g["horizons"]:
        k = "n%d/" % N
        p = _abi.default_params(num_horizon=int(N), num_ss_points=8, num_ss_iter=1)
        e = eng_mod.Engine(p, n_episodes=1, lap_capacity=g[k + "lap"].shape[0])
        e.add_lap(0, g[k + "lap"])
        e.set_obstacle(0, _abi.Obstacle(*[float(v) for v in g[k + "obs"]], int(g[k + "mode"]), 1))
        out = e.solve_step(g[k + "x"].reshape(1, 4))
        tr = e.step_trace(1)
        assert np.array_equal(tr["cand_iters"][0].ravel(), g[k + "c_iters"]), "N=%d" % N
        assert np.array_equal(tr["cand_lamb"][0].ravel(), g[k + "c_lamb"]), "N=%d" % N
        assert np.abs(out["x_pred"][0] - g[k + "x_pred"]).max() <= H.RTOL * max(1.0, np.abs(g[k + "x_pred"]).max())
        assert np.abs(out["u_pred"][0] - g[k + "u_pred"]).max() <= H.RTOL * 2
        e.close()


_VARIANT_SCRIPT = r"""
import sys, json, hashlib, numpy as np
sys.path.insert(0, {root!r})
from tests import helpers as H
from ilqr_iterative_tasks_b200 import engine
g = H.load("steps_k64.npz")
p = H.step_params(g)
e = engine.Engine(p, n_episodes=1, lap_capacity=128)
added, h, ok = 0, hashlib.sha256(), True
for si, s in list(enumerate(g["solve_ids"]))[::3]:
    n_laps = int(g["step/n_laps"][s])
    while added < n_laps:
        e.add_lap(0, g["lap%d" % added]); added += 1
    e.set_obstacle(0, H.step_obstacle(g, s))
    out = e.solve_step(g["step/x"][s].reshape(1, 4))
    tr = e.step_trace(1)
    ok = ok and np.array_equal(tr["sel_idx"][0], g["solve/sel_idx"][si]) and np.array_equal(tr["cand_iters"][0], g["solve/c_iters"][si])
    ok = ok and np.array_equal(np.nan_to_num(tr["cand_lamb"][0], nan=-1.0), np.nan_to_num(g["solve/c_lamb"][si], nan=-1.0))
    xr = g["solve/x_pred"][si]
    ok = ok and float(np.abs(out["x_pred"][0] - xr).max()) <= H.RTOL * max(1.0, float(np.abs(xr).max()))
    for k in ("u", "x_pred", "u_pred", "winner_idx"):
        h.update(np.ascontiguousarray(out[k]).tobytes())
    h.update(np.nan_to_num(tr["cand_cost"][0], nan=-1.0, posinf=-2.0).tobytes())
e.close()
print("RESULT" + json.dumps({{"ok": bool(ok), "digest": h.hexdigest()}}))
"""


def test_controller_step_kernel_variants_agree():
    """The controller-step kernel exists in several shapes chosen by load (rungs 1 / 2 / 4 per chain, 4 / 8 / 32
    lanes per chain, latency-oriented core or the generic one).  Each is forced through its tuning switch in a
    process of its own (the switches are read once), replays the K=64 golden steps, and must reproduce the
    reference's discrete trace; the latency-oriented shapes must agree with each other bit for bit, and so must
    the generic ones (the two cores differ by rounding only: same arithmetic, different FMA contraction)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    variants = [
        {"I2LQR_GROUP": "32"}, {"I2LQR_GROUP": "8"}, {"I2LQR_GROUP": "4"}, {"I2LQR_RUNGS": "2"}, {"I2LQR_RUNGS": "1"},
        {"I2LQR_CHAIN_LAT": "0"}, {"I2LQR_CHAIN_LAT": "0", "I2LQR_RUNGS": "1"}, {"I2LQR_CHAIN_LAT": "0", "I2LQR_SCRATCH": "1"},
    ]  # fmt: skip
    for env in variants:
        env["I2LQR_SPEC"] = "0"  # the three-kernel form of the step (one controller would otherwise take the one-launch form)
    # the speculative one-launch form (spec_step.cuh) with speculation sets of 1 / 8 / 24 (default) / 32 lambda values:
    # hits, waits and misses fall differently in each, the results may not
    variants += [{"I2LQR_SPEC_S": "1"}, {"I2LQR_SPEC_S": "8"}, {"I2LQR_SPEC_S": "24"}, {"I2LQR_SPEC_S": "32"}]
    # ... and with a whole warp per speculative candidate instead of four lanes (the shape small steps take by default)
    variants += [{"I2LQR_SPEC_WG": "32"}, {"I2LQR_SPEC_WG": "32", "I2LQR_SPEC_S": "8"}]
    for env in variants:
        full = dict(os.environ)
        full.update(env)
        out = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT.format(root=root)], capture_output=True, text=True, timeout=600, env=full)
        assert out.returncode == 0, (env, out.stderr[-1500:])
        res = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT")][0][6:])
        assert res["ok"], env
        digests[tuple(sorted(env.items()))] = res["digest"]
    lat = {d for k, d in digests.items() if ("I2LQR_CHAIN_LAT", "0") not in k}
    gen = {d for k, d in digests.items() if ("I2LQR_CHAIN_LAT", "0") in k}
    assert len(lat) == 1, digests
    assert len(gen) == 1, digests


def test_speculative_step_reports_hits(eng_mod):
    """The one-launch controller step serves most candidates from speculative solves once its lambda histogram has
    seen a few steps (K=64 recorded run); counters are consistent with the number of candidates walked."""
    g = H.load("steps_k64.npz")
    p = H.step_params(g)
    e = eng_mod.Engine(p, n_episodes=1, lap_capacity=128)
    added, walked = 0, 0
    for si, s in enumerate(g["solve_ids"]):
        while added < int(g["step/n_laps"][s]):
            e.add_lap(0, g["lap%d" % added])
            added += 1
        e.set_obstacle(0, H.step_obstacle(g, s))
        e.solve_step(g["step/x"][s].reshape(1, 4))
        walked += int((g["solve/c_iters"][si] >= 0).sum())
    st = e.spec_stats()
    e.close()
    print("[spec] %s" % st)
    if st["launches"] == 0:
        pytest.skip("speculative step disabled (I2LQR_SPEC=0)")
    assert st["launches"] == len(g["solve_ids"])
    n_first = 3 * sum(min(int(g["step/n_laps"][s]), p.num_ss_iter) for s in g["solve_ids"])  # rank-0 candidates: always solved in place
    assert st["hits"] + st["waited"] + st["self_solved"] + st["misses"] == walked - n_first
    assert st["hit_rate"] > 0.6, st


def test_obstacle_list_step_vs_oracle(eng_mod):
    """EXTENSION (obstacle lists): controller steps with one, two and three obstacles per episode against the
    oracle (whose member-sum is pinned to the composed reference by tests/test_oracle.py).  A one-element list must equal
    the single-obstacle call bit for bit, and an obstacle far from the path must not change a step."""
    from ilqr_iterative_tasks_b200 import _abi
    from oracle import oracle

    g = H.load("steps_stock.npz")
    p = H.step_params(g)
    laps = [g["lap0"], g["lap1"]]
    O = _abi.Obstacle
    near = [O(31.0, -3.0, 8.0, 6.0, 0.0, 0, 1), O(60.0, 2.5, 6.0, 3.0, 0.0, 0, 1), O(18.0, 4.0, 5.0, 2.5, 0.3, 1, 1), O(80.0, -2.0, 9.0, 4.0, 0.2, 2, 1)]
    e = eng_mod.Engine(p, n_episodes=1, lap_capacity=128)
    for lp in laps:
        e.add_lap(0, lp)
    states = [g["lap1"][i] + np.array([0.2, -0.1, 0.05, 0.01]) for i in (2, 9, 17, 30, 44)]
    for n in (1, 2, 3, 4):
        e.set_obstacle_list(0, near[:n])
        keep = oracle.set_extra_obstacles(near[1:n])
        try:
            for x in states:
                out = e.solve_step(x.reshape(1, 4))
                tr = e.step_trace(1)
                ref = oracle.calc_input(p, laps, x, near[0])
                assert np.array_equal(tr["cand_iters"][0], ref["cand_iters"]), n
                assert np.array_equal(np.nan_to_num(tr["cand_lamb"][0], nan=-1.0), np.nan_to_num(ref["cand_lamb"], nan=-1.0)), n
                assert int(out["winner_idx"][0]) == ref["winner_idx"] and int(out["winner_lap"][0]) == ref["winner_lap"]
                assert np.abs(out["x_pred"][0] - ref["x_pred"]).max() <= H.RTOL * max(1.0, np.abs(ref["x_pred"]).max())
                assert np.abs(out["u_pred"][0] - ref["u_pred"]).max() <= H.RTOL * 2
        finally:
            oracle.set_extra_obstacles(None)
        del keep
    # a list of one == the reference's single obstacle; a far-away second obstacle changes nothing
    x = states[1].reshape(1, 4)
    e.set_obstacle(0, near[0])
    single = e.solve_step(x)
    e.set_obstacle_list(0, near[:1])
    one = e.solve_step(x)
    e.set_obstacle_list(0, [near[0], O(5000.0, 5000.0, 3.0, 3.0, 0.0, 0, 1)])
    far = e.solve_step(x)
    for k in single:
        assert np.array_equal(single[k], one[k]), k
        assert np.array_equal(single[k], far[k]), k
    with pytest.raises(RuntimeError):
        e.set_obstacle_list(0, near + near[:1])  # five
    e.close()


def test_multi_episode_step_equals_single(eng_mod):
    """E episodes in one call == E single-episode calls (the batch axis must not couple episodes)."""
    g = H.load("steps_stock.npz")
    p = H.step_params(g)
    ids = [int(s) for s in g["solve_ids"][[3, 20, 40, 60, 61, 62]]]
    E = len(ids)
    big = eng_mod.Engine(p, n_episodes=E, lap_capacity=128)
    singles = []
    for k, s in enumerate(ids):
        one = eng_mod.Engine(p, n_episodes=1, lap_capacity=128)
        for l in range(int(g["step/n_laps"][s])):
            big.add_lap(k, g["lap%d" % l])
            one.add_lap(0, g["lap%d" % l])
        big.set_obstacle(k, H.step_obstacle(g, s))
        one.set_obstacle(0, H.step_obstacle(g, s))
        singles.append(one.solve_step(g["step/x"][s].reshape(1, 4)))
        one.close()
    out = big.solve_step(g["step/x"][ids])
    for k in range(E):
        for key in ("u", "x_pred", "u_pred", "winner_lap", "winner_rank", "winner_idx", "winner_cost", "shrink"):
            assert np.array_equal(out[key][k], singles[k][key][0]), (k, key)
    big.close()


@pytest.mark.parametrize("name,n_parts", [("k64", 2), ("k64", 3), ("stock", 2), ("moving_left", 4)])
def test_split_step_equals_single_handle(eng_mod, name, n_parts):
    """One controller step with its laps split over n_parts handles (the split a multi-GPU step
    uses, here with every part on this GPU and the all-gather replaced by a concatenation) is bit-identical to
    i2lqr_solve_step -- outputs and the per-candidate trace -- on recorded steps of the reference's runs."""
    from ilqr_iterative_tasks_b200 import dist as D

    g = H.load("steps_%s.npz" % name)
    p = H.step_params(g)
    whole = eng_mod.Engine(p, n_episodes=1, lap_capacity=128)
    parts = [eng_mod.Engine(p, n_episodes=1, lap_capacity=128) for _ in range(n_parts)]
    added = 0
    ids = list(g["solve_ids"])
    for s in ids[:: max(1, len(ids) // 12)]:
        n_laps = int(g["step/n_laps"][s])
        while added < n_laps:
            for e in [whole] + parts:
                e.add_lap(0, g["lap%d" % added])
            added += 1
        for e in [whole] + parts:
            e.set_obstacle(0, H.step_obstacle(g, s))
        x = g["step/x"][s].reshape(1, 4)
        ref = whole.solve_step(x)
        ref_tr = whole.step_trace(1)
        outs = D.split_step_local(parts, x)
        for part, out in enumerate(outs):
            for key in ref:
                assert np.array_equal(out[key], ref[key], equal_nan=True), (name, s, part, key)
            tr = parts[part].step_trace(1)
            for key in ref_tr:
                assert np.array_equal(tr[key], ref_tr[key], equal_nan=True), (name, s, part, key)
    for e in [whole] + parts:
        e.close()


def test_split_step_several_episodes(eng_mod):
    """The split form with more than one controller per handle: three episodes at different points of the same
    recorded run, two parts, against i2lqr_solve_step on one handle."""
    from ilqr_iterative_tasks_b200 import dist as D

    g = H.load("steps_add_up.npz")
    p = H.step_params(g)
    ids = [int(s) for s in g["solve_ids"] if int(g["step/n_laps"][s]) == 6][:40:13]
    assert len(ids) == 3
    E = len(ids)
    whole = eng_mod.Engine(p, n_episodes=E, lap_capacity=128)
    parts = [eng_mod.Engine(p, n_episodes=E, lap_capacity=128) for _ in range(2)]
    for k, s in enumerate(ids):
        for e in [whole] + parts:
            for l in range(6):
                e.add_lap(k, g["lap%d" % l])
            e.set_obstacle(k, H.step_obstacle(g, s))
    x = g["step/x"][ids]
    ref = whole.solve_step(x)
    ref_tr = whole.step_trace(E)
    for part, out in enumerate(D.split_step_local(parts, x)):
        for key in ref:
            assert np.array_equal(out[key], ref[key], equal_nan=True), (part, key)
        tr = parts[part].step_trace(E)
        for key in ref_tr:
            assert np.array_equal(tr[key], ref_tr[key], equal_nan=True), (part, key)
    for k, s in enumerate(ids):  # and the whole-handle result is the recorded reference step
        si = list(g["solve_ids"]).index(s)
        assert np.abs(ref["x_pred"][k] - g["solve/x_pred"][si]).max() <= H.RTOL * max(1.0, np.abs(g["solve/x_pred"][si]).max())
    for e in [whole] + parts:
        e.close()


def test_handle_recreation_is_race_free(eng_mod):
    """Regression: handle initialisation used legacy-stream cudaMemset calls that were not ordered against the
    handle's non-blocking stream, so the first add_lap could be wiped (episode 0 lost its lap ~3 % of the
    time when handles were created back to back).  Identical episodes must give identical answers."""
    g = H.load("steps_stock.npz")
    p = H.step_params(g)
    x = np.tile(g["step/x"][int(g["solve_ids"][2])], (3, 1))
    ref = None
    for _ in range(60):
        e = eng_mod.Engine(p, n_episodes=3, lap_capacity=128)
        for k in range(3):
            e.add_lap(k, g["lap0"])
            e.set_obstacle(k, H.step_obstacle(g, int(g["solve_ids"][2])))
        out = e.solve_step(x)
        e.close()
        assert np.array_equal(out["u"][0], out["u"][1]) and np.array_equal(out["u"][0], out["u"][2])
        assert out["winner_idx"][0] == out["winner_idx"][2]
        if ref is None:
            ref = out["u"].copy()
        assert np.array_equal(out["u"], ref)


def test_full_size_properties(eng_mod):
    """BASELINE config 5 at full size (2^20 candidates): size-independent properties.
    determinism (bit-identical reruns), permutation equivariance (results do not depend on which lane or in
    which order a candidate is scheduled), counters consistent, every solve terminated by a reference rule."""
    from ilqr_iterative_tasks_b200 import _abi, synth

    B = 1 << 20
    d = synth.flat_candidates(B, 6)
    p = _abi.default_params()
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    e = eng_mod.Engine(p)
    a = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    st = e.last_stats()
    assert st["solves"] == B and st["ilqr_iterations"] == int(a["iters"].sum())
    assert a["iters"].min() >= 1 and a["iters"].max() <= p.max_ilqr_iter
    assert np.isfinite(a["xvar"]).all() and np.isfinite(a["uvar"]).all()
    assert (np.abs(a["uvar"][:, 0]) <= p.a_max).all() and (np.abs(a["uvar"][:, 1]) <= p.delta_lim).all()
    assert np.array_equal(a["xvar"][:, :, 0], d["x0"])
    b = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    perm = np.random.default_rng(1).permutation(B)
    c = e.solve_batch(d["x0"][perm], d["x_term"][perm], d["lamb_in"][perm], obs[perm])
    for k in a:
        assert np.array_equal(a[k][perm], c[k]), k
    # the big batch runs as iteration-sliced passes (park / resume): check a sample, weighted towards the
    # candidates that were parked (>= 32 iterations), against the oracle
    from oracle import oracle

    long_ids = np.flatnonzero(a["iters"] >= 32)[:3000]
    sel = np.unique(np.concatenate([long_ids, np.random.default_rng(2).choice(B, 3000, replace=False)]))
    u, x, lamb, iters = oracle.ilqr_batch(p, d["x0"][sel], d["x_term"][sel], d["lamb_in"][sel], obs[sel], nthreads=8)
    same = H.same_discrete_path(a["iters"][sel], a["lamb_out"][sel], iters, lamb, "2^20-candidate batch, parked-heavy sample")
    assert H.rel_err(a["xvar"][sel][same], x[same]).max() <= H.RTOL
    assert H.rel_err(a["uvar"][sel][same], u[same]).max() <= H.RTOL
    assert (a["iters"][sel] >= 32).sum() >= 2000 and a["iters"].max() == p.max_ilqr_iter
    # lamb_out is lamb_in times a power of ten reached through the reference's /10, *10 sequence
    ratio = np.log10(a["lamb_out"])
    assert np.abs(ratio - np.round(ratio)).max() < 1e-9
    e.close()


@pytest.mark.parametrize("N", [6, 10, 25, 50])
def test_fp32_build_tolerance(eng_mod, N):
    """BASELINE config 4 gate for the optional FP32 build (params.precision = 1): on single solves from
    identical inputs, FP32 agrees with FP64 within 1e-4 relative wherever the discrete path (iteration count
    and lambda exponent) is the same; the fraction whose path diverged is reported, not hidden."""
    from ilqr_iterative_tasks_b200 import _abi, synth

    B = 4096
    d = synth.flat_candidates(B, N, seed=5 + N)
    obs = _abi.obstacle_array(d["obs"], d["mode"])
    e64 = eng_mod.Engine(_abi.default_params(num_horizon=N))
    e32 = eng_mod.Engine(_abi.default_params(num_horizon=N, precision=1))
    a = e64.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    b = e32.solve_batch(d["x0"], d["x_term"], d["lamb_in"], obs)
    same = (a["iters"] == b["iters"]) & (np.round(np.log10(a["lamb_out"])) == np.round(np.log10(b["lamb_out"])))
    err = np.maximum(H.rel_err(b["xvar"], a["xvar"]), H.rel_err(b["uvar"], a["uvar"]))
    within = (err[same] <= 1e-4).mean()
    print("FP32 N=%d: same discrete path %.1f%%, of those within 1e-4: %.2f%% (median %.1e)" % (N, 100 * same.mean(), 100 * within, np.median(err[same])))
    assert same.mean() >= 0.85
    assert within >= 0.99
    assert np.median(err[same]) <= 1e-5
    assert np.isfinite(b["xvar"]).all()
    e64.close()
    e32.close()


def test_fp32_controller_step_runs(eng_mod):
    """The FP32 build also drives the controller step (lambda chains); decisions may differ from FP64 on
    closed loops, so only sanity is checked here: same winner on an easy step, finite prediction."""
    g = H.load("steps_stock.npz")
    from ilqr_iterative_tasks_b200 import _abi

    p = _abi.default_params(num_ss_points=int(g["K"]), num_ss_iter=int(g["L"]), precision=1)
    e = eng_mod.Engine(p, n_episodes=1, lap_capacity=128)
    e.add_lap(0, g["lap0"])
    s = int(g["solve_ids"][5])
    e.set_obstacle(0, H.step_obstacle(g, s))
    out = e.solve_step(g["step/x"][s].reshape(1, 4))
    assert np.isfinite(out["x_pred"]).all() and out["winner_idx"][0] >= 0
    assert np.abs(out["x_pred"][0] - g["solve/x_pred"][5]).max() <= 1e-2 * max(1.0, np.abs(g["solve/x_pred"][5]).max())
    e.close()


def test_error_paths(eng_mod):
    from ilqr_iterative_tasks_b200 import _abi

    p = _abi.default_params()
    e = eng_mod.Engine(p, n_episodes=1, lap_capacity=32)
    with pytest.raises(RuntimeError):
        e.add_lap(0, np.zeros((6, 4)))  # not longer than the horizon (utils/base.py:348)
    with pytest.raises(RuntimeError):
        e.add_lap(0, np.zeros((64, 4)))  # over capacity
    with pytest.raises(RuntimeError):
        e.solve_step(np.zeros((1, 4)))  # no lap stored
    with pytest.raises(RuntimeError):
        e.set_obstacle(0, _abi.Obstacle(0, 0, 1, 1, 0.5, 0, 1))  # spd != 0 without moving_option
    bad = _abi.default_params(matrix_Q=np.array([[0, 1, 0, 0]] + [[0] * 4] * 3, dtype=float))
    with pytest.raises(RuntimeError):
        eng_mod.Engine(bad)
    e.close()
