"""CPU-side checks: the C-ABI library loads and exports every symbol include/i2lqr.h declares (no compute
calls without a GPU), struct layouts agree with the header, host-side logic (sharding, winner rule,
parameter flattening, obstacle validation) and the world_size-2 gloo path."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from ilqr_iterative_tasks_b200 import _abi, dist, engine, synth
from tests import helpers as H
from tests import host_exchange as X

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "i2lqr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(i2lqr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = C.CDLL(engine.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libi2lqr.so does not export %s" % name
    assert sorted(engine.EXPORTS) == declared
    assert lib.i2lqr_abi_version() == 1


def test_struct_layout_matches_header_defaults():
    lib = engine.lib()
    p = _abi.Params()
    lib.i2lqr_default_params(C.byref(p))
    q = _abi.default_params()
    assert bytes(p) == bytes(q)  # same size, same field order, same defaults
    assert C.sizeof(_abi.Obstacle) == 48


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        engine.Engine(_abi.default_params())


def test_params_from_reference_objects():
    class P:  # duck-typed iLqrParam
        num_horizon, num_ss_points, num_ss_iter = 10, 16, 3
        matrix_Q = np.diag([0.1, 0.1, 0.05, 0.01])
        matrix_R = np.diag([0.05, 0.05])
        matrix_Qterminal = 2 * np.diag([1.0, 1.0, 20.0, 0.02])
        lamb, lamb_factor, max_lamb, eps = 1, 10, 1000, 1e-2
        timestep = 1

    class S:
        a_max, delta_max = 2.0, np.pi / 2

    p = _abi.params_from_reference(P(), S())
    assert (p.num_horizon, p.num_ss_points, p.num_ss_iter) == (10, 16, 3)
    assert p.delta_lim == 1.57 and p.matrix_Q[5] == 0.1 and p.matrix_R[3] == 0.05 and p.timestep == 1.0


def test_obstacle_validation_mirrors_reference_failures():
    class O:
        x, y, width, height = 1.0, 2.0, 3.0, 4.0
        spd, moving_option = None, None

    o = O()
    s = _abi.obstacle_struct(o)
    assert s.present == 1 and s.moving_option == 0 and s.spd == 0.0
    o.spd, o.moving_option = 0.5, None
    with pytest.raises(UnboundLocalError):
        _abi.obstacle_struct(o)
    o.spd, o.moving_option = None, 1
    with pytest.raises(TypeError):
        _abi.obstacle_struct(o)
    assert _abi.obstacle_struct(None).present == 0


def test_shard_range_partitions():
    for n in (0, 1, 7, 4096, 1 << 20):
        for w in (1, 2, 3, 8):
            spans = [dist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_lexicographic_winner_rule():
    inf = np.inf
    # lexicographic lap choice differs from the global minimum
    costs = np.array([[106.0, 205.0, 104.0, inf], [106.0, 206.0, 100.0, 100.0]])
    assert X.lexicographic_winner(costs, [4, 4]) == (0, 2)
    # a prefix is smaller; first arg-min within the lap
    costs = np.array([[5.0, 5.0, 1.0, 1.0], [5.0, 5.0, 0.0, 0.0]])
    assert X.lexicographic_winner(costs, [4, 2]) == (1, 0)
    assert X.lexicographic_winner(np.array([[3.0, 2.0, 2.0]]), [3]) == (0, 1)
    # ties between laps: the first lap wins
    assert X.lexicographic_winner(np.array([[1.0, 2.0], [1.0, 2.0]]), [2, 2]) == (0, 0)


def test_winner_rule_agrees_with_oracle_on_golden_steps():
    from oracle import oracle

    g = H.load("steps_k64.npz")
    p = H.step_params(g)
    laps = [g["lap%d" % i] for i in range(int(g["n_laps"]))]
    for s in g["solve_ids"][-6:]:
        out = oracle.calc_input(p, laps[: int(g["step/n_laps"][s])], g["step/x"][s], H.step_obstacle(g, s))
        n_act = min(int(g["step/n_laps"][s]), p.num_ss_iter)
        counts = [(out["sel_idx"][2, l] >= 0).sum() for l in range(n_act)]
        lap, rank = X.lexicographic_winner(out["cand_cost"][2, :n_act], counts)
        first = max(0, int(g["step/n_laps"][s]) - p.num_ss_iter)
        assert (lap + first, rank) == (out["winner_lap"], out["winner_rank"])


_GLOO = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as td
from ilqr_iterative_tasks_b200 import dist, _abi
from oracle import oracle
from tests import helpers as H
from tests import host_exchange as X
td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
rank = td.get_rank()
g = H.load("steps_k64.npz")
p = H.step_params(g)
s = int(g["solve_ids"][-3])
laps = [g["lap%d" % i] for i in range(int(g["step/n_laps"][s]))]
full = oracle.calc_input(p, laps, g["step/x"][s], H.step_obstacle(g, s))
n_act = min(len(laps), p.num_ss_iter)
owners = X.lap_owner(n_act, 2)
# each rank "solves" only the laps it owns (the oracle stands in for the device here), then the exchange
local = np.full((n_act, p.num_ss_points), np.nan)
counts = np.zeros(n_act)
for l in range(n_act):
    if owners[l] == rank:
        local[l] = full["cand_cost"][2, l]
        counts[l] = (full["sel_idx"][2, l] >= 0).sum()
costs, cnt = X.gather_lap_costs(local, counts, owners, rank)
lap, r = X.lexicographic_winner(costs, cnt)
first = max(0, len(laps) - p.num_ss_iter)
assert (lap + first, r) == (full["winner_lap"], full["winner_rank"]), (lap, r)
lo, hi = dist.shard_range(1000, rank, 2)
t = torch.tensor([hi - lo]); td.all_reduce(t)
assert int(t) == 1000
td.destroy_process_group()
print("OK", rank)
"""


def test_world_size_2_gloo_exchange():
    """The only collective on the path: all-gather of per-lap cost rows when one step's laps span ranks."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, "-c", _GLOO.format(root=ROOT)], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for pr in procs:
        out, err = pr.communicate(timeout=300)
        assert pr.returncode == 0, err[-2000:]
        assert "OK" in out


def test_synth_generators_are_deterministic():
    a, b = synth.flat_candidates(100, 6), synth.flat_candidates(100, 6)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    g = H.load("ilqr_flat.npz")
    assert np.array_equal(synth.flat_candidates(256, 6)["x0"], g["n6/x0"])  # goldens were minted from this generator
    ep = synth.episode_scenarios(16)
    assert ep["obs"].shape == (16, 5) and set(np.unique(ep["mode"])) <= {0, 1, 2}
    assert (ep["obs"][ep["mode"] == 0, 4] == 0).all()
