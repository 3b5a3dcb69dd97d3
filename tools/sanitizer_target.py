"""Small compute-sanitizer target (racecheck / memcheck finish in minutes):

    compute-sanitizer --tool racecheck python tools/sanitizer_target.py [flat|step|episodes|all]

  step      K=8, L=2, three recorded controller steps of the stock scenario (chain_lat_kernel; set I2LQR_GROUP /
            I2LQR_RUNGS / I2LQR_CHAIN_LAT / I2LQR_SPEC in the environment to pick the variant under test)
  flat      8192 + 300 candidates (tensor-memory variant and the 32-thread shared-memory variant)
  episodes  8 closed-loop episodes, one lap (K4 kernels, compacted episode list)
Results are compared with the recorded reference outputs, so a race that changes a value fails the run as well."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqr_iterative_tasks_b200 import _abi, engine, episodes, synth  # noqa: E402
from tests import helpers as H  # noqa: E402


def step():
    g = H.load("steps_stock.npz")
    p = H.step_params(g)
    e = engine.Engine(p, n_episodes=1, lap_capacity=128)
    added = 0
    ids = [int(s) for s in g["solve_ids"] if int(g["step/n_laps"][s]) >= 2][:3]
    for s in ids:
        while added < int(g["step/n_laps"][s]):
            e.add_lap(0, g["lap%d" % added])
            added += 1
        e.set_obstacle(0, H.step_obstacle(g, s))
        out = e.solve_step(g["step/x"][s].reshape(1, 4))
        k = list(g["solve_ids"]).index(s)
        assert np.abs(out["x_pred"][0] - g["solve/x_pred"][k]).max() <= 1e-9 * max(1.0, np.abs(g["solve/x_pred"][k]).max())
        tr = e.step_trace(1)
        assert np.array_equal(tr["cand_iters"][0], g["solve/c_iters"][k])
    e.close()
    print("step ok (%d steps)" % len(ids))


def flat():
    for B in (8192, 300):
        d = synth.flat_candidates(B, 6, seed=11)
        p = _abi.default_params(max_ilqr_iter=40)
        e = engine.Engine(p)
        a = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], _abi.obstacle_array(d["obs"], d["mode"]))
        b = e.solve_batch(d["x0"], d["x_term"], d["lamb_in"], _abi.obstacle_array(d["obs"], d["mode"]))
        assert all(np.array_equal(a[k], b[k]) for k in a)
        e.close()
    print("flat ok")


def eps():
    sc = synth.episode_scenarios(8)
    p = _abi.default_params(num_ss_points=8, num_ss_iter=2, max_sim_steps=40)
    lap0, _ = synth.initial_lap()
    eb = episodes.EpisodeBatch(p, sc["x0"], obstacles=_abi.obstacle_array(sc["obs"], sc["mode"]), initial_lap=lap0)
    lens = eb.run_lap()
    eb.close()
    print("episodes ok", lens)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    for name, fn in (("step", step), ("flat", flat), ("episodes", eps)):
        if what in (name, "all"):
            fn()
