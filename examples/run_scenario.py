"""The reference's paper scenarios (iterative_ilqr/result/ilqr_test_*.py) on the GPU path.

    python examples/run_scenario.py --scenario add_moving_up --lap-number 6 --num-ss-iters 2 --num-ss-points 8
    python examples/run_scenario.py --scenario static --episodes 256          # E copies, device-resident

--episodes 1 (default) drives the drop-in classes exactly as the reference's scripts do (same `utils.base`
API, one controller step per `calc_input`); --episodes E > 1 runs E independent closed loops with
`EpisodeBatch` (plant, obstacle motion and safe sets stay on the device).

Scenario constants (reference file:line):
  no_obstacle      result/ilqr_test_no_obstacle.py
  static           result/ilqr_test_static_obstacle.py:36-39      Obstacle(100, -5, 20, 40)
  add_static       result/ilqr_test_add_static_obstacle.py:50-59  Obstacle(35, 0, 30, 30) on lap 5 only
  add_moving_up    result/ilqr_test_add_moving_obstacle.py:18-24  Obstacle(35, -16, 34, 34, spd=1, up) on lap 5 only
  add_moving_left  result/ilqr_test_add_moving_obstacle.py:25-31  Obstacle(50, -1, 35, 35, spd=0.2, left) on lap 5 only
  stock            iterative_ilqr/tests/ilqr_test.py:31-39        Obstacle(31, -3, 8, 6)
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ilqr_iterative_tasks_b200 as pkg  # noqa: E402

SCENARIOS = {
    "no_obstacle": lambda lap: None,
    "stock": lambda lap: (31, -3, 8, 6, None, None),
    "static": lambda lap: (100, -5, 20, 40, None, None),
    "add_static": lambda lap: (35, 0, 30, 30, None, None) if lap == 5 else None,
    "add_moving_up": lambda lap: (35, -16, 34, 34, 1, 1) if lap == 5 else None,
    "add_moving_left": lambda lap: (50, -1, 35, 35, 0.2, 2) if lap == 5 else None,
}


def run_single(args):
    pkg.install_dropin()
    from utils import base  # the drop-in mirror of the reference's module

    dt, plan = 1, SCENARIOS[args.scenario]
    ego = base.KineticBicycle(False, system_param=base.KineticBicycleParam())
    ego.set_state([0, 0, 0, 0])
    ego.set_timestep(dt)
    ego.get_traj()
    ego.set_zero_noise()
    param = base.iLqrParam(num_ss_points=args.num_ss_points, num_ss_iter=args.num_ss_iters, timestep=dt, num_horizon=6)
    ctrl = base.iLqr(param, system_param=base.KineticBicycleParam())
    ctrl.add_trajectory(ego.xcl, ego.ucl)
    ctrl.set_timestep(dt)
    ego.set_ctrl_policy(ctrl)
    sim = base.Simulator()
    sim.set_robotic(ego)
    sim.set_timestep(dt)
    sim.set_traj()
    for lap in range(args.lap_number):
        o = plan(lap)
        ctrl.obstacle = None if o is None else base.Obstacle(o[0], o[1], o[2], o[3], spd=o[4], timestep=dt, moving_option=o[5])
        t0 = time.perf_counter()
        sim.sim(lap)
        ego.data["state"][-1][-1, :] = ego.xcl[-1, :]
        ctrl.add_trajectory(ego.data["state"][-1], ego.data["input"][-1])
        steps = len(ego.data["timestamp"][-1])
        print("lap %d: %d steps, %.3f s wall (%.2f ms per controller step)" % (lap + 1, steps, time.perf_counter() - t0, 1e3 * (time.perf_counter() - t0) / steps))
    print("time at iteration 0 is", len(ego.xcl) * dt, " s")
    for i, ts in enumerate(ego.data["timestamp"]):
        print("time at iteration ", i + 1, " is ", len(ts) * dt, " s")
    if args.save:
        ego.save(args.save)


def run_batch(args):
    from ilqr_iterative_tasks_b200 import _abi, episodes, synth

    plan, E = SCENARIOS[args.scenario], args.episodes
    p = _abi.default_params(num_ss_points=args.num_ss_points, num_ss_iter=args.num_ss_iters)
    lap0, _ = synth.initial_lap()
    eb = episodes.EpisodeBatch(p, np.zeros((E, 4)), initial_lap=lap0, noise_seed=args.noise_seed)
    for lap in range(args.lap_number):
        o = plan(lap)
        if o is None:
            eb.set_obstacles(None)
        else:
            row = np.array([[o[0], o[1], o[2], o[3], o[4] or 0.0]] * E)
            eb.set_obstacles(_abi.obstacle_array(row, [o[5] or 0] * E))
        t0 = time.perf_counter()
        lens = eb.run_lap()
        st = eb.stats()
        print("lap %d: steps min/mean/max %d/%.1f/%d, %d solves, %.3f s wall, %.2f M solves/s"
              % (lap + 1, lens.min(), lens.mean(), lens.max(), st["solves"], time.perf_counter() - t0, st["solves"] / (time.perf_counter() - t0) / 1e6))
    eb.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="stock", choices=sorted(SCENARIOS))
    ap.add_argument("--lap-number", type=int, default=6)
    ap.add_argument("--num-ss-points", type=int, default=8)
    ap.add_argument("--num-ss-iters", type=int, default=2)
    ap.add_argument("--episodes", type=int, default=1)
    ap.add_argument("--noise-seed", type=int, default=0, help="> 0: plant noise of utils/base.py:207-211 (batched runs)")
    ap.add_argument("--save", default=None, help="pickle the plant object like the reference's .obj files")
    a = ap.parse_args()
    run_single(a) if a.episodes == 1 else run_batch(a)
