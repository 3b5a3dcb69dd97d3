"""ncu target: a few controller steps of E episodes through the three-kernel form (I2LQR_SPEC=0) so that the chain kernel
(GroupSolver, the code the speculative step's warps also run) can be sampled.  E, K, L from argv."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqr_iterative_tasks_b200 import _abi, engine, synth  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = _abi.default_params(num_ss_points=K, num_ss_iter=L)
lap0, _ = synth.initial_lap()
sc = synth.episode_scenarios(E)
eng = engine.Engine(p, n_episodes=E, lap_capacity=128)
eng.add_lap_all(lap0)
eng.set_obstacles(_abi.obstacle_array(sc["obs"], sc["mode"]))
x = lap0[10][None, :] + np.random.default_rng(0).normal(0, 0.2, (E, 4)) * np.array([1, 1, 0.3, 0.05])
for _ in range(3):
    out = eng.solve_step(x)
print("ok", eng.last_stats())
eng.close()
