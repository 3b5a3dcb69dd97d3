"""Many independent closed-loop episodes on one GPU (BASELINE.json config 3).

Replaces, for E episodes at once, the lap loop of the reference's scripts: `Simulator.sim`
(iterative_ilqr/utils/base.py:708-726), `KineticBicycle.forward_one_step` (:157-213), the obstacle motion
(:50-64) and the lap post-processing of iterative_ilqr/tests/ilqr_test.py:56-60.  Plant, safe sets and
controller state never leave the device between steps.
"""
import numpy as np

from . import engine, synth


class EpisodeBatch:
    def __init__(self, params, x0, obstacles=None, initial_lap=None, device=0, lap_capacity=None, noise_seed=0):
        """x0 (E,4) first-lap start states; obstacles: None or structured array (E) of i2lqr_obstacle;
        initial_lap (T,4): the feasible first lap every episode starts from (default: the reference's
        scripted lap, KineticBicycle.get_traj)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, 4)
        self.E = x0.shape[0]
        if initial_lap is None:
            initial_lap, _ = synth.initial_lap(params.timestep)
        self.initial_lap = np.ascontiguousarray(initial_lap, dtype=np.float64)
        cap = lap_capacity or max(self.initial_lap.shape[0], params.max_sim_steps + 1)
        self.engine = engine.Engine(params, device=device, n_episodes=self.E, lap_capacity=cap)
        self.engine.add_lap_all(self.initial_lap)  # one upload, transposed into every episode's ring on the device
        if obstacles is not None:
            self.engine.set_obstacles(obstacles)
        self.engine.episodes_begin(x0, self.initial_lap[-1])
        if noise_seed:  # plant noise of utils/base.py:207-211 (off in every reference script)
            self.engine.episodes_set_noise(noise_seed)
        self.laps_run = 0

    def set_obstacles(self, obstacles):
        """Per-lap obstacle plan (e.g. obstacle only on one lap, result/ilqr_test_add_moving_obstacle.py:63-75)."""
        self.engine.set_obstacles(obstacles)

    def run_lap(self, want_traj=False):
        """One lap for every episode.  Returns steps per episode (and the logged states / inputs)."""
        out = self.engine.episodes_run_lap(want_traj=want_traj)
        self.laps_run += 1
        return out

    def status(self):
        return self.engine.episodes_status()

    def stats(self):
        return self.engine.last_stats()

    def close(self):
        self.engine.close()
