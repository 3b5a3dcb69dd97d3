"""Synthetic workload generators for BASELINE.json configs 3, 4 and 5.

Pure NumPy, deterministic given the seed; shared by bench.py, the parity tests and
tests/golden/make_golden.py so that every arm sees byte-identical inputs.

Obstacle encoding used everywhere in this package (one row per candidate / episode):
    obs[:, 0:5] = x, y, width, height, spd        (float64)
    mode        = 0 static, 1 moving up, 2 moving left   (int32)
which mirrors `Obstacle(x, y, width, height, spd, timestep, moving_option)` of the
reference (iterative_ilqr/utils/base.py:23-35).
"""
import numpy as np

SEED = 20260101


def flat_candidates(B, num_horizon=6, dt=1.0, seed=SEED):
    """Config 4a / 5: B independent `ilqr()` problems.

    x0 ~ [0,0,2,0.1] + N(0, diag(1,1,.5,.1)^2); x_term = constant-velocity endpoint after
    `num_horizon` steps + N(0, diag(2,2,.5,.1)^2); obstacle near the path (centre between x0 and
    x_term + N(0, 2^2 I), axes U(2,6), mode in {static, up, left}, spd U(0,1) when moving);
    lamb_in = 1.
    """
    rng = np.random.default_rng(seed)
    x0 = np.array([0.0, 0.0, 2.0, 0.1]) + rng.standard_normal((B, 4)) * np.array([1.0, 1.0, 0.5, 0.1])
    reach = x0[:, 2] * (num_horizon * dt)
    end = x0.copy()
    end[:, 0] += np.cos(x0[:, 3]) * reach
    end[:, 1] += np.sin(x0[:, 3]) * reach
    x_term = end + rng.standard_normal((B, 4)) * np.array([2.0, 2.0, 0.5, 0.1])
    frac = rng.uniform(0.2, 1.0, size=(B, 1))
    centre = x0[:, :2] + frac * (x_term[:, :2] - x0[:, :2]) + rng.standard_normal((B, 2)) * 2.0
    axes = rng.uniform(2.0, 6.0, size=(B, 2))
    mode = rng.integers(0, 3, size=B).astype(np.int32)
    spd = rng.uniform(0.0, 1.0, size=B) * (mode != 0)
    obs = np.concatenate([centre, axes, spd[:, None]], axis=1)
    lamb_in = np.ones(B)
    return {
        "x0": np.ascontiguousarray(x0),
        "x_term": np.ascontiguousarray(x_term),
        "obs": np.ascontiguousarray(obs),
        "mode": mode,
        "lamb_in": lamb_in,
    }


def synthetic_lap(num_horizon, extra=64, seed=SEED):
    """Config 4b: one-lap safe set of T = num_horizon + extra points, tie-free.

    Constant-speed lap (2t cos0.1, 2t sin0.1, 2, 0.1) + N(0, 0.05^2) jitter; the controller state is
    lap[0] + N(0, diag(.5,.5,.2,.05)^2).  Returns (lap (T,4), u (T-1,2), x (4,)).
    """
    rng = np.random.default_rng(seed + 7 * num_horizon)
    T = num_horizon + extra
    t = np.arange(T, dtype=np.float64)
    lap = np.stack([2 * t * np.cos(0.1), 2 * t * np.sin(0.1), np.full(T, 2.0), np.full(T, 0.1)], axis=1)
    lap = lap + rng.standard_normal((T, 4)) * 0.05
    u = np.zeros((T - 1, 2))
    x = lap[0] + rng.standard_normal(4) * np.array([0.5, 0.5, 0.2, 0.05])
    return lap, u, x


def episode_scenarios(E, seed=SEED):
    """Config 3: per-episode obstacle on the 200 m track and perturbed initial state.

    x_o~U(20,180), y_o~U(-8,8), a,b~U(4,12), moving_option in {None,1,2}, spd~U(0,1);
    x0 = (U(-1,1), U(-1,1), U(0,1), U(-0.2,0.2)).
    """
    rng = np.random.default_rng(seed)
    obs = np.stack(
        [
            rng.uniform(20, 180, E),
            rng.uniform(-8, 8, E),
            rng.uniform(4, 12, E),
            rng.uniform(4, 12, E),
            rng.uniform(0, 1, E),
        ],
        axis=1,
    )
    mode = rng.integers(0, 3, size=E).astype(np.int32)
    obs[:, 4] *= mode != 0
    x0 = np.stack(
        [rng.uniform(-1, 1, E), rng.uniform(-1, 1, E), rng.uniform(0, 1, E), rng.uniform(-0.2, 0.2, E)], axis=1
    )
    return {"obs": np.ascontiguousarray(obs), "mode": mode, "x0": np.ascontiguousarray(x0)}


def initial_lap(dt=1.0):
    """The scripted first lap every reference script starts from.

    Restates `KineticBicycle.get_traj` (iterative_ilqr/utils/base.py:103-133): 120/dt steps of the
    kinematic bicycle under a scripted accel / steer sequence.  Returns (xcl (T+1,4), ucl (T,2)).
    """
    angle = np.pi / 6
    total = int(120 / dt)
    x = np.zeros(4)
    xs = [x.copy()]
    us = []
    for i in range(total):
        if i <= 1 / dt:
            a = 1.0
        elif (total - 4 / dt) <= i <= (total - 3 / dt):
            a = -1.0
        else:
            a = 0.0
        if 0 < i <= 1 / dt:
            d = angle
        elif (total / 2 - 2 / dt) <= i <= (total / 2 - 1 / dt):
            d = -angle
        else:
            d = 0.0
        s = x[2] * dt + (a * dt ** 2) / 2
        x = np.array([x[0] + np.cos(x[3]) * s, x[1] + np.sin(x[3]) * s, x[2] + a * dt, x[3] + d * dt])
        xs.append(x.copy())
        us.append([a, d])
    return np.array(xs), np.array(us)
