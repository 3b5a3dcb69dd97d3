"""Drop-in for the i2LQR part of the reference's `utils.base` (iterative_ilqr/utils/base.py).

Put this directory (`ilqr_iterative_tasks_b200/dropin`) on sys.path where the reference puts
`iterative_ilqr/` and the reference's own scripts (iterative_ilqr/tests/ilqr_test.py, result/ilqr_test_*.py,
tests/ilqr_test.py) run unchanged: same class names, constructor arguments, methods and read-back attributes.

What runs where
  * `iLqr.calc_input` / `iLqr.select_close_ss` -- the per-timestep control solve -- run on the GPU through
    libi2lqr.so (K-nearest selection, one iLQR per candidate with the lambda chain, relax-cost ladder,
    lexicographic winner).  There is no CPU fallback; failures raise RuntimeError, the exception
    `KineticBicycle.calc_ctrl_input` already handles (reference :146-155).
  * The plant (`KineticBicycle`), obstacle bookkeeping and `Simulator` are the same thin host loop as in the
    reference: they advance one simulated vehicle between controller calls and keep `ego.data` /
    `ego.diagnostics` in the reference's layout.  (Many-episode runs use the device-resident
    `ilqr_iterative_tasks_b200.episodes.EpisodeBatch` instead.)
  * LMPC / NLMPC (reference :482-690, control/nonlinear_lmpc.py) are out of scope and raise
    NotImplementedError.

Observable differences: the inner-loop prints ("k neighbors", "Convergence achieved") are dropped; the plotting
methods are no-ops; the LMPC-only attributes of `iLqr` (`ss_point_selected_id`, `x_sol`, `u_sol`, `cost_improve`,
`old_cost`, `old_iter`) are not kept; `ilqr.obstacle` may also be a list of up to four `Obstacle`s (extension: their barrier
terms add); `max_outloop_iter < 3` raises ValueError (the reference would then do fewer
refinements and never shrink the horizon).  K-nearest
ties (the stock initial lap ends in identical rows) come out in the order numpy.argsort's scalar introsort leaves
them -- what the reference's pinned numpy==1.19.5 does; a NumPy >= 1.25 on an AVX2 / AVX-512 host sorts float64
with x86-simd-sort and orders ties differently (run the reference with NPY_DISABLE_CPU_FEATURES="AVX512F
AVX512_SKX AVX2" to compare like with like).
"""
import datetime
import os
import pickle as pkl
import sys
from copy import deepcopy

import numpy as np

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.append(_PKG_PARENT)

from ilqr_iterative_tasks_b200 import _abi, engine  # noqa: E402
from systems.kinetic_bicycle import kinetic_bicycle  # noqa: E402
from utils.constants_kinetic_bicycle import *  # noqa: E402,F401,F403
from utils.constants_kinetic_bicycle import U_DIM, U_ID, X_DIM, X_ID  # noqa: E402


class KineticBicycleParam:
    """Input / speed limits (reference :15-20)."""

    def __init__(self, delta_max=np.pi / 2, a_max=2.0, v_max=10, v_min=0):
        self.delta_max = delta_max
        self.a_max = a_max
        self.v_max = v_max
        self.v_min = v_min


class Obstacle:
    """Elliptic obstacle, static or moving up (1) / left (2) (reference :23-64)."""

    def __init__(self, x, y, width, height, spd=None, timestep=None, moving_option=None):
        self.x0 = self.x = x
        self.y0 = self.y = y
        self.width = width
        self.height = height
        self.spd = spd
        self.timestep = timestep
        self.moving_option = moving_option
        self.data = {"state": []}
        self.states = np.array([self.x0, self.y0])

    def update_obstacle(self):
        if self.spd is not None:
            if self.moving_option == 1:
                self.y += self.spd * self.timestep
            if self.moving_option == 2:
                self.x -= self.spd * self.timestep
        self.states = np.vstack((self.states, [self.x, self.y]))

    def reset_obstacle(self):
        self.x, self.y = self.x0, self.y0
        self.data["state"].append(self.states)
        self.states = np.array([self.x0, self.y0])

    def plot_obstacle(self):
        """Plotting is outside the accelerated path; kept as a no-op so that scripts calling it run."""


class KineticBicycle:
    """Simulated vehicle + per-lap logs (reference :67-213)."""

    def __init__(self, direct_ctrl_policy=False, system_param=None):
        self.system_param = system_param
        self.direct_ctrl_policy = direct_ctrl_policy
        self.time = 0.0
        self.timestep = None
        self.x = None
        self.u = None
        self.delta_timer = None
        self.feasible = None
        self.zero_noise_flag = False
        self.states = self.inputs = self.timestamps = None
        self.solver_times = self.feasibility = None
        self.data = {"state": [], "input": [], "timestamp": []}
        self.diagnostics = {"solver_time": [], "feasibility": []}
        self.iters = 0
        self.ctrl_policy = None
        self.xcl = self.ucl = None

    def set_zero_noise(self):
        self.zero_noise_flag = True

    def set_timestep(self, dt):
        self.timestep = dt

    def set_state(self, x):
        self.x = x
        self.states = x
        self.timestamps = None
        self.inputs = None
        self.solver_times = None
        self.feasible = None

    def get_traj(self):
        """Scripted first lap (reference :103-138); also written to data/closed_loop_feasible.txt when a
        data/ directory exists, as the reference does."""
        if self.direct_ctrl_policy:
            self.xcl = np.loadtxt("data/closed_loop_multi_laps.txt")
            self.ucl = np.loadtxt("data/input_multi_laps.txt")
            return
        dt = self.timestep
        steps = int(120 / dt)
        turn = np.pi / 6
        xcl = np.zeros((steps + 1, X_DIM))
        ucl = np.zeros((steps, U_DIM))
        for i in range(steps):
            accel = 1 if i <= 1 / dt else (-1 if steps - 4 / dt <= i <= steps - 3 / dt else 0)
            steer = turn if 0 < i <= 1 / dt else (-turn if steps / 2 - 2 / dt <= i <= steps / 2 - 1 / dt else 0)
            ucl[i] = (accel, steer)
            xcl[i + 1] = kinetic_bicycle(xcl[i], ucl[i], dt)
        if os.path.isdir("data"):
            np.savetxt("data/closed_loop_feasible.txt", xcl, fmt="%f")
        if steps > 1:
            # the reference logs `ucl = u` on step 0 -- an alias of the array it mutates on step 1 -- so its first
            # logged input row shows step 1's input (reference :109-131); the states are unaffected
            ucl[0] = ucl[1]
        self.xcl, self.ucl = xcl, ucl

    def set_ctrl_policy(self, ctrl_policy):
        self.ctrl_policy = ctrl_policy

    def calc_ctrl_input(self):
        self.ctrl_policy.set_state(self.x)
        tic = datetime.datetime.now()
        try:
            self.ctrl_policy.calc_input()
            self.u = self.ctrl_policy.get_input()
            self.delta_timer = (datetime.datetime.now() - tic).total_seconds()
            self.feasible = 1
        except RuntimeError as err:
            self.feasible = 0
            print("solver fail to find the solution", err)

    def forward_one_step(self):
        self.calc_ctrl_input()
        self.forward_dynamics()
        self.ctrl_policy.set_state(self.x)
        self.update_memory()

    def forward_dynamics(self):
        x_next = kinetic_bicycle(self.x, self.u, self.timestep)
        if not self.zero_noise_flag:  # reference :207-211
            noise_v = np.maximum(-0.05, np.minimum(np.random.randn() * 0.01, 0.05))
            noise_theta = np.maximum(-0.05, np.minimum(np.random.randn() * 0.005, 0.05))
            x_next[X_ID["v"]] += 0.5 * noise_v
            x_next[X_ID["theta"]] += 0.5 * noise_theta
        self.x = x_next
        self.time += self.timestep

    def update_memory(self):
        def push(stack, item):
            return item if stack is None else np.vstack((stack, item))

        self.states = np.vstack((self.states, self.x))
        self.inputs = push(self.inputs, self.u)
        self.timestamps = push(self.timestamps, self.time)
        self.solver_times = push(self.solver_times, self.delta_timer)
        self.feasibility = push(self.feasibility, self.feasible)

    def update_memory_post_iter(self):
        self.data["state"].append(self.states)
        self.data["input"].append(self.inputs)
        self.data["timestamp"].append(self.timestamps)
        self.diagnostics["solver_time"].append(self.solver_times)
        self.diagnostics["feasibility"].append(self.feasibility)
        self.set_state(np.zeros((X_DIM,)))

    def save(self, path):
        """Pickle in the reference's `.obj` convention (result/ilqr_test_no_obstacle.py:61-62); the controller
        travels along (its GPU handle is dropped and rebuilt on first use, see iLqr.__getstate__)."""
        with open(path, "wb") as handle:
            pkl.dump(self, handle, protocol=pkl.HIGHEST_PROTOCOL)


class ControlBase:
    """Controller interface the plant talks to (reference :216-234)."""

    def __init__(self):
        self.time = 0.0
        self.timestep = None
        self.x = None
        self.u = None
        self.iters = 0

    def set_timestep(self, timestep):
        self.timestep = timestep

    def set_state(self, x):
        self.x = x

    def calc_input(self):
        pass

    def get_input(self):
        return self.u


class iLqrParam:
    """All tunables of the i2LQR controller, reference defaults (reference :242-302)."""

    def __init__(
        self,
        matrix_Q=0 * np.diag([0.0, 0.0, 0.0, 0.0]),
        matrix_R=0 * np.diag([0.05, 0.05]),
        matrix_Qterminal=2 * np.diag([1.0, 1.0, 20.0, 0.02]),
        num_ss_points=8,
        num_ss_iter=1,
        num_horizon=6,
        tuning_state_q1=1.0,
        tuning_state_q2=1.0,
        tuning_ctrl_q1=1.0,
        tuning_ctrl_q2=1.0,
        tuning_obs_q1=2.74,
        tuning_obs_q2=2.74,
        safety_margin=0.0,
        max_ilqr_iter=150,
        eps=1e-2,
        lamb=1,
        lamb_factor=10,
        max_lamb=1000,
        reach_error=1.0,
        max_relax_iter=55,
        max_outloop_iter=50,
        timestep=None,
        lap_number=None,
        time_ilqr=None,
        ss_option=None,
        all_ss_point=False,
        all_ss_iter=False,
    ):
        for name, value in list(locals().items()):
            if name != "self":
                setattr(self, name, value)


class iLqr(ControlBase):
    """i2LQR controller with the reference's API (reference :305-479), solved on the GPU."""

    LAP_CAPACITY = 256  # grows automatically when a longer lap is added

    def __init__(self, ilqr_param, obstacle=None, system_param=None):
        ControlBase.__init__(self)
        self.ilqr_param = ilqr_param
        self.system_param = system_param
        self.ss, self.u_ss, self.Qfun = [], [], []
        self.x_terminal_guess = None
        self.x_guess = None
        self.iter = 0
        self.iter_cost = []
        self.itCost = []
        self.cost = None
        self.u_old = self.x_pred = self.u_pred = None
        self.num_horizon = ilqr_param.num_horizon
        self.matrix_Qterminal = ilqr_param.matrix_Qterminal
        self.matrix_Q = ilqr_param.matrix_Q
        self.matrix_R = ilqr_param.matrix_R
        self.obstacle = obstacle
        self.winner = None  # (lap, rank, index, cost) of the last solve, for diagnostics
        self._engine = None
        self._engine_key = None
        self._capacity = self.LAP_CAPACITY

    # ---- GPU handle ------------------------------------------------------------------------------
    def _params(self):
        system = self.system_param if self.system_param is not None else KineticBicycleParam()
        dt = self.timestep if self.timestep is not None else self.ilqr_param.timestep
        return _abi.params_from_reference(self.ilqr_param, system, dt)

    def _sync_engine(self):
        params = self._params()
        key = (bytes(params), self._capacity)
        if self._engine is None or (self._engine_key[1] != self._capacity) or (
            self._engine.params.num_ss_iter != params.num_ss_iter
        ):
            if self._engine is not None:
                self._engine.close()
            self._engine = engine.Engine(params, n_episodes=1, lap_capacity=self._capacity)
            first = max(0, len(self.ss) - params.num_ss_iter)
            # keep absolute lap numbering: pad the ring position with the laps that are still active
            for lap in self.ss[first:]:
                self._engine.add_lap(0, np.ascontiguousarray(lap.T))
            self._lap_offset = first
        elif key != self._engine_key:
            self._engine.set_params(params)
        self._engine_key = key
        return self._engine

    # ---- pickling (result/ilqr_test_*.py end with `pkl.dump(ego, ...)`, and ego holds its controller) -----------
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engine"] = None  # the GPU handle stays behind; it is rebuilt from self.ss on the next call
        state["_engine_key"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    # ---- reference API ---------------------------------------------------------------------------
    def select_close_ss(self, iter, x0):
        eng = self._sync_engine()
        return eng.select_knn(0, iter - self._lap_offset, np.asarray(x0, dtype=np.float64), self.ilqr_param.num_ss_points)

    def add_trajectory(self, x, u):
        x = np.asarray(x, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        self.ss.append(deepcopy(x.T))
        self.u_ss.append(deepcopy(u.T))
        self.Qfun.append(np.arange(x.shape[0] - 1, -1, -1))
        self.num_horizon = self.ilqr_param.num_horizon
        self.x_terminal_guess = x.T[:, self.num_horizon]  # IndexError for laps shorter than the horizon, as :348
        self.cost = self.Qfun[-1][0]
        self.iter_cost.append(deepcopy(self.cost))
        self.iter += 1  # (the reference's LMPC-only bookkeeping, :353-369, is not kept: nothing on the i2LQR path reads it)
        if x.shape[0] > self._capacity:
            while x.shape[0] > self._capacity:
                self._capacity *= 2
            self._sync_engine()  # rebuilds the handle with the stored laps, this one included
        else:
            existed = self._engine is not None
            eng = self._sync_engine()
            if existed and eng.num_laps(0) + self._lap_offset < self.iter:
                eng.add_lap(0, x)

    def calc_input(self):
        param = self.ilqr_param
        if self.num_horizon < param.num_horizon:  # playback of the last plan (reference :377-382)
            self.u_pred = self.u_old
            self.u = self.u_pred[:, 0]
            self.u_old = self.u_pred[:, 1:]
            self.num_horizon -= 1
        else:
            if int(param.max_outloop_iter) < 3:
                # the reference loops range(max_outloop_iter) and only tests the horizon shrink / breaks at iter == 2
                # (:384,:472-478); the device step is exactly those three outer iterations
                raise ValueError("max_outloop_iter < 3 is not supported: the controller step runs the reference's three outer iterations")
            eng = self._sync_engine()
            if isinstance(self.obstacle, (list, tuple)):  # extension: several obstacles, their barrier terms add (i2lqr_set_obstacle_list)
                eng.set_obstacle_list(0, [_abi.obstacle_struct(o) for o in self.obstacle])
            else:
                eng.set_obstacle(0, _abi.obstacle_struct(self.obstacle) if self.obstacle is not None else None)
            self.x_guess = self.x
            out = eng.solve_step(np.asarray(self.x, dtype=np.float64).reshape(1, 4))
            self.u_pred = out["u_pred"][0]
            self.x_pred = out["x_pred"][0]
            self.u = self.u_pred[:, 0]
            self.x_terminal_guess = self.x_pred[:, -1]
            self.x_guess = self.x_terminal_guess
            if self.num_horizon > 1:
                self.u_old = self.u_pred[:, 1:]
            self.winner = (
                int(out["winner_lap"][0]) + self._lap_offset,
                int(out["winner_rank"][0]),
                int(out["winner_idx"][0]),
                float(out["winner_cost"][0]),
            )
            if out["shrink"][0]:
                self.num_horizon -= 1
        self.time += self.timestep


class Simulator:
    """Closed-loop lap runner (reference :693-762)."""

    def __init__(self):
        self.initial_traj = None
        self.robotic = None
        self.timestep = None

    def set_timestep(self, dt):
        self.timestep = dt

    def set_robotic(self, robotic):
        self.robotic = robotic

    def set_traj(self):
        self.initial_traj = self.robotic.xcl

    def sim(self, iter, sim_time=121.0):
        sim_time = 121  # the reference overrides its argument (:709)
        steps = int(sim_time / self.timestep)
        ego = self.robotic
        for i in range(steps):
            ego.forward_one_step()
            obstacle = ego.ctrl_policy.obstacle
            obstacles = list(obstacle) if isinstance(obstacle, (list, tuple)) else ([obstacle] if obstacle is not None else [])
            for o in obstacles:
                o.update_obstacle()
            reached = np.linalg.norm(ego.x - self.initial_traj[-1, :]) <= 0.8
            if reached or i == steps - 1:
                ego.update_memory_post_iter()
                for o in obstacles:
                    o.reset_obstacle()
                print("iteration: {}".format(iter) + (" finished" if reached else " not finished"))
                break

    def plot_inputs(self):
        """No-op: plotting is outside the accelerated path (tests/ilqr_test.py:57-58 calls it unconditionally)."""

    def plot_simulation(self):
        """No-op, see plot_inputs."""


class LMPCParam:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("LMPC / NLMPC (CasADi + IPOPT) is outside the i2LQR hot path this package accelerates")


class LMPC(ControlBase):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("LMPC / NLMPC (CasADi + IPOPT) is outside the i2LQR hot path this package accelerates")
