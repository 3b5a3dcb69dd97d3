"""ctypes mirror of include/i2lqr.h (structs only) and converters from the reference's Python objects."""
import ctypes as C

import numpy as np

OUTER_ITERS = 3


class Params(C.Structure):
    """`i2lqr_params`: iLqrParam + KineticBicycleParam flattened (iterative_ilqr/utils/base.py:15-20,242-302)."""

    _fields_ = [
        ("num_horizon", C.c_int32),
        ("num_ss_points", C.c_int32),
        ("num_ss_iter", C.c_int32),
        ("max_ilqr_iter", C.c_int32),
        ("max_relax_iter", C.c_int32),
        ("max_sim_steps", C.c_int32),
        ("obstacle_horizon", C.c_int32),
        ("precision", C.c_int32),
        ("matrix_Q", C.c_double * 16),
        ("matrix_R", C.c_double * 4),
        ("matrix_Qterminal", C.c_double * 16),
        ("tuning_ctrl_q1", C.c_double),
        ("tuning_ctrl_q2", C.c_double),
        ("tuning_obs_q1", C.c_double),
        ("tuning_obs_q2", C.c_double),
        ("safety_margin", C.c_double),
        ("eps", C.c_double),
        ("lamb", C.c_double),
        ("lamb_factor", C.c_double),
        ("max_lamb", C.c_double),
        ("a_max", C.c_double),
        ("delta_lim", C.c_double),
        ("timestep", C.c_double),
        ("xtarget", C.c_double * 4),
        ("goal_tol", C.c_double),
    ]


class Obstacle(C.Structure):
    """`i2lqr_obstacle` (iterative_ilqr/utils/base.py:23-35)."""

    _fields_ = [
        ("x", C.c_double),
        ("y", C.c_double),
        ("width", C.c_double),
        ("height", C.c_double),
        ("spd", C.c_double),
        ("moving_option", C.c_int32),
        ("present", C.c_int32),
    ]


OBSTACLE_DTYPE = np.dtype(
    [("x", "f8"), ("y", "f8"), ("width", "f8"), ("height", "f8"), ("spd", "f8"), ("moving_option", "i4"), ("present", "i4")]
)
assert OBSTACLE_DTYPE.itemsize == C.sizeof(Obstacle)


def default_params(**overrides):
    """The reference's defaults: iLqrParam() (utils/base.py:242-302), KineticBicycleParam() (:15-20)."""
    p = Params()
    p.num_horizon, p.num_ss_points, p.num_ss_iter = 6, 8, 1
    p.max_ilqr_iter, p.max_relax_iter, p.max_sim_steps = 150, 55, 121
    qf = 2 * np.diag([1.0, 1.0, 20.0, 0.02])
    p.matrix_Qterminal[:] = list(qf.ravel())
    p.tuning_ctrl_q1 = p.tuning_ctrl_q2 = 1.0
    p.tuning_obs_q1 = p.tuning_obs_q2 = 2.74
    p.safety_margin, p.eps = 0.0, 1e-2
    p.lamb, p.lamb_factor, p.max_lamb = 1.0, 10.0, 1000.0
    p.a_max, p.delta_lim = 2.0, round(np.pi / 2, 2)
    p.timestep, p.goal_tol = 1.0, 0.8
    for k, v in overrides.items():
        set_field(p, k, v)
    return p


def set_field(p, name, value):
    cur = getattr(p, name)
    if isinstance(cur, C.Array):
        flat = np.asarray(value, dtype=np.float64).ravel()
        if flat.size != len(cur):
            raise ValueError("%s needs %d values" % (name, len(cur)))
        cur[:] = list(flat)
    else:
        setattr(p, name, value)


def params_from_reference(ilqr_param, system_param=None, timestep=None):
    """Flatten an `iLqrParam`-like and a `KineticBicycleParam`-like object (duck-typed) into `Params`."""
    p = default_params()
    for name in (
        "num_horizon", "num_ss_points", "num_ss_iter", "max_ilqr_iter", "max_relax_iter",
        "tuning_ctrl_q1", "tuning_ctrl_q2", "tuning_obs_q1", "tuning_obs_q2", "safety_margin",
        "eps", "lamb", "lamb_factor", "max_lamb",
    ):  # fmt: skip
        if hasattr(ilqr_param, name):
            value = getattr(ilqr_param, name)
            # ctypes int fields reject floats
            setattr(p, name, int(value) if isinstance(getattr(p, name), int) else float(value))
    for name in ("matrix_Q", "matrix_R", "matrix_Qterminal"):
        if hasattr(ilqr_param, name):
            set_field(p, name, getattr(ilqr_param, name))
    if system_param is not None:
        p.a_max = float(system_param.a_max)
        p.delta_lim = round(float(system_param.delta_max), 2)  # control/iterative_ilqr.py:37-39
    dt = timestep if timestep is not None else getattr(ilqr_param, "timestep", None)
    if dt is not None:
        p.timestep = float(dt)
    return p


def obstacle_struct(obstacle):
    """`Obstacle`-like object or None -> `Obstacle` struct, with the reference's failure modes.

    The reference takes the static branch when `spd == 0 or spd is None`, the moving branches on
    `moving_option` 1 / 2 (control/ilqr_helper.py:34-43).  A non-zero spd without moving_option leaves
    its locals unbound (UnboundLocalError) and moving_option with spd None multiplies by None
    (TypeError); both are rejected here with the same exception types.
    """
    o = Obstacle()
    if obstacle is None:
        return o
    spd, mode = obstacle.spd, obstacle.moving_option
    if mode in (1, 2) and spd is None:
        raise TypeError("unsupported operand type(s) for *: 'int' and 'NoneType'")
    if mode not in (1, 2) and not (spd is None or spd == 0):
        raise UnboundLocalError("obstacle with spd != 0 needs moving_option 1 or 2")
    o.x, o.y, o.width, o.height = float(obstacle.x), float(obstacle.y), float(obstacle.width), float(obstacle.height)
    o.spd = 0.0 if spd is None else float(spd)
    o.moving_option = int(mode) if mode in (1, 2) else 0
    o.present = 1
    return o


def obstacle_array(obs_rows, modes):
    """(B,5) rows [x,y,w,h,spd] + (B,) modes (synth.py encoding) -> structured array of i2lqr_obstacle."""
    obs_rows = np.asarray(obs_rows, dtype=np.float64).reshape(-1, 5)
    out = np.zeros(obs_rows.shape[0], dtype=OBSTACLE_DTYPE)
    out["x"], out["y"], out["width"], out["height"], out["spd"] = obs_rows.T
    out["moving_option"] = np.asarray(modes, dtype=np.int32).reshape(-1)
    out["present"] = 1
    return out
