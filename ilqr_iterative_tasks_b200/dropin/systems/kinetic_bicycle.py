"""Drop-in for the reference module of the same name (iterative_ilqr/systems/kinetic_bicycle.py:10-52): the
kinematic bicycle model and its Jacobians, evaluated on the GPU by the same device code the solve kernels inline
(include/i2lqr.h: i2lqr_dynamics_batch, i2lqr_linearize_batch).  No CPU fallback: without libi2lqr.so or a GPU
these raise RuntimeError.  The batched simulator (episodes.py) never calls them -- its plant runs inside the
episode kernels; the single-controller simulator of utils.base advances its vehicle through `kinetic_bicycle`.
"""
import os
import sys

import numpy as np

from utils.constants_kinetic_bicycle import *  # noqa: F401,F403

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.append(_PKG_PARENT)

from ilqr_iterative_tasks_b200 import _abi, engine  # noqa: E402

_engines = {}


def _engine(dt, num_horizon=2):
    key = (float(dt), int(num_horizon))
    eng = _engines.get(key)
    if eng is None:
        if len(_engines) > 8:
            _engines.clear()
        eng = engine.Engine(_abi.default_params(timestep=float(dt), num_horizon=int(num_horizon)))
        _engines[key] = eng
    return eng


def kinetic_bicycle(x, u, dt):
    return _engine(dt).dynamics(np.asarray(x, dtype=np.float64).reshape(1, 4), np.asarray(u, dtype=np.float64).reshape(1, 2))[0]


def get_A_matrix(velocity_vals, theta, acceleration_vals, num_horizon, dt):
    A, _ = _engine(dt, num_horizon).linearize(velocity_vals, theta, acceleration_vals, want_B=False)
    return A[0]


def get_B_matrix(theta, num_horizon, dt):
    _, B = _engine(dt, num_horizon).linearize(None, theta, None, want_A=False)
    return B[0]
