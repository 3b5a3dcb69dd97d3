"""Host-side plant model for the single-controller drop-in (reference:
iterative_ilqr/systems/kinetic_bicycle.py:10-27).  The controller's copy of this model, its Jacobians
(get_A_matrix / get_B_matrix, :30-52) and everything else on the solve path run inside libi2lqr.so; this
function only advances the simulated vehicle between controller calls, as the reference's simulator does.
"""
import numpy as np

from utils.constants_kinetic_bicycle import *  # noqa: F401,F403


def kinetic_bicycle(x, u, dt):
    travel = x[2] * dt + (u[0] * dt ** 2) / 2
    return np.array([x[0] + np.cos(x[3]) * travel, x[1] + np.sin(x[3]) * travel, x[2] + u[0] * dt, x[3] + u[1] * dt])
