"""Drop-in for the reference module of the same name (iterative_ilqr/control/ilqr_helper.py): the cost
derivatives of the iLQR iteration, evaluated on the GPU by the device code the solve kernels inline
(include/i2lqr.h: i2lqr_cost_derivatives_batch, i2lqr_cost_final_batch).  `repelling_cost_function` (:59-64) has
no entry point of its own -- it is the three-line barrier formula inside those kernels -- and `select_points`
(:67-80) is dead code in the reference.
"""
import numpy as np

from control.iterative_ilqr import _abi, _engine_for


def _engine(ilqr_param, num_horizon, sys_param):
    params = _abi.params_from_reference(ilqr_param, sys_param, getattr(ilqr_param, "timestep", None))
    params.obstacle_horizon = int(ilqr_param.num_horizon)  # control/ilqr_helper.py:134-139
    params.num_horizon = int(num_horizon)
    return _engine_for(params)


def get_cost_derivation(ctrl_U, dX, ilqr_param, num_horizon, xvar, obstacle, sys_param):
    """(reference :9-56) -> l_u (2,N), l_uu (2,2,N), l_x (4,N), l_xx (4,4,N)"""
    eng = _engine(ilqr_param, num_horizon, sys_param)
    out = eng.cost_derivatives(np.asarray(ctrl_U, dtype=np.float64)[None], np.asarray(dX, dtype=np.float64)[None],
                               np.asarray(xvar, dtype=np.float64)[None], _abi.obstacle_struct(obstacle))  # fmt: skip
    return tuple(a[0] for a in out)


def add_control_constraint(u, ilqr_param, sys_param):
    """(reference :83-103) -> b_dot_ctrl (2,1), b_ddot_ctrl (2,2): the input-barrier part of l_u / l_uu."""
    params = _abi.params_from_reference(ilqr_param, sys_param, getattr(ilqr_param, "timestep", None))
    params.num_horizon = 2
    _abi.set_field(params, "matrix_Q", np.zeros(16))
    _abi.set_field(params, "matrix_R", np.zeros(4))
    eng = _engine_for(params)
    uvar = np.zeros((1, 2, 2))
    uvar[0, :, 0] = np.asarray(u, dtype=np.float64).reshape(2)
    l_u, l_uu, _, _ = eng.cost_derivatives(uvar, np.zeros((1, 4, 3)), np.zeros((1, 4, 3)), None)
    return l_u[0, :, 0].reshape(2, 1), l_uu[0, :, :, 0]


def get_cost_final(x, x_terminal, matrix_Qterminal, obstacle, ilqr_param):
    """(reference :106-150) -> l_x (4,), l_xx (4,4)"""
    x = np.asarray(x, dtype=np.float64)
    params = _abi.params_from_reference(ilqr_param, None, getattr(ilqr_param, "timestep", None))
    params.obstacle_horizon = int(ilqr_param.num_horizon)
    params.num_horizon = x.shape[1] - 1
    _abi.set_field(params, "matrix_Qterminal", np.asarray(matrix_Qterminal, dtype=np.float64))
    eng = _engine_for(params)
    l_x, l_xx = eng.cost_final(x[None], np.asarray(x_terminal, dtype=np.float64).reshape(1, 4), _abi.obstacle_struct(obstacle))
    return l_x[0], l_xx[0]
