"""Drop-in for `control.iterative_ilqr.ilqr` (reference: iterative_ilqr/control/iterative_ilqr.py:7-85).

Same signature and return value; the solve runs on the GPU through libi2lqr.so (no CPU fallback: a missing
extension or GPU raises RuntimeError).  Differences a caller can observe:
  * `uvar`, `xvar`, `dX` are not mutated in place (the reference clips uvar and rolls xvar / dX forward in
    the caller's arrays before rebinding); the returned arrays are fresh host copies;
  * nothing is printed ("Convergence achieved").
"""
import os
import sys

import numpy as np

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.append(_PKG_PARENT)

from ilqr_iterative_tasks_b200 import _abi, engine  # noqa: E402

_engines = {}


def _engine_for(params):
    key = bytes(params)
    eng = _engines.get(key)
    if eng is None:
        if len(_engines) > 8:
            _engines.clear()
        eng = engine.Engine(params)
        _engines[key] = eng
    return eng


def ilqr(ilqr_param, num_horizon, xtarget, timestep, obstacle, system_param, x_terminal, dX, uvar, xvar, lamb):
    params = _abi.params_from_reference(ilqr_param, system_param, timestep)
    params.obstacle_horizon = int(ilqr_param.num_horizon)  # control/ilqr_helper.py:134-139
    params.num_horizon = int(num_horizon)
    _abi.set_field(params, "xtarget", np.asarray(xtarget, dtype=np.float64).reshape(4))
    eng = _engine_for(params)
    xvar = np.asarray(xvar, dtype=np.float64)
    uinit = np.asarray(uvar, dtype=np.float64)
    out = eng.solve_batch(
        xvar[:, 0].reshape(1, 4),
        np.asarray(x_terminal, dtype=np.float64).reshape(1, 4),
        np.array([float(lamb)]),
        _abi.obstacle_struct(obstacle),
        want=("uvar", "xvar", "lamb_out"),
        uvar_init=uinit.reshape(1, 2, -1) if uinit.any() else None,
    )
    return out["uvar"][0], out["xvar"][0], float(out["lamb_out"][0])


def _component_engine(ilqr_param, num_horizon, timestep, sys_param):
    params = _abi.params_from_reference(ilqr_param, sys_param, timestep)
    params.obstacle_horizon = int(ilqr_param.num_horizon)  # control/ilqr_helper.py:134-139
    params.num_horizon = int(num_horizon)
    return _engine_for(params)


def backward_pass(xvar, uvar, x_terminal, dX, lamb, num_horizon, timestep, ilqr_param, obstacle, sys_param):
    """Drop-in for `backward_pass` (reference: iterative_ilqr/control/iterative_ilqr.py:88-130): the Riccati recursion
    on the GPU (i2lqr_backward_pass_batch), returning `(matrix_k (2,N), matrix_K (2,4,N))`.  `dX` must be what
    ilqr() maintains, xvar - xtarget with xtarget = 0 (utils/base.py:374); anything else raises ValueError."""
    xvar = np.asarray(xvar, dtype=np.float64)
    if not np.array_equal(np.asarray(dX, dtype=np.float64)[:, :num_horizon], xvar[:, :num_horizon]):
        raise ValueError("backward_pass: dX must equal xvar - xtarget (xtarget = 0), as ilqr() maintains it")
    eng = _component_engine(ilqr_param, num_horizon, timestep, sys_param)
    k, K = eng.backward_pass(xvar[None], np.asarray(uvar, dtype=np.float64)[None], np.asarray(x_terminal, dtype=np.float64).reshape(1, 4),
                             np.array([float(lamb)]), _abi.obstacle_struct(obstacle))  # fmt: skip
    return k[0], K[0]


def forward_pass(xvar, uvar, x_terminal, ilqr_param, timestep, num_horizon, matrix_k, matrix_K, sys_param):
    """Drop-in for `forward_pass` (reference: iterative_ilqr/control/iterative_ilqr.py:133-160): full-step rollout
    with clipping on the GPU (i2lqr_forward_pass_batch); returns `(xvar_new, uvar_new, cost_new)`."""
    eng = _component_engine(ilqr_param, num_horizon, timestep, sys_param)
    xn, un, cost = eng.forward_pass(np.asarray(xvar, dtype=np.float64)[None], np.asarray(uvar, dtype=np.float64)[None],
                                    np.asarray(x_terminal, dtype=np.float64).reshape(1, 4), np.asarray(matrix_k, dtype=np.float64)[None],
                                    np.asarray(matrix_K, dtype=np.float64)[None])  # fmt: skip
    return xn[0], un[0], float(cost[0])
