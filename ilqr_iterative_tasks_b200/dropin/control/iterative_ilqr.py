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
