"""The functions an iLQR iteration is made of, called one by one through the C ABI (include/i2lqr.h, "one by
one") and through the drop-in modules, against outputs of the reference's own functions on mid-solve
trajectories (tests/golden/components.npz, minted by make_golden.py components):
  kinetic_bicycle / get_A_matrix / get_B_matrix       iterative_ilqr/systems/kinetic_bicycle.py:10-52
  get_cost_derivation / add_control_constraint / get_cost_final   iterative_ilqr/control/ilqr_helper.py:9-150
  backward_pass / forward_pass                        iterative_ilqr/control/iterative_ilqr.py:88-160
Bar: 1e-9 relative (FP64 build), per problem, relative to max(1, max|reference array|)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from ilqr_iterative_tasks_b200 import _abi
from tests import helpers as H

pytestmark = pytest.mark.gpu
GROUPS = ["n6", "n6_qr", "n25"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng_mod():
    from ilqr_iterative_tasks_b200 import engine

    return engine


def _setup(eng_mod, g, grp):
    N = int(g[grp + "/N"])
    p = _abi.default_params(num_horizon=N, matrix_Q=g[grp + "/Q"], matrix_R=g[grp + "/R"])
    obs = _abi.obstacle_array(g[grp + "/obs"], g[grp + "/mode"])
    obs["present"] = g[grp + "/has_obs"]
    return eng_mod.Engine(p), obs, N


def _close(a, ref, tol=H.RTOL):
    err = H.rel_err(a, ref)
    assert err.max() <= tol, (err.max(), int(err.argmax()))


@pytest.mark.parametrize("grp", GROUPS)
def test_dynamics_and_linearisation(eng_mod, grp):
    g = H.load("components.npz")
    e, _, N = _setup(eng_mod, g, grp)
    x, u = g[grp + "/xvar"], g[grp + "/uvar"]
    B = x.shape[0]
    # kinetic_bicycle: every stage of every stored rollout is one (x_k, u_k) -> x_{k+1} sample
    xk = np.ascontiguousarray(x[:, :, :N].transpose(0, 2, 1)).reshape(-1, 4)
    uk = np.ascontiguousarray(u.transpose(0, 2, 1)).reshape(-1, 2)
    xn = e.dynamics(xk, uk).reshape(B, N, 4).transpose(0, 2, 1)
    _close(xn, x[:, :, 1:], 1e-13)
    A, Bm = e.linearize(x[:, 2, 1:], x[:, 3, 1:], u[:, 0, :])
    _close(A, g[grp + "/A"], 1e-13)
    _close(Bm, g[grp + "/B"], 1e-13)
    assert np.array_equal(A[:, [0, 1, 2, 3], [0, 1, 2, 3], :], np.ones((B, 4, N)))  # structural entries exact
    e.close()


@pytest.mark.parametrize("grp", GROUPS)
def test_cost_derivatives_and_final(eng_mod, grp):
    g = H.load("components.npz")
    e, obs, N = _setup(eng_mod, g, grp)
    x, u = g[grp + "/xvar"], g[grp + "/uvar"]
    l_u, l_uu, l_x, l_xx = e.cost_derivatives(u, x, x, obs)
    for a, name in ((l_u, "l_u"), (l_uu, "l_uu"), (l_x, "l_x"), (l_xx, "l_xx")):
        _close(a, g[grp + "/" + name])
    Vx, Vxx = e.cost_final(x, g[grp + "/x_term"], obs)
    _close(Vx, g[grp + "/Vx"])
    _close(Vxx, g[grp + "/Vxx"])
    e.close()


@pytest.mark.parametrize("grp", GROUPS)
def test_backward_and_forward_pass(eng_mod, grp):
    g = H.load("components.npz")
    e, obs, N = _setup(eng_mod, g, grp)
    x, u, xt = g[grp + "/xvar"], g[grp + "/uvar"], g[grp + "/x_term"]
    k, K = e.backward_pass(x, u, xt, g[grp + "/lamb"], obs)
    _close(k, g[grp + "/k"])
    _close(K, g[grp + "/K"])
    # forward pass from the reference's gains (so that an error of the backward pass cannot hide here) ...
    xn, un, cn = e.forward_pass(x, u, xt, g[grp + "/k"], g[grp + "/K"])
    _close(xn, g[grp + "/xvar_new"])
    _close(un, g[grp + "/uvar_new"])
    _close(cn[:, None], g[grp + "/cost_new"][:, None])
    # ... and chained from our own gains, as the solve kernels do
    xn2, un2, cn2 = e.forward_pass(x, u, xt, k, K)
    _close(xn2, g[grp + "/xvar_new"], 1e-8)
    _close(cn2[:, None], g[grp + "/cost_new"][:, None], 1e-8)
    e.close()


@pytest.mark.parametrize("grp", GROUPS)
def test_control_constraint_vs_reference_golden(eng_mod, grp):
    """add_control_constraint (control/ilqr_helper.py:83-103) against the reference's own outputs (`cc_b`, `cc_bdd`, one
    call per stage input): the input-barrier gradient / Hessian is what l_u / l_uu are when matrix_R = 0."""
    g = H.load("components.npz")
    N = int(g[grp + "/N"])
    e = eng_mod.Engine(_abi.default_params(num_horizon=N))  # matrix_Q = matrix_R = 0
    u = g[grp + "/uvar"]
    B = u.shape[0]
    l_u, l_uu, _, _ = e.cost_derivatives(u, np.zeros((B, 4, N + 1)), np.zeros((B, 4, N + 1)), None)
    _close(l_u.transpose(0, 2, 1), g[grp + "/cc_b"])
    _close(l_uu.transpose(0, 3, 1, 2), g[grp + "/cc_bdd"])
    e.close()


def test_component_error_paths(eng_mod):
    p = _abi.default_params()
    p.precision = 1  # the component entry points exist in FP64 only
    e = eng_mod.Engine(p)
    with pytest.raises(RuntimeError):
        e.dynamics(np.zeros((1, 4)), np.zeros((1, 2)))
    e.close()
    e = eng_mod.Engine(_abi.default_params())
    with pytest.raises(ValueError):
        e.backward_pass(np.zeros((2, 4, 7)), np.zeros((2, 2, 6)), np.zeros((2, 4)), None, np.zeros(3, dtype=_abi.OBSTACLE_DTYPE))
    e.close()


_DROPIN = r"""
import sys, numpy as np
sys.path.insert(0, r"{root}/ilqr_iterative_tasks_b200/dropin")
sys.path.insert(0, r"{root}")
from utils import base
from control.iterative_ilqr import backward_pass, forward_pass
from control.ilqr_helper import get_cost_derivation, get_cost_final, add_control_constraint
from systems.kinetic_bicycle import kinetic_bicycle, get_A_matrix, get_B_matrix
from tests import helpers as H
g = H.load("components.npz")
grp, tol = "n6_qr", 1e-9
N = int(g[grp + "/N"])
param = base.iLqrParam(num_ss_points=8, num_ss_iter=1, timestep=1, num_horizon=N, matrix_Q=g[grp + "/Q"], matrix_R=g[grp + "/R"])
sysp = base.KineticBicycleParam()
def close(a, ref):
    assert np.abs(np.asarray(a) - ref).max() <= tol * max(1.0, np.abs(ref).max()), (np.abs(np.asarray(a) - ref).max())
for b in (0, 1, 2, 7):
    x, u, xt = g[grp + "/xvar"][b], g[grp + "/uvar"][b], g[grp + "/x_term"][b]
    o, m = g[grp + "/obs"][b], int(g[grp + "/mode"][b])
    obstacle = None
    if g[grp + "/has_obs"][b]:
        obstacle = base.Obstacle(o[0], o[1], o[2], o[3]) if m == 0 else base.Obstacle(o[0], o[1], o[2], o[3], spd=o[4], timestep=1, moving_option=m)
    close(kinetic_bicycle(x[:, 2], u[:, 2], 1), x[:, 3])
    close(get_A_matrix(x[2, 1:], x[3, 1:], u[0, :], N, 1), g[grp + "/A"][b])
    close(get_B_matrix(x[3, 1:], N, 1), g[grp + "/B"][b])
    for a, name in zip(get_cost_derivation(u, x.copy(), param, N, x, obstacle, sysp), ("l_u", "l_uu", "l_x", "l_xx")):
        close(a, g[grp + "/" + name][b])
    Vx, Vxx = get_cost_final(x, xt, param.matrix_Qterminal, obstacle, param)
    close(Vx, g[grp + "/Vx"][b]); close(Vxx, g[grp + "/Vxx"][b])
    k, K = backward_pass(x, u, xt, x.copy(), float(g[grp + "/lamb"][b]), N, 1, param, obstacle, sysp)
    close(k, g[grp + "/k"][b]); close(K, g[grp + "/K"][b])
    xn, un, cn = forward_pass(x, u, xt, param, 1, N, g[grp + "/k"][b], g[grp + "/K"][b], sysp)
    close(xn, g[grp + "/xvar_new"][b]); close(un, g[grp + "/uvar_new"][b]); close(cn, g[grp + "/cost_new"][b])
    for k in (0, N - 1):
        bd, bdd = add_control_constraint(u[:, k], param, sysp)
        assert bd.shape == (2, 1) and bdd.shape == (2, 2)
        close(bd[:, 0], g[grp + "/cc_b"][b, k]); close(bdd, g[grp + "/cc_bdd"][b, k])  # the reference's own outputs
# the scripted first lap and its input log (utils/base.py:103-138), first-row aliasing quirk included
ego = base.KineticBicycle(False, system_param=sysp)
ego.set_state([0, 0, 0, 0]); ego.set_timestep(1); ego.get_traj()
assert np.array_equal(ego.ucl, g["traj/ucl"]), (ego.ucl[:3], g["traj/ucl"][:3])
close(ego.xcl, g["traj/xcl"])
print("OK")
"""


def test_dropin_component_functions():
    """The reference's module-level functions under their own names and signatures (drop-in modules)."""
    r = subprocess.run([sys.executable, "-c", _DROPIN.format(root=ROOT)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "OK" in r.stdout
