"""Import the UNMODIFIED Python reference in this container and instrument it from outside.

Only tests/golden/make_golden.py uses this module, and only in the build container where
/root/reference exists (it does not exist on the GPU box).  Nothing here is product code.

Recipe: put the inert casadi / matplotlib stand-ins and
<reference>/iterative_ilqr on sys.path, chdir to a scratch directory that holds a copy of the
reference's data/ (KineticBicycle.get_traj WRITES data/closed_loop_feasible.txt,
iterative_ilqr/utils/base.py:133), silence the reference's prints.

Instrumentation is harness-side monkeypatching only:
  * iteration counting: wrap `backward_pass` in control.iterative_ilqr (called once per iLQR iteration,
    iterative_ilqr/control/iterative_ilqr.py:50);
  * per-candidate trace: wrap the `ilqr` name that utils.base calls (iterative_ilqr/utils/base.py:414);
  * (unused since the goldens are minted with NumPy's scalar sort, see make_golden.py) optionally give utils.base
    a numpy proxy whose argsort is kind="stable" (iterative_ilqr/utils/base.py:339 uses the default kind).
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np

REF_ROOT = os.environ.get("I2LQR_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


class _StableNumpy:
    """numpy look-alike whose argsort is stable; everything else is numpy."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def argsort(a, *args, **kwargs):
        kwargs["kind"] = "stable"
        return np.argsort(a, *args, **kwargs)


class Reference:
    def __init__(self, stable_argsort=True):
        if not os.path.isdir(REF_ROOT):
            raise RuntimeError("reference tree not found at %s" % REF_ROOT)
        self.scratch = tempfile.mkdtemp(prefix="i2lqr_ref_")
        shutil.copytree(os.path.join(REF_ROOT, "data"), os.path.join(self.scratch, "data"))
        self._cwd = os.getcwd()
        os.chdir(self.scratch)
        sys.path.insert(0, os.path.join(REF_ROOT, "iterative_ilqr"))
        sys.path.insert(0, os.path.join(_HERE, "_stubs"))
        with self.quiet():
            from utils import base  # noqa: the reference's own module
            import control.iterative_ilqr as ref_ilqr_mod
        self.base = base
        self.ilqr_mod = ref_ilqr_mod
        self.iter_count = 0
        self.trace = None  # list of per-candidate dicts while a step is recorded
        # a process may build several Reference objects: always wrap the pristine functions
        if not hasattr(ref_ilqr_mod, "_pristine"):
            ref_ilqr_mod._pristine = (ref_ilqr_mod.backward_pass, ref_ilqr_mod.ilqr)
        orig_bp = ref_ilqr_mod._pristine[0]

        def counting_bp(*a, **k):
            self.iter_count += 1
            return orig_bp(*a, **k)

        ref_ilqr_mod.backward_pass = counting_bp
        self._orig_ilqr = ref_ilqr_mod._pristine[1]

        def traced_ilqr(*a, **k):
            n0 = self.iter_count
            lamb_in = a[10] if len(a) > 10 else k["lamb"]
            out = self._orig_ilqr(*a, **k)
            if self.trace is not None:
                self.trace.append(
                    {
                        "lamb_in": float(lamb_in),
                        "lamb_out": float(out[2]),
                        "iters": self.iter_count - n0,
                        "x_terminal": np.array(a[6], dtype=np.float64).copy(),
                        "xvar": np.array(out[1]).copy(),
                        "uvar": np.array(out[0]).copy(),
                    }
                )
            return out

        base.ilqr = traced_ilqr
        base.np = _StableNumpy() if stable_argsort else np

    @contextlib.contextmanager
    def quiet(self):
        with contextlib.redirect_stdout(io.StringIO()):
            yield

    def call_ilqr(self, param, N, x0, x_term, obstacle, lamb_in, dt=1, uvar_init=None):
        """One `ilqr()` call set up exactly as calc_input does (iterative_ilqr/utils/base.py:405-426)."""
        base = self.base
        uvar = np.zeros((2, N)) if uvar_init is None else np.array(uvar_init, dtype=np.float64).copy()
        xvar = np.zeros((4, N + 1))
        xvar[:, 0] = x0
        xtarget = np.array([0, 0, 0, 0])
        dX = np.zeros((4, N + 1))
        dX[:, 0] = xvar[:, 0] - xtarget
        n0 = self.iter_count
        with self.quiet():
            u, x, lamb = self._orig_ilqr(
                param, N, xtarget, dt, obstacle, base.KineticBicycleParam(), np.array(x_term), dX, uvar, xvar, lamb_in
            )
        return np.array(u), np.array(x), float(lamb), self.iter_count - n0

    def make_obstacle(self, row, mode, dt=1):
        """Obstacle from the package's (x,y,w,h,spd | mode) row encoding (synth.py)."""
        if row is None:
            return None
        x, y, w, h, spd = [float(v) for v in row]
        if mode == 0:
            return self.base.Obstacle(x, y, w, h)
        return self.base.Obstacle(x, y, w, h, spd=spd, timestep=dt, moving_option=int(mode))

    def close(self):
        os.chdir(self._cwd)
        shutil.rmtree(self.scratch, ignore_errors=True)
