"""B200-native i2LQR per-timestep control solve (drop-in for HybridRobotics/ilqr-iterative-tasks' hot path).

  engine.Engine            ctypes binding of the C ABI (include/i2lqr.h) -> libi2lqr.so (sm_100a kernels)
  episodes.EpisodeBatch    many independent closed-loop episodes, device-resident
  dist                     sharding of candidates / episodes over the GPUs of one node
  synth                    synthetic workload generators (BASELINE.json configs 3-5)
  dropin/                  put on sys.path: `utils.base.iLqr`, `control.iterative_ilqr.ilqr`, ... with the
                           reference's names and signatures
"""
import os

DROPIN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install_dropin():
    """Prepends the drop-in module tree to sys.path (what `package_dir = =iterative_ilqr` does for the reference)."""
    import sys

    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    return DROPIN_PATH
