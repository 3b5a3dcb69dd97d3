"""Inert stand-in for casadi, used only by tests/golden/make_golden.py.

The reference imports casadi at module load (iterative_ilqr/utils/base.py:7,
iterative_ilqr/control/ilqr_helper.py:3-4) but never calls it on the i2LQR path.
"""
