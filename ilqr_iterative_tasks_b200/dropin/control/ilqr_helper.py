"""Placeholder for the reference module of the same name (iterative_ilqr/control/ilqr_helper.py).

Its functions (get_cost_derivation, add_control_constraint, repelling_cost_function, get_cost_final) are
internal to the solve and live in libi2lqr.so (csrc/ilqr_core.cuh); `select_points` (:67-80) is dead code in
the reference.  Nothing on the controller API imports from here.
"""
