"""State / input dimensions and index maps (reference: iterative_ilqr/utils/constants_kinetic_bicycle.py:2-6)."""
X_DIM, U_DIM = 4, 2
X_ID = {"x": 0, "y": 1, "v": 2, "theta": 3}
U_ID = {"accel": 0, "delta": 1}
