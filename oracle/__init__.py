"""CPU oracle of the i2LQR control solve -- test infrastructure only (see oracle/i2lqr_oracle.c)."""
