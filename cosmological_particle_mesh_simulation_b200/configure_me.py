"""Simulation parameters -- same names, meaning and defaults as the reference's
src/configure_me.py:7-40.  The hot path reads N_CELLS, N_PARTS, N_CPU, STEPS, OMEGA_M0, OMEGA_K0,
OMEGA_LAMBDA0, H0, A_INIT, A_END; the rest is kept so a reference driver finds every name.

Unlike the reference (numba freezes these at first call) the values are read at call time, so a
process may change them between runs; see `_runtime.config()` for which module is consulted."""

###################################################
# General simulation settings
###################################################
N_PARTS          = 256   # particles per dimension
N_CELLS          = 512   # mesh cells per dimension
BOX_SIZE         = 100   # Mpc/h

N_CPU            = 16    # accepted for compatibility; the GPU path ignores it
RANDOM_SEED      = 38

STEPS            = 1000
N_SAVE_FILES     = 100
N_PLOTS          = 100

PLOT_STEPS       = False
PLOT_PROJECTIONS = False
PLOT_GRF         = False
SAVE_DATA        = True
SAVE_DENSITY     = False
PRINT_STATUS     = True

RESTART          = False
RESTART_FROM_N   = 0

###################################################
# Cosmology settings
###################################################
POWER                  = 1.00
LCDM_TRANSFER_FUNCTION = True

OMEGA_M0               = 0.31
OMEGA_B0               = 0.04
OMEGA_K0               = 0.00
OMEGA_LAMBDA0          = 0.69
H0                     = 0.68
A_INIT                 = 0.01
A_END                  = 1.00
