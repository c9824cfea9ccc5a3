"""Run parameters under the names the reference's modules import (src/configure_me.py:7-40).

A user keeps editing one file, as with the reference; the difference is that the GPU modules look the
values up when they are CALLED (`_runtime.config()` decides which configure_me is consulted), whereas
numba freezes the reference's at first call, so one process may change them between runs.

Every entry below is `NAME: (default, what it controls, who reads it)`; the defaults are the
reference's.  The loop `globals().update(...)` at the end turns the table into the module attributes
`from configure_me import N_CELLS, ...` expects.
"""

PARAMETERS = {
    # -- size of the run ---------------------------------------------------------------------------
    "N_PARTS": (256, "particles along one edge of the lattice (N_PARTS**3 in total)", "every stage"),
    "N_CELLS": (512, "mesh cells along one edge", "every stage"),
    "BOX_SIZE": (100, "comoving edge length, Mpc/h", "initial conditions, snapshot units, images"),
    "N_CPU": (16, "host threads of the reference; accepted and ignored by the GPU path", "host-side reference runs only"),
    "RANDOM_SEED": (38, "key of the Philox streams of the initial conditions", "gaussian_random_field, zeldovich"),
    # -- length of the run and its outputs ----------------------------------------------------------
    "STEPS": (1000, "the scale-factor interval [A_INIT, A_END] is cut into this many steps", "pmesh"),
    "N_SAVE_FILES": (100, "snapshots over that interval", "pmesh"),
    "N_PLOTS": (100, "images over that interval", "pmesh"),
    "PLOT_STEPS": (False, "write Data/snapshots_density{n}.png", "pmesh"),
    "PLOT_PROJECTIONS": (False, "write Data/projection_density{n}.png", "pmesh"),
    "PLOT_GRF": (False, "write Data/snapshot_grf.png", "pmesh"),
    "SAVE_DATA": (True, "write Data/data.{n}.hdf5", "pmesh"),
    "SAVE_DENSITY": (False, "include the density mesh in each snapshot", "save_data"),
    "PRINT_STATUS": (True, "one progress line per step", "pmesh"),
    "RESTART": (False, "start from snapshot RESTART_FROM_N instead of new initial conditions", "pmesh"),
    "RESTART_FROM_N": (0, "index of that snapshot", "pmesh"),
    # -- cosmology ------------------------------------------------------------------------------------
    "POWER": (1.00, "spectral index n of the primordial spectrum (1 = Harrison-Zel'dovich)", "gaussian_random_field"),
    "LCDM_TRANSFER_FUNCTION": (True, "apply the BBKS transfer function (meant for POWER = 1)", "gaussian_random_field"),
    "OMEGA_M0": (0.31, "matter density today", "Poisson factor, growth factor, initial conditions"),
    "OMEGA_B0": (0.04, "baryon density today", "transfer function"),
    "OMEGA_K0": (0.00, "curvature density today", "f(a), growth factor"),
    "OMEGA_LAMBDA0": (0.69, "dark-energy density today", "f(a), growth factor"),
    "H0": (0.68, "Hubble parameter today, units of 100 km/s/Mpc", "f(a) (see SURVEY Q1), snapshot units"),
    "A_INIT": (0.01, "scale factor of the initial conditions", "pmesh, initial conditions"),
    "A_END": (1.00, "scale factor at which the run ends", "pmesh"),
}

globals().update({name: entry[0] for name, entry in PARAMETERS.items()})
