"""Synthetic Lyman-alpha fields for tests and bench.py.

Lognormal density around mean_rhob, log-uniform temperature in [1e3, 1e7] K, momenta zero.
Cosmology of the reference's Exec/LyA/inputs:72-74; mean_rhob as in
Source/Initialization/Nyx_setup.cpp:162; unit constants evaluated in the order of
Source/Driver/constants_cosmo.H:7-50.  numpy only; no product code depends on this file.
"""
import math

import numpy as np

# Source/Driver/constants_cosmo.H (same expression order)
M_unit = 1.98848e33
L_unit = 3.0856776e24
V_unit = 1.0e5
T_unit = L_unit / V_unit
Gconst = 6.67408e-8 * M_unit * T_unit * T_unit / (L_unit * L_unit * L_unit)
k_B = 1.38064852e-16 * T_unit * T_unit / (M_unit * L_unit * L_unit)
m_proton = 1.672621e-24 / M_unit
mp_over_kb = m_proton / k_B

OMEGA_M = 0.275
OMEGA_B = 0.046
HUBBLE_H = 0.702
H_SPECIES = 0.76
GAMMA = 5.0 / 3.0

NCOMP_STATE = 6
DENSITY, XMOM, YMOM, ZMOM, EDEN, EINT = range(6)
TEMP, NE = 0, 1


def mean_rhob(omega_b=OMEGA_B, h=HUBBLE_H):
    return omega_b * 3.0 * (h * 100.0) * (h * 100.0) / (8.0 * math.pi * Gconst)


def step_dt(z, omega_m=OMEGA_M, h=HUBBLE_H, rel_change=0.01):
    """Full coarse dt = 0.01 * a / (da/dt) (Source/Driver/comoving.cpp:85-101, relative_max_change_a)."""
    a = 1.0 / (1.0 + z)
    oml = 1.0 - omega_m
    dadt = (100.0 * h) * math.sqrt(omega_m / a + oml * a * a)
    return rel_change * a / dadt


def a_after(z, dt, omega_m=OMEGA_M, h=HUBBLE_H):
    a = 1.0 / (1.0 + z)
    oml = 1.0 - omega_m
    dadt = (100.0 * h) * math.sqrt(omega_m / a + oml * a * a)
    return a + dt * dadt


SIGMA_OF_Z = {2.0: 1.4, 3.0: 1.0, 6.0: 0.6}


def sigma_for(z):
    if z in SIGMA_OF_Z:
        return SIGMA_OF_Z[z]
    return float(np.interp(z, [2.0, 3.0, 6.0], [1.4, 1.0, 0.6]))


def e_from_T(T, ne=1.0, h_species=H_SPECIES, gamma=GAMMA):
    """nyx_eos_given_RT (Source/EOS/eos_hc.H:222-231)."""
    Y = (1.0 - h_species) / (4.0 * h_species)
    mu = (1.0 + 4.0 * Y) / (1.0 + Y + ne)
    return T / ((gamma - 1.0) * mp_over_kb * mu)


def make_fab(shape_xyz, seed, z, sigma=None, t_lo=1.0e3, t_hi=1.0e7):
    """One FAB's worth of state (6 comps) and diag (2 comps), Fortran order: arrays are
    returned with shape (ncomp, nz, ny, nx) C-contiguous == (x fastest, component slowest)."""
    nx, ny, nz = shape_xyz
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    if sigma is None:
        sigma = sigma_for(z)
    g = rng.standard_normal((nz, ny, nx))
    u = rng.random((nz, ny, nx))
    rho = mean_rhob() * np.exp(sigma * g - 0.5 * sigma * sigma)
    T = 10.0 ** (math.log10(t_lo) + (math.log10(t_hi) - math.log10(t_lo)) * u)
    e = e_from_T(T)
    state = np.zeros((NCOMP_STATE, nz, ny, nx))
    state[DENSITY] = rho
    state[EINT] = rho * e
    state[EDEN] = rho * e
    diag = np.zeros((2, nz, ny, nx))
    diag[TEMP] = T
    diag[NE] = 1.0
    return state, diag
