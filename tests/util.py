"""Shared helpers for the parity tests."""
import ctypes as C
import os

import numpy as np

from nyx_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sdc_inputs(z, n, seed, src_scale=0.0):
    """A self-consistent SDC reaction step: S_new = update_state_with_sources(S_old, hydro_src) as in
    Source/TimeStep/Nyx_update_state_with_sources.cpp:43-71 (+ reset_src)."""
    a = 1.0 / (1.0 + z)
    dt = synth.step_dt(z)
    a_end = synth.a_after(z, dt)
    rng = np.random.default_rng(seed)
    s_old, diag = synth.make_fab((n, n, n), seed=seed, z=z)
    s_new = s_old.copy()
    hs = np.zeros_like(s_old)
    rs = np.zeros((1, n, n, n))
    ir = np.zeros((1, n, n, n))
    if src_scale:
        hs[0] = s_old[0] * src_scale * rng.standard_normal((n, n, n))
        hs[5] = s_old[5] * src_scale * rng.standard_normal((n, n, n))
        rs[0] = s_old[5] * src_scale * 0.1 * rng.standard_normal((n, n, n))
        s_new[0] = s_old[0] + hs[0]
        s_new[5] = (a * a * s_old[5] + hs[5]) / (a_end * a_end) + rs[0]
        s_new[4] = s_new[5].copy()
    return dict(a=a, a_end=a_end, dt=dt, s_old=s_old, s_new=s_new, diag=diag, hydro_src=hs, reset_src=rs, ir=ir)


def inhomo_inputs(z, n, seed, src_scale=0.05):
    """sdc_inputs with a third diag component: the per-cell hydrogen reionization redshift z_HI (nyx.inhomo_reion = 1,
    HC/f_rhs_struct.H:202-209,353-359): cells not yet reionized (z > z_HI: no UVB), cells reionized during this step
    (z_end <= z_HI < z: heat injection towards T_zHI) and cells reionized earlier."""
    d = sdc_inputs(z, n, seed, src_scale)
    z_end = 1.0 / d["a_end"] - 1.0
    rng = np.random.default_rng(seed + 1000)
    u = rng.random((n, n, n))
    zhi = np.where(u < 0.3, rng.uniform(z_end - 2.0, z_end - 1e-3, (n, n, n)),            # later: z > z_HI, JH = 0
                   np.where(u < 0.6, rng.uniform(z_end, z - 1e-9, (n, n, n)),             # during this step
                            rng.uniform(z + 1e-6, z + 3.0, (n, n, n))))                   # earlier
    d["diag"] = np.ascontiguousarray(np.concatenate([d["diag"], zhi[None]], axis=0))
    d["kw"] = dict(inhomo_reion=1, zhi_flash=6.0, T_zhi=2.0e4, zheii_flash=-1.0, T_zheii=0.0)
    return d


def eos_rows_inputs(n, seed, z=3.0):
    """A box for Nyx::compute_new_temp / reset_internal_energy: the synthetic LyA field plus momenta, a total energy that is (in
    parts of the box) inconsistent with rho e, cells with rho e <= 0, very hot cells (large_temp clipping) and stale diag(Ne)."""
    rng = np.random.default_rng(seed)
    state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
    rho = state[0]
    for c in (1, 2, 3):
        state[c] = rho * 300.0 * rng.standard_normal((n, n, n))          # momenta: ~300 km/s
    ke = 0.5 * (state[1] ** 2 + state[2] ** 2 + state[3] ** 2) / rho
    state[4] = state[5] + ke
    u = rng.random((n, n, n))
    state[4] = np.where(u < 0.10, state[5] * 0.5 + ke, state[4])          # rho E inconsistent with rho e
    state[4] = np.where((u >= 0.10) & (u < 0.15), ke * (1.0 + 1e-8), state[4])   # (e from E) below 1e-6 of E
    state[5] = np.where((u >= 0.15) & (u < 0.20), -np.abs(state[5]), state[5])   # negative rho e
    state[5] = np.where((u >= 0.20) & (u < 0.22), 0.0, state[5])
    state[5] = np.where((u >= 0.22) & (u < 0.25), state[5] * 1e3, state[5])      # very hot
    diag[1] = rng.uniform(0.0, 1.2, (n, n, n))                            # the cell's current ne
    reset_src = rng.standard_normal((1, n, n, n)) * 1e-3
    return state, diag, reset_src


SRC_BOXES = [(0, 0, 0, 15, 11, 9), (16, 0, 0, 27, 11, 9), (0, 12, 0, 27, 19, 9)]   # ragged: 16x12x10, 12x12x10, 28x8x10
SRC_NG = (4, 1, 4, 0, 1, 4)     # S_old_tmp, S_new, ext_src_old, hydro_src, grav_vector, reset_e_src (sdc_hydro.cpp:60-106)
SRC_NCOMP = (6, 6, 6, 6, 3, 1)


def sources_inputs(seed, z=3.0, low_density_cells=0, boxes=SRC_BOXES, ng=SRC_NG):
    """Inputs of Nyx::update_state_with_sources (SDC) on a multi-box level: per slot a list of per-box arrays (ncomp, nz, ny, nx) covering
    the box grown by ng[slot].  low_density_cells > 0: that many cells per box get a hydro source that drives the new density below
    small_dens (and a few exactly to it), so that enforce_minimum_density acts."""
    rng = np.random.default_rng(seed)
    a_old = 1.0 / (1.0 + z)
    dt = synth.step_dt(z)
    a_new = synth.a_after(z, dt)
    small_dens = 1.0e-2 * synth.mean_rhob()
    out = {k: [] for k in ("s_old", "s_new", "ext_src", "hydro_src", "grav", "reset_src")}
    for bi, bx in enumerate(boxes):
        shp = lambda g: (bx[5] - bx[2] + 1 + 2 * g, bx[4] - bx[1] + 1 + 2 * g, bx[3] - bx[0] + 1 + 2 * g)
        nz, ny, nx = shp(ng[0])
        s_old, _ = synth.make_fab((nx, ny, nz), seed=seed + bi, z=z)
        s_old[0] = np.maximum(s_old[0], 3.0 * small_dens)
        for c in (1, 2, 3):
            s_old[c] = s_old[0] * 300.0 * rng.standard_normal((nz, ny, nx))
        s_old[4] = s_old[5] + 0.5 * (s_old[1] ** 2 + s_old[2] ** 2 + s_old[3] ** 2) / s_old[0]
        s_new = rng.standard_normal((6,) + shp(ng[1]))                       # overwritten by the call
        ext = np.zeros((6,) + shp(ng[2]))
        inner = tuple(slice(ng[0] - ng[2], s - (ng[0] - ng[2])) for s in (nz, ny, nx))
        for c in range(6):
            ext[c] = s_old[(c,) + inner] * (0.0 if c == 0 else 0.05 / dt) * rng.standard_normal(ext.shape[1:])
        vin = tuple(slice(ng[0], s - ng[0]) for s in (nz, ny, nx))
        hs = np.stack([s_old[(c,) + vin] * 0.1 * rng.standard_normal(shp(0)) for c in range(6)])
        if low_density_cells:
            flat = hs[0].reshape(-1)
            pick = rng.choice(flat.size, low_density_cells, replace=False)
            rho_v = s_old[(0,) + vin].reshape(-1)
            flat[pick] = -rho_v[pick] * rng.uniform(0.9, 1.3, low_density_cells)       # new density near or below zero
            flat[pick[0]] = small_dens - rho_v[pick[0]]                              # new density == small_dens (up to rounding): not floored or floored, both sides must agree
        grav = rng.standard_normal((3,) + shp(ng[4])) * 1.0e3
        rs = np.zeros((1,) + shp(ng[5]))
        for k, v in zip(out, (s_old, s_new, ext, hs, grav, rs)):
            out[k].append(np.ascontiguousarray(v))
    out.update(boxes=[tuple(b) for b in boxes], ng=ng, dt=dt, a_old=a_old, a_new=a_new, small_dens=small_dens, small_temp=1.0e-2)
    return out


FLASH_CASES = {
    "none": {},
    "hi_now": dict(zhi_flash=5.98, T_zhi=2e4, zheii_flash=3.0, T_zheii=1.5e4),     # with z = 5.99: H flash inside the step
    "heii_now": dict(zhi_flash=6.0, T_zhi=2e4, zheii_flash=2.999, T_zheii=1.5e4),  # with z = 3.0: HeII flash inside the step
    "before": dict(zhi_flash=6.0, T_zhi=2e4, zheii_flash=3.0, T_zheii=1.5e4),      # with z = 7.0: J = 0, no UVB yet
}


class Harness:
    """tests/host_harness.cpp: the product's state machine compiled for the host."""

    def __init__(self, path):
        self.lib = lib = C.CDLL(path)
        fp, pp = C.POINTER(capi.HcFab), C.POINTER(capi.HcParams)
        lib.hh_tabulate_rates.argtypes = [C.c_char_p, C.c_double, C.c_void_p]
        lib.hh_integrate_vec.argtypes = [C.c_void_p, pp, fp, fp, capi.HcBox, C.c_double, C.c_double, C.c_void_p]
        lib.hh_integrate_struct.argtypes = [C.c_void_p, pp] + [fp] * 6 + [capi.HcBox, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]

    @staticmethod
    def params(**kw):
        p = capi.HcParams(rtol=1e-4, atol_factor=1e-4, h_species=0.76, gamma_minus_1=5.0 / 3.0 - 1.0, uvb_density_A=1.0,
                          uvb_density_B=0.0, zhi_flash=-1.0, zheii_flash=-1.0, T_zhi=0.0, T_zheii=0.0, max_steps=2000, old_max_steps=3)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def tabulate(self, treecool, mean_rhob):
        out = np.zeros(capi.RATES_DOUBLES)
        rc = self.lib.hh_tabulate_rates(treecool.encode(), mean_rhob, out.ctypes.data)
        return rc, out

    def integrate_vec(self, rates, state, diag, lo, hi, a, dt, params=None):
        n = int(np.prod([h - l + 1 for l, h in zip(lo, hi)]))
        cs = np.zeros(n, dtype=capi.CELLSTAT_DTYPE)
        p = params or self.params()
        self.lib.hh_integrate_vec(rates.ctypes.data, C.byref(p), C.byref(capi.fab_of_numpy(state, (0, 0, 0))),
                                  C.byref(capi.fab_of_numpy(diag, (0, 0, 0))), capi.make_box(lo, hi), a, dt, cs.ctypes.data)
        return cs

    def integrate_struct(self, rates, d, lo, hi, sdc_iter=0, params=None):
        n = int(np.prod([h - l + 1 for l, h in zip(lo, hi)]))
        cs = np.zeros(n, dtype=capi.CELLSTAT_DTYPE)
        p = params or self.params()
        fabs = [C.byref(capi.fab_of_numpy(d[k], (0, 0, 0))) for k in ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")]
        self.lib.hh_integrate_struct(rates.ctypes.data, C.byref(p), *fabs, capi.make_box(lo, hi), d["a"], d["a_end"], d["dt"], sdc_iter,
                                     cs.ctypes.data)
        return cs


def stats_equal(cs, port_stats):
    return all(np.array_equal(cs[f], port_stats[:, i]) for i, f in enumerate(capi.CELLSTAT_FIELDS))


def cons_inputs(n, seed, nlow=40):
    """a periodic n^3 state for the "conservative" variant of enforce_minimum_density: lognormal density with `nlow` cells pushed below
    small_dens -- isolated ones, adjacent pairs, cells on the box faces and one cell that one iteration cannot fill -- momenta and
    energies random.  Returns (s_new (6, n, n, n), small_dens)"""
    rng = np.random.default_rng(seed)
    s = np.empty((6, n, n, n))
    s[0] = np.exp(rng.normal(0.0, 0.7, (n, n, n)))
    small = 0.05
    s[0] = np.maximum(s[0], 1.2 * small)
    for c in (1, 2, 3):
        s[c] = s[0] * rng.normal(0.0, 1.0, (n, n, n))
    s[5] = s[0] * np.exp(rng.normal(0.0, 0.5, (n, n, n)))
    s[4] = s[5] + 0.5 * (s[1] ** 2 + s[2] ** 2 + s[3] ** 2) / s[0]
    idx = rng.integers(0, n, (nlow, 3))
    idx[:6, 0] = 0; idx[6:10, 1] = n - 1; idx[10:12, 2] = 0            # on the faces of the box: filled through the ghost cells
    for q, (k, j, i) in enumerate(idx):
        s[0, k, j, i] = small * rng.uniform(-0.5, 0.99)                # below small_dens (some negative)
        if q % 5 == 0:
            s[0, k, j, (i + 1) % n] = small * rng.uniform(0.1, 0.9)    # an adjacent pair
    k, j, i = n // 2, n // 2, n // 2
    s[0, k - 1:k + 2, j - 1:j + 2, i - 1:i + 2] = small * 2.0          # a cell whose neighbours cannot cover its need in one iteration
    s[0, k, j, i] = -0.5 * small                                       # (each gives a sixth of its excess over 1.01 small_dens): two iterations
    return s, small


def fill_border(valid, ng):
    """periodic FillPatch of one box that covers the whole domain: (ncomp, n, n, n) -> (ncomp, n + 2 ng, ...)"""
    return np.ascontiguousarray(np.pad(valid, ((0, 0), (ng, ng), (ng, ng), (ng, ng)), mode="wrap"))
