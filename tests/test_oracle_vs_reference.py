"""The oracle's C restatement against the REFERENCE ITSELF (oracle/_ref: Nyx HeatCool sources + SUNDIALS CVODE compiled from
/root/reference by oracle/Makefile), bit for bit.  Skipped where oracle/_ref is not built (then tests/test_oracle_golden.py,
which holds outputs of the same reference build, still pins the oracle)."""
import numpy as np
import pytest

from nyx_b200 import synth
from tests import util


def test_tables_and_uvb(reference, port):
    assert np.array_equal(reference.rates(), port.rates())
    from oracle import pyref
    for z in (0.0, 1.7, 2.999, 3.0, 6.0, 14.99, 20.0):
        a, b = np.zeros(6), np.zeros(6)
        reference.lib.nyxref_interp_to_this_z(float(z), a.ctypes.data_as(pyref._dp))
        port.lib.hco_interp_to_this_z(port.rp, float(z), b.ctypes.data_as(pyref._dp))
        assert np.array_equal(a, b), z


def test_ion_n_and_eos_probes(reference, port):
    rng = np.random.default_rng(3)
    for _ in range(2000):
        U = 10.0 ** rng.uniform(9.0, 16.0)
        nh = 10.0 ** rng.uniform(-8.0, -1.0)
        ne = 10.0 ** rng.uniform(-6.0, 0.1)
        z = rng.choice([2.0, 3.0, 6.0, 9.0, 16.0])
        JH, JHe = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        assert np.array_equal(reference.ion_n(JH, JHe, U, nh, ne, 2.0 / 3.0, 0.76, z), port.ion_n(JH, JHe, U, nh, ne, 2.0 / 3.0, 0.76, z))
    for _ in range(500):
        R = synth.mean_rhob() * np.exp(rng.normal(0, 2.0))
        e = synth.e_from_T(10.0 ** rng.uniform(0.5, 9.5))
        a = 1.0 / (1.0 + rng.uniform(1.5, 8.0))
        assert reference.eos_T_given_Re(1, 1, float(R), float(e), float(a), 2.0 / 3.0, 0.76) == port.eos_T_given_Re(1, 1, float(R), float(e), float(a), 2.0 / 3.0, 0.76)


@pytest.mark.parametrize("z,n,seed", [(3.0, 10, 301), (2.0, 10, 302), (6.0, 10, 303)])
def test_vec_percell_mode_bitwise(reference, port, z, n, seed):
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s1, d1 = state.copy(), diag.copy()
    reference.set("nyx.sundials_tile_size", "1 1 1")
    reference.stats_reset()
    try:
        reference.integrate_state_vec([lo + hi], [s1], [d1], a, dt)
    finally:
        reference.set("nyx.sundials_tile_size", "1024000 8 8")
    pst = port.integrate_state_vec(state, diag, lo, hi, a, dt)
    assert np.array_equal(s1, state) and np.array_equal(d1, diag)
    assert np.array_equal(reference.stats(), pst[:, :8])


def test_grownvec_ghost_cells_integrated(reference, port):
    """integrate_state_grownvec (HC/integrate_state_vec_3d.cpp:367-396) integrates the ghost cells too."""
    z, n, ng = 3.0, 6, 2
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    m = n + 2 * ng
    state, diag = synth.make_fab((m, m, m), seed=311, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s1, d1 = state.copy(), diag.copy()
    reference.set("fabarray.mfiter_tile_size", "1 1 1")      # grownvec tiles with the AMReX default tile size (:381), not sundials_tile_size
    try:
        reference.integrate_state_vec([lo + hi], [s1], [d1], a, dt, ng_state=ng, ng_diag=ng, grown=True)
    finally:
        reference.set("fabarray.mfiter_tile_size", "1024000 8 8")
    glo, ghi = (-ng,) * 3, (n - 1 + ng,) * 3
    orig = state.copy()
    port.integrate_state_vec(state, diag, glo, ghi, a, dt, fab_lo=glo)
    assert np.all(s1[5] != orig[5]) and np.all(state[5] != orig[5])           # every ghost cell was integrated by both
    # with tile size 1 the reference's EDGE tiles are grown (1+ng cells wide: one coupled CVODE instance), the interior tiles
    # are single cells: interior bit-identical, edge/ghost region within the 10 x rtol contract
    inner = (slice(ng + 1, m - ng - 1),) * 3
    assert np.array_equal(s1[5][inner], state[5][inner]) and np.array_equal(d1[0][inner], diag[0][inner])
    assert np.abs(s1[5] / state[5] - 1).max() < 1e-3 and np.abs(d1[0] / diag[0] - 1).max() < 1e-3


@pytest.mark.parametrize("z,seed,src,flash", [(2.0, 322, 0.05, "none"), (5.99, 324, 0.05, "hi_now"), (3.0, 325, 0.05, "heii_now")])
def test_struct_percell_mode_bitwise(reference, port, z, seed, src, flash):
    from tests.golden.make_golden import FLASH_KEYS
    n = 8
    d = util.sdc_inputs(z, n, seed, src)
    r = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in d.items()}
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    reference.set("nyx.sundials_tile_size", "1 1 1")
    for k, v in util.FLASH_CASES[flash].items():
        reference.set(FLASH_KEYS[k], v)
    reference.stats_reset()
    try:
        reference.integrate_state_struct([lo + hi], [r["s_old"]], [r["s_new"]], [r["diag"]], [r["hydro_src"]], [r["ir"]], [r["reset_src"]],
                                         d["a"], d["a_end"], d["dt"], 0)
    finally:
        reference.set("nyx.sundials_tile_size", "1024000 8 8")
        for k in util.FLASH_CASES[flash]:
            reference.unset(FLASH_KEYS[k])
    pst = port.integrate_state_struct(d["s_old"], d["s_new"], d["diag"], d["hydro_src"], d["reset_src"], d["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**util.FLASH_CASES[flash]))
    for k in ("s_old", "s_new", "diag", "ir"):
        assert np.array_equal(d[k], r[k]), k
    assert np.array_equal(reference.stats(), pst[:, :8])


@pytest.mark.parametrize("interp", [0, 1])
def test_reset_internal_energy_bitwise(reference, port, interp):
    """the port's reset_internal_e loop against the reference header Source/EOS/reset_internal_e.H."""
    n = 12
    state, diag, rs = util.eos_rows_inputs(n, 401)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s1, d1, r1 = state.copy(), diag.copy(), rs.copy()
    reference.reset_internal_energy(lo + hi, s1, d1, r1, 0.25, 1.0e-2, interp)
    port.reset_internal_energy(state, diag, rs, lo, hi, 1.0e-2, interp)
    assert np.array_equal(s1, state) and np.array_equal(d1, diag) and np.array_equal(r1, rs)
    assert not np.array_equal(state, util.eos_rows_inputs(n, 401)[0])        # something was reset


@pytest.mark.parametrize("max_temp_dt,large_temp", [(0, 1.0e9), (1, 3.0e6)])
def test_compute_new_temp_bitwise(reference, port, max_temp_dt, large_temp):
    """the port's compute_new_temp loop against the reference's EOS functions driven as Nyx.cpp:2473-2519 drives them."""
    n, z = 12, 3.0
    a = 1.0 / (1.0 + z)
    state, diag, _ = util.eos_rows_inputs(n, 402, z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s1, d1 = state.copy(), diag.copy()
    reference.compute_new_temp(lo + hi, s1, d1, a, 1.0e-2, large_temp, max_temp_dt)
    port.compute_new_temp(state, diag, lo, hi, a, 1.0e-2, large_temp, max_temp_dt)
    assert np.array_equal(s1, state) and np.array_equal(d1, diag)
    assert (diag[0] == 1.0e-2).any()                                           # the rho e <= 0 branch ran
    if max_temp_dt:
        assert (diag[0] == large_temp).any()                                   # and the clipping branch


def test_inhomogeneous_reionization_bitwise(reference, port):
    """nyx.inhomo_reion = 1: per-cell z_HI in diag component 2 decides the UVB switch and the instantaneous heating."""
    from tests.golden.make_golden import FLASH_KEYS
    n, z = 8, 5.5
    d = util.inhomo_inputs(z, n, 331)
    r = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in d.items()}
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    reference.set("nyx.sundials_tile_size", "1 1 1")
    reference.set("nyx.inhomo_reion", 1)
    for k in ("zhi_flash", "T_zhi"):
        reference.set(FLASH_KEYS[k], d["kw"][k])
    reference.stats_reset()
    try:
        reference.integrate_state_struct([lo + hi], [r["s_old"]], [r["s_new"]], [r["diag"]], [r["hydro_src"]], [r["ir"]], [r["reset_src"]],
                                         d["a"], d["a_end"], d["dt"], 0)
    finally:
        reference.set("nyx.sundials_tile_size", "1024000 8 8")
        reference.set("nyx.inhomo_reion", 0)
        reference.unset("nyx.inhomo_reion")
        for k in ("zhi_flash", "T_zhi"):
            reference.unset(FLASH_KEYS[k])
    pst = port.integrate_state_struct(d["s_old"], d["s_new"], d["diag"], d["hydro_src"], d["reset_src"], d["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**d["kw"]))
    for k in ("s_old", "s_new", "diag", "ir"):
        assert np.array_equal(d[k], r[k]), k
    assert np.array_equal(reference.stats(), pst[:, :8])
    assert not np.array_equal(d["s_new"][5], util.inhomo_inputs(z, n, 331)["s_new"][5])


@pytest.mark.parametrize("low", [0, 7])
def test_update_state_with_sources_bitwise(reference, port, low):
    """the reference's own Nyx_update_state_with_sources.cpp (+ floor_density) over a ragged three-box level with the
    production ghost widths, against the port: every component of S_new and hydro_src bit for bit, with (low > 0) and without cells
    below small_dens."""
    import copy
    d = util.sources_inputs(seed=811 + low, low_density_cells=low)
    r = copy.deepcopy(d)
    reference.update_state_with_sources(d["boxes"], r["s_old"], r["s_new"], r["ext_src"], r["hydro_src"], r["grav"], r["reset_src"],
                                        d["dt"], d["a_old"], d["a_new"], d["small_dens"], d["small_temp"], ng=d["ng"])
    p = copy.deepcopy(d)
    m = port.update_state_with_sources(d["boxes"], p["s_old"], p["s_new"], p["ext_src"], p["hydro_src"], p["grav"], d["dt"], d["a_old"],
                                       d["a_new"], d["small_dens"], d["small_temp"], ng=d["ng"][:5])
    assert (m < d["small_dens"]) == (low > 0)
    n_floor = 0
    for bi in range(len(d["boxes"])):
        v = (slice(None), slice(1, -1), slice(1, -1), slice(1, -1))       # valid region of S_new (one ghost cell)
        assert np.array_equal(r["s_new"][bi][v], p["s_new"][bi][v])
        assert np.array_equal(r["s_new"][bi], p["s_new"][bi])             # ghost cells untouched on both sides
        assert np.array_equal(r["hydro_src"][bi], p["hydro_src"][bi])
        assert np.array_equal(r["s_old"][bi], d["s_old"][bi]) and np.array_equal(p["s_old"][bi], d["s_old"][bi])
        n_floor += int((p["s_new"][bi][0][v[1:]] == d["small_dens"]).sum())
        if low:
            assert not np.array_equal(p["hydro_src"][bi][0], d["hydro_src"][bi][0])     # reset in every cell
        else:
            assert np.array_equal(p["hydro_src"][bi][0], d["hydro_src"][bi][0])
    assert (n_floor >= low) if low else n_floor == 0     # about 3/4 of the 3 x low prepared cells end below small_dens


@pytest.mark.parametrize("n,seed,ng_new", [(12, 701, 1), (9, 702, 0)])
def test_enforce_min_density_conservative_iterations_bitwise(reference, port, n, seed, ng_new):
    """nyx.enforce_min_density_type = "conservative": the port's restatement of one iteration of Nyx::enforce_minimum_density_cons against the
    reference's own per-cell functions (compute_mu_for_enforce_min / create_update_for_minimum through oracle/_ref), iterated with a periodic
    FillPatch until the density is enforced: every iteration bit for bit, total mass conserved"""
    s0, small = util.cons_inputs(n, seed)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    pad = lambda a, g: util.fill_border(a, g) if g else a.copy()   # noqa: E731
    sn_ref, sn_port = pad(s0, ng_new), pad(s0, ng_new)
    inner = (slice(None),) + (slice(ng_new, ng_new + n),) * 3
    rs_ref, rs_port = np.full((1, n, n, n), -3.0), np.full((1, n, n, n), -3.0)
    mass0 = s0[0].sum()
    m = s0[0].min()
    it = 0
    while m < small and it < 10:
        sb = util.fill_border(sn_port[inner], 2)
        m_ref = reference.enforce_min_cons_iter(sb.copy(), sn_ref, rs_ref, lo, hi, small, ng_new=ng_new)
        m, bad = port.enforce_min_cons_iter(sb, sn_port, rs_port, lo, hi, small, ng_new=ng_new)
        assert bad == 0
        assert m == m_ref and np.array_equal(sn_ref, sn_port) and np.array_equal(rs_ref, rs_port), it
        it += 1
    assert 2 <= it < 10 and m >= small                # the cluster needs more than one iteration
    assert abs(sn_port[inner][0].sum() / mass0 - 1) < 1e-12       # density only moves between cells
    assert (rs_port != -3.0).all()
