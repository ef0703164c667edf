"""GPU parity tests: the CUDA path, called through the C-ABI (include/nyx_hc.h), against the oracle on the same
seeded inputs.  Tolerance contract (BASELINE.json north_star): e and T within 10 x rtol = 1e-3 of the
reference; failed-cell count, per-cell substep count and ne matched exactly.  The oracle is the reference run
one CVODE instance per cell (pinned bitwise in test_oracle_vs_reference.py); libdevice log10/pow differ from
glibc in the last bit, so "exactly" is asserted as: identical failure flags and >= 99.9 % of cells with identical
(nst, netf, nfe, nni, ncfn, nsetups, nfeLS).  Cells with identical counters followed the same step sequence and must
agree to 1e-5 in e, T and ne (measured <= 1e-6: a last-bit difference can flip the |dne| < 1e-6 exit of the inner
ne Newton solve, which moves ne by up to ~1e-7; the median difference is 0, ~50 % of cells are bitwise equal);
the rare cell where a last-bit difference flips a step-size/convergence decision takes
a different (equally valid) step sequence and is held to the 10 x rtol contract only.
"""
import ctypes as C

import numpy as np
import pytest

from nyx_b200 import capi, synth
from tests import util

pytestmark = pytest.mark.gpu

E_T_TOL = 1e-3        # the contract: 10 x rtol
E_T_TIGHT = 1e-5      # cells that follow the same step sequence as the oracle (measured: <= 1e-6, median 0: ~50 % bitwise)
EXACT_FRACTION = 0.999      # struct (SDC) cases, 13824 cells each: measured 0.99928 .. 1.0 (profiles/r2_s12_pytest.log)
EXACT_FRACTION_VEC = 0.9999  # Strang cases: measured 1.0 in every case (3 cells of 32768 allowed)


def _torch():
    import torch
    return torch


def _cell_stats_buffer(n):
    torch = _torch()
    return torch.zeros(n * 8, dtype=torch.int32, device="cuda")


def _cs_to_numpy(buf):
    return buf.cpu().numpy().view(capi.CELLSTAT_DTYPE)


def _compare_counts(cs, pst, what, chaotic_cells=0, need=None):
    """chaotic_cells: how many cells may differ in WHETHER they fail.  0 for every physical input.  The one stress case with
    20 % random energy sources drives a handful of cells into e <= 0, where the reference's DBL_MIN clamp (f_rhs_struct.H:482)
    makes the RHS discontinuous and the integrator thrashes for ~2000 steps: which of those cells hits max_steps first is
    decided by last-bit differences of log10/pow (libdevice vs glibc), i.e. it also differs between two builds of the reference."""
    bad = np.flatnonzero(cs["flag"] != pst[:, 7])
    fail_diff = np.flatnonzero((cs["flag"] < 0) != (pst[:, 7] < 0))
    assert len(fail_diff) <= chaotic_cells, f"{what}: failed cells differ at {fail_diff[:8]}: gpu {cs['flag'][fail_diff[:8]]} oracle {pst[fail_diff[:8], 7]}"
    assert abs(int((cs["flag"] < 0).sum()) - int((pst[:, 7] < 0).sum())) <= chaotic_cells
    assert len(bad) <= max(1, chaotic_cells), f"{what}: CVODE flags differ in {len(bad)} cells: gpu {cs['flag'][bad[:8]]} oracle {pst[bad[:8], 7]}"
    same = np.ones(len(cs), dtype=bool)
    for i, f in enumerate(capi.CELLSTAT_FIELDS[:7]):
        same &= cs[f] == pst[:, i]
    frac = same.mean()
    # in the stress case the thrashing cells (dozens of error-test failures each) amplify every last-bit difference -- e.g. the
    # device takes x^(1/3) with cbrt(), glibc with pow(x, 0.333..) -- into different counters: 1 % of its cells
    need = need or (0.99 if chaotic_cells else EXACT_FRACTION)      # stress case measured: 0.9919
    print(f"\n[parity] {what}: identical counters {int(same.sum())}/{len(same)} ({frac:.6f}), flags differ in {len(bad)} cells, failed gpu {int((cs['flag'] < 0).sum())} oracle {int((pst[:, 7] < 0).sum())}")
    assert frac >= need, f"{what}: only {frac:.5f} of cells have identical counters"
    return same


def _rel(x, y):
    return np.abs(x / y - 1)


def _weighted_err(rhoe, rhoe_ref, rho, rhoe0, rho0, rtol=1e-4, atol_factor=1e-4):
    """|e - e_ref| in units of the integrator's own error weight rtol*|e| + atol, atol = atol_factor * e(t0)
    (cvEwtSetSV, cvode.c:4413-4441, with Nyx's abstol vector, HC/integrate_state_vec_3d.cpp:254-257).  For a cell that keeps its
    energy scale this is the relative difference / 2e-4; for a cell that cools by orders of magnitude the absolute tolerance is
    what both integrations were asked to meet."""
    e, e_ref, e0 = rhoe / rho, rhoe_ref / rho, rhoe0 / rho0
    return np.abs(e - e_ref) / (rtol * np.abs(e_ref) + atol_factor * np.abs(e0))


def test_fast_log10_on_device(hc_lib):
    """The RHS fast path's table-driven log10 on the GPU: within 1 ulp of numpy's over the temperatures the tables cover (it is
    0.504-ulp accurate by construction, tests/test_host_logic.py), and everything that is not a positive normal number is flagged."""
    rng = np.random.default_rng(3)
    x = np.concatenate([10.0 ** rng.uniform(0.0, 10.0, 400000), rng.uniform(0.9, 1.1, 20000), [1.0, 2.0, 10.0, 1e9, 1e-300, 1e300]])
    y, bad = hc_lib.selftest_log10(x)
    ref = np.log10(x)
    ulp = np.spacing(np.abs(ref))
    sel = np.abs(ref) > 0.3
    assert not bad.any()
    assert np.all(np.abs(y - ref)[sel] <= ulp[sel]) and np.mean(y[sel] == ref[sel]) > 0.99
    assert np.all(np.abs(y - ref)[~sel] <= 2.0 ** -52)          # |log10| small: absolute accuracy is what the table lookup needs
    _, bad = hc_lib.selftest_log10(np.array([0.0, -1.0, np.nan, np.inf, 5e-324, 2e-308]))
    assert bad.tolist() == [1, 1, 1, 1, 1, 1]


def test_div_delta_t_on_device(hc_lib):
    """the table-position quotient log10(T) / DELTA_T of the RHS fast path (three FP64 instructions, constant reciprocal + one FMA residual)
    equals IEEE division bit for bit: random arguments over the table range, and every table node +- 3 ulp"""
    rng = np.random.default_rng(5)
    d = 9.0 / 2000
    nodes = np.arange(1, 2001) * d
    near = [nodes]
    for k in range(3):
        near += [np.nextafter(near[-1] if k else nodes, 10.0)]
    lo = nodes
    for k in range(3):
        lo = np.nextafter(lo, -1.0); near.append(lo)
    x = np.concatenate([rng.uniform(0.5 * d, 9.0, 2000000), rng.uniform(0.5 * d, 0.02, 100000)] + near)
    y = hc_lib.selftest_div_delta_t(x)
    assert np.array_equal(y, x / d)


@pytest.mark.parametrize("z,n,seed", [(3.0, 32, 11), (2.0, 32, 12), (6.0, 32, 13), (3.0, 7, 14)])
def test_vec_matches_oracle(hc_lib, port, z, n, seed):
    torch = _torch()
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
    csb = _cell_stats_buffer(n ** 3)
    st = hc_lib.integrate_vec_batch([capi.fab_of_torch(s_dev, lo)], [capi.fab_of_torch(d_dev, lo)], [capi.make_box(lo, hi)], a, dt,
                                    cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    s_ref, d_ref = state.copy(), diag.copy()
    pst = port.integrate_state_vec(s_ref, d_ref, lo, hi, a, dt)
    s_gpu, d_gpu = s_dev.cpu().numpy(), d_dev.cpu().numpy()
    for comp in (0, 1, 2, 3):
        assert np.array_equal(s_gpu[comp], state[comp])          # untouched components stay bit-identical
    cs = _cs_to_numpy(csb)
    same = _compare_counts(cs, pst, f"vec z={z}", need=EXACT_FRACTION_VEC).reshape(n, n, n)
    e_rel, E_rel, T_rel = _rel(s_gpu[5], s_ref[5]), _rel(s_gpu[4], s_ref[4]), _rel(d_gpu[0], d_ref[0])
    ne_abs = np.abs(d_gpu[1] - d_ref[1])
    assert max(e_rel.max(), E_rel.max(), T_rel.max()) < E_T_TOL
    assert max(e_rel[same].max(), E_rel[same].max(), T_rel[same].max()) < E_T_TIGHT, (e_rel[same].max(), T_rel[same].max())
    assert ne_abs[same].max() < E_T_TIGHT and ne_abs.max() < E_T_TOL
    assert st.n_cells == n ** 3
    assert st.n_failed == int((pst[:, 7] < 0).sum())
    assert st.sum_nst == int(cs["nst"].sum()) and st.max_nst == int(cs["nst"].max())
    assert st.sum_nfe == int(cs["nfe"].sum()) and st.sum_nfe_ls == int(cs["nfe_ls"].sum())


@pytest.mark.parametrize("z,seed,src,flash", [(3.0, 21, 0.0, "none"), (2.0, 22, 0.05, "none"), (6.0, 23, 0.2, "none"),
                                               (5.99, 24, 0.05, "hi_now"), (3.0, 25, 0.05, "heii_now"), (7.0, 26, 0.0, "before")])
def test_struct_matches_oracle(hc_lib, port, z, seed, src, flash):
    torch = _torch()
    n = 24
    d = util.sdc_inputs(z, n, seed, src)
    kw = util.FLASH_CASES[flash]
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in names}
    csb = _cell_stats_buffer(n ** 3)
    st = hc_lib.integrate_struct_batch(*[[capi.fab_of_torch(dev[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0,
                                       params=hc_lib.default_params(**kw), cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    ref = {k: d[k].copy() for k in names}
    pst = port.integrate_state_struct(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**kw))
    out = {k: dev[k].cpu().numpy() for k in names}
    assert np.array_equal(out["s_old"], d["s_old"])     # SDC path does not touch S_old
    cs = _cs_to_numpy(csb)
    chaotic = 4 if src >= 0.2 else 0
    same = _compare_counts(cs, pst, f"struct z={z} {flash}", chaotic_cells=chaotic).reshape(n, n, n)
    ok3 = ((pst[:, 7] == 0) & (cs["flag"] == 0)).reshape(n, n, n)   # cells that fail are compared on flags/counters only
    e_rel = _rel(out["s_new"][5], ref["s_new"][5])
    # I_R carries the same information as the energy update, rho_out * de * a_end^2 / (dt * a_half) (f_rhs_struct.H:307): its difference is
    # converted back to an energy difference and weighed with the integrator's own error weight, like werr
    ahalf = 0.5 * (d["a"] + d["a_end"])
    de_ir = np.abs(out["ir"][0] - ref["ir"][0]) * d["dt"] * ahalf / (d["a_end"] ** 2) / ref["s_new"][0]
    ir_w = de_ir / (1e-4 * np.abs(ref["s_new"][5] / ref["s_new"][0]) + 1e-4 * np.abs(d["s_old"][5] / d["s_old"][0]))
    # all cells, whatever step sequence they took: 10 x the integrator's tolerance (relative OR absolute part, as CVODE weighs them)
    werr = _weighted_err(out["s_new"][5], ref["s_new"][5], ref["s_new"][0], d["s_old"][5], d["s_old"][0])
    # (the +-20 % source stress case drives a few cells to e <= 0, where the reference's DBL_MIN clamp makes the RHS discontinuous
    # and the step sequence chaotic: those `chaotic` cells may take different step sequences and land further apart, but within 20 x)
    assert int((werr[ok3] >= 10.0).sum()) <= chaotic and werr[ok3].max() < 20.0, (int((werr[ok3] >= 10.0).sum()), werr[ok3].max())
    # and the plain relative 10 x rtol bound in all but isolated strongly-cooled cells (cells whose energy dropped by orders of magnitude
    # are held by the ABSOLUTE tolerance atol = 1e-4 e(t0) in both integrations -- werr above; the +-20 % source stress case has ~0.6 % of them)
    assert np.mean(e_rel[ok3] > E_T_TOL) < (1e-2 if src >= 0.2 else 1e-3), np.mean(e_rel[ok3] > E_T_TOL)
    m = same & ok3
    # same step sequence: 1/10 of the tolerance itself (<= 1e-5 relative for cells that keep their energy scale)
    # (stress case: in cells driven towards e = 0 the trajectory amplifies last-bit differences of the RHS; half the tolerance there)
    tight = 0.5 if src >= 0.2 else 0.1
    assert werr[m].max() < tight and ir_w[m].max() < tight, (werr[m].max(), e_rel[m].max(), ir_w[m].max())
    # diag holds T, ne of the LAST RHS evaluation (f_rhs_struct.H:290-291), not T(e_out).  That evaluation is often the
    # finite-difference probe of the diagonal Jacobian at y + 0.1*rl1*(h*f - zn[1]) (cvode_diag.c:364), whose offset is a
    # cancellation residue: a last-bit difference in f moves it by O(1), so this diagnostic T differs by up to ~1e-4
    # (1e-2 in cells that cooled to a few K) between any two libm's even when every integrator decision is identical.
    # (measured: up to 3.7 % of the cells -- the H I flash case -- beyond 1e-3).  Held to: median agreement at round-off level; and the T the caller's next compute_new_temp
    # derives from the updated state (tested below through the EOS kernel) to the tight bound.
    T_rel = _rel(out["diag"][0], ref["diag"][0])
    ne_abs = np.abs(out["diag"][1] - ref["diag"][1])
    print(f"[parity] struct z={z} {flash}: diag T of the last RHS evaluation beyond 1e-3 in {np.mean(T_rel[m] > E_T_TOL):.4f} of the same-sequence cells, Ne in {np.mean(ne_abs[m] > E_T_TOL):.4f}")
    assert np.median(T_rel[m]) < 1e-9 and np.mean(T_rel[m] > E_T_TOL) < 0.06, (np.median(T_rel[m]), np.mean(T_rel[m] > E_T_TOL))
    assert np.median(ne_abs[m]) < 1e-9 and np.mean(ne_abs[m] > E_T_TOL) < 2e-3      # measured: T <= 3.7 % (H I flash case), Ne <= 1e-4
    # T, ne recomputed from the updated state (what Nyx::compute_new_temp does right after, Nyx_advance.cpp:374-376)
    hc_lib.eos_T_given_Re(capi.fab_of_torch(dev["s_new"], lo), capi.fab_of_torch(dev["diag"], lo), capi.make_box(lo, hi), d["a_end"])
    torch.cuda.synchronize()
    port.eos_box(ref["s_new"], ref["diag"], lo, hi, d["a_end"])
    d_gpu = dev["diag"].cpu().numpy()
    # (cells the random sources cooled below 100 K are left out: there the inner ne Newton solve itself stalls at its
    # 15-iteration cap and its result depends on last bits, in the reference as much as here)
    okT = m & (ref["s_new"][5] > 0) & (ref["diag"][0] > 1.0e2)
    assert _rel(d_gpu[0], ref["diag"][0])[okT].max() < E_T_TIGHT and np.abs(d_gpu[1] - ref["diag"][1])[okT].max() < E_T_TIGHT
    assert st.n_cells == n ** 3 and abs(st.n_failed - int((pst[:, 7] < 0).sum())) <= chaotic


@pytest.mark.parametrize("z,seed,src,flash", [(3.0, 25, 0.05, "heii_now"), (5.99, 24, 0.05, "hi_now"), (6.0, 23, 0.2, "none")])
def test_save_react_matches_oracle(hc_lib, port, z, seed, src, flash):
    """the SAVE_REACT overload (Nyx.H:571-580; ode_eos_save_react_arrays f_rhs_struct.H:213-267) against the port's restatement of it -- itself
    equal bit for bit to the reference built with USE_SAVE_REACT (tests/test_hctest_fixture.py) -- with sources, instantaneous reionization
    heating (two EOS solves in the finalize step) and the +-20 % source stress case (floored cells)"""
    torch = _torch()
    n = 24
    d = util.sdc_inputs(z, n, seed, src)
    kw = util.FLASH_CASES[flash]
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in names}
    rdev = [torch.zeros((c, n, n, n), dtype=torch.float64, device="cuda") for c in (7, 7, 9)]
    csb = _cell_stats_buffer(n ** 3)
    st = hc_lib.integrate_struct_react_batch(*[[capi.fab_of_torch(dev[k], lo)] for k in names], *[[capi.fab_of_torch(r, lo)] for r in rdev],
                                             [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0, params=hc_lib.default_params(**kw),
                                             cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    ref = {k: d[k].copy() for k in names}
    ri, ro, rw = np.zeros((7, n, n, n)), np.zeros((7, n, n, n)), np.zeros((9, n, n, n))
    pst = port.integrate_state_struct_react(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], ri, ro, rw,
                                            lo, hi, d["a"], d["a_end"], d["dt"], 0, params=port.params(**kw))
    cs = _cs_to_numpy(csb)
    chaotic = 4 if src >= 0.2 else 0
    same = _compare_counts(cs, pst, f"react z={z} {flash}", chaotic_cells=chaotic).reshape(n, n, n)
    gi, go, gw = (r.cpu().numpy() for r in rdev)
    # functions of the inputs only: bit for bit
    for c in range(7):
        assert np.array_equal(gi[c], ri[c]), c
    for c in (1, 5, 6):
        assert np.array_equal(go[c], ro[c]), c
    # the counters are the cell's own (those of cell_stats)
    for c, f in ((0, "nst"), (1, "netf"), (2, "nfe"), (3, "nni"), (4, "ncfn"), (5, "nsetups"), (8, "nfe_ls")):
        assert np.array_equal(gw[c].ravel(), cs[f].astype(np.float64)), f
    assert not gw[6].any() and not gw[7].any()
    m = same & ((pst[:, 7] == 0) & (cs["flag"] == 0)).reshape(n, n, n)
    # CVODE's solution before the finalize step (may be negative / floored afterwards): in units of the integrator's own error weight
    wgt = 1e-4 * np.abs(ro[0]) + ri[4]
    # (stress case: the RAW solution of cells driven to e <= 0, which the finalize step floors afterwards: within the tolerance itself)
    assert (np.abs(go[0] - ro[0]) / wgt)[m].max() < (1.0 if src >= 0.2 else 0.1)
    # T, ne of the finalize step's LAST EOS solve (cells cooled below 100 K left out: the ne iteration stalls at its cap there, see above)
    okT = m & (ro[2] > 1.0e2) & (np.abs(go[0] / ro[0] - 1) < 1e-7)
    assert okT.mean() > 0.3
    assert _rel(go[2], ro[2])[okT].max() < E_T_TIGHT and np.abs(go[3] - ro[3])[okT].max() < E_T_TIGHT
    # the estimated local error is a cancellation residue (acor): same sign and size where it matters, i.e. where it is not round-off of the weight
    big = m & (np.abs(ro[4]) > 1e-3 * wgt)
    if big.any():
        assert np.median(np.abs(go[4] / ro[4] - 1)[big]) < 1e-6
    assert (np.abs(go[4] - ro[4]) / wgt)[m].max() < (1.0 if src >= 0.2 else 0.1)
    # and the integration is the one of the plain entry point
    dev2 = {k: torch.from_numpy(d[k]).cuda() for k in names}
    st2 = hc_lib.integrate_struct_batch(*[[capi.fab_of_torch(dev2[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0,
                                        params=hc_lib.default_params(**kw))
    torch.cuda.synchronize()
    assert st2.as_dict() == st.as_dict()
    for k in ("s_new", "diag", "ir"):
        assert torch.equal(dev[k], dev2[k]), k


def test_grown_tiles_batch(hc_lib, port):
    """Two boxes with 4 ghost cells, tiles = grown boxes (integrate_state_grownvec, HC/integrate_state_vec_3d.cpp:367-396):
    ghost cells are integrated too, and state/diag FABs have different ghost widths."""
    torch = _torch()
    z, n, ng_s, ng_d = 3.0, 8, 4, 1
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    boxes = [((0, 0, 0), (n - 1, n - 1, n - 1)), ((n, 0, 0), (2 * n - 1, n - 1, n - 1))]
    fabs_s, fabs_d, tiles, keep, refs = [], [], [], [], []
    for b, (lo, hi) in enumerate(boxes):
        m = n + 2 * ng_s
        state, diag_big = synth.make_fab((m, m, m), seed=40 + b, z=z)
        cut = ng_s - ng_d
        diag = np.ascontiguousarray(diag_big[:, cut:m - cut, cut:m - cut, cut:m - cut])
        slo = tuple(l - ng_s for l in lo)
        dlo = tuple(l - ng_d for l in lo)
        tlo, thi = tuple(l - ng_d for l in lo), tuple(h + ng_d for h in hi)   # tile: valid box grown by 1 (inside both FABs)
        s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
        keep += [s_dev, d_dev]
        fabs_s.append(capi.fab_of_torch(s_dev, slo)); fabs_d.append(capi.fab_of_torch(d_dev, dlo)); tiles.append(capi.make_box(tlo, thi))
        s_ref, d_ref = state.copy(), diag.copy()
        port.integrate_state_vec(s_ref, d_ref, tlo, thi, a, dt, fab_lo=slo, diag_lo=dlo)
        refs.append((state, diag, s_ref, d_ref, s_dev, d_dev))
    st = hc_lib.integrate_vec_batch(fabs_s, fabs_d, tiles, a, dt)
    torch.cuda.synchronize()
    assert st.n_cells == 2 * (n + 2 * ng_d) ** 3
    for state, diag, s_ref, d_ref, s_dev, d_dev in refs:
        s_gpu, d_gpu = s_dev.cpu().numpy(), d_dev.cpu().numpy()
        touched = s_ref[5] != state[5]
        assert np.array_equal(s_gpu[5] != state[5], touched)        # exactly the tile's cells were updated
        assert np.abs(s_gpu[5] / s_ref[5] - 1).max() < E_T_TOL and np.median(np.abs(s_gpu[5] / s_ref[5] - 1)) < 1e-12
        assert np.abs(d_gpu[0] / d_ref[0] - 1).max() < E_T_TOL


def test_host_buffer_api(hc_lib, port):
    z, n = 3.0, 16
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=51, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s_ref, d_ref = state.copy(), diag.copy()
    port.integrate_state_vec(s_ref, d_ref, lo, hi, a, dt)
    st = hc_lib.integrate_vec_host([capi.fab_of_numpy(state, lo)], [capi.fab_of_numpy(diag, lo)], [capi.make_box(lo, hi)], a, dt)
    assert st.n_cells == n ** 3
    assert np.abs(state[5] / s_ref[5] - 1).max() < E_T_TOL and np.abs(diag[0] / d_ref[0] - 1).max() < E_T_TOL
    assert np.median(np.abs(state[5] / s_ref[5] - 1)) < 1e-12
    d = util.sdc_inputs(2.0, n, 52, 0.05)
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    ref = {k: d[k].copy() for k in names}
    port.integrate_state_struct(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], lo, hi,
                                d["a"], d["a_end"], d["dt"], 0)
    hc_lib.integrate_struct_host(*[[capi.fab_of_numpy(d[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0)
    assert np.abs(d["s_new"][5] / ref["s_new"][5] - 1).max() < E_T_TOL
    assert np.median(np.abs(d["ir"][0] - ref["ir"][0])) / np.abs(ref["ir"][0]).max() < 1e-9


def test_host_buffer_api_stages_only_the_tiles_of_ghosted_fabs(hc_lib):
    """host-buffer entry points on FABs with wide ghost regions (12^3 tiles inside 24^3 FABs: the tiles are 12.5 % of the FAB): only the
    bounding box of the tiles travels, so NaN-poisoned ghost cells of the inputs never reach the device and every ghost cell of every
    host FAB keeps its bytes; the valid cells equal those of the device-resident call, bit for bit.  Two tiles that share one FAB (sub-box =
    their union) and a tile in a FAB of its own, Strang and SDC."""
    torch = _torch()
    z, n, g = 3.0, 12, 6
    m = n + 2 * g
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    inner = (slice(None),) + (slice(g, g + n),) * 3

    def poisoned(arr):
        out = np.full_like(arr, np.nan)
        out[inner] = arr[inner]
        return out
    # ---- Strang: FAB A holds two tiles (split in x), FAB B one
    sa, da = synth.make_fab((m, m, m), seed=61, z=z)
    sb, db = synth.make_fab((m, m, m), seed=62, z=z)
    host = [poisoned(x) for x in (sa, da, sb, db)]
    before = [x.copy() for x in host]
    lo = (-g, -g, -g)
    tiles = [capi.make_box((0, 0, 0), (5, n - 1, n - 1)), capi.make_box((6, 0, 0), (n - 1, n - 1, n - 1)), capi.make_box((0, 0, 0), (n - 1, n - 1, n - 1))]
    fs = [capi.fab_of_numpy(host[0], lo), capi.fab_of_numpy(host[0], lo), capi.fab_of_numpy(host[2], lo)]
    fd = [capi.fab_of_numpy(host[1], lo), capi.fab_of_numpy(host[1], lo), capi.fab_of_numpy(host[3], lo)]
    st = hc_lib.integrate_vec_host(fs, fd, tiles, a, dt)
    assert st.n_cells == 2 * n ** 3 and st.n_failed == 0
    dev = [torch.from_numpy(x).cuda() for x in (sa, da, sb, db)]
    ds = [capi.fab_of_torch(dev[0], lo), capi.fab_of_torch(dev[0], lo), capi.fab_of_torch(dev[2], lo)]
    dd = [capi.fab_of_torch(dev[1], lo), capi.fab_of_torch(dev[1], lo), capi.fab_of_torch(dev[3], lo)]
    hc_lib.integrate_vec_batch(ds, dd, tiles, a, dt)
    torch.cuda.synchronize()
    ghost = np.ones((m, m, m), dtype=bool); ghost[g:g + n, g:g + n, g:g + n] = False
    for h, b0, d in zip(host, before, dev):
        assert h[inner].tobytes() == d.cpu().numpy()[inner].tobytes()
        assert h[:, ghost].tobytes() == b0[:, ghost].tobytes()              # NaNs and all: untouched
    assert not np.isnan(host[0][inner]).any() and (host[0][5][g:g + n, g:g + n, g:g + n] != before[0][5][g:g + n, g:g + n, g:g + n]).all()
    # ---- SDC, one tile per FAB
    d0 = util.sdc_inputs(z, m, 63, 0.05)
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    hh = {k: poisoned(d0[k]) for k in names}
    b1 = {k: v.copy() for k, v in hh.items()}
    gg = {k: torch.from_numpy(d0[k]).cuda() for k in names}
    t1 = [capi.make_box((0, 0, 0), (n - 1, n - 1, n - 1))]
    hc_lib.integrate_struct_host(*[[capi.fab_of_numpy(hh[k], lo)] for k in names], t1, d0["a"], d0["a_end"], d0["dt"], 0)
    hc_lib.integrate_struct_batch(*[[capi.fab_of_torch(gg[k], lo)] for k in names], t1, d0["a"], d0["a_end"], d0["dt"], 0)
    torch.cuda.synchronize()
    for k in names:
        assert hh[k][inner].tobytes() == gg[k].cpu().numpy()[inner].tobytes(), k
        assert hh[k][:, ghost].tobytes() == b1[k][:, ghost].tobytes(), k


def test_eos_kernel(hc_lib, port):
    torch = _torch()
    z, n = 3.0, 16
    a = 1.0 / (1.0 + z)
    state, diag = synth.make_fab((n, n, n), seed=61, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
    st = hc_lib.eos_T_given_Re(capi.fab_of_torch(s_dev, lo), capi.fab_of_torch(d_dev, lo), capi.make_box(lo, hi), a)
    torch.cuda.synchronize()
    d_ref = diag.copy()
    port.eos_box(state, d_ref, lo, hi, a)
    d_gpu = d_dev.cpu().numpy()
    assert st.n_cells == n ** 3
    assert np.abs(d_gpu[0] / d_ref[0] - 1).max() < E_T_TIGHT and np.abs(d_gpu[1] - d_ref[1]).max() < E_T_TIGHT
    assert np.median(np.abs(d_gpu[0] / d_ref[0] - 1)) < 1e-12


def test_failed_cells_are_counted_not_fatal(hc_lib, port):
    """max_steps = 3 makes every stiff cell return CV_TOO_MUCH_WORK; like the reference (which ignores the flag,
    HC/integrate_state_vec_3d.cpp:284) the call succeeds and the failures are only counted."""
    torch = _torch()
    z, n = 3.0, 12
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=71, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
    st = hc_lib.integrate_vec_batch([capi.fab_of_torch(s_dev, lo)], [capi.fab_of_torch(d_dev, lo)], [capi.make_box(lo, hi)], a, dt,
                                    params=hc_lib.default_params(max_steps=3))
    torch.cuda.synchronize()
    s_ref, d_ref = state.copy(), diag.copy()
    pst = port.integrate_state_vec(s_ref, d_ref, lo, hi, a, dt, params=port.params(max_steps=3))
    assert st.n_failed == int((pst[:, 7] < 0).sum()) and st.n_failed > 0
    assert np.abs(s_dev.cpu().numpy()[5] / s_ref[5] - 1).max() < E_T_TOL


@pytest.mark.parametrize("max_temp_dt,large_temp", [(0, 1.0e9), (1, 3.0e6)])
def test_compute_new_temp_matches_oracle(hc_lib, port, max_temp_dt, large_temp):
    """Nyx::compute_new_temp's cell loop over two boxes (hc_compute_new_temp_batch) against the oracle."""
    torch = _torch()
    n, z = 20, 3.0
    a = 1.0 / (1.0 + z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    outs, refs, sfab, dfab, keep = [], [], [], [], []
    for b in range(2):
        state, diag, _ = util.eos_rows_inputs(n, 500 + b, z)
        s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
        keep += [s_dev, d_dev]
        sfab.append(capi.fab_of_torch(s_dev, lo)); dfab.append(capi.fab_of_torch(d_dev, lo))
        port.compute_new_temp(state, diag, lo, hi, a, 1.0e-2, large_temp, max_temp_dt)
        refs.append((state, diag)); outs.append((s_dev, d_dev))
    st = hc_lib.compute_new_temp_batch(sfab, dfab, [capi.make_box(lo, hi)] * 2, a, 1.0e-2, large_temp, max_temp_dt)
    torch.cuda.synchronize()
    n_small = n_large = 0
    for (s_dev, d_dev), (s_ref, d_ref) in zip(outs, refs):
        s_gpu, d_gpu = s_dev.cpu().numpy(), d_dev.cpu().numpy()
        for comp in (0, 1, 2, 3):
            assert np.array_equal(s_gpu[comp], s_ref[comp])
        small, large = d_ref[0] == 1.0e-2, (d_ref[0] == large_temp) & bool(max_temp_dt)
        n_small += int(small.sum()); n_large += int(large.sum())
        assert np.array_equal(d_gpu[0][small | large], d_ref[0][small | large])
        # cells rewritten from (T, ne): plain arithmetic, identical up to the ne the EOS returned
        assert np.array_equal(s_gpu[5][small], s_ref[5][small]) and np.array_equal(s_gpu[4][small], s_ref[4][small])
        assert np.all(np.abs(s_gpu[5] - s_ref[5]) <= 1e-9 * np.abs(s_ref[5])) and np.all(np.abs(s_gpu[4] - s_ref[4]) <= 1e-9 * np.abs(s_ref[4]))
        ok = ~small & (d_ref[0] > 1.0e2)      # (below 100 K the inner ne Newton solve itself depends on last bits, see the SDC test)
        assert _rel(d_gpu[0], d_ref[0])[ok].max() < E_T_TIGHT and np.abs(d_gpu[1] - d_ref[1])[ok].max() < E_T_TIGHT
    assert st.n_cells == 2 * n ** 3 and st.n_floor == n_small and st.n_failed == n_large and n_small > 0


@pytest.mark.parametrize("interp", [0, 1])
def test_reset_internal_energy_matches_oracle_bitwise(hc_lib, port, interp):
    """Nyx::reset_internal_energy's cell loop (no transcendental: bit for bit), FABs with ghost cells."""
    torch = _torch()
    n, ng = 16, 2
    m = n + 2 * ng
    state, diag, rs = util.eos_rows_inputs(m, 510)
    flo, lo, hi = (-ng, -ng, -ng), (0, 0, 0), (n - 1, n - 1, n - 1)
    s_dev, d_dev, r_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda(), torch.from_numpy(rs).cuda()
    hc_lib.reset_internal_energy_batch([capi.fab_of_torch(s_dev, flo)], [capi.fab_of_torch(d_dev, flo)], [capi.fab_of_torch(r_dev, flo)],
                                       [capi.make_box(lo, hi)], 0.25, 1.0e-2, interp)
    torch.cuda.synchronize()
    from oracle import pyref
    p = port.params()
    sf, df, rf = pyref.fab_of(state, flo), pyref.fab_of(diag, flo), pyref.fab_of(rs, flo)
    l, h = port._box(lo, hi)
    import ctypes as C
    port.lib.hco_reset_internal_e_box.argtypes = [C.POINTER(pyref.HcoParams), C.POINTER(type(sf)), C.POINTER(type(sf)), C.POINTER(type(sf)), type(l), type(l),
                                                  C.c_double, C.c_int]
    port.lib.hco_reset_internal_e_box(C.byref(p), C.byref(sf), C.byref(df), C.byref(rf), l, h, 1.0e-2, interp)
    assert np.array_equal(s_dev.cpu().numpy(), state) and np.array_equal(r_dev.cpu().numpy(), rs) and np.array_equal(d_dev.cpu().numpy(), diag)


@pytest.mark.parametrize("n,z,path", [(128, 3.0, "vec"), (256, 2.0, "vec"), (128, 6.0, "struct")])
def test_full_size_properties(hc_lib, n, z, path):
    """BASELINE.json configs 2-5 at sizes the oracle cannot follow, through size-independent properties of the path:
    (a) determinism: two runs give bit-identical FABs and counters, whatever order the work queue hands the cells out in;
    (b) decomposition independence: every cell is integrated on its own, so the same field as ONE box, as 8 boxes or as 64 boxes
        (each box its own FAB) gives bit-identical cell values and identical summed counters -- this is what box sharding over
        GPUs relies on;
    (c) counters: all cells integrated, none failed, none floored, max_nst bounded, and the statistics add up over boxes."""
    torch = _torch()
    a, dt = 1.0 / (1.0 + z), synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=77, z=z)
    kw = dict(zhi_flash=6.0, T_zhi=2e4, zheii_flash=3.0, T_zheii=1.5e4) if path == "struct" else {}   # config 5: flash reionization at z = 6
    prm = hc_lib.default_params(**kw)

    def run(nsplit):
        m = n // nsplit
        keep, fabs, tiles = [], {k: [] for k in ("s", "d", "sn", "hs", "rs", "ir")}, []
        for kb in range(nsplit):
            for jb in range(nsplit):
                for ib in range(nsplit):
                    sl = (slice(None), slice(kb * m, (kb + 1) * m), slice(jb * m, (jb + 1) * m), slice(ib * m, (ib + 1) * m))
                    lo = (ib * m, jb * m, kb * m)
                    s = torch.from_numpy(np.ascontiguousarray(state[sl])).cuda()
                    d = torch.from_numpy(np.ascontiguousarray(diag[sl])).cuda()
                    ent = {"s": s, "d": d}
                    if path == "struct":
                        ent.update(sn=s.clone(), hs=torch.zeros_like(s), rs=torch.zeros((1, m, m, m), dtype=torch.float64, device="cuda"),
                                   ir=torch.zeros((1, m, m, m), dtype=torch.float64, device="cuda"))
                    keep.append(ent)
                    for k2, v in ent.items():
                        fabs[k2].append(capi.fab_of_torch(v, lo))
                    tiles.append(capi.make_box(lo, tuple(x + m - 1 for x in lo)))
        if path == "vec":
            st = hc_lib.integrate_vec_batch(fabs["s"], fabs["d"], tiles, a, 0.5 * dt, params=prm)
        else:
            st = hc_lib.integrate_struct_batch(fabs["s"], fabs["d"], fabs["sn"], fabs["hs"], fabs["rs"], fabs["ir"], tiles, a, synth.a_after(z, dt), dt, 0, params=prm)
        torch.cuda.synchronize()
        out_s = np.empty_like(state); out_d = np.empty_like(diag)
        b = 0
        for kb in range(nsplit):
            for jb in range(nsplit):
                for ib in range(nsplit):
                    sl = (slice(None), slice(kb * m, (kb + 1) * m), slice(jb * m, (jb + 1) * m), slice(ib * m, (ib + 1) * m))
                    out_s[sl] = keep[b]["sn" if path == "struct" else "s"].cpu().numpy()
                    out_d[sl] = keep[b]["d"].cpu().numpy()
                    b += 1
        return out_s, out_d, st.as_dict()

    s1, d1, st1 = run(1)
    s1b, d1b, st1b = run(1)
    assert np.array_equal(s1, s1b) and np.array_equal(d1, d1b) and st1 == st1b                     # (a)
    for nsplit in (2, 4):
        s2, d2, st2 = run(nsplit)
        assert np.array_equal(s1, s2) and np.array_equal(d1, d2), f"{nsplit}^3 boxes differ from one box"   # (b)
        assert st2 == st1
    assert st1["n_cells"] == n ** 3 and st1["n_failed"] == 0 and st1["n_floor"] == 0                # (c)
    assert 3 <= st1["max_nst"] <= 200 and st1["sum_nst"] >= 3 * n ** 3
    assert st1["sum_nfe"] + st1["sum_nfe_ls"] >= st1["sum_nni"] and st1["sum_attempts"] >= st1["sum_nst"]
    assert np.isfinite(s1).all() and np.isfinite(d1).all() and (d1[0] > 0).all() and (d1[1] >= 0).all() and (d1[1] <= 1.0 + 2.0 * 0.0789474 + 1e-9).all()
    for comp in (0, 1, 2, 3):
        assert np.array_equal(s1[comp], state[comp])


def test_ragged_and_empty_batches(hc_lib, port):
    """Edge cases of the batch interface: no tiles at all, an empty tile among the tiles (hi < lo: an empty MFIter tile), boxes of
    very different shapes in one launch (1 cell; a thin slab; rows longer than one work-queue chunk of 256 cells), and the
    per-cell statistics buffer laid out tile after tile."""
    torch = _torch()
    z = 3.0
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    st = hc_lib.integrate_vec_batch([], [], [], a, dt)
    assert st.n_cells == 0 and st.sum_nst == 0
    shapes = [(1, 1, 1), (5, 7, 3), (300, 2, 2), (33, 2, 9)]
    keep, fs, fd, tiles, refs = [], [], [], [], []
    for b, shp in enumerate(shapes):
        state, diag = synth.make_fab(shp, seed=80 + b, z=z)
        lo = (3 * b, -b, 7)
        hi = tuple(l + s - 1 for l, s in zip(lo, shp))
        s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
        keep += [s_dev, d_dev]
        fs.append(capi.fab_of_torch(s_dev, lo)); fd.append(capi.fab_of_torch(d_dev, lo)); tiles.append(capi.make_box(lo, hi))
        s_ref, d_ref = state.copy(), diag.copy()
        refs.append((s_ref, d_ref, port.integrate_state_vec(s_ref, d_ref, lo, hi, a, dt), s_dev, d_dev))
    # an empty tile in the middle (its FABs are those of box 1; it contributes no cells and no statistics rows)
    fs.insert(2, fs[1]); fd.insert(2, fd[1]); tiles.insert(2, capi.make_box((0, 0, 7), (-1, 5, 9)))
    ncell = sum(s[0] * s[1] * s[2] for s in shapes)
    csb = _cell_stats_buffer(ncell)
    st = hc_lib.integrate_vec_batch(fs, fd, tiles, a, dt, cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    assert st.n_cells == ncell and st.n_failed == 0
    cs = _cs_to_numpy(csb)
    off = 0
    for s_ref, d_ref, pst, s_dev, d_dev in refs:
        n = len(pst)
        same_nst = np.mean(cs["nst"][off:off + n] == pst[:, 0])
        print(f"[parity] ragged batch box of {n} cells: identical nst {same_nst:.6f}")
        assert same_nst >= EXACT_FRACTION_VEC or (n < 10000 and (1 - same_nst) * n <= 1)
        assert np.abs(s_dev.cpu().numpy()[5] / s_ref[5] - 1).max() < E_T_TOL and np.abs(d_dev.cpu().numpy()[0] / d_ref[0] - 1).max() < E_T_TOL
        off += n
    assert st.sum_nst == int(cs["nst"].sum())


def test_integrator_options_on_device(hc_lib, port):
    """nyx.use_sundials_constraint (CVodeSetConstraints y > 0) and nyx.use_typical_steps (CVodeSetMaxStep(dt / old_max_steps)) on the GPU."""
    torch = _torch()
    z, n = 2.0, 20
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    for kw in (dict(use_constraint=1), dict(use_typical_steps=1, old_max_steps=5), dict(rtol=1e-6, atol_factor=1e-6)):
        state, diag = synth.make_fab((n, n, n), seed=90, z=z)
        s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
        csb = _cell_stats_buffer(n ** 3)
        st = hc_lib.integrate_vec_batch([capi.fab_of_torch(s_dev, lo)], [capi.fab_of_torch(d_dev, lo)], [capi.make_box(lo, hi)], a, dt,
                                        params=hc_lib.default_params(**kw), cell_stats_ptr=csb.data_ptr())
        torch.cuda.synchronize()
        pst = port.integrate_state_vec(state, diag, lo, hi, a, dt, params=port.params(**kw))
        cs = _cs_to_numpy(csb)
        _compare_counts(cs, pst, f"options {kw}")
        tol = 10.0 * kw.get("rtol", 1e-4)
        assert np.abs(s_dev.cpu().numpy()[5] / state[5] - 1).max() < tol and np.abs(d_dev.cpu().numpy()[0] / diag[0] - 1).max() < tol
        if "old_max_steps" in kw:
            assert st.max_nst >= 5       # the step-size cap forces at least old_max_steps steps on every cell
            assert int(cs["nst"].min()) >= 5


def test_concurrent_calls_on_two_streams(hc_lib, port):
    """The C-ABI is re-entrant per (thread, stream) (the SDC caller may run its MFIter loop under OpenMP): two host threads,
    two streams, two different boxes at once; each result equals the one of a lone call bit for bit."""
    import threading
    torch = _torch()
    z, n = 3.0, 40
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    fields = [synth.make_fab((n, n, n), seed=95 + i, z=z) for i in range(2)]

    def run(i, stream, out):
        s_dev, d_dev = torch.from_numpy(fields[i][0]).cuda(), torch.from_numpy(fields[i][1]).cuda()
        torch.cuda.synchronize()
        st = hc_lib.integrate_vec_batch([capi.fab_of_torch(s_dev, lo)], [capi.fab_of_torch(d_dev, lo)], [capi.make_box(lo, hi)], a, dt,
                                        stream=stream.cuda_stream if stream is not None else None)
        torch.cuda.synchronize()
        out[i] = (s_dev.cpu().numpy(), d_dev.cpu().numpy(), st.as_dict())

    alone, together = {}, {}
    for i in range(2):
        run(i, None, alone)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    th = [threading.Thread(target=run, args=(i, streams[i], together)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in range(2):
        assert np.array_equal(alone[i][0], together[i][0]) and np.array_equal(alone[i][1], together[i][1]) and alone[i][2] == together[i][2]


def test_inhomogeneous_reionization_on_device(hc_lib, port):
    """nyx.inhomo_reion = 1: z_HI per cell from diag component 2; device pointers and the pipelined host entry point."""
    torch = _torch()
    n, z = 20, 5.5
    # (a pure reaction step: with random hydro sources a z ~ 5.5 box has ~0.5 % of cells whose forcing sits on the cooling equilibrium; there
    # the 1e-7 noise of the RHS -- the inner ne Newton exit -- defeats the finite-difference Jacobian, the integration thrashes through
    # dozens of convergence failures in the oracle and on the GPU alike, and the outcome depends on last bits: tests/diag/gpu_diag_inhomo.py)
    d = util.inhomo_inputs(z, n, 351, src_scale=0.0)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in names}
    csb = _cell_stats_buffer(n ** 3)
    st = hc_lib.integrate_struct_batch(*[[capi.fab_of_torch(dev[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0,
                                       params=hc_lib.default_params(**d["kw"]), cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    host = {k: d[k].copy() for k in names}
    hc_lib.integrate_struct_host(*[[capi.fab_of_numpy(host[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0,
                                 params=hc_lib.default_params(**d["kw"]))
    ref = {k: d[k].copy() for k in names}
    pst = port.integrate_state_struct(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**d["kw"]))
    out = {k: dev[k].cpu().numpy() for k in names}
    same = _compare_counts(_cs_to_numpy(csb), pst, "inhomo").reshape(n, n, n)
    werr = _weighted_err(out["s_new"][5], ref["s_new"][5], ref["s_new"][0], d["s_old"][5], d["s_old"][0])
    assert werr.max() < 10.0 and werr[same].max() < 0.1
    assert np.array_equal(out["diag"][2], d["diag"][2])                                   # z_HI is an input
    assert st.n_cells == n ** 3 and st.n_failed == int((pst[:, 7] < 0).sum())
    for k in ("s_new", "ir", "diag"):
        assert np.array_equal(host[k], out[k]), k                                         # host entry point == device entry point, bit for bit
    # the populations behave differently: cold cells reionized during the step were heated towards T_zHI = 2e4 K
    z_end = 1.0 / d["a_end"] - 1.0
    cold_during = (d["diag"][2] >= z_end) & (d["diag"][2] < z) & (d["diag"][0] < 5.0e3)
    assert cold_during.sum() > 50 and np.mean(out["s_new"][5][cold_during] > d["s_new"][5][cold_during]) > 0.95


# ---------------------------------------------------------------------------------------------- SDC source assembly
def _src_fabs(d, arrs, pinned_host=False):
    """per slot the HcFab list over arrs[slot][box] (torch device tensors or numpy arrays), each covering its box grown by ng[slot]"""
    fabs = {}
    for slot, g in zip(("s_old", "s_new", "ext_src", "hydro_src", "grav", "reset_src"), d["ng"]):
        mk = capi.fab_of_numpy if pinned_host else capi.fab_of_torch
        fabs[slot] = [mk(arrs[slot][bi], tuple(x - g for x in bx[:3])) for bi, bx in enumerate(d["boxes"])]
    tiles = [capi.make_box(bx[:3], bx[3:]) for bx in d["boxes"]]
    return fabs, tiles


def _src_oracle(port, d):
    import copy
    p = copy.deepcopy({k: d[k] for k in ("s_old", "s_new", "ext_src", "hydro_src", "grav")})
    m = port.update_state_with_sources(d["boxes"], p["s_old"], p["s_new"], p["ext_src"], p["hydro_src"], p["grav"], d["dt"], d["a_old"],
                                       d["a_new"], d["small_dens"], d["small_temp"], ng=d["ng"][:5])
    return p, m


@pytest.mark.parametrize("low", [0, 7])
@pytest.mark.parametrize("host", [False, True])
def test_update_state_with_sources_bitwise(hc_lib, port, low, host):
    """Nyx::update_state_with_sources + enforce_minimum_density(floor) + gravity as one fused sweep: a ragged three-box level with the
    production ghost widths, with and without cells below small_dens, device FABs and host FABs (pipelined entry point): every byte of
    S_new (ghost cells included: untouched), hydro_src and the inputs equals the oracle's, and so does the reported minimum."""
    torch = _torch()
    d = util.sources_inputs(seed=900 + low, low_density_cells=low)
    ref, m_ref = _src_oracle(port, d)
    prm = hc_lib.src_params(small_dens=d["small_dens"], small_temp=d["small_temp"])
    slots = ("s_old", "s_new", "ext_src", "hydro_src", "grav", "reset_src")
    if host:
        arrs = {k: [x.copy() for x in d[k]] for k in slots}
        fabs, tiles = _src_fabs(d, arrs, pinned_host=True)
    else:
        arrs = {k: [torch.from_numpy(x).cuda() for x in d[k]] for k in slots}
        fabs, tiles = _src_fabs(d, arrs)
    m = hc_lib.update_state_with_sources_batch(fabs["s_old"], fabs["s_new"], fabs["ext_src"], fabs["hydro_src"], fabs["grav"], tiles, d["dt"],
                                               d["a_old"], d["a_new"], prm, host=host)
    assert m == m_ref and (m < d["small_dens"]) == (low > 0)
    get = (lambda x: x) if host else (lambda x: x.cpu().numpy())
    for bi in range(len(d["boxes"])):
        for k in ("s_new", "hydro_src", "s_old", "ext_src", "grav"):
            assert np.array_equal(get(arrs[k][bi]), ref[k][bi]), (k, bi)


@pytest.mark.parametrize("host", [False, True])
def test_enforce_minimum_density_multi_rank_protocol(hc_lib, port, host):
    """Two 'ranks' (two disjoint tile sets): only rank 1 has cells below small_dens.  Each rank runs hc_update_state_with_sources_batch on
    its own tiles, the minima are reduced (the reference's S_new.min() is global), and rank 0 -- whose own minimum was fine -- then calls
    hc_enforce_minimum_density_batch.  The result equals the oracle run over all boxes at once, bit for bit."""
    torch = _torch()
    d = util.sources_inputs(seed=930, low_density_cells=5)
    for bi in (0, 1):      # rank 0 = boxes 0, 1: take their low-density cells away again
        d["hydro_src"][bi][0] = np.abs(d["hydro_src"][bi][0])
    ref, m_ref = _src_oracle(port, d)
    prm = hc_lib.src_params(small_dens=d["small_dens"], small_temp=d["small_temp"])
    slots = ("s_old", "s_new", "ext_src", "hydro_src", "grav", "reset_src")
    arrs = {k: [x.copy() if host else torch.from_numpy(x).cuda() for x in d[k]] for k in slots}
    fabs, tiles = _src_fabs(d, arrs, pinned_host=host)
    ranks = [[0, 1], [2]]
    mins = []
    for own in ranks:
        sub = lambda k: [fabs[k][i] for i in own]
        mins.append(hc_lib.update_state_with_sources_batch(sub("s_old"), sub("s_new"), sub("ext_src"), sub("hydro_src"), sub("grav"),
                                                           [tiles[i] for i in own], d["dt"], d["a_old"], d["a_new"], prm, host=host))
    assert mins[0] >= d["small_dens"] > mins[1] and min(mins) == m_ref
    for own, m in zip(ranks, mins):
        if min(mins) < d["small_dens"] and not (m < d["small_dens"]):
            sub = lambda k: [fabs[k][i] for i in own]
            hc_lib.enforce_minimum_density_batch(sub("s_old"), sub("s_new"), sub("ext_src"), sub("hydro_src"), sub("grav"), [tiles[i] for i in own],
                                                 d["dt"], d["a_old"], d["a_new"], prm, host=host)
    torch.cuda.synchronize()
    for bi in range(len(d["boxes"])):
        for k in ("s_new", "hydro_src"):
            assert np.array_equal(arrs[k][bi] if host else arrs[k][bi].cpu().numpy(), ref[k][bi]), (k, bi)


@pytest.mark.parametrize("host", [False, True])
def test_update_state_with_sources_conservative(hc_lib, port, host):
    """nyx.enforce_min_density_type = "conservative" (Nyx::enforce_minimum_density_cons): the update as the reference's three sweeps on a
    periodic 16^3 level cut into two boxes -- source update alone + minimum; iterations of the density redistribution, each after a FillPatch
    of the two-ghost-cell border copy (emulated here: periodic wrap of the assembled level); gravity + the SDC reset of hydro_src(rho).  Every
    array after every call equals the port's bit for bit (the port's iteration equals the reference's own per-cell functions,
    tests/test_oracle_vs_reference.py); cells on the box faces are filled through the ghost cells; mass is conserved."""
    import ctypes as C
    from oracle import pyref
    torch = _torch()
    n, z = 16, 3.0
    rng = np.random.default_rng(77)
    target, small = util.cons_inputs(n, 741)
    a_old = 1.0 / (1.0 + z); dt = synth.step_dt(z); a_new = synth.a_after(z, dt)
    s_old = target * rng.uniform(0.8, 1.2, target.shape); s_old[0] = np.abs(s_old[0]) + small
    ext = 0.01 * s_old / dt * rng.standard_normal(s_old.shape); ext[0] = 0.0
    hs = 0.05 * s_old * rng.standard_normal(s_old.shape)
    hs[0] = target[0] - s_old[0]                       # the source update lands on the crafted density (up to rounding)
    grav = 1.0e3 * rng.standard_normal((3, n, n, n))
    level = dict(s_old=s_old, s_new=rng.standard_normal(s_old.shape), ext_src=ext, hydro_src=hs, grav=grav, reset_src=np.full((1, n, n, n), -3.0))
    boxes = [((0, 0, 0), (7, n - 1, n - 1)), ((8, 0, 0), (n - 1, n - 1, n - 1))]
    cut = lambda a, b: np.ascontiguousarray(a[..., b[0][0]:b[1][0] + 1])       # noqa: E731
    # ---- the port, box by box
    P = {k: [cut(v, b) for b in boxes] for k, v in level.items()}
    lib = port.lib
    fp, l3 = C.POINTER(pyref.HcoFab), C.c_int * 3
    lib.hco_sources_apply_box.restype = C.c_double
    lib.hco_sources_apply_box.argtypes = [fp] * 4 + [l3, l3, C.c_double, C.c_double, C.c_double]
    lib.hco_sources_finish_box.argtypes = [C.POINTER(pyref.HcoParams)] + [fp] * 4 + [l3, l3] + [C.c_double] * 5 + [C.c_int, C.c_int]
    F = lambda k, bi: C.byref(pyref.fab_of(P[k][bi], boxes[bi][0]))             # noqa: E731
    m_port = min(lib.hco_sources_apply_box(F("s_old", bi), F("s_new", bi), F("ext_src", bi), F("hydro_src", bi), l3(*b[0]), l3(*b[1]), dt, a_old, a_new)
                 for bi, b in enumerate(boxes))
    assert m_port < small
    # ---- the CUDA path
    prm = hc_lib.src_params(small_dens=small, small_temp=1.0e-2, min_density_type=1)
    if host:
        G = {k: [cut(v, b) for b in boxes] for k, v in level.items()}
        mk = capi.fab_of_numpy
    else:
        G = {k: [torch.from_numpy(cut(v, b)).cuda() for b in boxes] for k, v in level.items()}
        mk = capi.fab_of_torch
    get = (lambda x: x) if host else (lambda x: x.cpu().numpy())
    fabs = {k: [mk(G[k][bi], boxes[bi][0]) for bi in range(2)] for k in G}
    tiles = [capi.make_box(*b) for b in boxes]
    five = [fabs[k] for k in ("s_old", "s_new", "ext_src", "hydro_src", "grav")]
    m = hc_lib.update_state_with_sources_batch(*five, tiles, dt, a_old, a_new, prm, host=host)
    assert m == m_port
    for bi in range(2):
        assert np.array_equal(get(G["s_new"][bi]), P["s_new"][bi])          # the source update alone (no gravity yet)
    with pytest.raises(Exception):                                          # the floor variant's second pass is not for this type
        hc_lib.enforce_minimum_density_batch(*five, tiles, dt, a_old, a_new, prm, host=host)
    mass0 = sum(x[0].sum() for x in P["s_new"])
    it = 0
    while m < small and it < 10:
        whole = np.concatenate([get(x) for x in G["s_new"]], axis=3)
        assert np.array_equal(whole, np.concatenate(P["s_new"], axis=3))
        wb = util.fill_border(whole, 2)                                     # FillPatch: the caller's
        sb = [np.ascontiguousarray(wb[..., b[0][0]:b[1][0] + 5]) for b in boxes]
        sbl = [tuple(x - 2 for x in b[0]) for b in boxes]
        ms = []
        for bi, b in enumerate(boxes):
            mm, bad = port.enforce_min_cons_iter(sb[bi], P["s_new"][bi], P["reset_src"][bi], b[0], b[1], small)
            assert bad == 0
            ms.append(mm)
        sbg = sb if host else [torch.from_numpy(x).cuda() for x in sb]
        m = hc_lib.enforce_min_density_cons_iter([mk(sbg[bi], sbl[bi]) for bi in range(2)], fabs["s_new"], fabs["reset_src"], tiles, prm, host=host)
        assert m == min(ms), it
        for bi in range(2):
            assert np.array_equal(get(G["s_new"][bi]), P["s_new"][bi]) and np.array_equal(get(G["reset_src"][bi]), P["reset_src"][bi]), (it, bi)
        it += 1
    assert 2 <= it < 10 and m >= small
    assert abs(sum(x[0].sum() for x in P["s_new"]) / mass0 - 1) < 1e-12
    # ---- gravity and the SDC reset of hydro_src(rho)
    pp = port.params()
    for bi, b in enumerate(boxes):
        P["hydro_src"][bi][0] = P["s_new"][bi][0] - P["s_old"][bi][0]       # Nyx_enforce_minimum_density.cpp:40-63
        lib.hco_sources_finish_box(C.byref(pp), F("s_old", bi), F("s_new", bi), F("hydro_src", bi), F("grav", bi), l3(*b[0]), l3(*b[1]), dt, a_old, a_new,
                                   small, 1.0e-2, 0, 1)
    hc_lib.finish_state_with_sources_batch(*five, tiles, dt, a_old, a_new, prm, True, host=host)
    if not host:
        torch.cuda.synchronize()
    for bi in range(2):
        for k in ("s_new", "hydro_src", "reset_src", "s_old", "ext_src", "grav"):
            assert np.array_equal(get(G[k][bi]), P[k][bi]), (k, bi)
    # a border copy with fewer than two ghost cells is refused
    with pytest.raises(Exception):
        hc_lib.enforce_min_density_cons_iter(fabs["s_new"], fabs["s_new"], fabs["reset_src"], tiles, prm, host=host)


def test_sources_argument_errors_and_async(hc_lib):
    """an unknown enforce_min_density_type, the floor variant's second pass asked for the conservative type and wrong component counts are
    rejected; want_min=False does not synchronise and gives the same S_new"""
    torch = _torch()
    d = util.sources_inputs(seed=940)
    slots = ("s_old", "s_new", "ext_src", "hydro_src", "grav", "reset_src")
    arrs = {k: [torch.from_numpy(x).cuda() for x in d[k]] for k in slots}
    fabs, tiles = _src_fabs(d, arrs)
    args = (fabs["s_old"], fabs["s_new"], fabs["ext_src"], fabs["hydro_src"], fabs["grav"], tiles, d["dt"], d["a_old"], d["a_new"])
    with pytest.raises(capi.HcError, match="enforce_min_density_type"):
        hc_lib.update_state_with_sources_batch(*args, hc_lib.src_params(small_dens=1.0, small_temp=1.0, min_density_type=2))
    with pytest.raises(capi.HcError, match="conservative"):
        hc_lib.enforce_minimum_density_batch(*args, hc_lib.src_params(small_dens=1.0, small_temp=1.0, min_density_type=1))
    with pytest.raises(capi.HcError, match="6 components"):
        hc_lib.update_state_with_sources_batch(fabs["s_old"], fabs["s_new"], fabs["ext_src"], fabs["hydro_src"], fabs["reset_src"], tiles, d["dt"], d["a_old"],
                                               d["a_new"], hc_lib.src_params())
    prm = hc_lib.src_params(small_dens=d["small_dens"], small_temp=d["small_temp"])
    hc_lib.update_state_with_sources_batch(*args, prm)
    first = [x.clone() for x in arrs["s_new"]]
    assert hc_lib.update_state_with_sources_batch(*args, prm, want_min=False) is None
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(first, arrs["s_new"]))
    assert hc_lib.update_state_with_sources_batch([], [], [], [], [], [], d["dt"], d["a_old"], d["a_new"], prm) == np.finfo(np.float64).max


def test_fab_copy_add_subtract(hc_lib):
    """MultiFab::Copy / Add / Subtract of a component range over valid boxes (sdc_hydro.cpp:83-84,94-95,112,135), FABs with different ghost widths"""
    torch = _torch()
    rng = np.random.default_rng(950)
    boxes = util.SRC_BOXES
    shp = lambda bx, g: (bx[5] - bx[2] + 1 + 2 * g, bx[4] - bx[1] + 1 + 2 * g, bx[3] - bx[0] + 1 + 2 * g)
    ext = [rng.standard_normal((6,) + shp(b, 4)) for b in boxes]
    ir = [rng.standard_normal((1,) + shp(b, 1)) for b in boxes]
    ext_d, ir_d = [torch.from_numpy(x).cuda() for x in ext], [torch.from_numpy(x).cuda() for x in ir]
    fe = [capi.fab_of_torch(x, tuple(c - 4 for c in b[:3])) for x, b in zip(ext_d, boxes)]
    fi = [capi.fab_of_torch(x, tuple(c - 1 for c in b[:3])) for x, b in zip(ir_d, boxes)]
    tiles = [capi.make_box(b[:3], b[3:]) for b in boxes]
    want = [x.copy() for x in ext]
    v4, v1 = (slice(4, -4),) * 3, (slice(1, -1),) * 3
    for op, comp in (("add", 4), ("add", 5), ("subtract", 4), ("copy", 1)):
        hc_lib.fab_op_batch(op, fe, comp, fi, 0, 1, tiles)
        for w, s in zip(want, ir):
            if op == "add": w[(comp,) + v4] = w[(comp,) + v4] + s[(0,) + v1]
            elif op == "subtract": w[(comp,) + v4] = w[(comp,) + v4] - s[(0,) + v1]
            else: w[(comp,) + v4] = s[(0,) + v1]
    torch.cuda.synchronize()
    for x, w in zip(ext_d, want):
        assert np.array_equal(x.cpu().numpy(), w)
    with pytest.raises(capi.HcError, match="component range"):
        hc_lib.fab_op_batch("copy", fe, 6, fi, 0, 1, tiles)


def test_sources_full_size_properties(hc_lib):
    """Rank-2 row at a full-size level (256^3 cells, beyond what the oracle follows in seconds), through size-independent properties:
    (a) identity: no sources, no gravity, a_new == a_old  ->  S_new == S_old bit for bit and the reported minimum is min(rho);
    (b) decomposition independence: the same field as one box and as 64 boxes (each its own FAB) gives bit-identical S_new and minimum;
    (c) the floor acts exactly on the cells below small_dens, which end at (small_dens, 0 momentum + gravity, floor energy), and hydro_src(rho)
        is S_new(rho) - S_old(rho) everywhere; (d) determinism."""
    torch = _torch()
    n = 256
    gen = torch.Generator(device="cuda").manual_seed(77)
    z = 3.0
    a_old, dt = 1.0 / (1.0 + z), synth.step_dt(z)
    a_new = synth.a_after(z, dt)
    rho_b = synth.mean_rhob()
    s_old = torch.randn((6, n, n, n), generator=gen, device="cuda", dtype=torch.float64)
    s_old[0] = rho_b * torch.exp(s_old[0].clamp(-3, 3))      # >= 0.05 rho_b: five times small_dens
    s_old[4:6] = s_old[4:6].abs() * 1e12 * s_old[0]
    zeros6 = torch.zeros_like(s_old)
    grav0 = torch.zeros((3, n, n, n), dtype=torch.float64, device="cuda")
    lo, hi = (0, 0, 0), (n - 1,) * 3
    prm = hc_lib.src_params(small_dens=1.0e-2 * rho_b, small_temp=1.0e-2)

    def run(s_in, ext, hs, grav, a1, nsplit=1):
        m = n // nsplit
        keep, fabs, tiles = [], [[] for _ in range(5)], []
        for kb in range(nsplit):
            for jb in range(nsplit):
                for ib in range(nsplit):
                    sl = (slice(None), slice(kb * m, (kb + 1) * m), slice(jb * m, (jb + 1) * m), slice(ib * m, (ib + 1) * m))
                    blo = (ib * m, jb * m, kb * m)
                    parts = [x[sl].contiguous() if nsplit > 1 else x for x in (s_in, None, ext, hs, grav) if x is not None]
                    parts.insert(1, torch.full_like(parts[0], float("nan")))
                    keep.append((sl, parts))
                    for slot, p in enumerate(parts):
                        fabs[slot].append(capi.fab_of_torch(p, blo))
                    tiles.append(capi.make_box(blo, tuple(c + m - 1 for c in blo)))
        mn = hc_lib.update_state_with_sources_batch(fabs[0], fabs[1], fabs[2], fabs[3], fabs[4], tiles, dt, a_old, a1, prm)
        out = torch.empty_like(s_in)
        hs_out = torch.empty_like(s_in)
        for sl, parts in keep:
            out[sl] = parts[1]
            hs_out[sl] = parts[3]
        return out, hs_out, mn

    out, _, mn = run(s_old, zeros6, zeros6.clone(), grav0, a_old)                               # (a)
    # a_old * u / a_old and a_old^2 * u / a_old^2 are exact only up to one rounding each way: the reference's expressions, not an identity in floating point
    assert torch.equal(out[0], s_old[0]) and mn == float(s_old[0].min())
    assert torch.allclose(out[1:], s_old[1:], rtol=4e-16, atol=0.0)
    ext = 0.05 / dt * s_old * torch.randn(s_old.shape, generator=gen, device="cuda", dtype=torch.float64)
    ext[0] = 0.0
    hs = 0.1 * s_old * torch.randn(s_old.shape, generator=gen, device="cuda", dtype=torch.float64).clamp(-3, 3)
    grav = 1.0e3 * torch.randn((3, n, n, n), generator=gen, device="cuda", dtype=torch.float64)
    o1, h1, m1 = run(s_old, ext, hs.clone(), grav, a_new)                                       # no cell below small_dens
    assert m1 > prm.small_dens and torch.equal(h1, hs)
    o1b, _, m1b = run(s_old, ext, hs.clone(), grav, a_new)
    assert torch.equal(o1, o1b) and m1 == m1b                                                   # (d)
    o4, h4, m4 = run(s_old, ext, hs.clone(), grav, a_new, nsplit=4)                             # (b)
    assert torch.equal(o1, o4) and torch.equal(h1, h4) and m1 == m4
    hs_low = hs.clone()
    hs_low[0, ::37, ::41, ::43] = -1.5 * s_old[0, ::37, ::41, ::43]                             # (c): these cells go below small_dens
    o2, h2, m2 = run(s_old, ext, hs_low.clone(), grav, a_new)
    o2s, h2s, m2s = run(s_old, ext, hs_low.clone(), grav, a_new, nsplit=4)
    assert m2 < prm.small_dens and m2 == m2s and torch.equal(o2, o2s) and torch.equal(h2, h2s)
    low = torch.zeros((n, n, n), dtype=torch.bool, device="cuda")
    low[::37, ::41, ::43] = True
    assert torch.equal(o2[0] == prm.small_dens, low)
    assert torch.equal(h2[0], o2[0] - s_old[0]) and torch.equal(h2[1:], hs_low[1:])
    assert torch.equal(o2[:, ~low], o1[:, ~low])                                                # untouched cells: same values as without the floor
    mom = s_old[0, low] * grav[0, low] * (dt / a_new)                                           # floored momentum 0 + gravity
    assert torch.equal(o2[1, low], mom) and bool((o2[5, low] == o2[5, low][0]).all())
