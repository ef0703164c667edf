"""The oracle's C restatement (oracle/hc_oracle.c) against golden vectors produced by the reference itself
(tests/golden/make_golden.py ran oracle/_ref, i.e. Nyx HeatCool + SUNDIALS CVODE compiled from /root/reference).
Runs anywhere (no GPU, no /root/reference).  Bar: bit-exact against the per-cell reference mode; within the
north-star tolerance 10 x rtol = 1e-3 in e and T against the coupled (default-tile) reference mode.
"""
import os

import numpy as np
import pytest

from nyx_b200 import synth
from tests import util
from tests.golden import make_golden as mg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "heatcool_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_rate_tables_bitwise(port, gold):
    assert np.array_equal(port.rates(), gold["rates"])


def test_uvb_interpolation_bitwise(port, gold):
    from oracle import pyref
    for z, want in zip(gold["uvb_z"], gold["uvb"]):
        got = np.zeros(6)
        port.lib.hco_interp_to_this_z(port.rp, float(z), got.ctypes.data_as(pyref._dp))
        assert np.array_equal(got, want), z
    assert np.all(gold["uvb"][-1] == 0.0)      # z beyond the TREECOOL table: UV background off (eos_hc.H:19-27)


def test_eos_known_answers_bitwise(port, gold):
    for R, e, want in zip(gold["eos_R"], gold["eos_e"], gold["eos_TNe"]):
        assert port.eos_T_given_Re(1, 1, float(R), float(e), 0.25, 2.0 / 3.0, 0.76) == tuple(want)


@pytest.mark.parametrize("name,z,n,seed", mg.VEC_CASES)
def test_vec_golden(port, gold, name, z, n, seed):
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    pst = port.integrate_state_vec(state, diag, lo, hi, a, dt)
    # per-cell reference mode: everything bit-identical, CVODE counters and flags included
    assert np.array_equal(state[4:6], gold[f"{name}.percell.state"])
    assert np.array_equal(diag, gold[f"{name}.percell.diag"])
    assert np.array_equal(pst[:, :8], gold[f"{name}.percell.stats"])
    # coupled reference mode (what production Nyx runs): the north-star tolerance
    assert np.abs(state[5] / gold[f"{name}.coupled.state"][1] - 1).max() < 1e-3
    assert np.abs(diag[0] / gold[f"{name}.coupled.diag"][0] - 1).max() < 1e-3


@pytest.mark.parametrize("name,z,n,seed,src,flash", mg.STRUCT_CASES)
def test_struct_golden(port, gold, name, z, n, seed, src, flash):
    d = util.sdc_inputs(z, n, seed, src)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    pst = port.integrate_state_struct(d["s_old"], d["s_new"], d["diag"], d["hydro_src"], d["reset_src"], d["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**util.FLASH_CASES[flash]))
    assert np.array_equal(d["s_new"][4:6], gold[f"{name}.percell.s_new"])
    assert np.array_equal(d["diag"], gold[f"{name}.percell.diag"])
    assert np.array_equal(d["ir"], gold[f"{name}.percell.ir"])
    assert np.array_equal(pst[:, :8], gold[f"{name}.percell.stats"])
    ok = (pst[:, 7] == 0).reshape(n, n, n)
    rel = np.abs(d["s_new"][5] / gold[f"{name}.coupled.s_new"][1] - 1)[ok]
    # The coupled mode shares one step size/order over a tile and tests the RMS error over it; the per-cell mode controls
    # the error cell by cell.  The two modes OF THE REFERENCE therefore differ by more than 10 x rtol in a few stiff cells
    # (up to 41 % in one cell of struct_z2_src).  Bulk within 10 x rtol, outliers rare:
    assert np.median(rel) < 3e-4 and np.mean(rel > 1e-3) < 0.04


@pytest.mark.parametrize("case", ["nofloor", "floor"])
def test_update_state_with_sources_golden(port, case):
    """the port against outputs of the reference's own Nyx_update_state_with_sources.cpp (tests/golden/make_sources_golden.py),
    bit for bit, with and without cells below small_dens"""
    import copy
    from tests.golden import make_sources_golden as msg
    gold = np.load(os.path.join(os.path.dirname(GOLD), "sources_golden.npz"))
    name, z, seed, low = [c for c in msg.CASES if c[0] == case][0]
    d = msg.inputs(z, seed, low)
    p = copy.deepcopy(d)
    m = port.update_state_with_sources(d["boxes"], p["s_old"], p["s_new"], p["ext_src"], p["hydro_src"], p["grav"], d["dt"], d["a_old"], d["a_new"],
                                       d["small_dens"], d["small_temp"], ng=d["ng"][:5])
    assert (m < d["small_dens"]) == (low > 0)
    for bi in range(len(msg.BOXES)):
        assert np.array_equal(p["s_new"][bi], gold[f"{case}.s_new.{bi}"])
        assert np.array_equal(p["hydro_src"][bi], gold[f"{case}.hydro_src.{bi}"])
