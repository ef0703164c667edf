"""CPU tests (no GPU): the product's per-cell state machine compiled for the host (tests/host_harness.cpp) against the oracle,
the host-side rate tables, the C-ABI library (loads, exports every declared symbol, fails loudly without a device), and the
box-sharding / statistics all-reduce logic (world_size 2, gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from nyx_b200 import capi, sharded, synth
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(built):
    return util.Harness(built.build_harness())


@pytest.fixture(scope="module")
def rates(harness):
    from oracle import pyref
    rc, r = harness.tabulate(pyref.TREECOOL, synth.mean_rhob())
    assert rc == 0
    return r


def test_rate_tables_match_oracle_bitwise(rates, port):
    assert np.array_equal(rates, port.rates())


def test_tabulate_rates_error_codes(harness, tmp_path):
    rc, _ = harness.tabulate(str(tmp_path / "missing"), 1.0)
    assert rc == -4                                   # HC_ERR_IO
    from oracle import pyref
    long_file = tmp_path / "TREECOOL_long"
    txt = open(pyref.TREECOOL).read()
    long_file.write_text(txt + txt.splitlines()[-1] + "\n")
    rc, _ = harness.tabulate(str(long_file), 1.0)
    assert rc == -5                                   # HC_ERR_TREECOOL_LEN: the reference aborts (atomic_rates.H:37-51)
    short_file = tmp_path / "TREECOOL_short"
    short_file.write_text("\n".join(txt.splitlines()[:100]) + "\n")
    rc, _ = harness.tabulate(str(short_file), 1.0)
    assert rc == -4


# (z = 20, 100: above the last TREECOOL row interp_to_this_z returns zeros (eos_hc.H:17-27) -- every photo-ionisation numerator is exactly 0 and
#  Compton cooling against the CMB dominates; z = 100 is the regime of BASELINE config 1)
@pytest.mark.parametrize("z,n,seed", [(3.0, 16, 5), (2.0, 16, 6), (6.0, 16, 7), (3.0, 5, 8), (20.0, 12, 9), (100.0, 12, 10)])
def test_state_machine_vec_bitwise(harness, rates, port, z, n, seed):
    """The staged per-lane BDF state machine of nyx_b200/csrc/hc_device.cuh == the oracle, bit for bit, counters included."""
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s_ref, d_ref = state.copy(), diag.copy()
    pst = port.integrate_state_vec(s_ref, d_ref, lo, hi, a, dt)
    cs = harness.integrate_vec(rates, state, diag, lo, hi, a, dt)
    assert np.array_equal(state, s_ref) and np.array_equal(diag, d_ref)
    assert util.stats_equal(cs, pst)


@pytest.mark.parametrize("z,seed,src,flash", [(3.0, 21, 0.0, "none"), (2.0, 22, 0.05, "none"), (6.0, 23, 0.2, "none"),
                                               (5.99, 24, 0.05, "hi_now"), (3.0, 25, 0.05, "heii_now"), (7.0, 26, 0.0, "before"),
                                               (20.0, 27, 0.02, "none"), (100.0, 28, 0.02, "none")])   # (a source-free S_new = S_old is not a consistent SDC input at z = 100: cells that cool to e_out < 0 escape the floor test and the reference's final EOS solve indexes its tables with log10 of a negative number)
def test_state_machine_struct_bitwise(harness, rates, port, z, seed, src, flash):
    n = 12
    kw = util.FLASH_CASES[flash]
    d = util.sdc_inputs(z, n, seed, src)
    r = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in d.items()}
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    pst = port.integrate_state_struct(r["s_old"], r["s_new"], r["diag"], r["hydro_src"], r["reset_src"], r["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**kw))
    cs = harness.integrate_struct(rates, d, lo, hi, 0, params=harness.params(**kw))
    for k in ("s_old", "s_new", "diag", "ir"):
        assert np.array_equal(d[k], r[k]), k
    assert util.stats_equal(cs, pst)


def test_state_machine_inhomogeneous_reionization_bitwise(harness, rates, port):
    n, z = 10, 5.5
    d = util.inhomo_inputs(z, n, 341)
    r = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in d.items()}
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    pst = port.integrate_state_struct(r["s_old"], r["s_new"], r["diag"], r["hydro_src"], r["reset_src"], r["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**d["kw"]))
    cs = harness.integrate_struct(rates, d, lo, hi, 0, params=harness.params(**d["kw"]))
    for k in ("s_old", "s_new", "diag", "ir"):
        assert np.array_equal(d[k], r[k]), k
    assert util.stats_equal(cs, pst)


@pytest.mark.parametrize("kw", [dict(use_constraint=1), dict(use_typical_steps=1, old_max_steps=3), dict(max_steps=4),
                                dict(rtol=1e-6, atol_factor=1e-6), dict(h_species=0.7)])
def test_state_machine_options_bitwise(harness, rates, port, kw):
    z, n = 3.0, 10
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=31, z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s_ref, d_ref = state.copy(), diag.copy()
    pst = port.integrate_state_vec(s_ref, d_ref, lo, hi, a, dt, params=port.params(**kw))
    cs = harness.integrate_vec(rates, state, diag, lo, hi, a, dt, params=harness.params(**kw))
    assert np.array_equal(state, s_ref) and np.array_equal(diag, d_ref) and util.stats_equal(cs, pst)
    if "max_steps" in kw:
        assert (pst[:, 7] == -1).any()                # CV_TOO_MUCH_WORK is reached and reproduced


def test_sdc_iter_negative_and_sub_box(harness, rates, port):
    """sdc_iter < 0 (pure reaction step on S_old) and a tile strictly inside its FABs."""
    z, n = 2.0, 10
    d = util.sdc_inputs(z, n, 33, 0.0)
    r = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in d.items()}
    lo, hi = (2, 1, 3), (7, 8, 6)
    pst = port.integrate_state_struct(r["s_old"], r["s_new"], r["diag"], r["hydro_src"], r["reset_src"], r["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], -1, los=[(0, 0, 0)] * 6)
    cs = harness.integrate_struct(rates, d, lo, hi, -1)
    for k in ("s_old", "s_new", "diag", "ir"):
        assert np.array_equal(d[k], r[k]), k
    assert util.stats_equal(cs, pst)
    untouched = np.ones((n, n, n), bool)
    untouched[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = False
    ref0 = util.sdc_inputs(z, n, 33, 0.0)
    assert np.array_equal(d["s_old"][5][untouched], ref0["s_old"][5][untouched])


# ------------------------------------------------------------------------------------------------ the C-ABI library
def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "nyx_hc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[A-Za-z0-9_]+)\s*\(", txt)))


def test_log10_table_and_algorithm(harness):
    """The table hc_tables_upload builds for the device's fast_log10, and the algorithm itself emulated in exact rational
    arithmetic (every FMA = exact product-sum, one rounding): max error well below 1 ulp against a 50-digit log10."""
    import math
    import struct
    from decimal import Decimal, getcontext
    from fractions import Fraction as F
    getcontext().prec = 50
    tab = np.zeros(128 * 4)
    harness.lib.hh_log10_table.argtypes = [C.POINTER(C.c_double)]
    harness.lib.hh_log10_table.restype = None
    harness.lib.hh_log10_table(tab.ctypes.data_as(C.POINTER(C.c_double)))
    tab = tab.reshape(128, 4)

    def dlog10(fr):
        return Decimal(fr.numerator).log10() - Decimal(fr.denominator).log10()

    for i in range(128):
        r, lhi, llo, pad = (float(v) for v in tab[i])
        assert r == float(1 / F(2 * 128 + 2 * i + 1, 2 * 128)) and pad == 0.0
        assert (lhi * 2.0 ** 42).is_integer() and 0.0 <= lhi < 0.31
        assert abs(Decimal(lhi) + Decimal(llo) + dlog10(F(r))) < Decimal(2) ** -62
    log2 = Decimal(2).log10()
    log2_hi, log2_lo = float.fromhex("0x1.34413509f7000p-2"), float.fromhex("0x1.3fde623e2566bp-43")
    assert (log2_hi * 2.0 ** 42).is_integer() and abs(Decimal(log2_hi) + Decimal(log2_lo) - log2) < Decimal(2) ** -95
    coef = [float.fromhex(h) for h in ("0x1.bcb7b1526e50ep-2", "-0x1.bcb7b1526e50ep-3", "0x1.287a7636f435fp-3", "-0x1.bcb7b1526e50ep-4",
                                      "0x1.63c62775250d8p-4", "-0x1.287a7636f435fp-4")]

    def fma(a, b, c):
        return float(F(a) * F(b) + F(c))

    def flog10(x):
        bits = struct.unpack("<q", struct.pack("<d", x))[0]
        e = ((bits >> 52) & 0x7ff) - 1023
        m = struct.unpack("<d", struct.pack("<q", (bits & ((1 << 52) - 1)) | (1023 << 52)))[0]
        r, lhi, llo, _ = (float(v) for v in tab[(bits >> 45) & 127])
        z = fma(m, r, -1.0)
        q = coef[5]
        for c in coef[4::-1]:
            q = fma(z, q, c)
        small = fma(z, q, fma(float(e), log2_lo, llo))
        big = fma(float(e), log2_hi, lhi)
        assert F(big) == F(float(e)) * F(log2_hi) + F(lhi)        # exact by construction
        return float(F(big) + F(small))

    rng = np.random.default_rng(5)
    worst = 0.0
    for x in 10.0 ** rng.uniform(0.35, 10.0, 3000):
        y, ref = flog10(float(x)), dlog10(F(float(x)))
        worst = max(worst, float(abs(Decimal(y) - ref) / Decimal(math.ulp(float(ref)))))
    assert worst < 0.52, worst


def test_cabi_exports_every_declared_symbol(built):
    lib = C.CDLL(built.build_cuda())
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libnyx_hc.so does not export {n}"
    capi.declare(lib)
    assert b"sm_100a" in lib.hc_version()


def test_cabi_struct_layouts_match_header():
    assert C.sizeof(capi.HcFab) == 8 + 3 * 8 + 6 * 4 + 8
    assert C.sizeof(capi.HcBox) == 24
    assert C.sizeof(capi.HcStats) == 14 * 8
    assert C.sizeof(capi.HcParams) == 10 * 8 + 2 * 8 + 4 * 4
    assert capi.CELLSTAT_DTYPE.itemsize == 32


def test_cabi_fails_loudly_without_device(built):
    """No CPU fallback: on a machine without a CUDA device every compute entry point returns HC_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from oracle import pyref
    hc = capi.NyxHC()
    rates = hc.tabulate_rates(pyref.TREECOOL, synth.mean_rhob())       # pure host code: works
    assert rates.shape == (capi.RATES_DOUBLES,)
    with pytest.raises(capi.HcError, match="-1"):
        hc.tables_upload(rates)
    state, diag = synth.make_fab((4, 4, 4), seed=1, z=3.0)
    with pytest.raises(capi.HcError):
        hc.integrate_vec_host([capi.fab_of_numpy(state, (0, 0, 0))], [capi.fab_of_numpy(diag, (0, 0, 0))],
                              [capi.make_box((0, 0, 0), (3, 3, 3))], 0.25, 1e-3)
    assert np.array_equal(state, synth.make_fab((4, 4, 4), seed=1, z=3.0)[0])     # and nothing was computed behind our back
    # the rows either side of the path likewise: host and device entry points
    d = util.sources_inputs(seed=3)
    before = [x.copy() for x in d["s_new"]]
    fabs = {k: [capi.fab_of_numpy(x, tuple(c - g for c in bx[:3])) for x, bx in zip(d[k], d["boxes"])]
            for k, g in zip(("s_old", "s_new", "ext_src", "hydro_src", "grav"), d["ng"])}
    tiles = [capi.make_box(bx[:3], bx[3:]) for bx in d["boxes"]]
    prm = hc.src_params(small_dens=d["small_dens"], small_temp=d["small_temp"])
    for host in (True, False):
        with pytest.raises(capi.HcError):
            hc.update_state_with_sources_batch(fabs["s_old"], fabs["s_new"], fabs["ext_src"], fabs["hydro_src"], fabs["grav"], tiles, d["dt"],
                                               d["a_old"], d["a_new"], prm, host=host)
        with pytest.raises(capi.HcError):
            hc.enforce_minimum_density_batch(fabs["s_old"], fabs["s_new"], fabs["ext_src"], fabs["hydro_src"], fabs["grav"], tiles, d["dt"],
                                             d["a_old"], d["a_new"], prm, host=host)
    with pytest.raises(capi.HcError):
        hc.fab_op_batch("add", fabs["ext_src"], 4, fabs["hydro_src"], 0, 1, tiles)
    with pytest.raises(capi.HcError):
        hc.compute_new_temp_batch([capi.fab_of_numpy(state, (0, 0, 0))], [capi.fab_of_numpy(diag, (0, 0, 0))], [capi.make_box((0, 0, 0), (3, 3, 3))],
                                  0.25, 1e-2, 1e9, 0)
    # round-2 entry points: the conservative variant's iteration / finish calls and the SAVE_REACT overload
    prm_c = hc.src_params(small_dens=d["small_dens"], small_temp=d["small_temp"], min_density_type=1)
    n = 6
    sn = np.ones((6, n, n, n)); sb = np.ones((6, n + 4, n + 4, n + 4)); rs = np.zeros((1, n, n, n))
    tl = [capi.make_box((0, 0, 0), (n - 1,) * 3)]
    for host in (True, False):
        with pytest.raises(capi.HcError):
            hc.enforce_min_density_cons_iter([capi.fab_of_numpy(sb, (-2, -2, -2))], [capi.fab_of_numpy(sn, (0, 0, 0))], [capi.fab_of_numpy(rs, (0, 0, 0))],
                                             tl, prm_c, host=host)
        with pytest.raises(capi.HcError):
            hc.finish_state_with_sources_batch(fabs["s_old"], fabs["s_new"], fabs["ext_src"], fabs["hydro_src"], fabs["grav"], tiles, d["dt"],
                                               d["a_old"], d["a_new"], prm_c, True, host=host)
    assert (sn == 1.0).all()
    sd = util.sdc_inputs(3.0, 4, 5, 0.05)
    six = [[capi.fab_of_numpy(sd[k], (0, 0, 0))] for k in ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")]
    react = [[capi.fab_of_numpy(np.zeros((c, 4, 4, 4)), (0, 0, 0))] for c in (7, 7, 9)]
    with pytest.raises(capi.HcError):
        hc.integrate_struct_react_host(*six, *react, [capi.make_box((0, 0, 0), (3, 3, 3))], sd["a"], sd["a_end"], sd["dt"], 0)
    with pytest.raises(capi.HcError):
        hc.integrate_struct_react_batch(*six, *react, [capi.make_box((0, 0, 0), (3, 3, 3))], sd["a"], sd["a_end"], sd["dt"], 0)
    assert all(np.array_equal(x, y) for x, y in zip(before, d["s_new"]))
    with pytest.raises(capi.HcError, match="missing"):
        capi.NyxHC(path="/nonexistent/libnyx_hc.so")


# ------------------------------------------------------------------------------------------------ box sharding
def test_box_list_and_distribution_map():
    boxes = sharded.box_list(512, 128)
    assert len(boxes) == 64 and boxes[0] == ((0, 0, 0), (127, 127, 127)) and boxes[-1][1] == (511, 511, 511)
    assert sum(sharded.box_cells(b) for b in boxes) == 512 ** 3
    ragged = sharded.box_list(100, 64)
    assert len(ragged) == 8 and sum(sharded.box_cells(b) for b in ragged) == 100 ** 3
    for world in (1, 2, 4, 8):
        owner = sharded.distribution_map(boxes, world)
        counts = np.bincount(owner, minlength=world)
        assert counts.min() == counts.max() == 64 // world
        mine = [sharded.local_boxes(boxes, world, r) for r in range(world)]
        assert sorted(sum(mine, [])) == list(range(64))
    owner = sharded.distribution_map(ragged, 2)
    cells = [sum(sharded.box_cells(ragged[i]) for i in range(8) if owner[i] == r) for r in (0, 1)]
    assert abs(cells[0] - cells[1]) <= 64 ** 3


def test_algorithmic_flops_formula():
    st = dict(zip(capi.STATS_FIELDS, [0] * len(capi.STATS_FIELDS)))
    st.update(sum_ne_iters=10, sum_nfe=3, sum_nfe_ls=1, sum_attempts=2, sum_eos=1)
    assert sharded.algorithmic_flops(st) == 186 * 10 + 234 * 4 + 60 * 2 + 99


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from nyx_b200 import capi, sharded
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
boxes = sharded.box_list(64, 32)
mine = sharded.local_boxes(boxes, 2, rank)
st = dict(zip(capi.STATS_FIELDS, [0] * len(capi.STATS_FIELDS)))
st["n_cells"] = sum(sharded.box_cells(boxes[i]) for i in mine)
st["sum_nst"] = 100 * (rank + 1)
st["max_nst"] = 7 + 5 * rank
g = sharded.allreduce_stats(st)
assert g["n_cells"] == 64 ** 3 and g["sum_nst"] == 300 and g["max_nst"] == 12, g
assert sorted(mine + sharded.local_boxes(boxes, 2, 1 - rank)) == list(range(8))
# SURVEY 8f rank 2: the global density-floor decision of enforce_minimum_density (MIN all-reduce), three cases
for case, mins, want_late in (("nobody below", (2.0, 3.0), (False, False)), ("rank 1 below", (2.0, 0.5), (True, False)), ("both below", (0.2, 0.5), (False, False))):
    calls = []
    lm, gm, late = sharded.update_state_with_sources_sharded(lambda: (calls.append("update"), mins[rank])[1], lambda: calls.append("enforce"), small_dens=1.0)
    assert lm == mins[rank] and gm == min(mins) and late == want_late[rank], (case, rank, lm, gm, late)
    assert calls == (["update", "enforce"] if want_late[rank] else ["update"]), (case, calls)
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_stats_allreduce_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in (0, 1)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_bench_reference_arm_line(built):
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line, the same metric / unit / config keys as the GPU
    arm's line (so that the driver's same-config check holds), min / median / max of the timed samples, host threads counted before libgomp binds"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "32", "--box", "16", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "heatcool_cell_updates_per_s" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["min"] <= cb["median"] <= cb["max"] and cb["value"] == d["value"]
    assert cb["cores"] == len(os.sched_getaffinity(0))      # torchrun's OMP_NUM_THREADS=1 does not pin the reference to one thread
    # the keys the GPU arm writes into `config` (bench.py: config_of)
    for k in ("workload", "cells_per_gpu", "boxes_per_gpu", "path", "z", "rtol", "atol_factor", "parallelism"):
        assert k in d["config"], k
    assert d["config"]["path"] == "vec" and d["config"]["cells_per_gpu"] == 32 ** 3
