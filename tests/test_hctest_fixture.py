"""BASELINE config 1 / SURVEY 8f rank 3: a hctest snapshot WRITTEN BY THE REFERENCE ITSELF (the unmodified Nyx + AMReX + SUNDIALS
build of tests/golden/make_hctest_fixture.sh: Exec/LyA inputs.rt, first coarse step, 32^3, z = 100 -> 99, UV background off), together with
the reference's own answer for the replayed step (Nyx::integrate_state_struct with nyx.hctest_example_read = 1, captured by
tests/golden/hctest_replay_ref.cpp linked against the reference's objects).

CPU: the reader of nyx_b200/hctest.py takes the reference's bytes (sdc_writeOn, Source/HeatCool/f_rhs_struct.H:587-655); the replay through
oracle/_ref (the reference's HeatCool translation units over the AMReX API shim) equals the real AMReX build BIT FOR BIT, which pins the
shim build on the real one; the per-cell port agrees within the 10 x rtol contract.
GPU: the CUDA path replays the snapshot (device entry point, host entry point, C++ drop-in) and is compared with the reference's answer and,
counter for counter, with the port.
"""
import hashlib
import os

import numpy as np
import pytest

from nyx_b200 import hctest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "hctest_lya32")
OUT_NAMES = ("s_new", "diag", "ir")   # what hctest_replay_ref.cpp writes after the step
EINT, EDEN = 5, 4


@pytest.fixture(scope="module")
def fx():
    f = hctest.read_fixture(FIX, 0)
    out, olos = hctest.read_fabs(os.path.join(FIX, "Chunk.0.out.0"), OUT_NAMES)
    f["ref_out"], f["ref_out_los"] = out, olos
    return f


def _valid(arr, lo, box):
    (blo, bhi) = box
    sl = tuple(slice(blo[d] - lo[d], bhi[d] - lo[d] + 1) for d in (2, 1, 0))
    return arr[(slice(None),) + sl]


def test_reference_bytes(fx):
    """the survey's facts about the reference-written chunk, now on the file itself"""
    sums = dict(line.split()[::-1] for line in open(os.path.join(FIX, "SHA256SUMS")))
    raw = hctest._read_bytes(os.path.join(FIX, "Chunk.0.0"))
    assert len(raw) == 7987301
    assert hashlib.sha256(raw).hexdigest() == sums["Chunk.0.0"]
    assert hashlib.sha256(hctest._read_bytes(os.path.join(FIX, "Chunk.0.out.0"))).hexdigest() == sums["Chunk.0.out.0"]
    assert open(os.path.join(FIX, "BADMAP.0")).read() == "(1 0\n((0,0,0) (31,31,31) (0,0,0))\n)(1\n0\n)"
    assert fx["boxes"] == [((0, 0, 0), (31, 31, 31))]
    fabs, los = fx["chunks"][0]
    shapes = {k: (fabs[k].shape, los[k]) for k in hctest.FAB_ORDER}
    assert shapes == {"s_old": ((6, 40, 40, 40), (-4, -4, -4)), "diag": ((2, 34, 34, 34), (-1, -1, -1)), "s_new": ((6, 34, 34, 34), (-1, -1, -1)),
                      "hydro_src": ((6, 32, 32, 32), (0, 0, 0)), "reset_src": ((1, 40, 40, 40), (-4, -4, -4)), "ir": ((1, 34, 34, 34), (-1, -1, -1))}
    assert fx["z"] == 100.0 and abs(fx["z_end"] - 99.0004) < 1e-12 and fx["dt"] == 2.67608e-07
    # the writer reproduces the reference's bytes
    import io
    buf = io.BytesIO()
    for k in hctest.FAB_ORDER:
        hctest.write_fab(buf, fabs[k], los[k])
    assert buf.getvalue() == raw


def _replay_lists(fx):
    fabs, los = fx["chunks"][0]
    return {k: v.copy() for k, v in fabs.items()}, los


def test_shim_build_equals_real_build(fx, built):
    """oracle/_ref (reference TUs + CVODE over the AMReX API shim, OpenMP NVector, one thread, the default 1024000 x 8 x 8 tiles) replays the
    snapshot to the same bits as the reference executable built with the real AMReX"""
    from oracle import pyref
    try:
        ref = pyref.Reference("omp")
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built")
    w, los = _replay_lists(fx)
    a, a_end = 1.0 / (1.0 + fx["z"]), 1.0 / (1.0 + fx["z_end"])
    ref.set("nyx.h_species", fx["inputs"]["nyx.h_species"])
    nthreads = ref.max_threads()
    ref.set("omp.num_threads", 1)   # the OpenMP NVector sums in thread order; the fixture was generated with one thread
    ref.stats_reset()
    ref.integrate_state_struct(fx["boxes"][0][0] + fx["boxes"][0][1], [w["s_old"]], [w["s_new"]], [w["diag"]], [w["hydro_src"]], [w["ir"]],
                               [w["reset_src"]], a, a_end, fx["dt"], 0, ng=(4, 1, 1, 0, 1, 4))
    box = fx["boxes"][0]
    for name in OUT_NAMES:
        got, want = w[name], fx["ref_out"][name]
        assert got.shape == want.shape
        assert np.array_equal(_valid(got, los[name], box), _valid(want, fx["ref_out_los"][name], box)), name
        assert got.tobytes() == want.tobytes(), name + " (ghost cells: uninitialised memory in the snapshot, compared as bytes)"
    st = ref.stats()
    ref.set("omp.num_threads", nthreads)
    assert st.shape[0] == 16 and (st[:, 0] == 3).all() and (st[:, 2] == 6).all() and (st[:, 6] == 3).all()   # 16 tiles x {nst 3, nfe 6, nfeLS 3}


def test_port_within_contract(fx, port):
    """the per-cell restatement against the reference's tile-coupled answer: 10 x rtol in e, and in T / Ne after the caller's compute_new_temp
    (the SDC path stores the last RHS evaluation's T, Ne: SURVEY 9.2)"""
    w, los = _replay_lists(fx)
    a, a_end = 1.0 / (1.0 + fx["z"]), 1.0 / (1.0 + fx["z_end"])
    lo, hi = fx["boxes"][0]
    order = ("s_old", "s_new", "diag", "hydro_src", "reset_src", "ir")
    pst = port.integrate_state_struct(*[w[k] for k in order], lo, hi, a, a_end, fx["dt"], 0,
                                      params=port.params(h_species=float(fx["inputs"]["nyx.h_species"])), los=[los[k] for k in order])
    assert (pst[:, 7] == 0).all()
    box = fx["boxes"][0]
    for comp in (EINT, EDEN):
        g = _valid(w["s_new"], los["s_new"], box)[comp]
        r = _valid(fx["ref_out"]["s_new"], fx["ref_out_los"]["s_new"], box)[comp]
        assert np.abs(g / r - 1).max() < 1e-3
    g = _valid(w["ir"], los["ir"], box)[0]
    r = _valid(fx["ref_out"]["ir"], fx["ref_out_los"]["ir"], box)[0]
    scale = np.abs(r).max()
    assert np.abs(g - r).max() < 1e-3 * scale


@pytest.mark.gpu
def test_gpu_replays_reference_snapshot(fx, hc_lib, port):
    """config 1 on the GPU: device-resident and host-buffer entry points replay the reference's snapshot; against the reference's own
    (tile-coupled) answer e, rho_E and I_R agree within 10 x rtol, against the per-cell port every CVODE counter is identical"""
    import torch
    from nyx_b200 import capi
    hc = hc_lib
    a, a_end = 1.0 / (1.0 + fx["z"]), 1.0 / (1.0 + fx["z_end"])
    lo, hi = fx["boxes"][0]
    box = fx["boxes"][0]
    prm = hc.default_params(**hctest.params_from_inputs(fx["inputs"]))
    order = ("s_old", "s_new", "diag", "hydro_src", "reset_src", "ir")
    # the port's answer (per cell)
    wp, los = _replay_lists(fx)
    pst = port.integrate_state_struct(*[wp[k] for k in order], lo, hi, a, a_end, fx["dt"], 0,
                                      params=port.params(h_species=float(fx["inputs"]["nyx.h_species"])), los=[los[k] for k in order])
    # device-resident FABs
    w, _ = _replay_lists(fx)
    dev = {k: torch.from_numpy(w[k]).cuda() for k in hctest.FAB_ORDER}
    csb = torch.zeros(32 ** 3 * 8, dtype=torch.int32, device="cuda")
    fabs = {k: [capi.fab_of_torch(dev[k], los[k])] for k in hctest.FAB_ORDER}
    st = hc.integrate_struct_batch(fabs["s_old"], fabs["diag"], fabs["s_new"], fabs["hydro_src"], fabs["reset_src"], fabs["ir"],
                                   [capi.make_box(lo, hi)], a, a_end, fx["dt"], 0, params=prm, cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    assert st.n_cells == 32 ** 3 and st.n_failed == 0
    cs = csb.cpu().numpy().view(capi.CELLSTAT_DTYPE)
    for i, f in enumerate(capi.CELLSTAT_FIELDS[:7]):
        assert np.array_equal(cs[f], pst[:, i]), f
    got = {k: dev[k].cpu().numpy() for k in OUT_NAMES}
    # host-buffer entry point: same bits as the device-resident call
    wh, _ = _replay_lists(fx)
    lists = {k: [capi.fab_of_numpy(wh[k], los[k])] for k in hctest.FAB_ORDER}
    sth = hc.integrate_struct_host(lists["s_old"], lists["diag"], lists["s_new"], lists["hydro_src"], lists["reset_src"], lists["ir"],
                                   [capi.make_box(lo, hi)], a, a_end, fx["dt"], 0, params=prm)
    assert sth.as_dict() == st.as_dict()
    for k in OUT_NAMES:
        assert wh[k].tobytes() == got[k].tobytes(), k
    # against the reference's answer and the port's
    for comp in (EINT, EDEN):
        g = _valid(got["s_new"], los["s_new"], box)[comp]
        r = _valid(fx["ref_out"]["s_new"], fx["ref_out_los"]["s_new"], box)[comp]
        p = _valid(wp["s_new"], los["s_new"], box)[comp]
        assert np.abs(g / r - 1).max() < 1e-3       # 10 x rtol against the coupled reference
        assert np.abs(g / p - 1).max() < 1e-5       # identical step sequences: libm last bits only
    g = _valid(got["ir"], los["ir"], box)[0]
    r = _valid(fx["ref_out"]["ir"], fx["ref_out_los"]["ir"], box)[0]
    assert np.abs(g - r).max() < 1e-3 * np.abs(r).max()
    # ghost cells of the outputs are untouched (compared as bytes: the snapshot holds uninitialised memory there)
    mask = np.ones((34, 34, 34), dtype=bool); mask[1:33, 1:33, 1:33] = False
    assert got["s_new"][:, mask].tobytes() == fx["chunks"][0][0]["s_new"][:, mask].tobytes()
    assert got["ir"][:, mask].tobytes() == fx["chunks"][0][0]["ir"][:, mask].tobytes()


# ---- SAVE_REACT (integrate_state_with_source_3d.cpp:126-183,602-631) -------------------------------------------------
def _react_reference():
    import json
    return json.load(open(os.path.join(FIX, "react_reference.json")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _port_react(fx, port, rr):
    w, los = _replay_lists(fx)
    lo, hi = fx["boxes"][0]
    a, a_end, dt = float.fromhex(rr["a"]), float.fromhex(rr["a_end"]), float.fromhex(rr["dt"])   # the in-situ values (inputs.0 holds them rounded to 6 digits)
    ri, ro, rw = np.zeros((7, 32, 32, 32)), np.zeros((7, 32, 32, 32)), np.zeros((9, 32, 32, 32))
    order = ("s_old", "s_new", "diag", "hydro_src", "reset_src", "ir")
    from oracle import pyref
    import ctypes as C
    p = port.params(h_species=float(fx["inputs"]["nyx.h_species"]))
    fabs = [pyref.fab_of(w[k], los[k]) for k in order] + [pyref.fab_of(x, lo) for x in (ri, ro, rw)]
    st = np.zeros((32 ** 3, len(pyref.STAT_FIELDS)), dtype=np.int64)
    l, h = port._box(lo, hi)
    port.lib.hco_integrate_state_struct_react(port.rp, C.byref(p), *[C.byref(f) for f in fabs], l, h, a, a_end, dt, 0, st.ctypes.data_as(C.c_void_p))
    return dict(react_in=ri, react_out=ro, react_out_work=rw, stats=st, a=a, a_end=a_end, dt=dt)


def test_port_save_react_equals_reference(fx, port):
    """the SAVE_REACT dumps of the port equal, BIT FOR BIT, those the reference built with USE_SAVE_REACT=TRUE wrote for this step
    (tests/golden/make_react_fixture.sh): all 7 + 7 components and the 7 assigned counters.  (In this smooth z = 100 field the tile-wide CVODE
    instance and the per-cell one take the same 3 steps.)  nje is compared with 0: the reference never assigns it (its plotfile holds stack garbage)"""
    rr = _react_reference()
    pr = _port_react(fx, port, rr)
    for c in range(7):
        assert _sha(pr["react_in"][c]) == rr["sha256"]["in"][c], rr["names"]["in"][c]
        assert _sha(pr["react_out"][c]) == rr["sha256"]["out"][c], rr["names"]["out"][c]
    for c in (0, 1, 2, 3, 4, 5, 7, 8):
        assert _sha(pr["react_out_work"][c]) == rr["sha256"]["out_work"][c], rr["names"]["out_work"][c]
    assert rr["unique"]["out_work"][0] == [3.0] and rr["unique"]["out_work"][2] == [6.0] and rr["unique"]["out_work"][8] == [3.0]
    assert not pr["react_out_work"][6].any()


@pytest.mark.gpu
def test_gpu_save_react_of_reference_snapshot(fx, hc_lib, port):
    """the SAVE_REACT entry points (device-resident and host-buffer) on the reference's snapshot: inputs-only components and counters bit for bit
    the reference's, CVODE's solution / T / the local error estimate against the port to round-off of the device's log10 / pow; the integration
    itself is unchanged by the dumps"""
    import torch
    from nyx_b200 import capi
    hc = hc_lib
    rr = _react_reference()
    pr = _port_react(fx, port, rr)
    a, a_end, dt = pr["a"], pr["a_end"], pr["dt"]
    lo, hi = fx["boxes"][0]
    prm = hc.default_params(**hctest.params_from_inputs(fx["inputs"]))
    w, los = _replay_lists(fx)
    dev = {k: torch.from_numpy(w[k]).cuda() for k in hctest.FAB_ORDER}
    rdev = {k: torch.full((n, 34, 34, 34), -7.0, dtype=torch.float64, device="cuda") for k, n in (("react_in", 7), ("react_out", 7), ("react_out_work", 9))}
    fabs = [[capi.fab_of_torch(dev[k], los[k])] for k in ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")]
    rfabs = [[capi.fab_of_torch(rdev[k], (-1, -1, -1))] for k in ("react_in", "react_out", "react_out_work")]
    st = hc.integrate_struct_react_batch(*fabs, *rfabs, [capi.make_box(lo, hi)], a, a_end, dt, 0, params=prm)
    torch.cuda.synchronize()
    assert st.n_cells == 32 ** 3 and st.n_failed == 0
    got = {k: v.cpu().numpy() for k, v in rdev.items()}
    inner = (slice(None), slice(1, 33), slice(1, 33), slice(1, 33))
    for k in got:   # ghost cells of the react FABs are not written
        mask = np.ones((34, 34, 34), dtype=bool); mask[1:33, 1:33, 1:33] = False
        assert (got[k][:, mask] == -7.0).all()
    gi, go, gw = got["react_in"][inner], got["react_out"][inner], got["react_out_work"][inner]
    for c in range(7):
        assert _sha(gi[c]) == rr["sha256"]["in"][c], rr["names"]["in"][c]
    for c in (1, 3, 5, 6):
        assert _sha(go[c]) == rr["sha256"]["out"][c], rr["names"]["out"][c]
    for c in (0, 1, 2, 3, 4, 5, 7, 8):
        assert _sha(gw[c]) == rr["sha256"]["out_work"][c], rr["names"]["out_work"][c]
    assert not gw[6].any()
    assert np.abs(go[0] / pr["react_out"][0] - 1).max() < 1e-9 and np.abs(go[2] / pr["react_out"][2] - 1).max() < 1e-9
    # the local error estimate is a cancellation residue of size ~1e-13 x abstol here: compared in units of the tolerance
    assert np.abs(go[4] - pr["react_out"][4]).max() < 1e-9 * gi[4].min()
    # same integration as the plain entry point
    w2, _ = _replay_lists(fx)
    dev2 = {k: torch.from_numpy(w2[k]).cuda() for k in hctest.FAB_ORDER}
    fabs2 = [[capi.fab_of_torch(dev2[k], los[k])] for k in ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")]
    st2 = hc.integrate_struct_batch(*fabs2, [capi.make_box(lo, hi)], a, a_end, dt, 0, params=prm)
    torch.cuda.synchronize()
    assert st2.as_dict() == st.as_dict()
    for k in OUT_NAMES:
        assert dev[k].cpu().numpy().tobytes() == dev2[k].cpu().numpy().tobytes(), k
    # host-buffer entry point: same bits
    wh, _ = _replay_lists(fx)
    rh = {k: np.full((n, 34, 34, 34), -7.0) for k, n in (("react_in", 7), ("react_out", 7), ("react_out_work", 9))}
    hl = [[capi.fab_of_numpy(wh[k], los[k])] for k in ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")]
    rl = [[capi.fab_of_numpy(rh[k], (-1, -1, -1))] for k in ("react_in", "react_out", "react_out_work")]
    sth = hc.integrate_struct_react_host(*hl, *rl, [capi.make_box(lo, hi)], a, a_end, dt, 0, params=prm)
    assert sth.as_dict() == st.as_dict()
    for k in rh:
        assert rh[k].tobytes() == got[k].tobytes(), k
    for k in OUT_NAMES:
        assert wh[k].tobytes() == dev[k].cpu().numpy().tobytes(), k
    # without the SDC sources the reference's dump code reads null pointers: refused
    with pytest.raises(Exception):
        hc.integrate_struct_react_batch(*fabs, *rfabs, [capi.make_box(lo, hi)], a, a_end, dt, -1, params=prm)
