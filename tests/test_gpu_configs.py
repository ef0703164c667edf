"""GPU parity at the sizes and regimes BASELINE.json's configs name: the CUDA path against BOTH forms of the oracle.

  * the per-cell port (oracle/hc_oracle.c, pinned bit for bit on the reference run with nyx.sundials_tile_size = 1 1 1), threaded over boxes
    on the host cores: per-cell CVODE counters, failure flags, e / T / ne;
  * oracle/_ref in the reference's DEFAULT mode (OpenMP NVector, one CVODE instance per 1024000 x 8 x 8 tile: the cells of a tile share step
    size and order, and the error test is an RMS over the tile): the "within 10 x rtol of the OpenMP reference" contract, with the cells
    beyond it counted -- and, for exactly those cells, a third integration at rtol = atol = 1e-9 that says which of the two is off.

config 2: 128^3, z = 3, Strang half-step.  config 3: 256^3, z = 2 (hot, stiff IGM).  UVB-off regime: z = 20 and z = 100 (interp_to_this_z returns
zeros above the last TREECOOL row, Source/EOS/eos_hc.H:17-27: every photo-ionisation numerator is exactly 0, Compton cooling dominates, <nst> ~ 40).
Config 1 (the reference's own 32^3 hctest snapshot) is tests/test_hctest_fixture.py; configs 4 / 5 are the same kernels on more boxes
(tests/test_gpu_parity.py::test_full_size_properties, bench.py).
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from nyx_b200 import capi, synth
from tests import util

pytestmark = pytest.mark.gpu
NTHREADS = max(1, min(32, os.cpu_count() or 1))


def _boxes(n, m):
    nb = n // m
    return [((i * m, j * m, k * m), (i * m + m - 1, j * m + m - 1, k * m + m - 1)) for k in range(nb) for j in range(nb) for i in range(nb)]


def _sub(arr, b):
    (lo, hi) = b
    return arr[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]


def _gpu_vec(hc, state, diag, boxes, a, dt, params=None):
    import torch
    s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
    ncell = sum(int(np.prod([h - l + 1 for l, h in zip(*b)])) for b in boxes)
    csb = torch.zeros(ncell * 8, dtype=torch.int32, device="cuda")
    sf, df = capi.fab_of_torch(s_dev, (0, 0, 0)), capi.fab_of_torch(d_dev, (0, 0, 0))
    st = hc.integrate_vec_batch([sf] * len(boxes), [df] * len(boxes), [capi.make_box(*b) for b in boxes], a, dt, params=params,
                                cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    return st, s_dev.cpu().numpy(), d_dev.cpu().numpy(), csb.cpu().numpy().view(capi.CELLSTAT_DTYPE)


def _port_vec(port, state, diag, boxes, a, dt, **kw):
    """the port over the boxes on all host cores (ctypes releases the GIL); stats concatenated in box order like the GPU's cell_stats"""
    with ThreadPoolExecutor(NTHREADS) as ex:
        sts = list(ex.map(lambda b: port.integrate_state_vec(state, diag, b[0], b[1], a, dt, fab_lo=(0, 0, 0), **kw), boxes))
    return np.concatenate(sts) if sts[0] is not None else None


def _same_counters(cs, pst):
    same = np.ones(len(cs), dtype=bool)
    for i, f in enumerate(capi.CELLSTAT_FIELDS[:7]):
        same &= cs[f] == pst[:, i]
    return same


def _cellwise(arr, boxes):
    """values of a (nz, ny, nx) field in cell_stats order (tiles concatenated, x fastest within a tile)"""
    return np.concatenate([arr[b[0][2]:b[1][2] + 1, b[0][1]:b[1][1] + 1, b[0][0]:b[1][0] + 1].ravel() for b in boxes])


@pytest.mark.parametrize("n,z,m,seed,coupled_boxes", [(128, 3.0, 64, 301, None), (256, 2.0, 64, 302, 16)])
def test_config_size_vec_vs_both_oracles(hc_lib, port, n, z, m, seed, coupled_boxes):
    a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
    boxes = _boxes(n, m)
    st, s_gpu, d_gpu, cs = _gpu_vec(hc_lib, state, diag, boxes, a, dt)
    # ---- (1) per-cell port: exact counters
    t0 = time.time()
    s_p, d_p = state.copy(), diag.copy()
    pst = _port_vec(port, s_p, d_p, boxes, a, dt)
    t_port = time.time() - t0
    same = _same_counters(cs, pst)
    assert np.array_equal(cs["flag"], pst[:, 7])
    assert st.n_cells == n ** 3 and st.n_failed == int((pst[:, 7] < 0).sum()) == 0
    e_rel = _cellwise(np.abs(s_gpu[5] / s_p[5] - 1), boxes)
    T_rel = _cellwise(np.abs(d_gpu[0] / d_p[0] - 1), boxes)
    ne_abs = _cellwise(np.abs(d_gpu[1] - d_p[1]), boxes)
    print(f"\n[config {n}^3 z={z}] port on {NTHREADS} threads {t_port:.1f} s; identical counters {same.sum()}/{len(same)} ({same.mean():.6f}); "
          f"same-sequence cells: max e {e_rel[same].max():.2e} T {T_rel[same].max():.2e} ne {ne_abs[same].max():.2e}; all cells: e {e_rel.max():.2e} T {T_rel.max():.2e}")
    assert same.mean() >= 0.9999
    assert max(e_rel[same].max(), T_rel[same].max(), ne_abs[same].max()) < 1e-5
    assert max(e_rel.max(), T_rel.max()) < 1e-3 and ne_abs.max() < 1e-3
    assert st.sum_nst == int(pst[:, 0].sum()) or same.mean() < 1.0
    for comp in (0, 1, 2, 3):
        assert np.array_equal(s_gpu[comp], state[comp])
    # ---- (2) the reference in its default tile-coupled OpenMP mode (a subsample of the boxes at 256^3, stated)
    from oracle import pyref
    try:
        ref = pyref.Reference("omp")
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built")
    ref.set("omp.num_threads", NTHREADS)
    pick = boxes if coupled_boxes is None else boxes[::max(1, len(boxes) // coupled_boxes)][:coupled_boxes]
    sb = [np.ascontiguousarray(_sub(state, b)) for b in pick]
    db = [np.ascontiguousarray(_sub(diag, b)) for b in pick]
    t0 = time.time()
    ref.integrate_state_vec([b[0] + b[1] for b in pick], sb, db, a, dt)
    t_ref = time.time() - t0
    eg = np.concatenate([np.abs(_sub(s_gpu, b)[5] / s[5] - 1).ravel() for s, b in zip(sb, pick)])
    Tg = np.concatenate([np.abs(_sub(d_gpu, b)[0] / d[0] - 1).ravel() for d, b in zip(db, pick)])
    ncell = eg.size
    beyond = np.flatnonzero((eg > 1e-3) | (Tg > 1e-3))
    print(f"[config {n}^3 z={z}] coupled OpenMP reference ({NTHREADS} threads, {len(pick)} of {len(boxes)} boxes, {ncell} cells) {t_ref:.1f} s = "
          f"{ncell / t_ref:.3e} cells/s; GPU vs it: max e {eg.max():.2e} T {Tg.max():.2e}; cells beyond 10 x rtol: {len(beyond)} ({len(beyond) / ncell:.2e})")
    # the bulk is inside the contract; the few cells beyond it are cells the tile-wide RMS error test under-resolves IN THE REFERENCE:
    assert len(beyond) / ncell < 5e-4 and np.median(eg) < 1e-5
    if len(beyond):
        # truth for exactly those cells: the per-cell port at rtol = atol = 1e-9
        cells = []
        off = 0
        for b in pick:
            mm = [h - l + 1 for l, h in zip(*b)]
            cnt = mm[0] * mm[1] * mm[2]
            for idx in beyond[(beyond >= off) & (beyond < off + cnt)] - off:
                k, j, i = idx // (mm[0] * mm[1]), (idx // mm[0]) % mm[1], idx % mm[0]
                cells.append((b, (b[0][0] + i, b[0][1] + j, b[0][2] + k)))
            off += cnt
        s_t, d_t = state.copy(), diag.copy()
        tight = port.params(rtol=1e-9, atol_factor=1e-9, max_steps=1000000)
        for _, c in cells:
            port.integrate_state_vec(s_t, d_t, c, c, a, dt, params=tight, fab_lo=(0, 0, 0), want_stats=False)
        err_gpu = np.array([abs(s_gpu[5][c[2], c[1], c[0]] / s_t[5][c[2], c[1], c[0]] - 1) for _, c in cells])
        sb_of = {id(b): s for s, b in zip(sb, pick)}
        err_ref = np.array([abs(sb_of[id(b)][5][c[2] - b[0][2], c[1] - b[0][1], c[0] - b[0][0]] / s_t[5][c[2], c[1], c[0]] - 1) for b, c in cells])
        print(f"[config {n}^3 z={z}] those {len(cells)} cells against an rtol = 1e-9 integration: GPU max {err_gpu.max():.2e}, coupled reference max {err_ref.max():.2e}")
        # (measured on B200: 128^3 z=3: 117 cells beyond, GPU max 2.4e-4 / reference max 1.75e-3 against the truth; 256^3 z=2, 16 of 64 boxes: 448 cells,
        #  GPU max 1.28e-3 / reference max 1.19e-2 -- the global error of rtol = 1e-4 can itself reach ~10 x rtol in the stiffest cells)
        assert err_gpu.max() < 2e-3 and (err_ref > err_gpu).mean() > 0.9 and err_ref.max() > err_gpu.max()


@pytest.mark.parametrize("path", ["vec", "struct"])
@pytest.mark.parametrize("z,seed", [(20.0, 311), (100.0, 312)])
def test_uvb_off_regime(hc_lib, port, path, z, seed):
    """z >= 15: the UV background is off.  Synthetic 1e3..1e7 K field: Compton cooling against the CMB makes the hot cells very stiff
    (<nst> ~ 40 at z = 100, a regime the z = 2..6 tests never enter)."""
    import torch
    n = 48
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    boxes = _boxes(n, 24)
    if path == "vec":
        a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
        state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
        st, s_gpu, d_gpu, cs = _gpu_vec(hc_lib, state, diag, boxes, a, dt)
        s_p, d_p = state.copy(), diag.copy()
        pst = _port_vec(port, s_p, d_p, boxes, a, dt)
        e_gpu, e_ref, rho, e0 = s_gpu[5], s_p[5], state[0], state[5]
        T_gpu, T_ref, ne_gpu, ne_ref = d_gpu[0], d_p[0], d_gpu[1], d_p[1]
    else:
        d = util.sdc_inputs(z, n, seed, 0.02)   # consistent S_new (see tests/test_host_logic.py on the source-free case at z = 100)
        names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
        dev = {k: torch.from_numpy(d[k]).cuda() for k in names}
        csb = torch.zeros(n ** 3 * 8, dtype=torch.int32, device="cuda")
        fabs = [[capi.fab_of_torch(dev[k], lo)] * len(boxes) for k in names]
        st = hc_lib.integrate_struct_batch(*fabs, [capi.make_box(*b) for b in boxes], d["a"], d["a_end"], d["dt"], 0, cell_stats_ptr=csb.data_ptr())
        torch.cuda.synchronize()
        cs = csb.cpu().numpy().view(capi.CELLSTAT_DTYPE)
        r = {k: d[k].copy() for k in names}
        with ThreadPoolExecutor(NTHREADS) as ex:
            sts = list(ex.map(lambda b: port.integrate_state_struct(r["s_old"], r["s_new"], r["diag"], r["hydro_src"], r["reset_src"], r["ir"], b[0], b[1],
                                                                    d["a"], d["a_end"], d["dt"], 0, los=[lo] * 6), boxes))
        pst = np.concatenate(sts)
        out = {k: dev[k].cpu().numpy() for k in names}
        assert np.array_equal(out["s_old"], d["s_old"])
        e_gpu, e_ref, rho, e0 = out["s_new"][5], r["s_new"][5], r["s_new"][0], d["s_old"][5] / d["s_old"][0] * r["s_new"][0]
        T_gpu, T_ref, ne_gpu, ne_ref = out["diag"][0], r["diag"][0], out["diag"][1], r["diag"][1]
        ir_rel = _cellwise(np.abs(out["ir"][0] - r["ir"][0]) / np.abs(r["ir"][0]).max(), boxes)
    same = _same_counters(cs, pst)
    ok = (pst[:, 7] == 0) & (cs["flag"] == 0)
    # e in units of the integrator's own error weight rtol |e| + atol (atol = 1e-4 e(t0)): cells that cool by orders of magnitude are held by
    # the ABSOLUTE tolerance in both integrations
    werr = _cellwise(np.abs(e_gpu - e_ref) / rho / (1e-4 * np.abs(e_ref / rho) + 1e-4 * np.abs(e0 / rho)), boxes)
    e_rel = _cellwise(np.abs(e_gpu / e_ref - 1), boxes)
    print(f"\n[UVB off z={z} {path}] <nst> {pst[:, 0].mean():.1f} max {pst[:, 0].max()}; identical counters {same.sum()}/{len(same)} ({same.mean():.6f}); flags equal "
          f"{np.array_equal(cs['flag'], pst[:, 7])}; failed {int((pst[:, 7] < 0).sum())}; same-sequence cells: max rel e {e_rel[same & ok].max():.2e}, weighted {werr[same & ok].max():.2e}; "
          f"all cells: weighted {werr[ok].max():.2e}")
    assert np.array_equal(cs["flag"] < 0, pst[:, 7] < 0) and st.n_failed == int((pst[:, 7] < 0).sum())
    if z <= 20.0:
        # measured: 110592/110592 identical counters on both paths; same-sequence cells agree to 7e-8 (Strang) / 7e-6 (SDC) in e
        assert same.mean() >= 0.9999
        assert werr[same & ok].max() < 0.1          # 1/10 of the tolerance on cells that followed the oracle's step sequence
        assert e_rel[same & ok].max() < 1e-5
        assert werr[ok].max() < 10.0                # 10 x the tolerance everywhere
        if path == "vec":
            T_rel = _cellwise(np.abs(T_gpu / T_ref - 1), boxes)
            ne_abs = _cellwise(np.abs(ne_gpu - ne_ref), boxes)
            assert T_rel[same & ok].max() < 1e-5 and ne_abs[same & ok].max() < 1e-5
        else:
            assert ir_rel[same & ok].max() < 1e-5
    else:
        # z = 100 with 1e3..1e7 K cells (not a physical Lyman-alpha state: the reference's own z = 100 step, 200 K gas, is the hctest fixture and
        # matches counter for counter): the hot cells are Compton-cooled by three orders of magnitude, h * df/de reaches ~1e3, and the
        # ~1e-7 noise of the right-hand side (the inner ne Newton solve stops at |dne| < 1e-6, so a last-bit difference of log10 can flip its
        # iteration count) is amplified to the level of the error tolerance itself.  Measured on B200: identical counters in 99.94 % (Strang) /
        # 99.96 % (SDC) of the cells, no failed cell on either side, same-sequence cells within 1.5 x the integrator's tolerance, all cells within 8 x.
        assert same.mean() >= 0.999
        assert werr[same & ok].max() < 3.0
        assert werr[ok].max() < 20.0
