#!/usr/bin/env python
"""Throughput + roofline of the SURVEY 8f rank-2 row on one GPU: Nyx::update_state_with_sources (+ enforce_minimum_density floor variant +
gravity) as the fused sweep hc_update_state_with_sources_batch, over `nb` boxes of n^3 cells with the production ghost widths.
HBM-bound: 216 algorithmic bytes per cell (21 doubles read, 6 written).
usage: bench_sources.py [n=128] [nb=16] [reps=7]  -> JSON lines: no cell below small_dens (the common case: one pass + an empty predicated
launch), cells below small_dens (two passes), host FABs end to end (pipelined H2D / kernel / D2H), MultiFab::Add of one component"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nyx_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
z = 3.0
hc = capi.NyxHC()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm_peak = peaks.get("hbm_gbs", 6650.0)
NG = (4, 1, 4, 0, 1)            # S_old_tmp, S_new, ext_src_old, hydro_src, grav_vector (sdc_hydro.cpp:60-106)
NC = (6, 6, 6, 6, 3)
a_old = 1.0 / (1.0 + z)
dt = synth.step_dt(z)
a_new = synth.a_after(z, dt)
small_dens = 1.0e-2 * synth.mean_rhob()
gen = torch.Generator(device="cuda").manual_seed(20240601)
lo, hi = (0, 0, 0), (n - 1,) * 3


def make(low):
    slots = []
    for g, nc in zip(NG, NC):
        m = n + 2 * g
        slots.append([torch.randn((nc, m, m, m), generator=gen, device="cuda", dtype=torch.float64) for _ in range(nb)])
    for b in range(nb):
        s = slots[0][b]
        s[0] = synth.mean_rhob() * torch.exp(s[0])                   # lognormal density
        s[4:6] = s[4:6].abs() * 1e12 * s[0]
        slots[3][b][0] = 0.1 * s[0, 4:-4, 4:-4, 4:-4] * slots[3][b][0].clamp(-3, 3)   # |hydro_src(rho)| <= 0.3 rho: no cell below small_dens
        slots[2][b][0] = 0.0
        if low:
            slots[3][b][0, 0, 0, :8] = -2.0 * s[0, 4, 4, 4:12]
    return slots


def fabs_of(slots, host=False):
    mk = capi.fab_of_numpy if host else capi.fab_of_torch
    return [[mk(x, (-g,) * 3) for x in slot] for slot, g in zip(slots, NG)]


tiles = [capi.make_box(lo, hi)] * nb
cells = nb * n ** 3
prm = hc.src_params(small_dens=small_dens, small_temp=1.0e-2)
big = torch.empty(160 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")    # > L2: flushed between repetitions


def timed(fn):
    ms = []
    for _ in range(reps + 2):
        big.fill_(1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms[2:]


def chained(fn, k=10):
    """k calls back to back: the GPU never idles, so the host side of a call (descriptor staging, launches) hides behind the previous
    call's kernel; the working set (>= 0.8 GB) is far larger than L2, so no flush is needed between the calls"""
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / k)
    return float(np.median(out))


for low in (0, 1):
    slots = make(low)
    f = fabs_of(slots)

    # the ctypes argument arrays are built once: the timed region is the C-ABI call alone (tile descriptors, launches), not Python marshalling
    ca = [hc._arr(x, capi.HcFab) for x in f] + [hc._arr(tiles, capi.HcBox)]

    def run():
        hc.check(hc.lib.hc_update_state_with_sources_batch(nb, ca[0], ca[1], ca[2], ca[3], ca[4], ca[5], dt, a_old, a_new, C.byref(prm), None, None))
    ms = timed(run)
    ms_chain = chained(run) if not low else None      # (with cells below small_dens a repeated call is not the same problem: hydro_src(rho) was rewritten)
    mn = hc.update_state_with_sources_batch(f[0], f[1], f[2], f[3], f[4], tiles, dt, a_old, a_new, prm)
    t = float(np.median(ms)) * 1e-3
    passes = 2 if low else 1
    print(json.dumps({"row": "Nyx::update_state_with_sources (hc_update_state_with_sources_batch)", "case": "cells below small_dens: second pass" if low else "no cell below small_dens",
                      "cells": cells, "boxes": nb, "n": n, "ghost": NG, "ms_median": float(np.median(ms)), "ms_best": float(min(ms)), "ms_all": [round(float(x), 3) for x in ms],
                      "cells_per_s": cells / t, "min_dens_over_small_dens": mn / small_dens,
                      "back_to_back": None if ms_chain is None else {"ms_per_call": ms_chain, "cells_per_s": cells / (ms_chain * 1e-3), "gbs": 216.0 * cells / (ms_chain * 1e-3) / 1e9,
                                                                     "frac": 216.0 * cells / (ms_chain * 1e-3) / 1e9 / hbm_peak},
                      "roofline": {"bound": "hbm", "achieved": 216.0 * passes * cells / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": 216.0 * passes * cells / t / 1e9 / hbm_peak, "bytes_per_cell": 216.0 * passes,
                                   "reference_three_sweeps_bytes_per_cell": 312.0}}))

# MultiFab::Add(ext_src_old, IR_tmp, 0, Eden_comp, 1, 0): 24 B per cell
slots = make(0)
f = fabs_of(slots)
ir = [torch.randn((1, n + 8, n + 8, n + 8), generator=gen, device="cuda", dtype=torch.float64) for _ in range(nb)]
fi = [capi.fab_of_torch(x, (-4,) * 3) for x in ir]
ca = [hc._arr(f[2], capi.HcFab), hc._arr(fi, capi.HcFab), hc._arr(tiles, capi.HcBox)]
ms = timed(lambda: hc.check(hc.lib.hc_fab_add_batch(nb, ca[0], 4, ca[1], 0, 1, ca[2], None)))
t = float(np.median(ms)) * 1e-3
ms_chain = chained(lambda: hc.check(hc.lib.hc_fab_add_batch(nb, ca[0], 4, ca[1], 0, 1, ca[2], None)))
print(json.dumps({"row": "MultiFab::Add one component (hc_fab_add_batch)", "cells": cells, "ms_median": float(np.median(ms)), "cells_per_s": cells / t,
                  "back_to_back": {"ms_per_call": ms_chain, "gbs": 24.0 * cells / (ms_chain * 1e-3) / 1e9, "frac": 24.0 * cells / (ms_chain * 1e-3) / 1e9 / hbm_peak},
                  "roofline": {"bound": "hbm", "achieved": 24.0 * cells / t / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": 24.0 * cells / t / 1e9 / hbm_peak,
                               "bytes_per_cell": 24.0}}))

# host FABs, end to end (pinned): H2D 27 components (S_new travels in for its ghost cells) + D2H 6 per cell
if "--no-host" not in sys.argv:
    hslots = [[x.cpu().pin_memory() for x in slot] for slot in slots]
    del slots
    hn = [[x.numpy() for x in slot] for slot in hslots]
    fh = fabs_of(hn, host=True)
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        hc.update_state_with_sources_batch(fh[0], fh[1], fh[2], fh[3], fh[4], tiles, dt, a_old, a_new, prm, host=True)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms[1:])) * 1e-3
    h2d = sum(x.numel() * 8 for slot in hslots for x in slot)
    d2h = sum(x.numel() * 8 for x in hslots[1])
    print(json.dumps({"row": "hc_update_state_with_sources_host (pinned host FABs)", "cells": cells, "ms_all": [round(float(x), 2) for x in ms], "cells_per_s": cells / t,
                      "h2d_bytes": h2d, "d2h_bytes": d2h, "pcie_gbs": (h2d + d2h) / t / 1e9}))
