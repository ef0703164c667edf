#!/usr/bin/env python
"""BASELINE.json config 5 on one GPU: the redshift sweep z = 6 ... 2 through reionization heating on the SDC path
(Nyx::integrate_state_struct, flash reionization zHI = 6 / T = 2e4 K, zHeII = 3 / T = 1.5e4 K as in Exec/LyA/inputs:83-86), synthetic
lognormal field of n^3 cells with the survey's sigma(z).  One JSON line per redshift: device-resident cell-updates/s (CUDA events, median
of `reps` after one warm-up), work per cell, failed cells, fraction of the measured DFMA peak.
usage: redshift_sweep.py [n=256] [box=128] [reps=3]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nyx_b200 import capi, sharded, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
box = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
hc = capi.NyxHC(os.environ["HC_LIB"]) if os.environ.get("HC_LIB") else capi.NyxHC()   # HC_LIB: build-variant experiments
hc.tables_upload(hc.tabulate_rates(os.path.join(ROOT, "tests", "golden", "TREECOOL_middle"), synth.mean_rhob()))
peak = hc.measure_fp64_peak()
boxes = sharded.box_list(n, box)
prm = hc.default_params(zhi_flash=6.0, T_zhi=2.0e4, zheii_flash=3.0, T_zheii=1.5e4)
for z in (6.0, 5.5, 5.0, 4.5, 4.0, 3.5, 3.0, 2.5, 2.0):
    a, dt = 1.0 / (1.0 + z), synth.step_dt(z)
    a_end = synth.a_after(z, dt)
    S0, D0 = [], []
    for b, (lo, hi) in enumerate(boxes):
        s, d = synth.make_fab(tuple(h - l + 1 for l, h in zip(lo, hi)), seed=20240601 + b, z=z)
        S0.append(torch.from_numpy(s).cuda()); D0.append(torch.from_numpy(d).cuda())
    tiles = [capi.make_box(lo, hi) for lo, hi in boxes]
    ms, st = [], None
    for _ in range(reps + 1):
        S, Sn, D = [x.clone() for x in S0], [x.clone() for x in S0], [x.clone() for x in D0]
        H = [torch.zeros_like(x) for x in S0]
        R = [torch.zeros((1,) + tuple(x.shape[1:]), dtype=torch.float64, device="cuda") for x in S0]
        IR = [torch.zeros_like(x) for x in R]
        f = [[capi.fab_of_torch(x, lo) for x, (lo, hi) in zip(arrs, boxes)] for arrs in (S, D, Sn, H, R, IR)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = hc.integrate_struct_batch(f[0], f[1], f[2], f[3], f[4], f[5], tiles, a, a_end, dt, 0, params=prm)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms[1:])) * 1e-3
    flops = 186.0 * st.sum_ne_iters + 234.0 * (st.sum_nfe + st.sum_nfe_ls) + 60.0 * st.sum_attempts + 99.0 * st.sum_eos   # SURVEY 8d hand count
    Tn = torch.cat([x[0].flatten() for x in D])
    print(json.dumps({"z": z, "z_end": 1.0 / a_end - 1.0, "cells": st.n_cells, "ms_median": 1e3 * t, "ms_all": [round(float(x), 2) for x in ms[1:]],
                      "cell_updates_per_s": st.n_cells / t, "n_failed": st.n_failed, "n_floor": st.n_floor, "nst_per_cell": st.sum_nst / st.n_cells,
                      "max_nst": st.max_nst, "rhs_per_cell": (st.sum_nfe + st.sum_nfe_ls) / st.n_cells, "netf_per_cell": st.sum_netf / st.n_cells,
                      "flops_per_cell": flops / st.n_cells, "frac_of_dfma_peak": flops / t / peak, "T_median_after": float(Tn.median()),
                      "flash": "H" if (z > 6.0 - 1e-12 and 1.0 / a_end - 1.0 <= 6.0) or (1.0 / a_end - 1.0 <= 6.0 < z) else ("HeII" if 1.0 / a_end - 1.0 <= 3.0 < z or z == 3.0 else "")}), flush=True)
    del S0, D0, S, Sn, D, H, R, IR
    torch.cuda.empty_cache()
