#!/usr/bin/env python
"""Throughput + roofline of the SURVEY 8f rank-1 rows on one GPU: Nyx::compute_new_temp (FP64-bound: one ionization-equilibrium
solve per cell) and Nyx::reset_internal_energy (HBM-bound streaming) over `nb` boxes of n^3 cells of the synthetic LyA field.
usage: bench_eos_rows.py [n=128] [nb=16] [reps=5]   -> one JSON line per row (device-resident, CUDA events; throughput from the median of reps)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nyx_b200 import capi, synth  # noqa: E402
from tests import util  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
z = 3.0
a = 1.0 / (1.0 + z)
hc = capi.NyxHC()
hc.tables_upload(hc.tabulate_rates(os.path.join(ROOT, "tests", "golden", "TREECOOL_middle"), synth.mean_rhob()))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm_peak = peaks.get("hbm_gbs", 6650.0)
fp64_peak = hc.measure_fp64_peak()
lo, hi = (0, 0, 0), (n - 1,) * 3
S0, D0, R0 = [], [], []
for b in range(nb):
    s, d, r = util.eos_rows_inputs(n, 900 + b, z)
    S0.append(torch.from_numpy(s).cuda()); D0.append(torch.from_numpy(d).cuda()); R0.append(torch.from_numpy(r).cuda())
tiles = [capi.make_box(lo, hi)] * nb
cells = nb * n ** 3


def timed(fn):
    ms = []
    st = None
    for _ in range(reps + 1):
        S, D, R = [x.clone() for x in S0], [x.clone() for x in D0], [x.clone() for x in R0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = fn(S, D, R)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms[1:], st


ms, st = timed(lambda S, D, R: hc.compute_new_temp_batch([capi.fab_of_torch(x, lo) for x in S], [capi.fab_of_torch(x, lo) for x in D], tiles,
                                                        a, 1.0e-2, 1.0e9, 0))
# algorithmic flops (SURVEY 8d hand count): iterate_ne(k) = 146 k + 69 flops + (2k+1) transcendentals of weight 20, + 10 per cell
flops = 186.0 * st.sum_ne_iters + 99.0 * st.sum_eos
t = float(np.median(ms)) * 1e-3
print(json.dumps({"row": "Nyx::compute_new_temp (hc_compute_new_temp_batch)", "cells": cells, "ms_median": float(np.median(ms)), "ms_best": float(min(ms)), "ms_all": [round(float(x), 3) for x in ms],
                  "cells_per_s": cells / t, "roofline": {"bound": "fp64", "achieved_tflops": flops / t / 1e12, "peak_tflops": fp64_peak / 1e12,
                                                        "frac": flops / t / fp64_peak, "flops_per_cell": flops / cells},
                  "hbm_gbs": 48.0 * cells / t / 1e9, "n_small_temp": st.n_floor, "ne_iters_per_cell": st.sum_ne_iters / max(st.sum_eos, 1)}))
ms, _ = timed(lambda S, D, R: hc.reset_internal_energy_batch([capi.fab_of_torch(x, lo) for x in S], [capi.fab_of_torch(x, lo) for x in D],
                                                            [capi.fab_of_torch(x, lo) for x in R], tiles, a, 1.0e-2, 0))
torch.cuda.synchronize()
t = float(np.median(ms)) * 1e-3
# algorithmic bytes per cell: reads rho, 3 momenta, rho E, rho e, ne, reset (8 x 8 B); writes reset + one or two of (rho e, rho E): ~2 x 8 B
bytes_cell = 80.0
print(json.dumps({"row": "Nyx::reset_internal_energy (hc_reset_internal_energy_batch)", "cells": cells, "ms_median": float(np.median(ms)), "ms_best": float(min(ms)), "ms_all": [round(float(x), 3) for x in ms],
                  "cells_per_s": cells / t, "roofline": {"bound": "hbm", "achieved": bytes_cell * cells / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                                        "frac": bytes_cell * cells / t / 1e9 / hbm_peak, "bytes_per_cell": bytes_cell}}))
