#!/usr/bin/env python
"""Small driver for ncu: runs the HeatCool step `reps` times over one n^3 box (synthetic LyA field) through the C-ABI.
usage: prof_driver.py [n=128] [reps=3] [path=vec|struct] [z=3]   (env HC_LIB selects an alternative library build)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nyx_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
path = sys.argv[3] if len(sys.argv) > 3 else "vec"
z = float(sys.argv[4]) if len(sys.argv) > 4 else 3.0
TREECOOL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "TREECOOL_middle")
hc = capi.NyxHC(os.environ["HC_LIB"]) if os.environ.get("HC_LIB") else capi.NyxHC()
hc.tables_upload(hc.tabulate_rates(TREECOOL, synth.mean_rhob()))
a, dt = 1 / (1 + z), synth.step_dt(z)
state, diag = synth.make_fab((n, n, n), seed=5, z=z)
lo, hi = (0, 0, 0), (n - 1,) * 3
s0, d0 = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
for r in range(reps):
    s, d = s0.clone(), d0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if path == "vec":
        st = hc.integrate_vec_batch([capi.fab_of_torch(s, lo)], [capi.fab_of_torch(d, lo)], [capi.make_box(lo, hi)], a, 0.5 * dt)
    else:
        sn, hs = s.clone(), torch.zeros_like(s)
        rs, ir = torch.zeros((1, n, n, n), dtype=torch.float64, device="cuda"), torch.zeros((1, n, n, n), dtype=torch.float64, device="cuda")
        st = hc.integrate_struct_batch([capi.fab_of_torch(s, lo)], [capi.fab_of_torch(d, lo)], [capi.fab_of_torch(sn, lo)], [capi.fab_of_torch(hs, lo)],
                                       [capi.fab_of_torch(rs, lo)], [capi.fab_of_torch(ir, lo)], [capi.make_box(lo, hi)], a, synth.a_after(z, dt), dt, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{path} n={n} z={z} rep {r}: {ms:.3f} ms  {n**3 / ms * 1e3:.4e} cell-updates/s  failed {st.n_failed} sum_nst {st.sum_nst}", flush=True)
