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
if hasattr(hc.lib, "hc_debug_phase"):
    import ctypes
    buf = (ctypes.c_ulonglong * 48)()
    hc.lib.hc_debug_phase(buf)
    v = [int(x) for x in buf]
    rw = max(v[0], 1)
    tot = max(v[7], 1)
    print(f"phase timing (sum over {reps} reps): warp-rounds {v[0]}  cycles per warp-round {v[7] / rw:.0f}")
    v[1] += v[20] + v[21] + v[22]   # the sort phase: keys + counts, barrier, bases + order, barrier
    names = {1: "sort", 16: "B load", 17: "B resume+store", 18: "B refill", 19: "B writeback", 3: "B barrier wait", 4: "R work", 5: "R barrier wait"}
    for k, nm in names.items():
        print(f"  {nm:18s} {100.0 * v[k] / tot:6.2f} %   {v[k] / rw:9.0f} cycles per warp-round")
    print(f"  sort split: keys+counts {v[20] / rw:.0f}, barrier {v[21] / rw:.0f}, bases+order {v[22] / rw:.0f}, final barrier (+ first instructions of phase B) {(v[1] - v[20] - v[21] - v[22]) / rw:.0f} cycles per warp-round")
    print(f"  active lanes at R per warp-round: {v[6] / rw:.2f} / 32")
    keys = ["NEWTON", "SETUP_REQ", "LSETUP", "HIN", "INIT", "ETEST", "FINAL", "IDLE"]
    print("  lanes per key per warp-round: " + ", ".join(f"{k}={v[8 + i] / rw:.2f}" for i, k in enumerate(keys)))
    print("  B work of a warp by the sort key of its first lane (cycles per such warp-round, share of warp-rounds, active lanes in resume):")
    for i, k in enumerate(keys):
        if v[32 + i]:
            print(f"    {k:10s} {v[24 + i] / v[32 + i]:9.0f} cycles   {100.0 * v[32 + i] / rw:5.1f} % of warp-rounds   {v[40 + i] / v[32 + i]:5.1f} lanes")
if hasattr(hc.lib, "hc_debug_stage"):
    import ctypes
    buf = (ctypes.c_ulonglong * 16)()
    hc.lib.hc_debug_stage(buf)
    v = [int(x) for x in buf]
    tot = max(sum(v), 1)
    nm = ["stage 0 (handlers)", "HIN request", "Newton iteration", "Newton error", "NLS success + error test", "complete_step", "prepare_next_step",
          "cvStep tail rest", "after-hin", "STEP_TOP", "handle nflag", "predict", "set_coeffs", "attempt rest", "Newton top", "done/finalize"]
    print("resume() of the all-LSETUP warps, share of cycles per stage:")
    for i in range(16):
        print(f"    {nm[i]:26s} {100.0 * v[i] / tot:6.2f} %")
if hasattr(hc.lib, "hc_debug_mix"):
    import ctypes
    buf = (ctypes.c_ulonglong * 256)()
    hc.lib.hc_debug_mix(buf)
    v = [int(x) for x in buf]
    keys = ["NEWTON", "SETUP_REQ", "LSETUP", "HIN", "INIT", "ETEST", "FINAL", "IDLE"]
    nwr, nslow = max(sum(v[64:128]), 1), max(sum(v[128:192]), 1)
    print("bookkeeping phase by the classes of a warp's (first, last) lane: share of warp-rounds, mean cycles | share of the rounds in which it was the CTA's slowest warp, mean cycles then")
    for i in range(64):
        if v[64 + i] * 200 > nwr or v[128 + i] * 100 > nslow:
            a, b = keys[i // 8], keys[i % 8]
            slow = f"{100.0 * v[128 + i] / nslow:5.1f} %  {v[192 + i] / max(v[128 + i], 1):8.0f}" if v[128 + i] else "    -"
            print(f"    {a:10s} {b:10s} {100.0 * v[64 + i] / nwr:5.1f} %  {v[i] / max(v[64 + i], 1):8.0f} | {slow}")
    print(f"  mean cycles of the slowest warp of a CTA-round: {sum(v[192:256]) / nslow:.0f}")
