#!/usr/bin/env python
"""TEST INFRASTRUCTURE (diagnostics; the only place besides tests/, smoke() and bench.py's CPU arms that touches oracle/).  diagnostic: inhomogeneous-reionization SDC step, device vs oracle, where do the per-cell counters differ?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nyx_b200 import capi, synth
from oracle import pyref
from tests import util
hc = capi.NyxHC()
hc.tables_upload(hc.tabulate_rates(pyref.TREECOOL, synth.mean_rhob()))
port = pyref.Port()
n, z = 20, 5.5
src = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
for inh in (1, 0):
    d = util.inhomo_inputs(z, n, 351, src_scale=src)
    kw = dict(d["kw"])
    if not inh:
        kw = {}
        d["diag"] = np.ascontiguousarray(d["diag"][:2])
    lo, hi = (0, 0, 0), (n - 1,) * 3
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in names}
    csb = torch.zeros(n ** 3 * 8, dtype=torch.int32, device="cuda")
    st = hc.integrate_struct_batch(*[[capi.fab_of_torch(dev[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0,
                                   params=hc.default_params(**kw), cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    ref = {k: d[k].copy() for k in names}
    pst = port.integrate_state_struct(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], lo, hi,
                                      d["a"], d["a_end"], d["dt"], 0, params=port.params(**kw))
    cs = csb.cpu().numpy().view(capi.CELLSTAT_DTYPE)
    same = np.ones(n ** 3, dtype=bool)
    for i, f in enumerate(capi.CELLSTAT_FIELDS[:7]):
        same &= cs[f] == pst[:, i]
    bad = np.flatnonzero(~same)
    e_gpu = (dev["s_new"].cpu().numpy()[5] / ref["s_new"][0]).ravel(); e_ref = (ref["s_new"][5] / ref["s_new"][0]).ravel()
    T0 = d["diag"][0].ravel(); rho = d["s_old"][0].ravel() / synth.mean_rhob()
    print(f"inhomo={inh}: {len(bad)} of {n**3} cells differ in counters; max rel e diff overall {np.abs(e_gpu/e_ref-1).max():.2e}, in same-counter cells {np.abs(e_gpu/e_ref-1)[same].max():.2e}")
    for b in bad[:12]:
        print("  cell", b, "gpu", [int(cs[f][b]) for f in capi.CELLSTAT_FIELDS], "oracle", pst[b, :8].tolist(), f"rel e diff {e_gpu[b]/e_ref[b]-1:.2e} T0 {T0[b]:.3g} rho/mean {rho[b]:.3g}",
              ("zhi %.3f" % d["diag"][2].ravel()[b]) if inh else "")
