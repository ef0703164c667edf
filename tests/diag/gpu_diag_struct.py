"""TEST INFRASTRUCTURE (diagnostics; the only place besides tests/, smoke() and bench.py's CPU arms that touches oracle/).  GPU diagnostic: worst same-sequence cells of the struct parity case (z=3, seed 21)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nyx_b200 import capi, synth
from oracle import pyref
from tests import util
hc = capi.NyxHC(); hc.tables_upload(hc.tabulate_rates(pyref.TREECOOL, synth.mean_rhob()))
port = pyref.Port()
for (z, seed, src) in [(3.0, 21, 0.0), (2.0, 22, 0.05)]:
    n = 24
    d = util.sdc_inputs(z, n, seed, src)
    lo, hi = (0, 0, 0), (n - 1,) * 3
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in names}
    csb = torch.zeros(n ** 3 * 8, dtype=torch.int32, device="cuda")
    hc.integrate_struct_batch(*[[capi.fab_of_torch(dev[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0, cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    ref = {k: d[k].copy() for k in names}
    pst = port.integrate_state_struct(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], lo, hi, d["a"], d["a_end"], d["dt"], 0)
    cs = csb.cpu().numpy().view(capi.CELLSTAT_DTYPE)
    same = np.ones(n ** 3, bool)
    for i, f in enumerate(capi.CELLSTAT_FIELDS[:7]):
        same &= cs[f] == pst[:, i]
    out = {k: dev[k].cpu().numpy() for k in names}
    T_rel = np.abs(out["diag"][0] / ref["diag"][0] - 1).ravel()
    e_rel = np.abs(out["s_new"][5] / ref["s_new"][5] - 1).ravel()
    ne_abs = np.abs(out["diag"][1] - ref["diag"][1]).ravel()
    print(f"z={z}: same {same.mean():.5f}; same-seq max T_rel {T_rel[same].max():.3e} e_rel {e_rel[same].max():.3e} ne_abs {ne_abs[same].max():.3e}; n(T_rel>1e-6 & same) {(same & (T_rel > 1e-6)).sum()}")
    idx = np.argsort(-(T_rel * same))[:6]
    for i in idx:
        print("   cell", i, "T gpu/ref", out["diag"][0].ravel()[i], ref["diag"][0].ravel()[i], "ne", out["diag"][1].ravel()[i], ref["diag"][1].ravel()[i],
              "e_rel", e_rel[i], "rho/mean", d["s_old"][0].ravel()[i] / synth.mean_rhob(), "T0", d["diag"][0].ravel()[i], "stats", cs[i], pst[i, :8])
