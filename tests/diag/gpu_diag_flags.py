"""TEST INFRASTRUCTURE (diagnostics; the only place besides tests/, smoke() and bench.py's CPU arms that touches oracle/).  GPU diagnostic: cells whose CVODE flag differs from the oracle in the struct parity cases."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nyx_b200 import capi, synth
from oracle import pyref
from tests import util
hc = capi.NyxHC(); hc.tables_upload(hc.tabulate_rates(pyref.TREECOOL, synth.mean_rhob()))
port = pyref.Port()
for (z, seed, src) in [(2.0, 22, 0.05), (6.0, 23, 0.2)]:
    n = 24
    d = util.sdc_inputs(z, n, seed, src)
    lo, hi = (0, 0, 0), (n - 1,) * 3
    names = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in names}
    csb = torch.zeros(n ** 3 * 8, dtype=torch.int32, device="cuda")
    st = hc.integrate_struct_batch(*[[capi.fab_of_torch(dev[k], lo)] for k in names], [capi.make_box(lo, hi)], d["a"], d["a_end"], d["dt"], 0, cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    ref = {k: d[k].copy() for k in names}
    pst = port.integrate_state_struct(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], lo, hi, d["a"], d["a_end"], d["dt"], 0)
    cs = csb.cpu().numpy().view(capi.CELLSTAT_DTYPE)
    bad = np.flatnonzero(cs["flag"] != pst[:, 7])
    print(f"z={z}: failed gpu {int((cs['flag']<0).sum())} oracle {int((pst[:,7]<0).sum())} stats.n_failed {st.n_failed}; differing flags {len(bad)}")
    for i in bad[:10]:
        print("   cell", i, "gpu", cs[i], "oracle", pst[i, :8], "e0", d["s_old"][5].ravel()[i] / d["s_old"][0].ravel()[i], "hs_e/rhoe", d["hydro_src"][5].ravel()[i] / d["s_old"][5].ravel()[i],
              "e_new gpu/ref", dev["s_new"].cpu().numpy()[5].ravel()[i], ref["s_new"][5].ravel()[i])
