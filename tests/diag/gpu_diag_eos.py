import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nyx_b200 import capi, synth
from oracle import pyref
from tests import util
hc = capi.NyxHC(); hc.tables_upload(hc.tabulate_rates(pyref.TREECOOL, synth.mean_rhob()))
port = pyref.Port()
z, seed, src, n = 2.0, 22, 0.05, 24
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
e_rel = np.abs(out["s_new"][5] / ref["s_new"][5] - 1).ravel()
hc.eos_T_given_Re(capi.fab_of_torch(dev["s_new"], lo), capi.fab_of_torch(dev["diag"], lo), capi.make_box(lo, hi), d["a_end"])
torch.cuda.synchronize()
port.eos_box(ref["s_new"], ref["diag"], lo, hi, d["a_end"])
dg = dev["diag"].cpu().numpy()
rel = np.abs(dg[0] / ref["diag"][0] - 1).ravel() * same
for i in np.argsort(-rel)[:6]:
    print(i, "Trel", rel[i], "T gpu/ref", dg[0].ravel()[i], ref["diag"][0].ravel()[i], "e_rel", e_rel[i], "rhoe gpu/ref", out["s_new"][5].ravel()[i], ref["s_new"][5].ravel()[i],
          "rho", ref["s_new"][0].ravel()[i], "stats", cs[i])
