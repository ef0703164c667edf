"""Prints the parity statistics quoted in DESIGN.md (run on the GPU box: python tests/gpu_report.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nyx_b200 import capi, synth  # noqa: E402
from oracle import pyref  # noqa: E402

hc = capi.NyxHC()
hc.tables_upload(hc.tabulate_rates(pyref.TREECOOL, synth.mean_rhob()))
port = pyref.Port()
for z in (2.0, 3.0, 6.0):
    n = 48
    a, dt = 1 / (1 + z), 0.5 * synth.step_dt(z)
    state, diag = synth.make_fab((n, n, n), seed=100 + int(z), z=z)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    s_dev, d_dev = torch.from_numpy(state).cuda(), torch.from_numpy(diag).cuda()
    csb = torch.zeros(n ** 3 * 8, dtype=torch.int32, device="cuda")
    st = hc.integrate_vec_batch([capi.fab_of_torch(s_dev, lo)], [capi.fab_of_torch(d_dev, lo)], [capi.make_box(lo, hi)], a, dt,
                                cell_stats_ptr=csb.data_ptr())
    torch.cuda.synchronize()
    s_ref, d_ref = state.copy(), diag.copy()
    pst = port.integrate_state_vec(s_ref, d_ref, lo, hi, a, dt)
    cs = csb.cpu().numpy().view(capi.CELLSTAT_DTYPE)
    same = np.ones(n ** 3, dtype=bool)
    for i, f in enumerate(capi.CELLSTAT_FIELDS[:7]):
        same &= cs[f] == pst[:, i]
    e_rel = np.abs(s_dev.cpu().numpy()[5] / s_ref[5] - 1).ravel()
    T_rel = np.abs(d_dev.cpu().numpy()[0] / d_ref[0] - 1).ravel()
    ne_abs = np.abs(d_dev.cpu().numpy()[1] - d_ref[1]).ravel()
    print(f"z={z} n={n}^3: identical counters {same.sum()}/{n**3} ({same.mean():.6f}); nst equal {np.mean(cs['nst'] == pst[:, 0]):.6f}; "
          f"flags equal {np.array_equal(cs['flag'], pst[:, 7])}; same-seq cells: max e {e_rel[same].max():.2e} T {T_rel[same].max():.2e} "
          f"ne {ne_abs[same].max():.2e}; all cells: max e {e_rel.max():.2e} T {T_rel.max():.2e} ne {ne_abs.max():.2e}; "
          f"bitwise-equal e {np.mean(s_dev.cpu().numpy()[5] == s_ref[5]):.4f}; stats {st.as_dict()}")
