#!/usr/bin/env python
"""Replay a reference `hctest` snapshot (nyx.hctest_example_write = 1; format: nyx_b200/hctest.py) through the CUDA path -- and, with
--oracle, through the per-cell oracle -- the way Exec/HeatCoolTests replays it through Nyx::integrate_state_struct
(Source/HeatCool/integrate_state_with_source_3d.cpp:95-125).   usage: hctest_replay.py <dir> <step> [--prefix P] [--oracle] [--sdc-iter K]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nyx_b200 import capi, hctest, synth  # noqa: E402


def replay(hc, fx, sdc_iter=0):
    """-> (HcStats, chunks after the step); FABs are staged through the pipelined host entry point."""
    a, a_end = 1.0 / (1.0 + fx["z"]), 1.0 / (1.0 + fx["z_end"])
    prm = hc.default_params(**hctest.params_from_inputs(fx["inputs"]))
    out = [({k: v.copy() for k, v in fabs.items()}, los) for fabs, los in fx["chunks"]]
    lists = {k: [capi.fab_of_numpy(fabs[k], los[k]) for fabs, los in out] for k in hctest.FAB_ORDER}
    tiles = [capi.make_box(lo, hi) for lo, hi in fx["boxes"]]
    st = hc.integrate_struct_host(lists["s_old"], lists["diag"], lists["s_new"], lists["hydro_src"], lists["reset_src"], lists["ir"], tiles,
                                  a, a_end, fx["dt"], sdc_iter, params=prm)
    return st, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir"); ap.add_argument("step", type=int)
    ap.add_argument("--prefix", default=""); ap.add_argument("--oracle", action="store_true"); ap.add_argument("--sdc-iter", type=int, default=0)
    args = ap.parse_args()
    fx = hctest.read_fixture(args.dir, args.step, args.prefix)
    treecool = fx["inputs"].get("nyx.path_to_treecool", os.path.join(ROOT, "tests", "golden", "TREECOOL_middle")).strip('"')
    if not os.path.exists(treecool):
        treecool = os.path.join(ROOT, "tests", "golden", "TREECOOL_middle")
    hc = capi.NyxHC()
    hc.tables_upload(hc.tabulate_rates(treecool, synth.mean_rhob()))
    st, out = replay(hc, fx, args.sdc_iter)
    print(f"z {fx['z']:.6g} -> {fx['z_end']:.6g}, dt {fx['dt']:.6g}, {len(fx['boxes'])} box(es): {st.as_dict()}")
    if args.oracle:
        from oracle import pyref
        port = pyref.Port()
        a, a_end = 1.0 / (1.0 + fx["z"]), 1.0 / (1.0 + fx["z_end"])
        prm = port.params(**hctest.params_from_inputs(fx["inputs"]))
        for b, ((lo, hi), (fabs, los)) in enumerate(zip(fx["boxes"], fx["chunks"])):
            ref = {k: v.copy() for k, v in fabs.items()}
            port.integrate_state_struct(ref["s_old"], ref["s_new"], ref["diag"], ref["hydro_src"], ref["reset_src"], ref["ir"], lo, hi,
                                        a, a_end, fx["dt"], args.sdc_iter, params=prm, los=[los[k] for k in ("s_old", "s_new", "diag", "hydro_src", "reset_src", "ir")],
                                        want_stats=False)
            got = out[b][0]
            tgt = "s_new" if args.sdc_iter >= 0 else "s_old"
            den = np.where(ref[tgt][5] != 0, np.abs(ref[tgt][5]), 1.0)
            print(f"  box {b}: max rel diff rho_e {np.max(np.abs(got[tgt][5] - ref[tgt][5]) / den):.2e}, I_R median abs diff {np.median(np.abs(got['ir'][0] - ref['ir'][0])):.2e}")


if __name__ == "__main__":
    main()
