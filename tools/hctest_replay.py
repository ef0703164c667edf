#!/usr/bin/env python
"""Replay a reference `hctest` snapshot (nyx.hctest_example_write = 1; format: nyx_b200/hctest.py) through the CUDA path, the way
Exec/HeatCoolTests replays it through Nyx::integrate_state_struct (Source/HeatCool/integrate_state_with_source_3d.cpp:95-125), and
optionally write the result back as a snapshot.   usage: hctest_replay.py <dir> <step> [--prefix P] [--sdc-iter K] [--out DIR]
(The comparison with the oracle lives in tests/test_hctest_format.py: product-side tools do not touch oracle/.)"""
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
    ap.add_argument("--prefix", default=""); ap.add_argument("--sdc-iter", type=int, default=0); ap.add_argument("--out", default=None)
    args = ap.parse_args()
    fx = hctest.read_fixture(args.dir, args.step, args.prefix)
    treecool = fx["inputs"].get("nyx.path_to_treecool", os.path.join(ROOT, "tests", "golden", "TREECOOL_middle")).strip('"')
    if not os.path.exists(treecool):
        treecool = os.path.join(ROOT, "tests", "golden", "TREECOOL_middle")
    hc = capi.NyxHC()
    hc.tables_upload(hc.tabulate_rates(treecool, synth.mean_rhob()))
    st, out = replay(hc, fx, args.sdc_iter)
    print(f"z {fx['z']:.6g} -> {fx['z_end']:.6g}, dt {fx['dt']:.6g}, {len(fx['boxes'])} box(es): {st.as_dict()}")
    for b, ((fabs, _), (got, _)) in enumerate(zip(fx["chunks"], out)):
        tgt = "s_new" if args.sdc_iter >= 0 else "s_old"
        de = got[tgt][5] - fabs[tgt][5]
        print(f"  box {b}: rho_e changed in {int((de != 0).sum())} cells, T range after {got['diag'][0].min():.4g} .. {got['diag'][0].max():.4g} K")
    if args.out:
        hctest.write_fixture(args.out, args.step, fx["boxes"], out, fx["inputs"], args.prefix)


if __name__ == "__main__":
    main()
