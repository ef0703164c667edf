"""Generates tests/golden/sources_golden.npz from the REFERENCE ITSELF (oracle/_ref): Nyx::update_state_with_sources of the reference's own
Source/TimeStep/Nyx_update_state_with_sources.cpp (compiled unmodified with -DSDC; enforce_minimum_density / floor variant restated around
the reference's floor_density, oracle/ref_driver.cpp) on a ragged three-box level with the production ghost widths.  Run in the build container:

    python tests/golden/make_sources_golden.py

Inputs are regenerated from seeds by tests.util.sources_inputs; stored are the reference's outputs S_new and hydro_src per box, for a case
without and a case with cells below small_dens."""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402
from tests import util  # noqa: E402

BOXES = [(0, 0, 0, 7, 5, 3), (8, 0, 0, 13, 5, 3), (0, 6, 0, 13, 8, 3)]
CASES = [("nofloor", 3.0, 701, 0), ("floor", 2.0, 702, 5)]


def inputs(z, seed, low):
    return util.sources_inputs(seed=seed, z=z, low_density_cells=low, boxes=BOXES)


def main():
    ref = pyref.Reference("ser")
    out = {}
    for name, z, seed, low in CASES:
        d = inputs(z, seed, low)
        r = copy.deepcopy(d)
        ref.update_state_with_sources(d["boxes"], r["s_old"], r["s_new"], r["ext_src"], r["hydro_src"], r["grav"], r["reset_src"], d["dt"],
                                      d["a_old"], d["a_new"], d["small_dens"], d["small_temp"], ng=d["ng"])
        for bi in range(len(BOXES)):
            out[f"{name}.s_new.{bi}"] = r["s_new"][bi]
            out[f"{name}.hydro_src.{bi}"] = r["hydro_src"][bi]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sources_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
