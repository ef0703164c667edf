"""TEST INFRASTRUCTURE: reduces the reference's SAVE_REACT plotfiles (see make_react_fixture.sh) to tests/golden/hctest_lya32/react_reference.json.
usage: react_reference_digest.py <run directory of the SAVE_REACT build> <fixture directory>"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from nyx_b200 import hctest  # noqa: E402

run, fix = sys.argv[1], sys.argv[2]
old, new = hctest.read_fixture(fix, 0), hctest.read_fixture(os.path.join(run, "hctest"), 0)
assert old["boxes"] == new["boxes"]
(lo, hi) = old["boxes"][0]
for k in hctest.FAB_ORDER:   # same step: the valid regions of the two snapshots agree (ghost cells hold uninitialised memory)
    a, la = old["chunks"][0][0][k], old["chunks"][0][1][k]
    b, lb = new["chunks"][0][0][k], new["chunks"][0][1][k]
    va = a[:, lo[2] - la[2]:hi[2] - la[2] + 1, lo[1] - la[1]:hi[1] - la[1] + 1, lo[0] - la[0]:hi[0] - la[0] + 1]
    vb = b[:, lo[2] - lb[2]:hi[2] - lb[2] + 1, lo[1] - lb[1]:hi[1] - lb[1] + 1, lo[0] - lb[0]:hi[0] - lb[0] + 1]
    assert np.array_equal(va, vb), k
out = {"source": "Exec/LyA inputs.rt max_step=1, reference built with USE_SAVE_REACT=TRUE, OMP_NUM_THREADS=1", "names": {}, "sha256": {}, "unique": {}}
for nm in ("in", "out", "out_work"):
    d = os.path.join(run, f"plt_react_{nm}00000")
    header = open(os.path.join(d, "Header")).read().split("\n")
    ncomp = int(header[1])
    out["names"][nm] = header[2:2 + ncomp]
    buf = open(os.path.join(d, "Level_0", "Cell_D_00000"), "rb").read()
    arr, flo, pos = hctest.read_fab(buf, 0)
    assert pos == len(buf) and tuple(flo) == tuple(lo) and arr.shape == (ncomp, 32, 32, 32)
    out["sha256"][nm] = [hashlib.sha256(np.ascontiguousarray(arr[c]).tobytes()).hexdigest() for c in range(ncomp)]
    out["unique"][nm] = [(np.unique(arr[c]).tolist() if len(np.unique(arr[c])) <= 4 else None) for c in range(ncomp)]
    if nm == "in":
        out["a"] = float(arr[5].flat[0]).hex()
    if nm == "out":
        out["a_end"], out["dt"] = float(arr[5].flat[0]).hex(), float(arr[6].flat[0]).hex()
json.dump(out, open(os.path.join(fix, "react_reference.json"), "w"), indent=1)
print("wrote", os.path.join(fix, "react_reference.json"))
