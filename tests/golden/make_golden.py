"""Generates tests/golden/heatcool_golden.npz from the REFERENCE ITSELF (oracle/_ref: Nyx HeatCool sources +
SUNDIALS CVODE compiled from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

The reference stores no known-answer vectors for this path, so these are the pinned
ones: inputs are regenerated from seeds by nyx_b200.synth / tests.util, outputs are what the unmodified
reference produced (serial-NVector build: norms are bit-reproducible), in both of its modes:
  percell : nyx.sundials_tile_size = 1 1 1  (one CVODE instance per cell -- the semantic twin of the CUDA path)
  coupled : default tile 1024000 x 8 x 8    (one CVODE instance per tile, shared step size/order)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nyx_b200 import synth  # noqa: E402
from oracle import pyref  # noqa: E402
from tests import util  # noqa: E402

VEC_CASES = [("vec_z3", 3.0, 8, 101), ("vec_z2", 2.0, 8, 102), ("vec_z6", 6.0, 8, 103)]
STRUCT_CASES = [("struct_z3", 3.0, 8, 201, 0.0, "none"), ("struct_z2_src", 2.0, 8, 202, 0.05, "none"),
                ("struct_hi_flash", 5.99, 8, 203, 0.05, "hi_now"), ("struct_heii_flash", 3.0, 8, 204, 0.05, "heii_now"),
                ("struct_before_reion", 7.0, 8, 205, 0.0, "before")]
FLASH_KEYS = {"zhi_flash": "nyx.reionization_zHI_flash", "zheii_flash": "nyx.reionization_zHeII_flash",
              "T_zhi": "nyx.reionization_T_zHI", "T_zheii": "nyx.reionization_T_zHeII"}


def main():
    ref = pyref.Reference("ser")
    out = {}
    rates = ref.rates()
    out["rates"] = rates
    out["uvb_z"] = np.array([0.0, 2.0, 2.999, 3.0, 5.5, 6.0, 9.9, 14.9, 15.5])
    uvb = np.zeros((len(out["uvb_z"]), 6))
    for i, z in enumerate(out["uvb_z"]):
        ref.lib.nyxref_interp_to_this_z(float(z), uvb[i].ctypes.data_as(pyref._dp))
    out["uvb"] = uvb
    # EOS known answers
    rng = np.random.default_rng(7)
    R = synth.mean_rhob() * np.exp(rng.standard_normal(64))
    T = 10.0 ** rng.uniform(1.0, 8.5, 64)
    e = synth.e_from_T(T)
    eos = np.zeros((64, 2))
    for i in range(64):
        eos[i] = ref.eos_T_given_Re(1, 1, float(R[i]), float(e[i]), 0.25, 2.0 / 3.0, 0.76)
    out["eos_R"], out["eos_e"], out["eos_TNe"] = R, e, eos

    for mode, tile in (("percell", "1 1 1"), ("coupled", "1024000 8 8")):
        ref.set("nyx.sundials_tile_size", tile)
        for name, z, n, seed in VEC_CASES:
            a, dt = 1.0 / (1.0 + z), 0.5 * synth.step_dt(z)
            state, diag = synth.make_fab((n, n, n), seed=seed, z=z)
            lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
            ref.stats_reset()
            ref.integrate_state_vec([lo + hi], [state], [diag], a, dt)
            out[f"{name}.{mode}.state"] = state[4:6].copy()
            out[f"{name}.{mode}.diag"] = diag.copy()
            out[f"{name}.{mode}.stats"] = ref.stats()
        for name, z, n, seed, src, flash in STRUCT_CASES:
            d = util.sdc_inputs(z, n, seed, src)
            for k, v in util.FLASH_CASES[flash].items():
                ref.set(FLASH_KEYS[k], v)
            lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
            ref.stats_reset()
            ref.integrate_state_struct([lo + hi], [d["s_old"]], [d["s_new"]], [d["diag"]], [d["hydro_src"]], [d["ir"]], [d["reset_src"]],
                                       d["a"], d["a_end"], d["dt"], 0)
            for k in util.FLASH_CASES[flash]:
                ref.unset(FLASH_KEYS[k])
            out[f"{name}.{mode}.s_new"] = d["s_new"][4:6].copy()
            out[f"{name}.{mode}.diag"] = d["diag"].copy()
            out[f"{name}.{mode}.ir"] = d["ir"].copy()
            out[f"{name}.{mode}.stats"] = ref.stats()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "heatcool_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
