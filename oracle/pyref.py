"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref (the reference compiled from
/root/reference by oracle/Makefile) and to oracle/_build/libhc_oracle.so (the C restatement).
May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never from the product package nyx_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TREE = "/root/reference"
TREECOOL = os.path.join(os.path.dirname(HERE), "tests", "golden", "TREECOOL_middle")

_dp = C.POINTER(C.c_double)
_dpp = C.POINTER(_dp)


def build(what="all"):
    subprocess.run(["make", "-s", "-C", HERE, what], check=True)


def _ptrs(arrs):
    a = (_dp * len(arrs))()
    for i, x in enumerate(arrs):
        assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
        a[i] = x.ctypes.data_as(_dp)
    return a


class Reference:
    """The real reference (Nyx HeatCool + SUNDIALS CVODE) behind ref_driver.cpp."""

    def __init__(self, variant="ser", treecool=TREECOOL, mean_rhob=None):
        path = os.path.join(HERE, "_ref", f"libnyxhc_ref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        lib.nyxref_init.argtypes = [C.c_char_p, C.c_double]
        lib.nyxref_rates.restype = _dp
        lib.nyxref_rates.argtypes = [C.POINTER(C.c_long)]
        lib.nyxref_set.argtypes = [C.c_char_p, C.c_char_p]
        lib.nyxref_unset.argtypes = [C.c_char_p]
        lib.nyxref_stats_get.argtypes = [C.POINTER(C.c_long)]
        lib.nyxref_stats_count.restype = C.c_long
        lib.nyxref_get_max_steps.restype = C.c_long
        lib.nyxref_integrate_state_vec.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, _dpp, _dpp,
                                                   C.c_double, C.c_double, C.c_int]
        lib.nyxref_integrate_state_struct.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int] + [_dpp] * 6 + \
            [C.c_double, C.c_double, C.c_double, C.c_int]
        lib.nyxref_ion_n.argtypes = [C.c_int, C.c_int] + [C.c_double] * 6 + [_dp]
        lib.nyxref_eos_T_given_Re.argtypes = [C.c_int, C.c_int] + [C.c_double] * 5 + [_dp, _dp]
        lib.nyxref_interp_to_this_z.argtypes = [C.c_double, _dp]
        if mean_rhob is None:
            from nyx_b200 import synth
            mean_rhob = synth.mean_rhob()
        self.mean_rhob = mean_rhob
        lib.nyxref_init(treecool.encode(), mean_rhob)

    def set(self, key, value):
        self.lib.nyxref_set(key.encode(), str(value).encode())

    def unset(self, key):
        self.lib.nyxref_unset(key.encode())

    def max_threads(self):
        return self.lib.nyxref_max_threads()

    def rates(self):
        n = C.c_long()
        p = self.lib.nyxref_rates(C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def stats_reset(self):
        self.lib.nyxref_stats_reset()

    def stats(self):
        """(n_instances, 8): nst, netf, nfe, nni, ncfn, nsetups, nfeLS, CVode flag — MFIter order."""
        n = self.lib.nyxref_stats_count()
        out = np.zeros((n, 8), dtype=np.int64)
        if n:
            self.lib.nyxref_stats_get(out.ctypes.data_as(C.POINTER(C.c_long)))
        return out

    @staticmethod
    def _boxes(boxes):
        b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int32).reshape(-1, 6))
        return b, b.ctypes.data_as(C.POINTER(C.c_int))

    def integrate_state_vec(self, boxes, state, diag, a, dt, ng_state=0, ng_diag=0, grown=False):
        b, bp = self._boxes(boxes)
        ncd = diag[0].shape[0]
        return self.lib.nyxref_integrate_state_vec(len(b), bp, ng_state, ng_diag, ncd, _ptrs(state), _ptrs(diag), a, dt, int(grown))

    def integrate_state_struct(self, boxes, s_old, s_new, d_old, hydro_src, ir, reset_src, a, a_end, dt, sdc_iter=0,
                               ng=(0, 0, 0, 0, 0, 0)):
        b, bp = self._boxes(boxes)
        ngarr = (C.c_int * 6)(*ng)
        ncd = d_old[0].shape[0]
        return self.lib.nyxref_integrate_state_struct(len(b), bp, ngarr, ncd, _ptrs(s_old), _ptrs(s_new), _ptrs(d_old),
                                                      _ptrs(hydro_src), _ptrs(ir), _ptrs(reset_src), a, a_end, dt, sdc_iter)

    def ion_n(self, JH, JHe, U, nh, ne, gm1, hsp, z):
        out = np.zeros(4)
        self.lib.nyxref_ion_n(JH, JHe, U, nh, ne, gm1, hsp, z, out.ctypes.data_as(_dp))
        return out

    def eos_T_given_Re(self, JH, JHe, R, e, a, gm1, hsp):
        T = C.c_double(0.0)
        Ne = C.c_double(0.0)
        self.lib.nyxref_eos_T_given_Re(JH, JHe, R, e, a, gm1, hsp, C.byref(T), C.byref(Ne))
        return T.value, Ne.value
