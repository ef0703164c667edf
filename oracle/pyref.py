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


def _p(x):
    """pointer to the data of a numpy array, or of a torch tensor (device memory: the GPU flavour of the drop-in's test build takes device FABs)"""
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous() and x.element_size() == 8
        return C.cast(x.data_ptr(), _dp)
    assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
    return x.ctypes.data_as(_dp)


def _ptrs(arrs):
    a = (_dp * len(arrs))()
    for i, x in enumerate(arrs):
        a[i] = _p(x)
    return a


class Reference:
    """The real reference (Nyx HeatCool + SUNDIALS CVODE) behind ref_driver.cpp."""

    def __init__(self, variant="ser", treecool=TREECOOL, mean_rhob=None, path=None):
        # `path`: any library exporting the same nyxref_* driver entry points (tests/dropin_driver.cpp links them to the product)
        path = path or os.path.join(HERE, "_ref", f"libnyxhc_ref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        lib.nyxref_init.argtypes = [C.c_char_p, C.c_double]
        if hasattr(lib, "nyxref_rates"):
            lib.nyxref_rates.restype = _dp
            lib.nyxref_rates.argtypes = [C.POINTER(C.c_long)]
        lib.nyxref_set.argtypes = [C.c_char_p, C.c_char_p]
        lib.nyxref_unset.argtypes = [C.c_char_p]
        if hasattr(lib, "nyxref_stats_get"):
            lib.nyxref_stats_get.argtypes = [C.POINTER(C.c_long)]
            lib.nyxref_stats_count.restype = C.c_long
        lib.nyxref_get_max_steps.restype = C.c_long
        lib.nyxref_integrate_state_vec.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, _dpp, _dpp,
                                                   C.c_double, C.c_double, C.c_int]
        lib.nyxref_integrate_state_struct.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int] + [_dpp] * 6 + \
            [C.c_double, C.c_double, C.c_double, C.c_int]
        if hasattr(lib, "nyxref_ion_n"):
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
        if hasattr(self.lib, "nyxref_stats_reset"):
            self.lib.nyxref_stats_reset()

    def last_stats(self):
        """drop-in only: the HcStats counters of the last call"""
        out = (C.c_longlong * 14)()
        self.lib.nyxref_last_stats(out)
        return list(out)

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

    def compute_new_temp(self, box, state, diag, a, small_temp, large_temp, max_temp_dt, ng_state=0, ng_diag=0):
        """cell loop of Nyx::compute_new_temp over one box (restated around the reference's EOS functions, ref_driver.cpp)"""
        b, bp = self._boxes([box])
        self.lib.nyxref_compute_new_temp.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int]
        self.lib.nyxref_compute_new_temp(bp, ng_state, ng_diag, diag.shape[0], _p(state), _p(diag), a,
                                         small_temp, large_temp, max_temp_dt)

    def reset_internal_energy(self, box, state, diag, reset_src, a, small_temp, interp=0, ng_state=0, ng_diag=0, ng_reset=0):
        """cell loop of Nyx::reset_internal_energy over one box: the reference's reset_internal_e.H"""
        b, bp = self._boxes([box])
        self.lib.nyxref_reset_internal_energy.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int]
        self.lib.nyxref_reset_internal_energy(bp, ng_state, ng_diag, diag.shape[0], ng_reset, _p(state), _p(diag),
                                              _p(reset_src), a, small_temp, interp)

    def update_state_with_sources(self, boxes, s_old, s_new, ext_src, hydro_src, grav, reset_src, dt, a_old, a_new, small_dens, small_temp,
                                  ng=(0, 0, 0, 0, 0, 0)):
        """Nyx::update_state_with_sources (SDC signature): the reference's own Nyx_update_state_with_sources.cpp over a multi-box MultiFab;
        enforce_minimum_density (floor variant) restated around the reference's floor_density, see ref_driver.cpp"""
        b, bp = self._boxes(boxes)
        ngarr = (C.c_int * 6)(*ng)
        self.lib.nyxref_update_state_with_sources.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)] + [_dpp] * 6 + [C.c_double] * 5
        self.lib.nyxref_update_state_with_sources(len(b), bp, ngarr, _ptrs(s_old), _ptrs(s_new), _ptrs(ext_src), _ptrs(hydro_src), _ptrs(grav),
                                                  _ptrs(reset_src), dt, a_old, a_new, small_dens, small_temp)

    def enforce_min_cons_iter(self, sborder, s_new, reset_src, lo, hi, small_dens, ng_new=0, ng_rs=0, sdc=1):
        """one iteration of Nyx::enforce_minimum_density_cons on one box through the reference's own per-cell functions (ref_driver.cpp)"""
        box = (C.c_int * 6)(*lo, *hi)
        self.lib.nyxref_enforce_min_cons_iter.restype = C.c_double
        self.lib.nyxref_enforce_min_cons_iter.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_int]
        return self.lib.nyxref_enforce_min_cons_iter(box, ng_new, ng_rs, sborder.ctypes.data_as(_dp), s_new.ctypes.data_as(_dp),
                                                     _p(reset_src), small_dens, sdc)

    def ion_n(self, JH, JHe, U, nh, ne, gm1, hsp, z):
        out = np.zeros(4)
        self.lib.nyxref_ion_n(JH, JHe, U, nh, ne, gm1, hsp, z, out.ctypes.data_as(_dp))
        return out

    def eos_T_given_Re(self, JH, JHe, R, e, a, gm1, hsp):
        T = C.c_double(0.0)
        Ne = C.c_double(0.0)
        self.lib.nyxref_eos_T_given_Re(JH, JHe, R, e, a, gm1, hsp, C.byref(T), C.byref(Ne))
        return T.value, Ne.value


# --------------------------------------------------------------------------- the C restatement
class HcoFab(C.Structure):
    _fields_ = [("p", _dp), ("jstride", C.c_long), ("kstride", C.c_long), ("nstride", C.c_long),
                ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("ncomp", C.c_int)]


class HcoParams(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol_factor", C.c_double), ("h_species", C.c_double), ("gamma_minus_1", C.c_double),
                ("max_steps", C.c_long), ("use_constraint", C.c_int), ("use_typical_steps", C.c_int), ("old_max_steps", C.c_long),
                ("uvb_density_A", C.c_double), ("uvb_density_B", C.c_double), ("zhi_flash", C.c_double), ("zheii_flash", C.c_double),
                ("T_zhi", C.c_double), ("T_zheii", C.c_double), ("inhomo_reion", C.c_int)]


STAT_FIELDS = ("nst", "netf", "nfe", "nni", "ncfn", "nsetups", "nfeLS", "flag", "ne_iters", "attempts")
RATES_DOUBLES = 1 + 7 * 301 + 15 * 2001


def fab_of(arr, lo, cls=HcoFab):
    """arr: C-contiguous (ncomp, nz, ny, nx) float64 covering [lo, lo+shape-1]."""
    nc, nz, ny, nx = arr.shape
    f = cls()
    f.p = arr.ctypes.data_as(_dp)
    f.jstride, f.kstride, f.nstride = nx, nx * ny, nx * ny * nz
    f.lo[:] = list(lo)
    f.hi[:] = [lo[0] + nx - 1, lo[1] + ny - 1, lo[2] + nz - 1]
    f.ncomp = nc
    return f


class Port:
    """oracle/hc_oracle.c (scalar per-cell CVODE restatement)."""

    def __init__(self, treecool=TREECOOL, mean_rhob=None):
        path = os.path.join(HERE, "_build", "libhc_oracle.so")
        if not os.path.exists(path):
            build("port")
        self.lib = lib = C.CDLL(path)
        if mean_rhob is None:
            from nyx_b200 import synth
            mean_rhob = synth.mean_rhob()
        self.rates_buf = np.zeros(RATES_DOUBLES)
        self.rp = self.rates_buf.ctypes.data_as(C.c_void_p)
        lib.hco_tabulate_rates.argtypes = [C.c_char_p, C.c_double, C.c_void_p]
        rc = lib.hco_tabulate_rates(treecool.encode(), mean_rhob, self.rp)
        if rc != 0:
            raise RuntimeError(f"hco_tabulate_rates -> {rc}")
        lib.hco_default_params.argtypes = [C.POINTER(HcoParams)]
        lib.hco_ion_n.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 6 + [_dp]
        lib.hco_iterate_ne.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 5 + [_dp]
        lib.hco_eos_T_given_Re.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
        lib.hco_interp_to_this_z.argtypes = [C.c_void_p, C.c_double, _dp]
        lib.hco_f_rhs_rpar.restype = C.c_double
        lib.hco_f_rhs_rpar.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        fp = C.POINTER(HcoFab)
        ip = C.POINTER(C.c_int)
        lib.hco_integrate_state_vec.argtypes = [C.c_void_p, C.POINTER(HcoParams), fp, fp, ip, ip, C.c_double, C.c_double, C.c_void_p]
        lib.hco_integrate_state_struct_react.argtypes = [C.c_void_p, C.POINTER(HcoParams)] + [fp] * 9 + [ip, ip, C.c_double, C.c_double, C.c_double,
                                                                                                         C.c_int, C.c_void_p]
        lib.hco_integrate_state_struct.argtypes = [C.c_void_p, C.POINTER(HcoParams)] + [fp] * 6 + [ip, ip, C.c_double, C.c_double, C.c_double,
                                                                                                     C.c_int, C.c_void_p]
        lib.hco_eos_box.argtypes = [C.c_void_p, C.POINTER(HcoParams), fp, fp, ip, ip, C.c_double]

    def params(self, **kw):
        p = HcoParams()
        self.lib.hco_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def rates(self):
        return self.rates_buf.copy()

    def ion_n(self, JH, JHe, U, nh, ne, gm1, hsp, z):
        out = np.zeros(4)
        self.lib.hco_ion_n(self.rp, JH, JHe, U, nh, ne, gm1, hsp, z, out.ctypes.data_as(_dp))
        return out

    def eos_T_given_Re(self, JH, JHe, R, e, a, gm1, hsp):
        T = C.c_double(0.0)
        Ne = C.c_double(0.0)
        self.lib.hco_eos_T_given_Re(self.rp, gm1, hsp, JH, JHe, R, e, a, C.byref(T), C.byref(Ne))
        return T.value, Ne.value

    @staticmethod
    def _box(lo, hi):
        return (C.c_int * 3)(*lo), (C.c_int * 3)(*hi)

    def integrate_state_vec(self, state, diag, lo, hi, a, dt, params=None, fab_lo=None, diag_lo=None, want_stats=True):
        """state/diag: (ncomp, nz, ny, nx) arrays whose first cell is fab_lo/diag_lo (default lo)."""
        p = params or self.params()
        sf = fab_of(state, fab_lo or lo)
        df = fab_of(diag, diag_lo or fab_lo or lo)
        n = (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)
        st = np.zeros((n, len(STAT_FIELDS)), dtype=np.int64) if want_stats else None
        l, h = self._box(lo, hi)
        self.lib.hco_integrate_state_vec(self.rp, C.byref(p), C.byref(sf), C.byref(df), l, h, a, dt,
                                         st.ctypes.data_as(C.c_void_p) if want_stats else None)
        return st

    def integrate_state_struct(self, s_old, s_new, diag, hydro_src, reset_src, ir, lo, hi, a, a_end, dt, sdc_iter=0, params=None,
                               los=None, want_stats=True):
        p = params or self.params()
        los = los or [lo] * 6
        fabs = [fab_of(x, l) for x, l in zip((s_old, s_new, diag, hydro_src, reset_src, ir), los)]
        n = (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)
        st = np.zeros((n, len(STAT_FIELDS)), dtype=np.int64) if want_stats else None
        l, h = self._box(lo, hi)
        self.lib.hco_integrate_state_struct(self.rp, C.byref(p), *[C.byref(f) for f in fabs], l, h, a, a_end, dt, sdc_iter,
                                            st.ctypes.data_as(C.c_void_p) if want_stats else None)
        return st

    def integrate_state_struct_react(self, s_old, s_new, diag, hydro_src, reset_src, ir, react_in, react_out, react_out_work, lo, hi, a, a_end, dt,
                                     sdc_iter=0, params=None):
        """the SAVE_REACT build: integrate_state_struct + the three diagnostic FABs (7, 7, 9 components)"""
        p = params or self.params()
        fabs = [fab_of(x, lo) for x in (s_old, s_new, diag, hydro_src, reset_src, ir, react_in, react_out, react_out_work)]
        n = (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)
        st = np.zeros((n, len(STAT_FIELDS)), dtype=np.int64)
        l, h = self._box(lo, hi)
        self.lib.hco_integrate_state_struct_react(self.rp, C.byref(p), *[C.byref(f) for f in fabs], l, h, a, a_end, dt, sdc_iter,
                                                  st.ctypes.data_as(C.c_void_p))
        return st

    def compute_new_temp(self, state, diag, lo, hi, a, small_temp, large_temp, max_temp_dt, params=None):
        p = params or self.params()
        sf, df = fab_of(state, lo), fab_of(diag, lo)
        l, h = self._box(lo, hi)
        self.lib.hco_compute_new_temp_box.argtypes = [C.c_void_p, C.POINTER(HcoParams), C.POINTER(type(sf)), C.POINTER(type(sf)), type(l), type(l),
                                                      C.c_double, C.c_double, C.c_double, C.c_int]
        self.lib.hco_compute_new_temp_box(self.rp, C.byref(p), C.byref(sf), C.byref(df), l, h, a, small_temp, large_temp, max_temp_dt)

    def reset_internal_energy(self, state, diag, reset_src, lo, hi, small_temp, interp=0, params=None):
        p = params or self.params()
        sf, df, rf = fab_of(state, lo), fab_of(diag, lo), fab_of(reset_src, lo)
        l, h = self._box(lo, hi)
        self.lib.hco_reset_internal_e_box.argtypes = [C.POINTER(HcoParams), C.POINTER(type(sf)), C.POINTER(type(sf)), C.POINTER(type(sf)), type(l), type(l),
                                                      C.c_double, C.c_int]
        self.lib.hco_reset_internal_e_box(C.byref(p), C.byref(sf), C.byref(df), C.byref(rf), l, h, small_temp, interp)

    def update_state_with_sources(self, boxes, s_old, s_new, ext_src, hydro_src, grav, dt, a_old, a_new, small_dens, small_temp,
                                  ng=(0, 0, 0, 0, 0), sdc=1, params=None, global_min=None):
        """lists of per-box arrays (ncomp, nz, ny, nx), each covering its box grown by ng[slot]; slots: s_old, s_new, ext_src, hydro_src, grav.
        Returns the minimum of the new density over all boxes (before the floor); global_min overrides it for the decision (multi-rank)."""
        p = params or self.params()
        lib = self.lib
        fp = C.POINTER(HcoFab)
        l3 = C.c_int * 3
        lib.hco_sources_apply_box.restype = C.c_double
        lib.hco_sources_apply_box.argtypes = [fp] * 4 + [l3, l3, C.c_double, C.c_double, C.c_double]
        lib.hco_sources_finish_box.argtypes = [C.POINTER(HcoParams)] + [fp] * 4 + [l3, l3] + [C.c_double] * 5 + [C.c_int, C.c_int]
        fabs = []
        for bi, bx in enumerate(boxes):
            lo, hi = tuple(bx[:3]), tuple(bx[3:])
            f = [fab_of(arrs[bi], tuple(x - g for x in lo)) for arrs, g in zip((s_old, s_new, ext_src, hydro_src, grav), ng)]
            fabs.append((f, l3(*lo), l3(*hi)))
        m = min(lib.hco_sources_apply_box(C.byref(f[0]), C.byref(f[1]), C.byref(f[2]), C.byref(f[3]), lo, hi, dt, a_old, a_new)
                for f, lo, hi in fabs)
        decide = m if global_min is None else global_min
        for f, lo, hi in fabs:
            lib.hco_sources_finish_box(C.byref(p), C.byref(f[0]), C.byref(f[1]), C.byref(f[3]), C.byref(f[4]), lo, hi, dt, a_old, a_new,
                                       small_dens, small_temp, int(decide < small_dens), sdc)
        return m

    def enforce_min_cons_iter(self, sborder, s_new, reset_src, lo, hi, small_dens, ng_new=0, ng_rs=0, sdc=1):
        """one iteration of Nyx::enforce_minimum_density_cons on one box; sborder covers the box grown by 2 (filled), s_new / reset_src the box
        grown by ng_new / ng_rs.  Returns (new minimum density, faces with a negative coefficient)"""
        l3 = C.c_int * 3
        fp = C.POINTER(HcoFab)
        self.lib.hco_enforce_min_cons_iter_box.restype = C.c_double
        self.lib.hco_enforce_min_cons_iter_box.argtypes = [fp, fp, fp, l3, l3, C.c_double, C.c_int, C.POINTER(C.c_long)]
        bad = C.c_long(0)
        m = self.lib.hco_enforce_min_cons_iter_box(C.byref(fab_of(sborder, tuple(x - 2 for x in lo))), C.byref(fab_of(s_new, tuple(x - ng_new for x in lo))),
                                                   C.byref(fab_of(reset_src, tuple(x - ng_rs for x in lo))), l3(*lo), l3(*hi), small_dens, sdc, C.byref(bad))
        return m, bad.value

    def init_zhi(self, diag, diag_lo, zhi, zhi_lo, lo, hi, ratio):
        l3 = C.c_int * 3
        fp = C.POINTER(HcoFab)
        self.lib.hco_init_zhi_box.argtypes = [fp, fp, l3, l3, C.c_int]
        self.lib.hco_init_zhi_box(C.byref(fab_of(diag, diag_lo)), C.byref(fab_of(zhi, zhi_lo)), l3(*lo), l3(*hi), ratio)

    def eos_box(self, state, diag, lo, hi, a, params=None):
        p = params or self.params()
        sf, df = fab_of(state, lo), fab_of(diag, lo)
        l, h = self._box(lo, hi)
        self.lib.hco_eos_box(self.rp, C.byref(p), C.byref(sf), C.byref(df), l, h, a)
