"""ctypes binding of the C-ABI in include/nyx_hc.h (libnyx_hc.so, built in-tree by __graft_entry__.build()).

This is the host-side handle tests and bench.py use; a C++ host application links the same
library directly. There is no fallback: if the library is missing or no
CUDA device is present, every compute call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NYX_HC_LIB") or os.path.join(HERE, "csrc", "libnyx_hc.so")   # NYX_HC_LIB: build-variant experiments only
RATES_DOUBLES = 1 + 7 * 301 + 15 * 2001

_dp = C.POINTER(C.c_double)


class HcFab(C.Structure):
    _fields_ = [("p", C.c_void_p), ("jstride", C.c_longlong), ("kstride", C.c_longlong), ("nstride", C.c_longlong),
                ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("ncomp", C.c_int), ("pad_", C.c_int)]


class HcBox(C.Structure):
    _fields_ = [("lo", C.c_int * 3), ("hi", C.c_int * 3)]


class HcParams(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol_factor", C.c_double), ("h_species", C.c_double), ("gamma_minus_1", C.c_double),
                ("uvb_density_A", C.c_double), ("uvb_density_B", C.c_double), ("zhi_flash", C.c_double), ("zheii_flash", C.c_double),
                ("T_zhi", C.c_double), ("T_zheii", C.c_double), ("max_steps", C.c_longlong), ("old_max_steps", C.c_longlong),
                ("use_typical_steps", C.c_int), ("use_constraint", C.c_int), ("inhomo_reion", C.c_int), ("pad_", C.c_int)]


class HcSrcParams(C.Structure):
    _fields_ = [("small_dens", C.c_double), ("small_temp", C.c_double), ("gamma_minus_1", C.c_double), ("h_species", C.c_double),
                ("min_density_type", C.c_int), ("sdc", C.c_int)]


STATS_FIELDS = ("n_cells", "n_failed", "n_floor", "sum_nst", "max_nst", "sum_nfe", "sum_nfe_ls", "sum_netf", "sum_nni",
                "sum_ncfn", "sum_nsetups", "sum_ne_iters", "sum_attempts", "sum_eos")


class HcStats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in STATS_FIELDS]

    def as_dict(self):
        return {n: getattr(self, n) for n in STATS_FIELDS}


CELLSTAT_FIELDS = ("nst", "netf", "nfe", "nni", "ncfn", "nsetups", "nfe_ls", "flag")
CELLSTAT_DTYPE = np.dtype([(n, np.int32) for n in CELLSTAT_FIELDS])


def make_fab(ptr, lo, shape_xyz, ncomp):
    """HcFab over a buffer holding (ncomp, nz, ny, nx) doubles, x fastest, first cell = lo."""
    nx, ny, nz = shape_xyz
    f = HcFab()
    f.p = ptr
    f.jstride, f.kstride, f.nstride = nx, nx * ny, nx * ny * nz
    f.lo[:] = list(lo)
    f.hi[:] = [lo[0] + nx - 1, lo[1] + ny - 1, lo[2] + nz - 1]
    f.ncomp = ncomp
    return f


def fab_of_numpy(arr, lo):
    nc, nz, ny, nx = arr.shape
    assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
    return make_fab(arr.ctypes.data, lo, (nx, ny, nz), nc)


def fab_of_torch(t, lo):
    nc, nz, ny, nx = t.shape
    assert t.is_contiguous() and t.element_size() == 8
    return make_fab(t.data_ptr(), lo, (nx, ny, nz), nc)


def make_box(lo, hi):
    b = HcBox()
    b.lo[:] = list(lo)
    b.hi[:] = list(hi)
    return b


def default_params_struct(setter):
    p = HcParams()
    setter(C.byref(p))
    return p


def declare(lib, prefix="hc_"):
    fp, bp, pp, sp = C.POINTER(HcFab), C.POINTER(HcBox), C.POINTER(HcParams), C.POINTER(HcStats)
    lib.hc_last_error.restype = C.c_char_p
    lib.hc_version.restype = C.c_char_p
    lib.hc_default_params.argtypes = [pp]
    lib.hc_tabulate_rates.argtypes = [C.c_char_p, C.c_double, _dp]
    lib.hc_tables_upload.argtypes = [_dp, C.c_size_t]
    lib.hc_uvb_at_z.argtypes = [C.c_double, _dp]
    lib.hc_integrate_vec.argtypes = [fp, fp, HcBox, C.c_double, C.c_double, pp, sp, C.c_void_p, C.c_void_p]
    lib.hc_integrate_vec_batch.argtypes = [C.c_int, fp, fp, bp, C.c_double, C.c_double, pp, sp, C.c_void_p, C.c_void_p]
    lib.hc_integrate_struct.argtypes = [fp] * 6 + [HcBox, C.c_double, C.c_double, C.c_double, C.c_int, pp, sp, C.c_void_p, C.c_void_p]
    lib.hc_integrate_struct_batch.argtypes = [C.c_int] + [fp] * 6 + [bp, C.c_double, C.c_double, C.c_double, C.c_int, pp, sp,
                                                                      C.c_void_p, C.c_void_p]
    lib.hc_eos_T_given_Re.argtypes = [fp, fp, HcBox, C.c_double, pp, sp, C.c_void_p]
    lib.hc_compute_new_temp_batch.argtypes = [C.c_int, fp, fp, bp, C.c_double, pp, C.c_double, C.c_double, C.c_int, sp, C.c_void_p]
    lib.hc_reset_internal_energy_batch.argtypes = [C.c_int, fp, fp, fp, bp, C.c_double, pp, C.c_double, C.c_int, C.c_void_p]
    lib.hc_compute_new_temp_host.argtypes = [C.c_int, fp, fp, bp, C.c_double, pp, C.c_double, C.c_double, C.c_int, sp]
    lib.hc_reset_internal_energy_host.argtypes = [C.c_int, fp, fp, fp, bp, C.c_double, pp, C.c_double, C.c_int]
    lib.hc_integrate_vec_host.argtypes = [C.c_int, fp, fp, bp, C.c_double, C.c_double, pp, sp]
    lib.hc_integrate_struct_host.argtypes = [C.c_int] + [fp] * 6 + [bp, C.c_double, C.c_double, C.c_double, C.c_int, pp, sp]
    lib.hc_integrate_struct_react_batch.argtypes = [C.c_int] + [fp] * 9 + [bp, C.c_double, C.c_double, C.c_double, C.c_int, pp, sp,
                                                                            C.c_void_p, C.c_void_p]
    lib.hc_integrate_struct_react_host.argtypes = [C.c_int] + [fp] * 9 + [bp, C.c_double, C.c_double, C.c_double, C.c_int, pp, sp]
    qp = C.POINTER(HcSrcParams)
    lib.hc_default_src_params.argtypes = [qp]
    lib.hc_update_state_with_sources_batch.argtypes = [C.c_int] + [fp] * 5 + [bp, C.c_double, C.c_double, C.c_double, qp, _dp, C.c_void_p]
    lib.hc_enforce_minimum_density_batch.argtypes = [C.c_int] + [fp] * 5 + [bp, C.c_double, C.c_double, C.c_double, qp, C.c_void_p]
    lib.hc_enforce_minimum_density_host.argtypes = [C.c_int] + [fp] * 5 + [bp, C.c_double, C.c_double, C.c_double, qp]
    lib.hc_update_state_with_sources_host.argtypes = [C.c_int] + [fp] * 5 + [bp, C.c_double, C.c_double, C.c_double, qp, _dp]
    lib.hc_enforce_min_density_cons_iter_batch.argtypes = [C.c_int, fp, fp, fp, bp, qp, _dp, C.c_void_p]
    lib.hc_enforce_min_density_cons_iter_host.argtypes = [C.c_int, fp, fp, fp, bp, qp, _dp]
    lib.hc_finish_state_with_sources_batch.argtypes = [C.c_int] + [fp] * 5 + [bp, C.c_double, C.c_double, C.c_double, qp, C.c_int, C.c_void_p]
    lib.hc_finish_state_with_sources_host.argtypes = [C.c_int] + [fp] * 5 + [bp, C.c_double, C.c_double, C.c_double, qp, C.c_int]
    for name in ("hc_fab_copy_batch", "hc_fab_add_batch", "hc_fab_subtract_batch"):
        getattr(lib, name).argtypes = [C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, bp, C.c_void_p]
    lib.hc_init_zhi_batch.argtypes = [C.c_int, fp, fp, C.c_int, bp, C.c_void_p]
    lib.hc_init_zhi_host.argtypes = [C.c_int, fp, fp, C.c_int, bp]
    lib.hc_measure_fp64_peak.argtypes = [_dp]
    lib.hc_selftest_log10.argtypes = [_dp, _dp, C.POINTER(C.c_int), C.c_longlong]
    lib.hc_selftest_div_delta_t.argtypes = [_dp, _dp, C.c_longlong]
    lib.hc_last_launch_timing.argtypes = [_dp, _dp]
    lib.hc_sync.argtypes = [C.c_void_p]
    return lib


class HcError(RuntimeError):
    pass


class NyxHC:
    """Thin handle on libnyx_hc.so."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise HcError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        self.lib = declare(C.CDLL(path))

    def check(self, rc):
        if rc != 0:
            raise HcError(f"nyx_hc error {rc}: {self.lib.hc_last_error().decode()}")

    def default_params(self, **kw):
        p = HcParams()
        self.lib.hc_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def tabulate_rates(self, treecool, mean_rhob):
        out = np.zeros(RATES_DOUBLES)
        self.check(self.lib.hc_tabulate_rates(treecool.encode(), mean_rhob, out.ctypes.data_as(_dp)))
        return out

    def tables_upload(self, rates):
        rates = np.ascontiguousarray(rates, dtype=np.float64)
        self.check(self.lib.hc_tables_upload(rates.ctypes.data_as(_dp), rates.size))

    def uvb_at_z(self, z):
        out = np.zeros(6)
        self.check(self.lib.hc_uvb_at_z(z, out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def _arr(items, cls):
        a = (cls * len(items))()
        for i, x in enumerate(items):
            a[i] = x
        return a

    def integrate_vec_batch(self, state_fabs, diag_fabs, tiles, a, dt, params=None, cell_stats_ptr=None, stream=None, want_stats=True):
        p = params or self.default_params()
        st = HcStats()
        n = len(tiles)
        self.check(self.lib.hc_integrate_vec_batch(n, self._arr(state_fabs, HcFab), self._arr(diag_fabs, HcFab), self._arr(tiles, HcBox),
                                                   a, dt, C.byref(p), C.byref(st) if want_stats else None, cell_stats_ptr, stream))
        return st

    def integrate_struct_batch(self, s_old, diag, s_new, hydro_src, reset_src, ir, tiles, a, a_end, dt, sdc_iter=0, params=None,
                               cell_stats_ptr=None, stream=None, want_stats=True):
        p = params or self.default_params()
        st = HcStats()
        n = len(tiles)
        arrs = [self._arr(x, HcFab) for x in (s_old, diag, s_new, hydro_src, reset_src, ir)]
        self.check(self.lib.hc_integrate_struct_batch(n, *arrs, self._arr(tiles, HcBox), a, a_end, dt, sdc_iter, C.byref(p),
                                                      C.byref(st) if want_stats else None, cell_stats_ptr, stream))
        return st

    def integrate_struct_react_batch(self, s_old, diag, s_new, hydro_src, reset_src, ir, react_in, react_out, react_out_work, tiles, a, a_end, dt,
                                     sdc_iter=0, params=None, cell_stats_ptr=None, stream=None, want_stats=True):
        """the SAVE_REACT overload (Nyx.H:571-580): integrate_struct_batch + the three diagnostic FABs (7, 7, 9 components)"""
        p = params or self.default_params()
        st = HcStats()
        arrs = [self._arr(x, HcFab) for x in (s_old, diag, s_new, hydro_src, reset_src, ir, react_in, react_out, react_out_work)]
        self.check(self.lib.hc_integrate_struct_react_batch(len(tiles), *arrs, self._arr(tiles, HcBox), a, a_end, dt, sdc_iter, C.byref(p),
                                                            C.byref(st) if want_stats else None, cell_stats_ptr, stream))
        return st

    def integrate_struct_react_host(self, s_old, diag, s_new, hydro_src, reset_src, ir, react_in, react_out, react_out_work, tiles, a, a_end, dt,
                                    sdc_iter=0, params=None):
        p = params or self.default_params()
        st = HcStats()
        arrs = [self._arr(x, HcFab) for x in (s_old, diag, s_new, hydro_src, reset_src, ir, react_in, react_out, react_out_work)]
        self.check(self.lib.hc_integrate_struct_react_host(len(tiles), *arrs, self._arr(tiles, HcBox), a, a_end, dt, sdc_iter, C.byref(p),
                                                           C.byref(st)))
        return st

    def integrate_vec_host(self, state_fabs, diag_fabs, tiles, a, dt, params=None):
        p = params or self.default_params()
        st = HcStats()
        self.check(self.lib.hc_integrate_vec_host(len(tiles), self._arr(state_fabs, HcFab), self._arr(diag_fabs, HcFab),
                                                  self._arr(tiles, HcBox), a, dt, C.byref(p), C.byref(st)))
        return st

    def integrate_struct_host(self, s_old, diag, s_new, hydro_src, reset_src, ir, tiles, a, a_end, dt, sdc_iter=0, params=None):
        p = params or self.default_params()
        st = HcStats()
        arrs = [self._arr(x, HcFab) for x in (s_old, diag, s_new, hydro_src, reset_src, ir)]
        self.check(self.lib.hc_integrate_struct_host(len(tiles), *arrs, self._arr(tiles, HcBox), a, a_end, dt, sdc_iter, C.byref(p),
                                                     C.byref(st)))
        return st

    def eos_T_given_Re(self, state_fab, diag_fab, tile, a, params=None, stream=None):
        p = params or self.default_params()
        st = HcStats()
        self.check(self.lib.hc_eos_T_given_Re(C.byref(state_fab), C.byref(diag_fab), tile, a, C.byref(p), C.byref(st), stream))
        return st

    def measure_fp64_peak(self):
        v = C.c_double()
        self.check(self.lib.hc_measure_fp64_peak(C.byref(v)))
        return v.value

    def compute_new_temp_batch(self, state_fabs, diag_fabs, tiles, a, small_temp, large_temp, max_temp_dt, params=None, stream=None):
        p = params or self.default_params()
        st = HcStats()
        self.check(self.lib.hc_compute_new_temp_batch(len(tiles), self._arr(state_fabs, HcFab), self._arr(diag_fabs, HcFab), self._arr(tiles, HcBox),
                                                      a, C.byref(p), small_temp, large_temp, max_temp_dt, C.byref(st), stream))
        return st

    def reset_internal_energy_batch(self, state_fabs, diag_fabs, reset_fabs, tiles, a, small_temp, interp=0, params=None, stream=None):
        p = params or self.default_params()
        self.check(self.lib.hc_reset_internal_energy_batch(len(tiles), self._arr(state_fabs, HcFab), self._arr(diag_fabs, HcFab),
                                                           self._arr(reset_fabs, HcFab), self._arr(tiles, HcBox), a, C.byref(p), small_temp, interp, stream))

    def src_params(self, **kw):
        p = HcSrcParams()
        self.lib.hc_default_src_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def update_state_with_sources_batch(self, s_old, s_new, ext_src, hydro_src, grav, tiles, dt, a_old, a_new, params, want_min=True,
                                        stream=None, host=False):
        """Nyx::update_state_with_sources over all tiles; returns the minimum new density before the floor (None if not wanted)"""
        arrs = [self._arr(x, HcFab) for x in (s_old, s_new, ext_src, hydro_src, grav)]
        m = C.c_double(0.0)
        mp = C.byref(m) if want_min else None
        if host:
            self.check(self.lib.hc_update_state_with_sources_host(len(tiles), *arrs, self._arr(tiles, HcBox), dt, a_old, a_new, C.byref(params), mp))
        else:
            self.check(self.lib.hc_update_state_with_sources_batch(len(tiles), *arrs, self._arr(tiles, HcBox), dt, a_old, a_new, C.byref(params), mp,
                                                                   stream))
        return m.value if want_min else None

    def enforce_minimum_density_batch(self, s_old, s_new, ext_src, hydro_src, grav, tiles, dt, a_old, a_new, params, stream=None, host=False):
        arrs = [self._arr(x, HcFab) for x in (s_old, s_new, ext_src, hydro_src, grav)]
        if host:
            self.check(self.lib.hc_enforce_minimum_density_host(len(tiles), *arrs, self._arr(tiles, HcBox), dt, a_old, a_new, C.byref(params)))
            return
        self.check(self.lib.hc_enforce_minimum_density_batch(len(tiles), *arrs, self._arr(tiles, HcBox), dt, a_old, a_new, C.byref(params), stream))

    def enforce_min_density_cons_iter(self, sborder, s_new, reset_src, tiles, params, stream=None, host=False):
        """one iteration of Nyx::enforce_minimum_density_cons on a border-filled copy of S_new (two ghost cells); returns the new minimum density"""
        arrs = [self._arr(sborder, HcFab), self._arr(s_new, HcFab), self._arr(reset_src, HcFab) if reset_src is not None else None]
        m = C.c_double(0.0)
        if host:
            self.check(self.lib.hc_enforce_min_density_cons_iter_host(len(tiles), *arrs, self._arr(tiles, HcBox), C.byref(params), C.byref(m)))
        else:
            self.check(self.lib.hc_enforce_min_density_cons_iter_batch(len(tiles), *arrs, self._arr(tiles, HcBox), C.byref(params), C.byref(m), stream))
        return m.value

    def finish_state_with_sources_batch(self, s_old, s_new, ext_src, hydro_src, grav, tiles, dt, a_old, a_new, params, density_enforced,
                                        stream=None, host=False):
        """sweep (3) of Nyx::update_state_with_sources after the conservative iterations"""
        arrs = [self._arr(x, HcFab) for x in (s_old, s_new, ext_src, hydro_src, grav)]
        if host:
            self.check(self.lib.hc_finish_state_with_sources_host(len(tiles), *arrs, self._arr(tiles, HcBox), dt, a_old, a_new, C.byref(params),
                                                                  int(density_enforced)))
        else:
            self.check(self.lib.hc_finish_state_with_sources_batch(len(tiles), *arrs, self._arr(tiles, HcBox), dt, a_old, a_new, C.byref(params),
                                                                   int(density_enforced), stream))

    def fab_op_batch(self, op, dst, dcomp, src, scomp, ncomp, tiles, stream=None):
        """op in ("copy", "add", "subtract"): MultiFab::Copy / Add / Subtract of a component range over the tiles"""
        fn = getattr(self.lib, f"hc_fab_{op}_batch")
        self.check(fn(len(tiles), self._arr(dst, HcFab), dcomp, self._arr(src, HcFab), scomp, ncomp, self._arr(tiles, HcBox), stream))

    def init_zhi_batch(self, diag, zhi, ratio, tiles, stream=None):
        """Nyx::init_zhi cell loop: diag(i,j,k,2) = zhi(i/ratio, j/ratio, k/ratio)"""
        self.check(self.lib.hc_init_zhi_batch(len(tiles), self._arr(diag, HcFab), self._arr(zhi, HcFab), ratio, self._arr(tiles, HcBox), stream))

    def init_zhi_host(self, diag, zhi, ratio, tiles):
        """the same with HOST FABs (CPU build of the host application): staged through the device"""
        self.check(self.lib.hc_init_zhi_host(len(tiles), self._arr(diag, HcFab), self._arr(zhi, HcFab), ratio, self._arr(tiles, HcBox)))

    def selftest_log10(self, x):
        """log10 of a float64 array through the kernels' table-driven fast path -> (y, bad)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        bad = np.empty(x.size, dtype=np.int32)
        self.check(self.lib.hc_selftest_log10(x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), bad.ctypes.data_as(C.POINTER(C.c_int)), x.size))
        return y, bad

    def selftest_div_delta_t(self, x):
        """x / DELTA_T through the kernels' constant-divisor fast path"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.check(self.lib.hc_selftest_div_delta_t(x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), x.size))
        return y

    def last_launch_timing(self):
        """(kernel_ms, drain_ms) of this thread's last integrate_* launch that returned statistics"""
        k, d = C.c_double(0.0), C.c_double(0.0)
        self.check(self.lib.hc_last_launch_timing(C.byref(k), C.byref(d)))
        return k.value, d.value

    def sync(self, stream=None):
        self.check(self.lib.hc_sync(stream))
