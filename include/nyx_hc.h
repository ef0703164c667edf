/* nyx_hc.h -- C-ABI of the B200 heating-cooling (HeatCool) reaction integrator.
 *
 * This is the drop-in boundary of the hot path of AMReX-Astro/Nyx Lyman-alpha runs
 * (SURVEY.md section 8b). Plain C types only; FAB memory is described the way
 * amrex::Array4<Real> describes it, so a host shim can pass `MultiFab::array(mfi)` through
 * unchanged. Reference interfaces replaced (paths relative to the reference tree):
 *
 *   hc_tabulate_rates / hc_tables_upload   <- tabulate_rates()              Source/EOS/atomic_rates.H:10-166
 *                                             (called from Nyx::heatcool_setup, Source/Initialization/Nyx_setup.cpp:157-166)
 *   hc_integrate_vec[_batch]               <- Nyx::integrate_state_vec_mfin  Source/HeatCool/integrate_state_vec_3d.cpp:72-365
 *                                             (callers integrate_state_vec :44-70, integrate_state_grownvec :367-396)
 *   hc_integrate_struct[_batch]            <- Nyx::integrate_state_struct_mfin Source/HeatCool/integrate_state_with_source_3d.cpp:187-709
 *   hc_eos_T_given_Re                      <- nyx_eos_T_given_Re_device over a box  Source/EOS/eos_hc.H:190-220
 *   hc_compute_new_temp_batch              <- Nyx::compute_new_temp                 Source/Driver/Nyx.cpp:2435-2522
 *   hc_reset_internal_energy_batch         <- Nyx::reset_internal_energy            Source/Driver/Nyx.cpp:2356-2385, Source/EOS/reset_internal_e.H:16-68
 *   hc_update_state_with_sources_batch     <- Nyx::update_state_with_sources (SDC) Source/TimeStep/Nyx_update_state_with_sources.cpp:9-121
 *   hc_enforce_minimum_density_batch       <- Nyx::enforce_minimum_density, floor   Source/TimeStep/Nyx_enforce_minimum_density.cpp:8-107,
 *                                             floor_density                         Source/TimeStep/Nyx_enforce_minimum_density.H:8-58
 *   hc_enforce_min_density_cons_iter_batch <- Nyx::enforce_minimum_density_cons, one iteration    Nyx_enforce_minimum_density.cpp:190-236,
 *   hc_finish_state_with_sources_batch        compute_mu_for_enforce_min / create_update_for_minimum  Nyx_enforce_minimum_density.H:60-200
 *   hc_integrate_struct_react_batch        <- the SAVE_REACT overload of integrate_state_struct_mfin  Source/Driver/Nyx.H:571-580
 *   hc_fab_copy|add|subtract_batch         <- MultiFab::Copy / Add / Subtract around the call  Source/Hydro/sdc_hydro.cpp:83-84,94-95,112,135
 *   hc_init_zhi_batch                      <- Nyx::init_zhi, the cell loop            Source/Initialization/Nyx_initdata.cpp:163-213
 *   HcParams                               <- the nyx.* run-time flags of the path, Source/Driver/Nyx.cpp:116-181,
 *                                             Source/HeatCool/f_rhs_struct.H:45-101
 *   HcStats                                <- CVodeGetNum* counters (integrate_state_with_source_3d.cpp:755-790) and the
 *                                             CVode() return flag the reference ignores (:585), reduced over cells
 *
 * All pointers in HcFab are DEVICE pointers unless the function name ends in _host.
 * Every function returns 0 on success and a negative HC_ERR_* code otherwise; like the
 * reference (which ignores the CVode flag) a per-cell integrator failure is NOT an error:
 * it is counted in HcStats.n_failed. There is no CPU fallback: without a CUDA device every
 * compute entry point returns HC_ERR_CUDA.
 */
#ifndef NYX_HC_H
#define NYX_HC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HC_NCOOLFILE 301   /* rows of the TREECOOL UV-background file (atomic_rates_data.H:9)  */
#define HC_NCOOLTAB 2000   /* temperature table intervals, 2001 entries (atomic_rates_data.H:10) */
/* number of doubles in the reference's struct AtomicRates: mean_rhob, 7 x 301 UVB columns, 15 x 2001 rate tables */
#define HC_RATES_DOUBLES (1 + 7 * HC_NCOOLFILE + 15 * (HC_NCOOLTAB + 1))

enum {
    HC_OK = 0,
    HC_ERR_CUDA = -1,        /* no device / CUDA runtime error (message via hc_last_error) */
    HC_ERR_ARG = -2,         /* bad argument */
    HC_ERR_NO_TABLES = -3,   /* hc_tables_upload has not been called on this device */
    HC_ERR_IO = -4,          /* TREECOOL file missing or malformed */
    HC_ERR_TREECOOL_LEN = -5 /* more than HC_NCOOLFILE rows (the reference aborts, atomic_rates.H:37-51) */
};

/* amrex::Array4<Real> (AMReX_Array4.H:59-93): element (i,j,k,n) lives at
 * p[(i-lo[0]) + (j-lo[1])*jstride + (k-lo[2])*kstride + n*nstride]; hi is inclusive. */
typedef struct HcFab {
    double* p;
    long long jstride, kstride, nstride;
    int lo[3], hi[3];
    int ncomp;
    int pad_;
} HcFab;

/* amrex::Box, inclusive bounds: the tile the reference passes as `tbx` */
typedef struct HcBox {
    int lo[3], hi[3];
} HcBox;

typedef struct HcParams {
    double rtol;              /* nyx.sundials_reltol (1e-4) */
    double atol_factor;       /* nyx.sundials_abstol (1e-4): abstol_i = atol_factor * e_i(t0) */
    double h_species;         /* nyx.h_species (0.76) */
    double gamma_minus_1;     /* nyx.gamma - 1; used by the SDC path only (Strang hard-wires 2/3, f_rhs.H:133) */
    double uvb_density_A;     /* nyx.uvb_density_A (1.0), SDC path */
    double uvb_density_B;     /* nyx.uvb_density_B (0.0), SDC path */
    double zhi_flash;         /* nyx.reionization_zHI_flash  (-1 = off) */
    double zheii_flash;       /* nyx.reionization_zHeII_flash (-1 = off) */
    double T_zhi;             /* nyx.reionization_T_zHI */
    double T_zheii;           /* nyx.reionization_T_zHeII */
    long long max_steps;      /* CVodeSetMaxNumSteps (2000) */
    long long old_max_steps;  /* with use_typical_steps: CVodeSetMaxStep(dt / old_max_steps) */
    int use_typical_steps;    /* nyx.use_typical_steps (0) */
    int use_constraint;       /* nyx.use_sundials_constraint (0): CVodeSetConstraints(y > 0) */
    int inhomo_reion;         /* nyx.inhomo_reion (0): per-cell z_HI from diag component 2 */
    int pad_;
} HcParams;

typedef struct HcStats {
    long long n_cells;        /* cells integrated */
    long long n_failed;       /* cells whose CVODE-equivalent integration returned a flag < 0 */
    long long n_floor;        /* cells that hit the negative-energy floor (T = 10 K, ne = 0) */
    long long sum_nst;        /* internal BDF steps, summed / max over cells */
    long long max_nst;
    long long sum_nfe;        /* RHS evaluations (CVodeGetNumRhsEvals) */
    long long sum_nfe_ls;     /* RHS evaluations by the diagonal Jacobian setup (CVDiagGetNumRhsEvals) */
    long long sum_netf;       /* error-test failures */
    long long sum_nni;        /* Newton iterations */
    long long sum_ncfn;       /* nonlinear convergence failures (CVodeGetNumNonlinSolvConvFails) */
    long long sum_nsetups;    /* Jacobian setups */
    long long sum_ne_iters;   /* iterate_ne Newton iterations over all RHS evaluations and finalize solves */
    long long sum_attempts;   /* step attempts (predict + nonlinear solve) */
    long long sum_eos;        /* finalize EOS solves */
} HcStats;

/* optional per-cell record for parity checks, x-fastest over the tile (tiles concatenated for _batch) */
typedef struct HcCellStat {
    int nst, netf, nfe, nni, ncfn, nsetups, nfe_ls, flag;
} HcCellStat;

const char* hc_last_error(void);
const char* hc_version(void);

void hc_default_params(HcParams* p);

/* A1: build the reference's AtomicRates image (HC_RATES_DOUBLES doubles, same member order) on the host.
 * Pure host code, no device needed. */
int hc_tabulate_rates(const char* treecool_file, double mean_rhob, double* rates_out);
/* Upload a rates image to the current device (interleaved for the kernels); once per device. */
int hc_tables_upload(const double* rates, size_t n_doubles);
/* interp_to_this_z (eos_hc.H:10-49) on the host copy kept by hc_tables_upload: out6 = ggh0,gghe0,gghep,eh0,ehe0,ehep */
int hc_uvb_at_z(double z, double* out6);

/* Strang path: integrate de/dt over [0, dt] for every cell of `tile`, update state(Eint,Eden) and diag(Temp,Ne)
 * in place. stats may be NULL; cell_stats (device pointer) may be NULL. stream is a cudaStream_t (NULL = default). */
int hc_integrate_vec(const HcFab* state, const HcFab* diag, HcBox tile, double a, double dt, const HcParams* prm,
                     HcStats* stats, HcCellStat* cell_stats, void* stream);
int hc_integrate_vec_batch(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, double dt,
                           const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream);

/* SDC path (argument order of integrate_state_struct_mfin: state, diag, state_n, hydro_src, reset_src, IR) */
int hc_integrate_struct(const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                        const HcFab* reset_src, const HcFab* ir, HcBox tile, double a, double a_end, double dt, int sdc_iter,
                        const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream);
int hc_integrate_struct_batch(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                              const HcFab* reset_src, const HcFab* ir, const HcBox* tiles, double a, double a_end, double dt,
                              int sdc_iter, const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream);

/* The SAVE_REACT overload of Nyx::integrate_state_struct_mfin (Source/Driver/Nyx.H:571-580, integrate_state_with_source_3d.cpp:126-183,
 * 602-631; compiled with USE_SAVE_REACT = TRUE, Exec/Make.Nyx:51-52): the same integration, and per cell the three diagnostic FABs that
 * ode_eos_save_react_arrays (f_rhs_struct.H:213-267) fills right after the finalize step:
 *   react_in (7 components)        e(t0), rho(t0), rhoe_src, e_src, abstol, a, 0
 *   react_out (7)                  CVODE's solution (before the finalize step's floor / heating), rho(t0) + dt * rho_src, T and ne after the
 *                                  finalize step's last EOS solve, CVodeGetEstLocalErrors, a_end, dt
 *   react_out_work (9)             nst, netf, nfe, nni, ncfn, nsetups, 0, 0, nfeLS -- of the CELL (the reference reports the counters of the
 *                                  tile-wide CVODE instance in every cell of the tile; its nje / ncfl are never assigned)
 * Needs the SDC sources (sdc_iter >= 0: without them the reference reads null pointers here).  The caller writes the plotfiles. */
int hc_integrate_struct_react_batch(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                                    const HcFab* reset_src, const HcFab* ir, const HcFab* react_in, const HcFab* react_out,
                                    const HcFab* react_out_work, const HcBox* tiles, double a, double a_end, double dt, int sdc_iter,
                                    const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream);
int hc_integrate_struct_react_host(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                                   const HcFab* reset_src, const HcFab* ir, const HcFab* react_in, const HcFab* react_out,
                                   const HcFab* react_out_work, const HcBox* tiles, double a, double a_end, double dt, int sdc_iter,
                                   const HcParams* prm, HcStats* stats);

/* compute_new_temp core: diag(Temp,Ne) = EOS(state(Density), state(Eint)/state(Density), a) with JH = JHe = 1 */
int hc_eos_T_given_Re(const HcFab* state, const HcFab* diag, HcBox tile, double a, const HcParams* prm, HcStats* stats,
                      void* stream);

/* Nyx::compute_new_temp, the cell loop (Source/Driver/Nyx.cpp:2435-2522): for every cell of every tile, rho_e > 0: T, ne from the EOS at
 * e = rho_e * (1 / rho), clipped to large_temp when max_temp_dt == 1 (then rho_e, rho_E are rewritten); rho_e <= 0: T = small_temp, e from
 * nyx_eos_given_RT with the cell's current ne, rho_e and rho_E rewritten.  stats (may be NULL): n_cells, sum_eos, sum_ne_iters,
 * n_floor = cells reset to small_temp, n_failed = cells clipped at large_temp. */
int hc_compute_new_temp_batch(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, const HcParams* prm,
                              double small_temp, double large_temp, int max_temp_dt, HcStats* stats, void* stream);
/* Nyx::reset_internal_energy, the cell loop (Source/Driver/Nyx.cpp:2356-2385 with reset_internal_e, Source/EOS/reset_internal_e.H:16-68);
 * reset_src is the one-component reset_e_src FAB */
int hc_reset_internal_energy_batch(int ntiles, const HcFab* state, const HcFab* diag, const HcFab* reset_src, const HcBox* tiles, double a,
                                   const HcParams* prm, double small_temp, int interp, void* stream);

/* Host-buffer variants for CPU builds of the host application (and the end-to-end bench): same semantics,
 * HcFab.p are HOST pointers; the call stages H2D, runs, and stages the mutated components D2H. */
int hc_integrate_vec_host(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, double dt,
                          const HcParams* prm, HcStats* stats);
int hc_integrate_struct_host(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                             const HcFab* reset_src, const HcFab* ir, const HcBox* tiles, double a, double a_end, double dt,
                             int sdc_iter, const HcParams* prm, HcStats* stats);

int hc_compute_new_temp_host(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, const HcParams* prm,
                             double small_temp, double large_temp, int max_temp_dt, HcStats* stats);
int hc_reset_internal_energy_host(int ntiles, const HcFab* state, const HcFab* diag, const HcFab* reset_src, const HcBox* tiles, double a,
                                  const HcParams* prm, double small_temp, int interp);

/* ---- the SDC source assembly either side of sdc_reactions -------------------------------------------------------- */
enum { HC_MIN_DENSITY_FLOOR = 0, HC_MIN_DENSITY_CONSERVATIVE = 1 };
typedef struct HcSrcParams {
    double small_dens;        /* nyx.small_dens */
    double small_temp;        /* nyx.small_temp */
    double gamma_minus_1;     /* nyx.gamma - 1 */
    double h_species;         /* nyx.h_species */
    int min_density_type;     /* nyx.enforce_min_density_type: "floor" (default) = HC_MIN_DENSITY_FLOOR; "conservative" = HC_MIN_DENSITY_CONSERVATIVE moves
                                 density between neighbour cells and needs the host framework's FillPatch between its iterations: the update then
                                 runs as three calls (see hc_enforce_min_density_cons_iter_batch) */
    int sdc;                  /* 1: the reference's SDC build (enforce_minimum_density also resets hydro_src(rho)); 0: the non-SDC build */
} HcSrcParams;
void hc_default_src_params(HcSrcParams* p);

/* Nyx::update_state_with_sources (Source/TimeStep/Nyx_update_state_with_sources.cpp:9-121) for all tiles of this rank in one call:
 * S_new = source update of S_old (:33-76); Nyx::enforce_minimum_density, floor variant, decided by the minimum of the new density over
 * THE TILES OF THIS CALL (:79-84); gravity update (:90-120).  Six-component state FABs (CONST_SPECIES), grav 3 components; argument order
 * of the reference (reset_e_src is not touched by the floor variant and is not an argument).
 * min_dens (may be NULL): the minimum new density before the floor.  NULL: nothing is read back, the call is asynchronous on `stream`.
 * Several ranks: pass min_dens, reduce it over ranks (the reference's S_new.min() is a global reduction) and, where the global minimum is
 * below small_dens but the local one was not, call hc_enforce_minimum_density_batch with the same arguments. */
int hc_update_state_with_sources_batch(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                       const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                       double* min_dens, void* stream);
/* The enforce branch unconditionally: every cell recomputed from the (untouched) inputs with floor_density between the source and the
 * gravity updates; hydro_src(rho) = S_new(rho) - S_old(rho) when prm->sdc.  At most once per update (it rewrites hydro_src(rho)). */
int hc_enforce_minimum_density_batch(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                     const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                     void* stream);
int hc_update_state_with_sources_host(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                      const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                      double* min_dens);
int hc_enforce_minimum_density_host(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                    const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm);
/* nyx.enforce_min_density_type = "conservative" (Nyx::enforce_minimum_density_cons, Source/TimeStep/Nyx_enforce_minimum_density.cpp:101-330; per-cell
 * functions compute_mu_for_enforce_min / create_update_for_minimum, Nyx_enforce_minimum_density.H:60-200).  With prm->min_density_type =
 * HC_MIN_DENSITY_CONSERVATIVE the update runs as the reference's three sweeps, because the middle one exchanges ghost cells:
 *   1. hc_update_state_with_sources_*            sweep (1) only: S_new = source update of S_old; *min_dens = minimum of the new density
 *   2. while (the minimum over all ranks < small_dens and fewer than 10 iterations -- :179):
 *          FillPatch(Sborder: the 6 state components of S_new with TWO filled ghost cells)          <- the caller (AMReX)
 *          hc_enforce_min_density_cons_iter_*    ONE iteration (:190-236): S_new += div(mu grad Sborder) in the valid cells, reset_e_src =
 *                                                the (rho e) part of it (SDC build); *min_dens_after = the new minimum of this rank.
 *                                                A negative face coefficient (the reference aborts: "mu_x(i+1,j,k) < 0") returns HC_ERR_ARG
 *      (still below small_dens after the loop: the reference aborts, "Not able to enforce small_dens this way after all" -- the caller's decision)
 *   3. hc_finish_state_with_sources_*            sweep (3): gravity; density_enforced != 0 (step 2 ran) and prm->sdc: hydro_src(rho) = S_new(rho) - S_old(rho)
 * reset_src may be NULL when prm->sdc == 0. */
int hc_enforce_min_density_cons_iter_batch(int ntiles, const HcFab* sborder, const HcFab* s_new, const HcFab* reset_src, const HcBox* tiles,
                                           const HcSrcParams* prm, double* min_dens_after, void* stream);
int hc_enforce_min_density_cons_iter_host(int ntiles, const HcFab* sborder, const HcFab* s_new, const HcFab* reset_src, const HcBox* tiles,
                                          const HcSrcParams* prm, double* min_dens_after);
int hc_finish_state_with_sources_batch(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                       const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                       int density_enforced, void* stream);
int hc_finish_state_with_sources_host(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                      const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                      int density_enforced);
/* SURVEY 8f rank 4 -- Nyx::init_zhi, the cell loop (Source/Initialization/Nyx_initdata.cpp:198-209): diag(i,j,k,Zhi_comp = 2) =
 * zhi(i/ratio, j/ratio, k/ratio), zhi[t] = the one-component coarse reionization-redshift FAB that covers tile t coarsened by ratio
 * (the reference fills it with VisMF::Read + ParallelCopy; nyx_b200/nyxio.py reads the VisMF file) */
int hc_init_zhi_batch(int ntiles, const HcFab* diag, const HcFab* zhi, int ratio, const HcBox* tiles, void* stream);
int hc_init_zhi_host(int ntiles, const HcFab* diag, const HcFab* zhi, int ratio, const HcBox* tiles);
/* MultiFab::Copy / Add / Subtract (dst, src, scomp, dcomp, ncomp, nghost = tiles): dst(dcomp + n) {=, +=, -=} src(scomp + n) over the tiles */
int hc_fab_copy_batch(int ntiles, const HcFab* dst, int dcomp, const HcFab* src, int scomp, int ncomp, const HcBox* tiles, void* stream);
int hc_fab_add_batch(int ntiles, const HcFab* dst, int dcomp, const HcFab* src, int scomp, int ncomp, const HcBox* tiles, void* stream);
int hc_fab_subtract_batch(int ntiles, const HcFab* dst, int dcomp, const HcFab* src, int scomp, int ncomp, const HcBox* tiles, void* stream);

/* measured FP64 FMA throughput of the current device in FLOP/s (2 flops per DFMA), for roofline denominators */
int hc_measure_fp64_peak(double* flops_per_s);

/* self-test of the kernels' table-driven log10 (hc_device.cuh: fast_log10): y[i] = log10(x[i]) for n HOST doubles, evaluated on the device
 * by the same inline function the RHS fast path uses; bad[i] != 0 where the fast path would hand over to the library log10 */
int hc_selftest_log10(const double* x, double* y, int* bad, long long n);

/* self-test of the kernels' division by the table spacing (hc_device.cuh: div_delta_t, a three-instruction correctly rounded quotient by the
 * compile-time constant (TCOOLMAX - TCOOLMIN) / NCOOLTAB): y[i] = x[i] / DELTA_T for n HOST doubles, evaluated on the device */
int hc_selftest_div_delta_t(const double* x, double* y, long long n);

/* Device-side timing of the calling thread's most recent hc_integrate_*[_batch] launch that returned statistics (stats != NULL), from
 * %globaltimer: kernel_ms = first CTA start -> last CTA exit; drain_ms = first time a warp found the work queue empty -> last CTA exit (the tail
 * in which the SMs run out of cells one after the other: what strong scaling over many GPUs exposes) */
int hc_last_launch_timing(double* kernel_ms, double* drain_ms);

/* blocks until work queued on `stream` is done (cudaStreamSynchronize) */
int hc_sync(void* stream);

#ifdef __cplusplus
}
#endif
#endif
