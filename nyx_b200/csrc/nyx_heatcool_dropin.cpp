// nyx_heatcool_dropin.cpp -- host side of the B200 HeatCool path: the definitions of
//
//     Nyx::integrate_state_vec        Nyx::integrate_state_grownvec      Nyx::integrate_state_vec_mfin
//     Nyx::integrate_state_struct     Nyx::integrate_state_struct_mfin
//
// with the reference's exact signatures (Source/Driver/Nyx.H:549-580).  This translation unit REPLACES
// Source/HeatCool/integrate_state_vec_3d.cpp and integrate_state_with_source_3d.cpp in a Nyx build, the same way the
// reference's *_stubs.cpp files replace them (Source/HeatCool/Make.package:11-22); the callers strang_first_step /
// strang_second_step / sdc_reactions / advance_heatcool compile unchanged.  No SUNDIALS, no AMReX ParallelFor: the MFIter
// loop only COLLECTS Array4 views, and one C-ABI call (include/nyx_hc.h) integrates all local boxes in one persistent launch.
//
// Differences a maintainer should know:
//   * every cell is integrated by its own BDF instance (== the reference run with nyx.sundials_tile_size = 1 1 1): the
//     CVODE tile size and nyx.sundials_use_tiling no longer influence results;
//   * the CVode flag the reference ignores (integrate_state_vec_3d.cpp:284) is counted per cell (nyx_hc_last_stats());
//   * nyx.use_sundials_fused / sundials_alloc_type / sundials_atomic_reductions are accepted and ignored;
//   * FAB memory must be device-accessible when AMReX is built with AMREX_USE_GPU; in a CPU build of AMReX the host-buffer
//     entry points stage the touched components through pinned memory.
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#include <Nyx.H>

#include <algorithm>
#include <vector>

#include "nyx_hc.h"

using namespace amrex;

namespace {

HcStats g_last_stats;   // diagnostics of the most recent call on this rank

HcFab to_fab(Array4<Real> const& a) {
    HcFab f{};
    f.p = a.p;
    f.jstride = a.jstride; f.kstride = a.kstride; f.nstride = a.nstride;
    f.lo[0] = a.begin.x; f.lo[1] = a.begin.y; f.lo[2] = a.begin.z;
    f.hi[0] = a.end.x - 1; f.hi[1] = a.end.y - 1; f.hi[2] = a.end.z - 1;   // Array4::end is exclusive
    f.ncomp = a.ncomp;
    return f;
}

HcBox to_box(const Box& b) {
    HcBox r;
    for (int d = 0; d < 3; ++d) { r.lo[d] = b.smallEnd(d); r.hi[d] = b.bigEnd(d); }
    return r;
}

// the nyx.* run-time flags of the path: Nyx statics (Source/Driver/Nyx.cpp:116-181) + what ode_eos_setup re-parses on every
// tile call in the reference (Source/HeatCool/f_rhs_struct.H:45-101); parsed once per call here
HcParams params_from_nyx(long int old_max_steps) {
    HcParams p;
    hc_default_params(&p);
    p.rtol = Nyx::sundials_reltol;
    p.atol_factor = Nyx::sundials_abstol;
    p.h_species = Nyx::h_species;
    p.gamma_minus_1 = Nyx::gamma - 1.0;
    p.max_steps = 2000;                               // CVodeSetMaxNumSteps(cvode_mem, 2000)
    p.use_typical_steps = Nyx::use_typical_steps;
    p.old_max_steps = old_max_steps;                  // CVodeSetMaxStep(cvode_mem, delta_time / old_max_steps)
    p.use_constraint = Nyx::use_sundials_constraint;
    ParmParse pp_nyx("nyx");
    pp_nyx.query("inhomo_reion", p.inhomo_reion);
    pp_nyx.query("uvb_density_A", p.uvb_density_A);
    pp_nyx.query("uvb_density_B", p.uvb_density_B);
    pp_nyx.query("reionization_zHI_flash", p.zhi_flash);
    pp_nyx.query("reionization_zHeII_flash", p.zheii_flash);
    pp_nyx.query("reionization_T_zHI", p.T_zhi);
    pp_nyx.query("reionization_T_zHeII", p.T_zheii);
    return p;
}

int check(int rc) {
    if (rc != HC_OK) amrex::Abort(std::string("nyx_hc: ") + hc_last_error());
    return 0;   // like the reference: per-cell integrator failures are not errors (they are counted in HcStats)
}

void finish(const HcStats& st, long int& new_max_steps) {
    g_last_stats = st;
    if (Nyx::use_typical_steps) new_max_steps = std::max<long int>(st.max_nst, new_max_steps);   // integrate_state_vec_3d.cpp:285-290
}

int vec_batch(std::vector<HcFab>& s, std::vector<HcFab>& d, std::vector<HcBox>& t, Real a, Real dt, long int old_max, long int& new_max) {
    if (t.empty()) return 0;
    const HcParams p = params_from_nyx(old_max);
    HcStats st{};
#ifdef AMREX_USE_GPU
    const int rc = hc_integrate_vec_batch((int)t.size(), s.data(), d.data(), t.data(), a, dt, &p, &st, nullptr, nullptr);
#else
    const int rc = hc_integrate_vec_host((int)t.size(), s.data(), d.data(), t.data(), a, dt, &p, &st);
#endif
    finish(st, new_max);
    return check(rc);
}

}  // namespace

extern "C" const HcStats* nyx_hc_last_stats() { return &g_last_stats; }

// Replaces the call `tabulate_rates(file_in, mean_rhob)` in Nyx::heatcool_setup (Source/Initialization/Nyx_setup.cpp:157-166):
// builds the same AtomicRates image on the host and uploads it to the current device.
extern "C" int nyx_hc_setup(const char* treecool_path, double mean_rhob)
{
    std::vector<double> rates(HC_RATES_DOUBLES);
    int rc = hc_tabulate_rates(treecool_path, mean_rhob, rates.data());
    if (rc == HC_OK) rc = hc_tables_upload(rates.data(), rates.size());
    if (rc != HC_OK) amrex::Abort(std::string("nyx_hc_setup: ") + hc_last_error());
    return rc;
}

// HC/integrate_state_vec_3d.cpp:44-70: valid cells of every local box
int Nyx::integrate_state_vec(MultiFab& S_old, MultiFab& D_old, const Real& a, const Real& delta_time)
{
    const long int store_steps = new_max_sundials_steps;
    std::vector<HcFab> s, d; std::vector<HcBox> t;
    for (MFIter mfi(S_old); mfi.isValid(); ++mfi) {
        s.push_back(to_fab(S_old.array(mfi))); d.push_back(to_fab(D_old.array(mfi))); t.push_back(to_box(mfi.validbox()));
    }
    return vec_batch(s, d, t, a, delta_time, store_steps, new_max_sundials_steps);
}

// HC/integrate_state_vec_3d.cpp:367-396: valid cells AND the ghost cells of S_old (growntilebox of an untiled MFIter)
int Nyx::integrate_state_grownvec(MultiFab& S_old, MultiFab& D_old, const Real& a, const Real& delta_time)
{
    // the FIRST Strang half-step works on the OTHER counter (integrate_state_vec_3d.cpp:376,392): the step cap is dt / old_max and the
    // largest step count goes back into old_max_sundials_steps (the one the checkpoint files carry, Nyx_output.cpp:295-317)
    const long int store_steps = old_max_sundials_steps;
    std::vector<HcFab> s, d; std::vector<HcBox> t;
    for (MFIter mfi(S_old); mfi.isValid(); ++mfi) {
        s.push_back(to_fab(S_old.array(mfi))); d.push_back(to_fab(D_old.array(mfi))); t.push_back(to_box(mfi.growntilebox()));
    }
    return vec_batch(s, d, t, a, delta_time, store_steps, old_max_sundials_steps);
}

// HC/integrate_state_vec_3d.cpp:72-365: one tile
int Nyx::integrate_state_vec_mfin(Array4<Real> const& state4, Array4<Real> const& diag_eos4, const Box& tbx, const Real& a,
                                  const Real& delta_time, long int& old_max_steps, long int& new_max_steps)
{
    std::vector<HcFab> s{to_fab(state4)}, d{to_fab(diag_eos4)}; std::vector<HcBox> t{to_box(tbx)};
    return vec_batch(s, d, t, a, delta_time, old_max_steps, new_max_steps);
}

namespace {
int struct_batch(std::vector<HcFab> f[6], std::vector<HcBox>& t, Real a, Real a_end, Real dt, int sdc_iter, long int old_max, long int& new_max) {
    if (t.empty()) return 0;
    const HcParams p = params_from_nyx(old_max);
    HcStats st{};
#ifdef AMREX_USE_GPU
    const int rc = hc_integrate_struct_batch((int)t.size(), f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(),
                                             t.data(), a, a_end, dt, sdc_iter, &p, &st, nullptr, nullptr);
#else
    const int rc = hc_integrate_struct_host((int)t.size(), f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(),
                                            t.data(), a, a_end, dt, sdc_iter, &p, &st);
#endif
    finish(st, new_max);
    return check(rc);
}
}  // namespace

// HC/integrate_state_with_source_3d.cpp:50-185 (the hctest dump/replay hooks :82-125 stay with the reference's I/O layer)
int Nyx::integrate_state_struct(MultiFab& S_old, MultiFab& S_new, MultiFab& D_old, MultiFab& hydro_src, MultiFab& IR, MultiFab& reset_src,
                                const Real& a, const Real& a_end, const Real& delta_time, const int sdc_iter)
{
    const long int store_steps = new_max_sundials_steps;
    std::vector<HcFab> f[6]; std::vector<HcBox> t;
    for (MFIter mfi(S_old); mfi.isValid(); ++mfi) {
        // C-ABI order == integrate_state_struct_mfin's: state, diag, state_n, hydro_src, reset_src, IR
        f[0].push_back(to_fab(S_old.array(mfi))); f[1].push_back(to_fab(D_old.array(mfi))); f[2].push_back(to_fab(S_new.array(mfi)));
        f[3].push_back(to_fab(hydro_src.array(mfi))); f[4].push_back(to_fab(reset_src.array(mfi))); f[5].push_back(to_fab(IR.array(mfi)));
        t.push_back(to_box(mfi.validbox()));
    }
    return struct_batch(f, t, a, a_end, delta_time, sdc_iter, store_steps, new_max_sundials_steps);
}

// HC/integrate_state_with_source_3d.cpp:187-709: one tile
int Nyx::integrate_state_struct_mfin(Array4<Real> const& state4, Array4<Real> const& diag_eos4, Array4<Real> const& state_n4,
                                     Array4<Real> const& hydro_src4, Array4<Real> const& reset_src4, Array4<Real> const& IR4,
                                     const Box& tbx, const Real& a, const Real& a_end, const Real& delta_time,
                                     long int& old_max_steps, long int& new_max_steps, const int sdc_iter)
{
    std::vector<HcFab> f[6] = {{to_fab(state4)}, {to_fab(diag_eos4)}, {to_fab(state_n4)}, {to_fab(hydro_src4)}, {to_fab(reset_src4)}, {to_fab(IR4)}};
    std::vector<HcBox> t{to_box(tbx)};
    return struct_batch(f, t, a, a_end, delta_time, sdc_iter, old_max_steps, new_max_steps);
}


// ---- next rows of the path: the cell loops of Nyx::compute_new_temp (Source/Driver/Nyx.cpp:2435-2522) and
// Nyx::reset_internal_energy (:2356-2385).  Both methods live in Nyx.cpp next to unrelated code, so they are offered as free functions
// that the method bodies call (INTEGRATION.md shows the two-line patch); `a` = get_comoving_a(state[State_Type].curTime()),
// small_temp / large_temp / max_temp_dt are the Nyx statics of the same names.
void nyx_hc_compute_new_temp(MultiFab& S_new, MultiFab& D_new, Real a, Real small_temp, Real large_temp, int max_temp_dt)
{
    std::vector<HcFab> s, d; std::vector<HcBox> t;
    for (MFIter mfi(S_new); mfi.isValid(); ++mfi) {
        s.push_back(to_fab(S_new.array(mfi))); d.push_back(to_fab(D_new.array(mfi))); t.push_back(to_box(mfi.validbox()));
    }
    if (t.empty()) return;
    const HcParams p = params_from_nyx(Nyx::old_max_sundials_steps);
    HcStats st{};
#ifdef AMREX_USE_GPU
    check(hc_compute_new_temp_batch((int)t.size(), s.data(), d.data(), t.data(), a, &p, small_temp, large_temp, max_temp_dt, &st, nullptr));
#else
    check(hc_compute_new_temp_host((int)t.size(), s.data(), d.data(), t.data(), a, &p, small_temp, large_temp, max_temp_dt, &st));
#endif
    g_last_stats = st;
}

void nyx_hc_reset_internal_energy(MultiFab& S_new, MultiFab& D_new, MultiFab& reset_e_src, Real a, Real small_temp, int interp)
{
    std::vector<HcFab> s, d, r; std::vector<HcBox> t;
    for (MFIter mfi(S_new); mfi.isValid(); ++mfi) {
        s.push_back(to_fab(S_new.array(mfi))); d.push_back(to_fab(D_new.array(mfi))); r.push_back(to_fab(reset_e_src.array(mfi)));
        t.push_back(to_box(mfi.validbox()));
    }
    if (t.empty()) return;
    const HcParams p = params_from_nyx(Nyx::old_max_sundials_steps);
#ifdef AMREX_USE_GPU
    check(hc_reset_internal_energy_batch((int)t.size(), s.data(), d.data(), r.data(), t.data(), a, &p, small_temp, interp, nullptr));
#else
    check(hc_reset_internal_energy_host((int)t.size(), s.data(), d.data(), r.data(), t.data(), a, &p, small_temp, interp));
#endif
}


// ---- state that rides on the path across restarts.
// nyx.use_typical_steps = 1 caps the BDF step of the next call at dt / old_max_sundials_steps (HcParams.old_max_steps) and records the
// largest step count (finish() above); the reference carries the two numbers through a checkpoint in the files first_max_steps /
// second_max_steps (written Source/IO/Nyx_output.cpp:295-317 -- BOTH receive old_max_sundials_steps -- and read back by
// Nyx::typical_values_post_restart, Source/Driver/Nyx.cpp:1625-1657, which ParallelDescriptor::Bcast's them).  Same files, same quirk;
// the caller broadcasts.
#include <fstream>
int nyx_hc_write_typical_steps(const std::string& dir)
{
    if (!Nyx::use_typical_steps) return 0;
    for (const char* fn : {"/first_max_steps", "/second_max_steps"}) {
        std::ofstream File((dir + fn).c_str(), std::ios::out | std::ios::trunc);
        if (!File.good()) return -1;
        File.precision(15);
        File << Nyx::old_max_sundials_steps << '\n';
    }
    return 0;
}

int nyx_hc_read_typical_steps(const std::string& restart_file)
{
    if (!Nyx::use_typical_steps) return 0;
    {
        std::ifstream File((restart_file + "/first_max_steps").c_str(), std::ios::in);
        if (!File.good()) return -1;
        File >> Nyx::old_max_sundials_steps;
    }
    {
        std::ifstream File((restart_file + "/second_max_steps").c_str(), std::ios::in);
        if (!File.good()) return -1;
        File >> Nyx::new_max_sundials_steps;
    }
    return 0;
}

// Nyx::init_zhi's cell loop (Source/Initialization/Nyx_initdata.cpp:198-209); `zhi` is the coarse MultiFab the reference builds with
// VisMF::Read + ParallelCopy just before it (:181-191, AMReX I/O and communication: unchanged), ratio = prob_res / nyx.inhomo_grid
void nyx_hc_init_zhi(MultiFab& D_new, MultiFab& zhi, int ratio)
{
    if (D_new.nComp() <= 2) return;
    std::vector<HcFab> d, z; std::vector<HcBox> t;
    for (MFIter mfi(D_new); mfi.isValid(); ++mfi) {
        d.push_back(to_fab(D_new.array(mfi))); z.push_back(to_fab(zhi.array(mfi))); t.push_back(to_box(mfi.validbox()));
    }
    if (t.empty()) return;
#ifdef AMREX_USE_GPU
    check(hc_init_zhi_batch((int)t.size(), d.data(), z.data(), ratio, t.data(), nullptr));
    check(hc_sync(nullptr));
#else
    // CPU build of AMReX: the same kernel through the host-buffer staging path (no CPU loop on the product path)
    check(hc_init_zhi_host((int)t.size(), d.data(), z.data(), ratio, t.data()));
#endif
}
