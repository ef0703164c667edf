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
//
// This file compiles against the reference's REAL headers (Source/Driver/Nyx.H + AMReX 23.04): tests/test_dropin_real.py builds it with
// the flags of the reference's own GNUmake build, links it with the UNMODIFIED strang_reactions.cpp, sdc_reactions.cpp, Nyx_advance.cpp, ...
// objects in place of the two replaced translation units, and runs the resulting Nyx executable on the GPU.  The protected Nyx statics
// (sundials_reltol, use_typical_steps, ...) are therefore only read inside Nyx:: member functions.
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#ifdef SAVE_REACT
#include <AMReX_PlotFileUtil.H>
#endif
#include <Nyx.H>
#if __has_include(<atomic_rates_data.H>)
#include <atomic_rates_data.H>   // the reference's AtomicRates image, filled by its own tabulate_rates (Nyx::heatcool_setup, unchanged)
#define NYX_HC_HAVE_ATOMIC_RATES 1
#endif

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
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

// the nyx.* run-time flags of the path: Nyx statics (Source/Driver/Nyx.cpp:116-181; the protected ones are handed in by the member
// functions) + what ode_eos_setup re-parses on every tile call in the reference (Source/HeatCool/f_rhs_struct.H:45-101); parsed once per call
struct NyxFlags {
    Real reltol, abstol;
    int use_typical_steps, use_constraint;
};
// inside a Nyx:: member function
#define NYX_HC_FLAGS() NyxFlags{sundials_reltol, sundials_abstol, use_typical_steps, use_sundials_constraint}

HcParams make_params(const NyxFlags& fl, long int old_max_steps) {
    HcParams p;
    hc_default_params(&p);
    p.rtol = fl.reltol;
    p.atol_factor = fl.abstol;
    p.h_species = Nyx::h_species;
    p.gamma_minus_1 = Nyx::gamma - 1.0;
    p.max_steps = 2000;                               // CVodeSetMaxNumSteps(cvode_mem, 2000)
    p.use_typical_steps = fl.use_typical_steps;
    p.old_max_steps = old_max_steps;                  // CVodeSetMaxStep(cvode_mem, delta_time / old_max_steps)
    p.use_constraint = fl.use_constraint;
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
// for the rows next to the path (EOS only: no integrator flags needed)
HcParams eos_params() { return make_params(NyxFlags{1e-4, 1e-4, 0, 0}, 3); }

// GPU build of AMReX: launch on AMReX's current stream, so that the kernels are ordered with the host application's own (no device-wide
// synchronisation); the statistics read-back -- a stream synchronisation -- only where somebody uses it (nyx.use_typical_steps, or
// NYX_HC_STATS=1 for nyx_hc_last_stats())
#ifdef AMREX_USE_GPU
void* nyx_stream() { return (void*)amrex::Gpu::gpuStream(); }
bool want_stats(const NyxFlags& fl) {
    static const bool env = std::getenv("NYX_HC_STATS") != nullptr;
    return env || fl.use_typical_steps != 0;
}
#endif

int check(int rc) {
    if (rc != HC_OK) amrex::Abort(std::string("nyx_hc: ") + hc_last_error());
    return 0;   // like the reference: per-cell integrator failures are not errors (they are counted in HcStats)
}

// The rate tables.  A host application that calls nyx_hc_setup() from Nyx::heatcool_setup has uploaded them already; otherwise (callers
// and set-up code COMPLETELY unchanged) the image the reference's own tabulate_rates left in atomic_rates_glob is uploaded on first use.
bool g_tables_ready = false;
void ensure_tables() {
    if (g_tables_ready) return;
#if defined(NYX_HC_HAVE_ATOMIC_RATES)
    static_assert(sizeof(AtomicRates) == sizeof(double) * HC_RATES_DOUBLES, "AtomicRates image layout (Source/EOS/atomic_rates_data.H:22-49)");
    if (atomic_rates_glob == nullptr) amrex::Abort("nyx_hc: atomic_rates_glob is not allocated (heat_cool_type != 11?)");
    std::vector<double> img(HC_RATES_DOUBLES);
#ifdef AMREX_USE_GPU
    amrex::Gpu::dtoh_memcpy(img.data(), atomic_rates_glob, sizeof(AtomicRates));
#else
    std::memcpy(img.data(), atomic_rates_glob, sizeof(AtomicRates));
#endif
    check(hc_tables_upload(img.data(), img.size()));
    g_tables_ready = true;
#else
    amrex::Abort("nyx_hc: call nyx_hc_setup(treecool_file, mean_rhob) from Nyx::heatcool_setup first");
#endif
}

void finish(const HcStats& st, int use_typical, long int& new_max_steps) {
    g_last_stats = st;
    if (use_typical) new_max_steps = std::max<long int>(st.max_nst, new_max_steps);   // integrate_state_vec_3d.cpp:285-290
}

int vec_batch(std::vector<HcFab>& s, std::vector<HcFab>& d, std::vector<HcBox>& t, Real a, Real dt, const NyxFlags& fl, long int old_max, long int& new_max) {
    if (t.empty()) return 0;
    ensure_tables();
    const HcParams p = make_params(fl, old_max);
    HcStats st{};
#ifdef AMREX_USE_GPU
    const int rc = hc_integrate_vec_batch((int)t.size(), s.data(), d.data(), t.data(), a, dt, &p, want_stats(fl) ? &st : nullptr, nullptr, nyx_stream());
#else
    const int rc = hc_integrate_vec_host((int)t.size(), s.data(), d.data(), t.data(), a, dt, &p, &st);
#endif
    finish(st, fl.use_typical_steps, new_max);
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
    g_tables_ready = true;
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
    return vec_batch(s, d, t, a, delta_time, NYX_HC_FLAGS(), store_steps, new_max_sundials_steps);
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
    return vec_batch(s, d, t, a, delta_time, NYX_HC_FLAGS(), store_steps, old_max_sundials_steps);
}

// HC/integrate_state_vec_3d.cpp:72-365: one tile
int Nyx::integrate_state_vec_mfin(Array4<Real> const& state4, Array4<Real> const& diag_eos4, const Box& tbx, const Real& a,
                                  const Real& delta_time, long int& old_max_steps, long int& new_max_steps)
{
    std::vector<HcFab> s{to_fab(state4)}, d{to_fab(diag_eos4)}; std::vector<HcBox> t{to_box(tbx)};
    return vec_batch(s, d, t, a, delta_time, NYX_HC_FLAGS(), old_max_steps, new_max_steps);
}

namespace {
// react (SAVE_REACT builds): the three diagnostic FAB lists react_in / react_out / react_out_work, or nullptr
int struct_batch(std::vector<HcFab> f[6], std::vector<HcBox>& t, Real a, Real a_end, Real dt, int sdc_iter, const NyxFlags& fl, long int old_max,
                 long int& new_max, std::vector<HcFab>* react = nullptr) {
    if (t.empty()) return 0;
    ensure_tables();
    const HcParams p = make_params(fl, old_max);
    HcStats st{};
    if (react) {
#ifdef AMREX_USE_GPU
        const int rc = hc_integrate_struct_react_batch((int)t.size(), f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(),
                                                       react[0].data(), react[1].data(), react[2].data(), t.data(), a, a_end, dt, sdc_iter, &p,
                                                       want_stats(fl) ? &st : nullptr, nullptr, nyx_stream());
#else
        const int rc = hc_integrate_struct_react_host((int)t.size(), f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(),
                                                      react[0].data(), react[1].data(), react[2].data(), t.data(), a, a_end, dt, sdc_iter, &p, &st);
#endif
        finish(st, fl.use_typical_steps, new_max);
        return check(rc);
    }
#ifdef AMREX_USE_GPU
    const int rc = hc_integrate_struct_batch((int)t.size(), f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(),
                                             t.data(), a, a_end, dt, sdc_iter, &p, want_stats(fl) ? &st : nullptr, nullptr, nyx_stream());
#else
    const int rc = hc_integrate_struct_host((int)t.size(), f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(),
                                            t.data(), a, a_end, dt, sdc_iter, &p, &st);
#endif
    finish(st, fl.use_typical_steps, new_max);
    return check(rc);
}
}  // namespace

#if !defined(NYX_HC_SHIM_BUILD)
// ---- the hctest isolation-test hooks of the SDC path (HC/integrate_state_with_source_3d.cpp:82-125; file layout of sdc_writeOn / sdc_readFrom,
// HC/f_rhs_struct.H:587-697): nyx.hctest_example_write = 1 dumps the six MultiFabs of a call (+ the ParmParse table and the replay keys,
// + BoxArray / DistributionMapping) before it runs, nyx.hctest_example_read = 1 loads them instead of the caller's data -- what
// Exec/HeatCoolTests replays.  Host-side I/O through AMReX's own writers, so the files are interchangeable with the reference's.
namespace {
void hctest_write(MultiFab& S_old, MultiFab& S_new, MultiFab& D_old, MultiFab& hydro_src, MultiFab& IR, MultiFab& reset_src, int n_tiles,
                  Real a, Real a_end, Real delta_time, int index, const std::string& f_inputs, const std::string& f_badmap, const std::string& f_chunk)
{
    {
        std::ofstream ofs_inputs(f_inputs.c_str());
        ParmParse::dumpTable(ofs_inputs, true);
        ofs_inputs << "nyx.initial_z = " << 1 / a - 1 << std::endl;
        ofs_inputs << "nyx.final_z = " << 1 / a_end - 1 << std::endl;
        ofs_inputs << "nyx.fixed_dt = " << delta_time << std::endl;
        ofs_inputs << "nyx.hctest_filename_inputs = " << f_inputs << std::endl;
        ofs_inputs << "nyx.hctest_filename_badmap = " << f_badmap << std::endl;
        ofs_inputs << "nyx.hctest_filename_chunk = " << f_chunk << std::endl;
        ofs_inputs << "nyx.hctest_endIndex = " << n_tiles << std::endl;
        ofs_inputs << "nyx.hctest_example_write = 0" << std::endl;
        ofs_inputs << "nyx.hctest_example_read = 1" << std::endl;
        ofs_inputs << "nyx.do_dm_particles = 0" << std::endl;
        ofs_inputs << "nyx.do_hydro = 0" << std::endl;
        ofs_inputs << "nyx.hctest_example_index = " << index << std::endl;
    }
    {
        std::ofstream ofs(f_badmap.c_str());
        S_old.boxArray().writeOn(ofs);
        S_old.DistributionMap().writeOn(ofs);
    }
    // one chunk per local FAB, the six FABs in the reference's order
    for (MFIter mfi(S_old); mfi.isValid(); ++mfi) {
        std::ofstream ofs((f_chunk + std::to_string(mfi.index())).c_str());
        S_old[mfi].writeOn(ofs); D_old[mfi].writeOn(ofs); S_new[mfi].writeOn(ofs);
        hydro_src[mfi].writeOn(ofs); reset_src[mfi].writeOn(ofs); IR[mfi].writeOn(ofs);
    }
}

void hctest_read(MultiFab& S_old, MultiFab& S_new, MultiFab& D_old, MultiFab& hydro_src, MultiFab& IR, MultiFab& reset_src,
                 const std::string& f_badmap, const std::string& f_chunk)
{
    BoxArray grids;
    DistributionMapping dmap;
    std::ifstream ifs(f_badmap.c_str());
    grids.readFrom(ifs);
    dmap.readFrom(ifs);
    AMREX_ALWAYS_ASSERT(S_old.boxArray().CellEqual(grids) && dmap == S_old.DistributionMap());
    for (MFIter mfi(S_old); mfi.isValid(); ++mfi) {
        std::ifstream ifc((f_chunk + std::to_string(mfi.index())).c_str());
        S_old[mfi].readFrom(ifc); D_old[mfi].readFrom(ifc); S_new[mfi].readFrom(ifc);
        hydro_src[mfi].readFrom(ifc); reset_src[mfi].readFrom(ifc); IR[mfi].readFrom(ifc);
    }
}
}  // namespace
#endif

// HC/integrate_state_with_source_3d.cpp:50-185
int Nyx::integrate_state_struct(MultiFab& S_old, MultiFab& S_new, MultiFab& D_old, MultiFab& hydro_src, MultiFab& IR, MultiFab& reset_src,
                                const Real& a, const Real& a_end, const Real& delta_time, const int sdc_iter)
{
    const long int store_steps = new_max_sundials_steps;
#if !defined(NYX_HC_SHIM_BUILD)
    {   // :82-125, same keys, same defaults, same order (write before read)
        ParmParse pp_nyx("nyx");
        long writeProc = ParallelDescriptor::IOProcessor() ? ParallelDescriptor::MyProc() : -1;
        int hctest_example_write = 0, hctest_example_read = 0, hctest_example_index = 0, hctest_example_proc = 0;
        pp_nyx.query("hctest_example_write", hctest_example_write);
        if (pp_nyx.query("hctest_example_write_proc", writeProc)) hctest_example_proc = (ParallelDescriptor::MyProc() == writeProc);
        else hctest_example_proc = 1;
        if (hctest_example_write) hctest_example_index = nStep();
        pp_nyx.query("hctest_example_index", hctest_example_index);
        pp_nyx.query("hctest_example_read", hctest_example_read);
        if (hctest_example_write != 0 || hctest_example_read != 0) {
            std::string f_inputs = "hctest/inputs." + std::to_string(hctest_example_index);
            std::string f_badmap = "hctest/BADMAP." + std::to_string(hctest_example_index);
            std::string f_chunk = "hctest/Chunk." + std::to_string(hctest_example_index) + ".";
            int directory_overwrite = pp_nyx.query("hctest_filename_inputs", f_inputs);
            directory_overwrite += pp_nyx.query("hctest_filename_badmap", f_badmap);
            directory_overwrite += pp_nyx.query("hctest_filename_chunk", f_chunk);
            if (hctest_example_read == 0 && hctest_example_write != 0) {
                if (directory_overwrite == 0) amrex::UtilCreateCleanDirectory("hctest", true);
                else amrex::Print() << "Using the following paths for hctest:\n" << f_inputs << "\n" << f_badmap << "\n" << f_chunk << std::endl;
            }
            if (hctest_example_proc && hctest_example_write != 0) {
                // nyx.hctest_endIndex = number of tiles of the reference's MFIter (a CVODE instance each there)
                const auto tiling = (TilingIfNotGPU() && sundials_use_tiling) ? MFItInfo().EnableTiling(sundials_tile_size) : MFItInfo();
                int n_tiles = 0;
                { MFIter mfi(S_old, tiling); n_tiles = mfi.length(); }   // (its own scope: AMReX allows one active MFIter at a time)
                hctest_write(S_old, S_new, D_old, hydro_src, IR, reset_src, n_tiles, a, a_end, delta_time,
                             hctest_example_index, f_inputs, f_badmap, f_chunk);
            }
            if (hctest_example_proc && hctest_example_read != 0) hctest_read(S_old, S_new, D_old, hydro_src, IR, reset_src, f_badmap, f_chunk);
        }
    }
#endif
    std::vector<HcFab> f[6]; std::vector<HcBox> t;
#ifdef SAVE_REACT
    // :126-135, the same three MultiFabs with the same component names
    const amrex::Vector<std::string> react_in_names {"eptr-idx", "f_rhs_data-ptr-rho_init_vode-idx", "f_rhs_data-ptr-rhoe_src_vode-idx", "f_rhs_data-ptr-e_src_vode-idx", "abstol_ptr-idx", "f_rhs_data-ptr-a", "time_in"};
    const amrex::Vector<std::string> react_out_names {"dptr-idx", "f_rhs_data-ptr-rho_vode-idx", "f_rhs_data-ptr-T_vode-idx", "f_rhs_data-ptr-ne_vode-idx", "abstol_achieve_ptr-idx", "a_end", "delta_time"};
    const amrex::Vector<std::string> react_out_work_names {"nst", "netf", "nfe", "nni", "ncfn", "nsetups", "nje", "ncfl", "nfeLS"};
    MultiFab react_in(grids, dmap, react_in_names.size(), NUM_GROW);
    MultiFab react_out(grids, dmap, react_out_names.size(), NUM_GROW);
    MultiFab react_out_work(grids, dmap, react_out_work_names.size(), NUM_GROW);
    std::vector<HcFab> react[3];
#endif
    for (MFIter mfi(S_old); mfi.isValid(); ++mfi) {
        // C-ABI order == integrate_state_struct_mfin's: state, diag, state_n, hydro_src, reset_src, IR
        f[0].push_back(to_fab(S_old.array(mfi))); f[1].push_back(to_fab(D_old.array(mfi))); f[2].push_back(to_fab(S_new.array(mfi)));
        f[3].push_back(to_fab(hydro_src.array(mfi))); f[4].push_back(to_fab(reset_src.array(mfi))); f[5].push_back(to_fab(IR.array(mfi)));
        t.push_back(to_box(mfi.validbox()));
#ifdef SAVE_REACT
        react[0].push_back(to_fab(react_in.array(mfi))); react[1].push_back(to_fab(react_out.array(mfi))); react[2].push_back(to_fab(react_out_work.array(mfi)));
#endif
    }
#ifdef SAVE_REACT
    const int rc = struct_batch(f, t, a, a_end, delta_time, sdc_iter, NYX_HC_FLAGS(), store_steps, new_max_sundials_steps, react);
    {   // :165-182
#ifdef NO_HYDRO
        Real cur_time = state[PhiGrav_Type].curTime();
#else
        Real cur_time = state[State_Type].curTime();
#endif
        auto plotfilename = Concatenate("plt_react_in", nStep(), 5);
        WriteSingleLevelPlotfile(plotfilename, react_in, react_in_names, Geom(), cur_time, nStep());
        plotfilename = Concatenate("plt_react_out", nStep(), 5);
        WriteSingleLevelPlotfile(plotfilename, react_out, react_out_names, Geom(), cur_time, nStep());
        plotfilename = Concatenate("plt_react_out_work", nStep(), 5);
        WriteSingleLevelPlotfile(plotfilename, react_out_work, react_out_work_names, Geom(), cur_time, nStep());
    }
    return rc;
#else
    return struct_batch(f, t, a, a_end, delta_time, sdc_iter, NYX_HC_FLAGS(), store_steps, new_max_sundials_steps);
#endif
}

// HC/integrate_state_with_source_3d.cpp:187-709: one tile
int Nyx::integrate_state_struct_mfin(Array4<Real> const& state4, Array4<Real> const& diag_eos4, Array4<Real> const& state_n4,
                                     Array4<Real> const& hydro_src4, Array4<Real> const& reset_src4, Array4<Real> const& IR4,
#ifdef SAVE_REACT
                                     Array4<Real> const& react_in_arr, Array4<Real> const& react_out_arr, Array4<Real> const& react_out_work_arr,
#endif
                                     const Box& tbx, const Real& a, const Real& a_end, const Real& delta_time,
                                     long int& old_max_steps, long int& new_max_steps, const int sdc_iter)
{
    std::vector<HcFab> f[6] = {{to_fab(state4)}, {to_fab(diag_eos4)}, {to_fab(state_n4)}, {to_fab(hydro_src4)}, {to_fab(reset_src4)}, {to_fab(IR4)}};
    std::vector<HcBox> t{to_box(tbx)};
#ifdef SAVE_REACT
    std::vector<HcFab> react[3] = {{to_fab(react_in_arr)}, {to_fab(react_out_arr)}, {to_fab(react_out_work_arr)}};
    return struct_batch(f, t, a, a_end, delta_time, sdc_iter, NYX_HC_FLAGS(), old_max_steps, new_max_steps, react);
#else
    return struct_batch(f, t, a, a_end, delta_time, sdc_iter, NYX_HC_FLAGS(), old_max_steps, new_max_steps);
#endif
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
    ensure_tables();
    const HcParams p = eos_params();
    HcStats st{};
#ifdef AMREX_USE_GPU
    check(hc_compute_new_temp_batch((int)t.size(), s.data(), d.data(), t.data(), a, &p, small_temp, large_temp, max_temp_dt, nullptr, nyx_stream()));
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
    ensure_tables();
    const HcParams p = eos_params();
#ifdef AMREX_USE_GPU
    check(hc_reset_internal_energy_batch((int)t.size(), s.data(), d.data(), r.data(), t.data(), a, &p, small_temp, interp, nyx_stream()));
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
// (Nyx::old_max_sundials_steps / new_max_sundials_steps / use_typical_steps are protected statics: the calling Nyx members pass them)
int nyx_hc_write_typical_steps(const std::string& dir, int use_typical_steps, long int old_max_sundials_steps)
{
    if (!use_typical_steps) return 0;
    for (const char* fn : {"/first_max_steps", "/second_max_steps"}) {
        std::ofstream File((dir + fn).c_str(), std::ios::out | std::ios::trunc);
        if (!File.good()) return -1;
        File.precision(15);
        File << old_max_sundials_steps << '\n';
    }
    return 0;
}

int nyx_hc_read_typical_steps(const std::string& restart_file, int use_typical_steps, long int& old_max_sundials_steps, long int& new_max_sundials_steps)
{
    if (!use_typical_steps) return 0;
    {
        std::ifstream File((restart_file + "/first_max_steps").c_str(), std::ios::in);
        if (!File.good()) return -1;
        File >> old_max_sundials_steps;
    }
    {
        std::ifstream File((restart_file + "/second_max_steps").c_str(), std::ios::in);
        if (!File.good()) return -1;
        File >> new_max_sundials_steps;
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
    check(hc_init_zhi_batch((int)t.size(), d.data(), z.data(), ratio, t.data(), nyx_stream()));
#else
    // CPU build of AMReX: the same kernel through the host-buffer staging path (no CPU loop on the product path)
    check(hc_init_zhi_host((int)t.size(), d.data(), z.data(), ratio, t.data()));
#endif
}
