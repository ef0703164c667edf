// TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called from the product path.
//
// C-ABI driver around the UNMODIFIED reference translation units
//   /root/reference/Source/HeatCool/integrate_state_vec_3d.cpp
//   /root/reference/Source/HeatCool/integrate_state_with_source_3d.cpp
// (+ the reference's EOS/HeatCool headers and the vendored SUNDIALS 6.3.0 CVODE sources),
// all compiled where they lie by oracle/Makefile into oracle/_ref/. It plays the part of
// Exec/HeatCoolTests/nyx_main.cpp:97-104 + Source/Initialization/Nyx_setup.cpp:157-166:
// build the rate tables, build MultiFabs over caller memory, call the public
// Nyx::integrate_state_vec / integrate_state_grownvec / integrate_state_struct.
//
// CVode() and CVodeFree() are intercepted with -DCVode=nyxref_hook_CVode
// -DCVodeFree=nyxref_hook_CVodeFree on the two reference TUs only, so that the return flag
// the reference ignores (integrate_state_vec_3d.cpp:284) and the per-instance counters can be
// recorded without touching reference sources.
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#include <Nyx.H>

#include <cvode/cvode.h>
#include <cvode/cvode_diag.h>

using namespace amrex;
#include <atomic_rates.H>   // reference: tabulate_rates (non-inline, include in exactly one TU)

// ---- definitions the reference expects from Source/Driver/Nyx.cpp:102-201 (same defaults)
AtomicRates* atomic_rates_glob = nullptr;
Real Nyx::gamma = 5.0 / 3.0;
Real Nyx::h_species = 0.76;
int Nyx::verbose = 0;
int Nyx::strang_grown_box = 1;
int Nyx::heat_cool_type = 11;
int Nyx::sundials_atomic_reductions = -1;
int Nyx::sundials_alloc_type = 0;
int Nyx::use_typical_steps = 0;
int Nyx::use_sundials_constraint = 0;
int Nyx::use_sundials_fused = 0;
bool Nyx::sundials_use_tiling = true;
Real Nyx::sundials_reltol = 1e-4;
Real Nyx::sundials_abstol = 1e-4;
// Nyx.cpp:555-568: unset nyx.sundials_tile_size falls back to fabarray.mfiter_tile_size
IntVect Nyx::sundials_tile_size(1024000, 8, 8);
int Nyx::inhomo_reion = 0;
long int Nyx::old_max_sundials_steps = 3;
long int Nyx::new_max_sundials_steps = 3;

namespace {
struct InstanceStats { long v[8]; };
std::vector<InstanceStats> g_stats;
std::vector<int> g_flags;
bool g_record = true;
}

extern "C" {

int nyxref_hook_CVode(void* cvode_mem, realtype tout, N_Vector yout, realtype* tret, int itask) {
    int flag = CVode(cvode_mem, tout, yout, tret, itask);
    if (g_record) g_flags.push_back(flag);
    return flag;
}

void nyxref_hook_CVodeFree(void** cvode_mem) {
    if (g_record && cvode_mem && *cvode_mem) {
        InstanceStats s{};
        CVodeGetNumSteps(*cvode_mem, &s.v[0]);
        CVodeGetNumErrTestFails(*cvode_mem, &s.v[1]);
        CVodeGetNumRhsEvals(*cvode_mem, &s.v[2]);
        CVodeGetNumNonlinSolvIters(*cvode_mem, &s.v[3]);
        CVodeGetNumNonlinSolvConvFails(*cvode_mem, &s.v[4]);
        CVodeGetNumLinSolvSetups(*cvode_mem, &s.v[5]);
        CVDiagGetNumRhsEvals(*cvode_mem, &s.v[6]);
        s.v[7] = g_flags.empty() ? 0 : g_flags.back();
        g_stats.push_back(s);
    }
    CVodeFree(cvode_mem);
}

// Nyx::heatcool_setup (Source/Initialization/Nyx_setup.cpp:157-166)
int nyxref_init(const char* treecool_path, double mean_rhob) {
    if (!atomic_rates_glob) atomic_rates_glob = (AtomicRates*)The_Arena()->alloc(sizeof(AtomicRates));
    tabulate_rates(std::string(treecool_path), mean_rhob);
    return 0;
}

// raw view of the reference's AtomicRates (1 + 7*301 + 15*2001 doubles)
const double* nyxref_rates(long* n_doubles) {
    if (n_doubles) *n_doubles = long(sizeof(AtomicRates) / sizeof(double));
    return reinterpret_cast<const double*>(atomic_rates_glob);
}

// nyx.* ParmParse entries (read by ode_eos_setup, f_rhs_struct.H:45-101) and the Nyx statics
// parsed in Nyx::read_hydro_params (Source/Driver/Nyx.cpp:474-593)
int nyxref_set(const char* key, const char* value) {
    std::string k(key), v(value);
    std::istringstream is(v);
    if (k == "nyx.gamma") is >> Nyx::gamma;
    else if (k == "nyx.h_species") is >> Nyx::h_species;
    else if (k == "nyx.v") is >> Nyx::verbose;
    else if (k == "nyx.use_typical_steps") is >> Nyx::use_typical_steps;
    else if (k == "nyx.use_sundials_constraint") is >> Nyx::use_sundials_constraint;
    else if (k == "nyx.sundials_use_tiling") is >> Nyx::sundials_use_tiling;
    else if (k == "nyx.sundials_reltol") is >> Nyx::sundials_reltol;
    else if (k == "nyx.sundials_abstol") is >> Nyx::sundials_abstol;
    else if (k == "nyx.sundials_tile_size") { is >> Nyx::sundials_tile_size[0] >> Nyx::sundials_tile_size[1] >> Nyx::sundials_tile_size[2]; }
    else if (k == "nyx.inhomo_reion") { is >> Nyx::inhomo_reion; ParmParse::table()[k] = v; }
    else if (k == "nyx.old_max_sundials_steps") is >> Nyx::old_max_sundials_steps;
    else if (k == "nyx.new_max_sundials_steps") is >> Nyx::new_max_sundials_steps;
    else if (k == "fabarray.mfiter_tile_size") { is >> FabArrayBase::mfiter_tile_size[0] >> FabArrayBase::mfiter_tile_size[1] >> FabArrayBase::mfiter_tile_size[2]; }
    else if (k == "omp.num_threads") {
#ifdef _OPENMP
        int n; is >> n; omp_set_num_threads(n);
#endif
    }
    else ParmParse::table()[k] = v;
    return 0;
}

int nyxref_unset(const char* key) { ParmParse::table().erase(key); return 0; }

int nyxref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void nyxref_stats_reset(void) { g_stats.clear(); g_flags.clear(); }
void nyxref_stats_record(int on) { g_record = (on != 0); }
long nyxref_stats_count(void) { return long(g_stats.size()); }
// out[8*i + {0..7}] = nst, netf, nfe, nni, ncfn, nsetups, nfeLS, CVode() flag of instance i (MFIter order)
void nyxref_stats_get(long* out) { for (size_t i = 0; i < g_stats.size(); ++i) std::memcpy(out + 8 * i, g_stats[i].v, sizeof(long) * 8); }
long nyxref_get_max_steps(int which) { return which ? Nyx::new_max_sundials_steps : Nyx::old_max_sundials_steps; }

static BoxArray make_ba(int nboxes, const int* boxes) {
    std::vector<Box> b;
    for (int i = 0; i < nboxes; ++i) b.emplace_back(IntVect(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]), IntVect(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]));
    return BoxArray(b);
}

// boxes: nboxes x {lo[3], hi[3]} valid boxes; state[i]/diag[i]: FAB i = (box grown by ng) x ncomp, Fortran order.
// grown=0 -> Nyx::integrate_state_vec (HC/integrate_state_vec_3d.cpp:44); grown=1 -> integrate_state_grownvec (:367)
int nyxref_integrate_state_vec(int nboxes, const int* boxes, int ng_state, int ng_diag, int ncomp_diag,
                               double* const* state, double* const* diag, double a, double dt, int grown) {
    BoxArray ba = make_ba(nboxes, boxes);
    MultiFab S, D;
    S.defineAlias(ba, 6, ng_state, state);
    D.defineAlias(ba, ncomp_diag, ng_diag, diag);
    Nyx nyx;
    return grown ? nyx.integrate_state_grownvec(S, D, a, dt) : nyx.integrate_state_vec(S, D, a, dt);
}

// Nyx::integrate_state_struct (HC/integrate_state_with_source_3d.cpp:50); ng[6]/ptrs in the
// order S_old, S_new, D_old, hydro_src, IR, reset_src (the reference's argument order)
int nyxref_integrate_state_struct(int nboxes, const int* boxes, const int* ng, int ncomp_diag,
                                  double* const* s_old, double* const* s_new, double* const* d_old,
                                  double* const* hydro_src, double* const* ir, double* const* reset_src,
                                  double a, double a_end, double dt, int sdc_iter) {
    BoxArray ba = make_ba(nboxes, boxes);
    MultiFab S_old, S_new, D_old, H, IR, R;
    S_old.defineAlias(ba, 6, ng[0], s_old);
    S_new.defineAlias(ba, 6, ng[1], s_new);
    D_old.defineAlias(ba, ncomp_diag, ng[2], d_old);
    H.defineAlias(ba, 6, ng[3], hydro_src);
    IR.defineAlias(ba, 1, ng[4], ir);
    R.defineAlias(ba, 1, ng[5], reset_src);
    Nyx nyx;
    return nyx.integrate_state_struct(S_old, S_new, D_old, H, IR, R, a, a_end, dt, sdc_iter);
}

}  // extern "C"

// ---- single-function probes of the reference's device functions (unit parity of the port)
#include <eos_hc.H>
extern "C" {
void nyxref_ion_n(int JH, int JHe, double U, double nh, double ne, double gamma_minus_1, double h_species, double z, double* out4) {
    Real nhp, nhep, nhepp, t;
    ion_n_device(atomic_rates_glob, JH, JHe, U, nh, ne, nhp, nhep, nhepp, t, gamma_minus_1, h_species, z);
    out4[0] = nhp; out4[1] = nhep; out4[2] = nhepp; out4[3] = t;
}
void nyxref_eos_T_given_Re(int JH, int JHe, double R, double e, double a, double gamma_minus_1, double h_species, double* T, double* Ne) {
    nyx_eos_T_given_Re_device(atomic_rates_glob, gamma_minus_1, h_species, JH, JHe, T, Ne, R, e, a);
}
void nyxref_interp_to_this_z(double z, double* out6) {
    interp_to_this_z(atomic_rates_glob, z, out6[0], out6[1], out6[2], out6[3], out6[4], out6[5]);
}
}

// ---- the next rows of the hot-path table: Nyx::reset_internal_energy and Nyx::compute_new_temp.
// Both live in Source/Driver/Nyx.cpp, which cannot be compiled without the whole application.  reset_internal_energy's cell
// body is the reference header Source/EOS/reset_internal_e.H, called here exactly as Nyx.cpp:2377-2382 calls it.
// compute_new_temp's cell body is a lambda inside Nyx.cpp:2473-2519: its branch structure is RESTATED below around the
// reference's own nyx_eos_T_given_Re_device / nyx_eos_given_RT (eos_hc.H:204-231), statement by statement.
#include <reset_internal_e.H>
extern "C" {
void nyxref_reset_internal_energy(const int* box, int ng_state, int ng_diag, int ncomp_diag, int ng_reset, double* state, double* diag,
                                  double* reset_src, double a, double small_temp, int interp) {
    BoxArray ba = make_ba(1, box);
    MultiFab S, D, R;
    double* sp[1] = {state}; double* dp[1] = {diag}; double* rp[1] = {reset_src};
    S.defineAlias(ba, 6, ng_state, sp);
    D.defineAlias(ba, ncomp_diag, ng_diag, dp);
    R.defineAlias(ba, 1, ng_reset, rp);
    const Real gamma_minus_1 = Nyx::gamma - 1.0;
    for (MFIter mfi(S); mfi.isValid(); ++mfi) {
        const Box& bx = mfi.validbox();
        const auto fab = S.array(mfi);
        const auto fab_diag = D.array(mfi);
        const auto fab_reset = R.array(mfi);
        for (int k = bx.smallEnd(2); k <= bx.bigEnd(2); ++k) for (int j = bx.smallEnd(1); j <= bx.bigEnd(1); ++j) for (int i = bx.smallEnd(0); i <= bx.bigEnd(0); ++i)
            reset_internal_e(i, j, k, fab, fab_diag, fab_reset, atomic_rates_glob, a, gamma_minus_1, Nyx::h_species, small_temp, interp);
    }
}

void nyxref_compute_new_temp(const int* box, int ng_state, int ng_diag, int ncomp_diag, double* state, double* diag, double a,
                             double local_small_temp, double local_large_temp, int local_max_temp_dt) {
    BoxArray ba = make_ba(1, box);
    MultiFab S, D;
    double* sp[1] = {state}; double* dp[1] = {diag};
    S.defineAlias(ba, 6, ng_state, sp);
    D.defineAlias(ba, ncomp_diag, ng_diag, dp);
    const Real h_species_in = Nyx::h_species, gamma_minus_1_in = Nyx::gamma - 1.0;
    AtomicRates* atomic_rates = atomic_rates_glob;
    for (MFIter mfi(S); mfi.isValid(); ++mfi) {
        const Box& bx = mfi.validbox();
        const auto state_fab = S.array(mfi);
        const auto diag_eos_fab = D.array(mfi);
        for (int k = bx.smallEnd(2); k <= bx.bigEnd(2); ++k) for (int j = bx.smallEnd(1); j <= bx.bigEnd(1); ++j) for (int i = bx.smallEnd(0); i <= bx.bigEnd(0); ++i) {
            Real rhoInv = 1.0 / state_fab(i,j,k,Density_comp);
            Real eint = state_fab(i,j,k,Eint_comp) * rhoInv;
            if (state_fab(i,j,k,Eint_comp) > 0.0) {
                nyx_eos_T_given_Re_device(atomic_rates, gamma_minus_1_in, h_species_in, 1, 1,
                                          &diag_eos_fab(i,j,k,Temp_comp), &diag_eos_fab(i,j,k,Ne_comp),
                                          state_fab(i,j,k,Density_comp), state_fab(i,j,k,Eint_comp) * (1.0 / state_fab(i,j,k,Density_comp)), a);
                if (diag_eos_fab(i,j,k,Temp_comp) >= local_large_temp && local_max_temp_dt == 1) {
                    diag_eos_fab(i,j,k,Temp_comp) = local_large_temp;
                    Real dummy_pres = 0.0;
                    nyx_eos_given_RT(atomic_rates, gamma_minus_1_in, h_species_in, &eint, &dummy_pres,
                                     state_fab(i,j,k,Density_comp), diag_eos_fab(i,j,k,Temp_comp), diag_eos_fab(i,j,k,Ne_comp), a);
                    Real ke = 0.5e0 * (state_fab(i,j,k,Xmom_comp) * state_fab(i,j,k,Xmom_comp) + state_fab(i,j,k,Ymom_comp) * state_fab(i,j,k,Ymom_comp) +
                                       state_fab(i,j,k,Zmom_comp) * state_fab(i,j,k,Zmom_comp)) * rhoInv;
                    state_fab(i,j,k,Eint_comp) = state_fab(i,j,k,Density_comp) * eint;
                    state_fab(i,j,k,Eden_comp) = state_fab(i,j,k,Eint_comp) + ke;
                }
            } else {
                Real dummy_pres = 0.0;
                nyx_eos_given_RT(atomic_rates, gamma_minus_1_in, h_species_in, &eint, &dummy_pres,
                                 state_fab(i,j,k,Density_comp), local_small_temp, diag_eos_fab(i,j,k,Ne_comp), a);
                Real ke = 0.5e0 * (state_fab(i,j,k,Xmom_comp) * state_fab(i,j,k,Xmom_comp) + state_fab(i,j,k,Ymom_comp) * state_fab(i,j,k,Ymom_comp) +
                                   state_fab(i,j,k,Zmom_comp) * state_fab(i,j,k,Zmom_comp)) * rhoInv;
                diag_eos_fab(i,j,k,Temp_comp) = local_small_temp;
                state_fab(i,j,k,Eint_comp) = state_fab(i,j,k,Density_comp) * eint;
                state_fab(i,j,k,Eden_comp) = state_fab(i,j,k,Eint_comp) + ke;
            }
        }
    }
}
}

// ---- the SDC source assembly either side of sdc_reactions.
// Nyx::update_state_with_sources is the reference's own translation unit Source/TimeStep/Nyx_update_state_with_sources.cpp, compiled
// unmodified (with -DSDC) by oracle/Makefile.  It calls Nyx::enforce_minimum_density, whose file (Nyx_enforce_minimum_density.cpp) also
// holds the "conservative" variant that needs AMReX's FillPatch and cannot be built here: the top-level function (:8-65) and the floor
// variant (:67-107) are RESTATED below around the reference's own per-cell function floor_density (Nyx_enforce_minimum_density.H:8-58).
#include <Nyx_enforce_minimum_density.H>
Real Nyx::small_dens = -1.e200;   // Source/Driver/Nyx.cpp:98-99,199
Real Nyx::small_temp = -1.e200;
std::string Nyx::enforce_min_density_type = "floor";

void Nyx::enforce_minimum_density(MultiFab& S_old, MultiFab& S_new, MultiFab& hydro_source, MultiFab& reset_e_src, Real a_new) {
    if (S_new.min(Density_comp) < small_dens) {
        if (enforce_min_density_type == "floor") {
            enforce_minimum_density_floor(S_new, a_new);
        } else if (enforce_min_density_type == "conservative") {
            amrex::Abort("oracle/_ref: the conservative variant needs FillPatch (AMReX) and is not built");
        } else {
            amrex::Abort("Don't know this enforce_min_density_type");
        }
        for (MFIter mfi(hydro_source, TilingIfNotGPU()); mfi.isValid(); ++mfi) {
            const Box& bx = mfi.tilebox();
            auto const& hydro_src = hydro_source.array(mfi);
            auto const& uin = S_old.array(mfi);
            auto const& uout = S_new.array(mfi);
            amrex::ParallelFor(bx, [=](int i, int j, int k) noexcept {
                hydro_src(i,j,k,Density_comp) = uout(i,j,k,Density_comp) - uin(i,j,k,Density_comp);
            });
        }
    }
}

void Nyx::enforce_minimum_density_floor(MultiFab& S_new, Real a_new_in) {
    Real lsmall_dens = small_dens;
    Real lgamma_minus_1 = gamma - 1.0;
    Real lsmall_temp = small_temp;
    auto atomic_rates = atomic_rates_glob;
    Real l_h_species = h_species;
    for (MFIter mfi(S_new, TilingIfNotGPU()); mfi.isValid(); ++mfi) {
        const Box& bx = mfi.tilebox();
        auto const& uout = S_new.array(mfi);
        amrex::ParallelFor(bx, [=](int i, int j, int k) noexcept {
            floor_density(i, j, k, uout, atomic_rates, a_new_in, lgamma_minus_1, lsmall_dens, lsmall_temp, l_h_species);
        });
    }
}

// One iteration of Nyx::enforce_minimum_density_cons (Nyx_enforce_minimum_density.cpp:179-236) on ONE box: the loop body RESTATED around the
// reference's own per-cell functions compute_mu_for_enforce_min / create_update_for_minimum (Nyx_enforce_minimum_density.H:60-200, included
// above); FillPatch -- AMReX's ghost exchange, not built here -- is the caller's: sborder arrives with its two ghost cells filled.
extern "C" double nyxref_enforce_min_cons_iter(const int* box, int ng_new, int ng_rs, double* sborder, double* s_new, double* reset_src,
                                               double small_dens, int sdc) {
    const Box vbx(IntVect(box[0], box[1], box[2]), IntVect(box[3], box[4], box[5]));
    FArrayBox Sb(amrex::grow(vbx, 2), 6, sborder), Sn(amrex::grow(vbx, ng_new), 6, s_new), Rs(amrex::grow(vbx, ng_rs), 1, reset_src);
    // face-based coefficients with one ghost face (:122-124): stored on the cell box grown by two, which holds every face index touched
    FArrayBox mux(amrex::grow(vbx, 2), 1), muy(amrex::grow(vbx, 2), 1), muz(amrex::grow(vbx, 2), 1), upd(vbx, 6);
    mux.setVal(0.); muy.setVal(0.); muz.setVal(0.); upd.setVal(0.);
    auto const& sbord = Sb.array();
    auto const& mu_x_arr = mux.array(); auto const& mu_y_arr = muy.array(); auto const& mu_z_arr = muz.array();
    auto const& upd_arr = upd.array();
    const Real lsmall_dens = small_dens;
    amrex::ParallelFor(amrex::grow(vbx, 1), [=](int i, int j, int k) noexcept {
        compute_mu_for_enforce_min(i, j, k, Density_comp, sbord, mu_x_arr, mu_y_arr, mu_z_arr, lsmall_dens);
    });
    amrex::ParallelFor(vbx, [=](int i, int j, int k) noexcept {
        create_update_for_minimum(i, j, k, sbord, mu_x_arr, mu_y_arr, mu_z_arr, upd_arr);
    });
    auto const& sn = Sn.array(); auto const& rs = Rs.array();
    double mn = std::numeric_limits<double>::max();
    amrex::ParallelFor(vbx, [&](int i, int j, int k) noexcept {
        for (int n = 0; n < 6; ++n) sn(i, j, k, n) += upd_arr(i, j, k, n);       // S_new.plus(update, 0, nComp, 0)
        if (sdc) rs(i, j, k, 0) = upd_arr(i, j, k, Eint_comp);                    // MultiFab::Copy(reset_e_src, update, Eint_comp, 0, 1, 0)
        if (sn(i, j, k, Density_comp) < mn) mn = sn(i, j, k, Density_comp);
    });
    return mn;
}

extern "C" {
// ng[6] / pointers in the reference's argument order: S_old, S_new, ext_src_old, hydro_source, grav_vector, reset_e_src
// (6, 6, 6, 6, 3, 1 components)
void nyxref_update_state_with_sources(int nboxes, const int* boxes, const int* ng, double* const* s_old, double* const* s_new,
                                      double* const* ext_src, double* const* hydro_src, double* const* grav, double* const* reset_src,
                                      double dt, double a_old, double a_new, double small_dens, double small_temp) {
    BoxArray ba = make_ba(nboxes, boxes);
    MultiFab S_old, S_new, E, H, G, R;
    S_old.defineAlias(ba, 6, ng[0], s_old);
    S_new.defineAlias(ba, 6, ng[1], s_new);
    E.defineAlias(ba, 6, ng[2], ext_src);
    H.defineAlias(ba, 6, ng[3], hydro_src);
    G.defineAlias(ba, 3, ng[4], grav);
    R.defineAlias(ba, 1, ng[5], reset_src);
    Nyx::small_dens = small_dens;
    Nyx::small_temp = small_temp;
    Nyx nyx;
    nyx.update_state_with_sources(S_old, S_new, E, H, G, R, dt, a_old, a_new);
}
}
