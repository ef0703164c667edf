// TEST INFRASTRUCTURE ONLY: the same C entry points as oracle/ref_driver.cpp (nyxref_*), but linked against the PRODUCT's
// host drop-in (nyx_b200/csrc/nyx_heatcool_dropin.cpp) instead of the reference's translation units, so that one Python
// driver (oracle/pyref.py: Reference) can run the reference and the drop-in through identical calls.  AMReX is the thin
// API shim of oracle/shim (a CPU build: the drop-in takes its host-buffer path and stages through the C-ABI).
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#include <Nyx.H>

#include <sstream>

#include "nyx_hc.h"

using namespace amrex;

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
IntVect Nyx::sundials_tile_size(1024000, 8, 8);
int Nyx::inhomo_reion = 0;
long int Nyx::old_max_sundials_steps = 3;
long int Nyx::new_max_sundials_steps = 3;
Real Nyx::small_dens = -1.e200;
Real Nyx::small_temp = -1.e200;
std::string Nyx::enforce_min_density_type = "floor";

extern "C" const HcStats* nyx_hc_last_stats();
void nyx_hc_compute_new_temp(MultiFab& S_new, MultiFab& D_new, Real a, Real small_temp, Real large_temp, int max_temp_dt);
void nyx_hc_reset_internal_energy(MultiFab& S_new, MultiFab& D_new, MultiFab& reset_e_src, Real a, Real small_temp, int interp);
extern "C" int nyx_hc_setup(const char* treecool_path, double mean_rhob);
int nyx_hc_write_typical_steps(const std::string& dir, int use_typical_steps, long int old_max_sundials_steps);
int nyx_hc_read_typical_steps(const std::string& restart_file, int use_typical_steps, long int& old_max_sundials_steps, long int& new_max_sundials_steps);

extern "C" {

int nyxref_init(const char* treecool_path, double mean_rhob) { return nyx_hc_setup(treecool_path, mean_rhob); }

int nyxref_set(const char* key, const char* value) {
    std::string k(key), v(value);
    std::istringstream is(v);
    if (k == "nyx.gamma") is >> Nyx::gamma;
    else if (k == "nyx.h_species") is >> Nyx::h_species;
    else if (k == "nyx.use_typical_steps") is >> Nyx::use_typical_steps;
    else if (k == "nyx.use_sundials_constraint") is >> Nyx::use_sundials_constraint;
    else if (k == "nyx.sundials_reltol") is >> Nyx::sundials_reltol;
    else if (k == "nyx.sundials_abstol") is >> Nyx::sundials_abstol;
    else if (k == "nyx.sundials_tile_size") { is >> Nyx::sundials_tile_size[0] >> Nyx::sundials_tile_size[1] >> Nyx::sundials_tile_size[2]; }
    else if (k == "nyx.inhomo_reion") { is >> Nyx::inhomo_reion; ParmParse::table()[k] = v; }
    else if (k == "nyx.old_max_sundials_steps") is >> Nyx::old_max_sundials_steps;
    else if (k == "nyx.new_max_sundials_steps") is >> Nyx::new_max_sundials_steps;
    else ParmParse::table()[k] = v;
    return 0;
}
int nyxref_unset(const char* key) { ParmParse::table().erase(key); return 0; }
int nyxref_max_threads(void) { return 1; }
long nyxref_get_max_steps(int which) { return which ? Nyx::new_max_sundials_steps : Nyx::old_max_sundials_steps; }
// the 14 counters of HcStats of the last call (n_cells, n_failed, n_floor, sum_nst, max_nst, ...)
void nyxref_last_stats(long long* out14) { std::memcpy(out14, nyx_hc_last_stats(), sizeof(HcStats)); }

static BoxArray make_ba(int nboxes, const int* boxes) {
    std::vector<Box> b;
    for (int i = 0; i < nboxes; ++i) b.emplace_back(IntVect(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]), IntVect(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]));
    return BoxArray(b);
}

int nyxref_integrate_state_vec(int nboxes, const int* boxes, int ng_state, int ng_diag, int ncomp_diag, double* const* state, double* const* diag,
                               double a, double dt, int grown) {
    BoxArray ba = make_ba(nboxes, boxes);
    MultiFab S, D;
    S.defineAlias(ba, 6, ng_state, state);
    D.defineAlias(ba, ncomp_diag, ng_diag, diag);
    Nyx nyx;
    return grown ? nyx.integrate_state_grownvec(S, D, a, dt) : nyx.integrate_state_vec(S, D, a, dt);
}

int nyxref_integrate_state_struct(int nboxes, const int* boxes, const int* ng, int ncomp_diag, double* const* s_old, double* const* s_new,
                                  double* const* d_old, double* const* hydro_src, double* const* ir, double* const* reset_src, double a,
                                  double a_end, double dt, int sdc_iter) {
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

void nyxref_compute_new_temp(const int* box, int ng_state, int ng_diag, int ncomp_diag, double* state, double* diag, double a,
                             double small_temp, double large_temp, int max_temp_dt) {
    BoxArray ba = make_ba(1, box);
    MultiFab S, D;
    double* sp[1] = {state}; double* dp[1] = {diag};
    S.defineAlias(ba, 6, ng_state, sp);
    D.defineAlias(ba, ncomp_diag, ng_diag, dp);
    nyx_hc_compute_new_temp(S, D, a, small_temp, large_temp, max_temp_dt);
}

void nyxref_reset_internal_energy(const int* box, int ng_state, int ng_diag, int ncomp_diag, int ng_reset, double* state, double* diag,
                                  double* reset_src, double a, double small_temp, int interp) {
    BoxArray ba = make_ba(1, box);
    MultiFab S, D, R;
    double* sp[1] = {state}; double* dp[1] = {diag}; double* rp[1] = {reset_src};
    S.defineAlias(ba, 6, ng_state, sp);
    D.defineAlias(ba, ncomp_diag, ng_diag, dp);
    R.defineAlias(ba, 1, ng_reset, rp);
    nyx_hc_reset_internal_energy(S, D, R, a, small_temp, interp);
}

int nyxref_write_typical_steps(const char* dir) { return nyx_hc_write_typical_steps(dir, Nyx::use_typical_steps, Nyx::old_max_sundials_steps); }
int nyxref_read_typical_steps(const char* dir) {
    return nyx_hc_read_typical_steps(dir, Nyx::use_typical_steps, Nyx::old_max_sundials_steps, Nyx::new_max_sundials_steps);
}

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

}  // extern "C"
