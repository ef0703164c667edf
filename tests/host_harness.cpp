// TEST INFRASTRUCTURE ONLY: runs the product's per-cell state machine (nyx_b200/csrc/hc_device.cuh), compiled for
// the HOST, one lane at a time, so that its control flow and arithmetic can be compared bit-for-bit with the
// oracle in a container without a GPU. The product itself never runs this path (no CPU fallback): the C-ABI
// library only launches the CUDA kernels.
#include <cstring>
#include <vector>

#include "../nyx_b200/csrc/hc_host.hpp"

using namespace hc;

namespace {
// array storage policy of the lane on the host: one lane at a time, a plain static array
struct ArrHost {
    double a[ARR_DOUBLES];
    double& at(int slot) { return a[slot]; }
};
struct View {
    double* p; long long js, ks, ns; int lo[3];
    double& operator()(int i, int j, int k, int n) const { return p[(i - lo[0]) + (j - lo[1]) * js + (k - lo[2]) * ks + n * ns]; }
};
View view(const HcFab* f) { return View{f->p, f->jstride, f->kstride, f->nstride, {f->lo[0], f->lo[1], f->lo[2]}}; }
// the persistent slots of a lane between rounds (shared memory in the kernel)
struct HostIO {
    double dd[64]; unsigned ww[16];
    double& d(int s) { return dd[s]; }
    unsigned& w(int s) { return ww[s]; }
};
// One cell through the ROUND structure of the kernel: after every resume() the lane is saved to its slots, every transient member is
// destroyed (all-ones bytes: NaNs / garbage), the evaluation phase works on its own view of the slots, and the bookkeeping phase reloads
// the lane.  Any state the kernel would lose between rounds is lost here too.
template <class LaneT, class Fetch>
void run_cell(LaneT& ln, const Tables& tb, const Consts& k, Fetch fetch_finalize_data) {
    HostIO io;
    std::memset(&io, 0xff, sizeof io);
    while (ln.active()) {
        ln.save(io);
        LaneT ev;
        std::memset((void*)&ev, 0xff, sizeof ev);
        ev.load_request(io);
        const bool was_eos = (ev.pc == PC_FINAL_EOS);
        const double fv = ev.eval_request(tb, k);
        ev.save_result(io, fv, was_eos);
        LaneT bk;
        std::memset((void*)&bk, 0xff, sizeof bk);
        bk.arr = ln.arr;
        double f;
        bk.load(io, k, f);
        if (bk.pc == PC_FINAL_EOS) fetch_finalize_data(bk);
        bk.resume(k, f, 0u);
        if (bk.fin_pending) { fetch_finalize_data(bk); bk.begin_finalize(k); }
        ln = bk;
    }
}
void put_stats(HcCellStat* cs, long idx, int nst, int netf, int nfe, int nni, int nnf, int nsetups, int nfe_ls, int flag) {
    if (!cs) return;
    cs[idx] = HcCellStat{nst, netf, nfe, nni, nnf, nsetups, nfe_ls, flag};
}
}  // namespace

extern "C" {

// the table of the device's fast_log10, as hc_tables_upload builds it (128 entries x {r, Lhi, Llo, 0})
void hh_log10_table(double* out) {
    std::vector<double> t;
    build_log10_table(t);
    for (size_t i = 0; i < t.size(); ++i) out[i] = t[i];
}

int hh_tabulate_rates(const char* file, double mean_rhob, double* out) { return tabulate_rates(file, mean_rhob, out); }

int hh_integrate_vec(const double* rates, const HcParams* prm, const HcFab* state, const HcFab* diag, HcBox tile, double a, double dt,
                     HcCellStat* cs) {
    std::vector<double> ionx, iony, cool;
    interleave_tables(rates, ionx, iony, cool);
    Tables tb{ionx.data(), iony.data(), cool.data(), nullptr};
    const Consts k = make_consts_vec(rates, *prm, a, dt);
    View S = view(state), D = view(diag);
    long idx = 0;
    for (int kk = tile.lo[2]; kk <= tile.hi[2]; ++kk) for (int j = tile.lo[1]; j <= tile.hi[1]; ++j) for (int i = tile.lo[0]; i <= tile.hi[0]; ++i, ++idx) {
        Lane<PATH_VEC, ArrHost> ln;
        ln.rho = S(i, j, kk, 0);
        ln.e0 = S(i, j, kk, 5) / ln.rho;
        ln.abstol = nv_scale(k.atol_factor, ln.e0);
        ln.jh = 1.0;
        ln.lastT = D(i, j, kk, 0); ln.lastNe = D(i, j, kk, 1);
        ln.start(k);
        run_cell(ln, tb, k, [](Lane<PATH_VEC, ArrHost>&) {});
        D(i, j, kk, 0) = ln.outT; D(i, j, kk, 1) = ln.outNe;
        S(i, j, kk, 5) += S(i, j, kk, 0) * (ln.e_final - ln.e0);
        S(i, j, kk, 4) += S(i, j, kk, 0) * (ln.e_final - ln.e0);
        put_stats(cs, idx, ln.nst, ln.netf, ln.nfe, ln.nni, ln.nnf, ln.nsetups, ln.nfe_ls, ln.flag);
    }
    return 0;
}

int hh_integrate_struct(const double* rates, const HcParams* prm, const HcFab* s_old, const HcFab* diag, const HcFab* s_new,
                        const HcFab* hydro_src, const HcFab* reset_src, const HcFab* ir, HcBox tile, double a, double a_end, double dt,
                        int sdc_iter, HcCellStat* cs) {
    std::vector<double> ionx, iony, cool;
    interleave_tables(rates, ionx, iony, cool);
    Tables tb{ionx.data(), iony.data(), cool.data(), nullptr};
    const Consts k = make_consts_struct(rates, *prm, a, a_end, dt, sdc_iter);
    View S = view(s_old), D = view(diag), N = view(s_new), H = view(hydro_src), R = view(reset_src), I = view(ir);
    long idx = 0;
    for (int kk = tile.lo[2]; kk <= tile.hi[2]; ++kk) for (int j = tile.lo[1]; j <= tile.hi[1]; ++j) for (int i = tile.lo[0]; i <= tile.hi[0]; ++i, ++idx) {
        Lane<PATH_STRUCT, ArrHost> ln;
        ln.rho = S(i, j, kk, 0);
        const double rhoe0 = S(i, j, kk, 5);
        ln.e0 = rhoe0 / ln.rho;
        ln.abstol = nv_scale(k.atol_factor, ln.e0);
        ln.jh = (double)k.JH0;
        ln.lastT = D(i, j, kk, 0); ln.lastNe = D(i, j, kk, 1);
        ln.rho_src = ln.rhoe_src = ln.e_src = ln.reset_src = 0.0; ln.zhi = 0.0;
        if (k.sdc_has_src) {   // ode_eos_initialize_arrays f_rhs_struct.H:186-193
            ln.rho_src = H(i, j, kk, 0) / dt;
            ln.rhoe_src = H(i, j, kk, 5) / dt;
            ln.reset_src = R(i, j, kk, 0);
            ln.e_src = (((k.asq * rhoe0 + dt * ln.rhoe_src) / k.aendsq + ln.reset_src) / (ln.rho + dt * ln.rho_src) - ln.e0) / dt;
        }
        if (k.inhomo) { ln.zhi = D(i, j, kk, 2); ln.jh = (k.z > ln.zhi) ? 0.0 : 1.0; }
        ln.rho_out = N(i, j, kk, 0); ln.rhoe_new = N(i, j, kk, 5);
        ln.start(k);
        // the cell data only the finalize step reads is not kept in the lane: fetched again when needed (as the kernel does)
        const double c_rhoe_src = ln.rhoe_src, c_reset_src = ln.reset_src, c_rho_out = ln.rho_out, c_rhoe_new = ln.rhoe_new;
        run_cell(ln, tb, k, [&](Lane<PATH_STRUCT, ArrHost>& l) { l.rhoe_src = c_rhoe_src; l.reset_src = c_reset_src; l.rho_out = c_rho_out; l.rhoe_new = c_rhoe_new; });
        D(i, j, kk, 0) = ln.outT; D(i, j, kk, 1) = ln.outNe;
        if (k.sdc_has_src) {
            I(i, j, kk, 0) = ln.IR;
            N(i, j, kk, 5) = N(i, j, kk, 5) + dt * k.ahalf * ln.IR / k.aendsq;
            N(i, j, kk, 4) = N(i, j, kk, 4) + dt * k.ahalf * ln.IR / k.aendsq;
        } else {
            S(i, j, kk, 5) += S(i, j, kk, 0) * (ln.e_final - ln.e0);
            S(i, j, kk, 4) += S(i, j, kk, 0) * (ln.e_final - ln.e0);
        }
        put_stats(cs, idx, ln.nst, ln.netf, ln.nfe, ln.nni, ln.nnf, ln.nsetups, ln.nfe_ls, ln.flag);
    }
    return 0;
}
}
