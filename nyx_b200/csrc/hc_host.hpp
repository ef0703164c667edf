// hc_host.hpp -- host side of the HeatCool path: rate tables (A1), UV-background interpolation (A4),
// per-launch constants. Pure C++ (no CUDA), shared by the C-ABI layer (hc_api.cu) and the host unit harness.
//
// Reference behaviour:
//   tabulate_rates            Source/EOS/atomic_rates.H:10-166 (Lukic et al. fits, Katz96 = 0 branch)
//   struct AtomicRates        Source/EOS/atomic_rates_data.H:22-49 (image layout used by hc_tabulate_rates)
//   interp_to_this_z          Source/EOS/eos_hc.H:10-49
//   ode_eos_setup / _initialize_single   Source/HeatCool/f_rhs_struct.H:45-153 (flash flags, JH/JHe)
#ifndef NYXB200_HC_HOST_HPP
#define NYXB200_HC_HOST_HPP

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nyx_hc.h"
#include "hc_device.cuh"

namespace hc {

constexpr int NCOOLFILE = HC_NCOOLFILE;
constexpr int NTAB = HC_NCOOLTAB + 1;

// offsets (in doubles) inside the AtomicRates image
enum RatesOffset : int {
    OFF_MEAN_RHOB = 0,
    OFF_LZR = 1, OFF_RGGH0 = OFF_LZR + NCOOLFILE, OFF_RGGHE0 = OFF_RGGH0 + NCOOLFILE, OFF_RGGHEP = OFF_RGGHE0 + NCOOLFILE,
    OFF_REH0 = OFF_RGGHEP + NCOOLFILE, OFF_REHE0 = OFF_REH0 + NCOOLFILE, OFF_REHEP = OFF_REHE0 + NCOOLFILE,
    OFF_TAB0 = OFF_REHEP + NCOOLFILE   // AlphaHp, AlphaHep, AlphaHepp, Alphad, GammaeH0, GammaeHe0, GammaeHep, BetaH0, BetaHe0,
                                       // BetaHep, Betaff1, Betaff4, RecHp, RecHep, RecHepp: 15 x NTAB
};

inline int tabulate_rates(const char* file, double mean_rhob, double* R) {
    FILE* fp = std::fopen(file, "r");
    if (!fp) return HC_ERR_IO;
    R[OFF_MEAN_RHOB] = mean_rhob;
    const int cols[7] = {OFF_LZR, OFF_RGGH0, OFF_RGGHE0, OFF_RGGHEP, OFF_REH0, OFF_REHE0, OFF_REHEP};
    for (int i = 0; i < NCOOLFILE; ++i)
        for (int c = 0; c < 7; ++c)
            if (std::fscanf(fp, "%lf", &R[cols[c] + i]) != 1) { std::fclose(fp); return HC_ERR_IO; }
    int extra = 0; double dummy;
    while (std::fscanf(fp, "%lf", &dummy) == 1) ++extra;
    std::fclose(fp);
    if (extra >= 7) return HC_ERR_TREECOOL_LEN;

    double* AlphaHp = R + OFF_TAB0;
    double* AlphaHep = AlphaHp + NTAB; double* AlphaHepp = AlphaHep + NTAB; double* Alphad = AlphaHepp + NTAB;
    double* GammaeH0 = Alphad + NTAB; double* GammaeHe0 = GammaeH0 + NTAB; double* GammaeHep = GammaeHe0 + NTAB;
    double* BetaH0 = GammaeHep + NTAB; double* BetaHe0 = BetaH0 + NTAB; double* BetaHep = BetaHe0 + NTAB;
    double* Betaff1 = BetaHep + NTAB; double* Betaff4 = Betaff1 + NTAB;
    double* RecHp = Betaff4 + NTAB; double* RecHep = RecHp + NTAB; double* RecHepp = RecHep + NTAB;

    const double deltaT = (TCOOLMAX - TCOOLMIN) / NCOOLTAB;
    const double tfac = std::pow(10.0, deltaT);
    double t = std::pow(10.0, TCOOLMIN);
    for (int i = 0; i < NTAB; ++i, t = t * tfac) {   // temperature advanced multiplicatively, atomic_rates.H:157
        const double st = std::sqrt(t), lt = std::log(t), l10 = std::log10(t);
        Alphad[i] = 1.90e-03 / (t * st) * std::exp(-4.7e5 / t) * (1.0e0 + 0.3e0 * std::exp(-9.4e4 / t));
        AlphaHp[i] = 7.982e-11 / (std::sqrt(t / 3.148e0) * std::pow((1.0e0 + std::sqrt(t / 3.148e0)), 0.252) * std::pow((1.0e0 + std::sqrt(t / 7.036e5)), 1.748));
        AlphaHep[i] = (t <= 1.0e6)
            ? 3.294e-11 / (std::sqrt(t / 15.54e0) * std::pow((1.0e0 + std::sqrt(t / 15.54e0)), 0.309) * std::pow((1.0e0 + std::sqrt(t / 3.676e7)), 1.691))
            : 9.356e-10 / (std::sqrt(t / 4.266e-2) * std::pow((1.0e0 + std::sqrt(t / 4.266e-2)), 0.2108) * std::pow((1.0e0 + std::sqrt(t / 4.677e6)), 1.7892));
        AlphaHepp[i] = 1.891e-10 / (std::sqrt(t / 9.37e0) * std::pow((1.0e0 + std::sqrt(t / 9.37e0)), 0.2476) * std::pow((1.0e0 + std::sqrt(t / 2.774e6)), 1.7524));
        double U = 1.16045e4 * 13.6e0 / t;
        GammaeH0[i] = 2.91e-8 * std::pow(U, 0.39) * std::exp(-U) / (0.232e0 + U);
        U = 1.16045e4 * 24.6e0 / t;
        GammaeHe0[i] = 1.75e-8 * std::pow(U, 0.35) * std::exp(-U) / (0.18e0 + U);
        U = 1.16045e4 * 54.4e0 / t;
        GammaeHep[i] = 2.05e-9 * (1.0e0 + std::sqrt(U)) * std::pow(U, 0.25) * std::exp(-U) / (0.265e0 + U);
        const double corr = 1.e0 / (1.0e0 + st / std::sqrt(5.0e7));
        const double y = lt;
        BetaH0[i] = (t <= 1.0e5)
            ? 1.0e-20 * std::exp(2.137913e2 - 1.139492e2 * y + 2.506062e1 * y * y - 2.762755e0 * y * y * y + 1.515352e-1 * y * y * y * y - 3.290382e-3 * y * y * y * y * y - 1.18415e5 / t)
            : 1.0e-20 * std::exp(2.7125446e2 - 9.8019455e1 * y + 1.400728e1 * y * y - 9.780842e-1 * y * y * y + 3.356289e-2 * y * y * y * y - 4.553323e-4 * y * y * y * y * y - 1.18415e5 / t);
        BetaHe0[i] = 9.38e-22 * st * std::exp(-285335.4e0 / t) * corr;
        BetaHep[i] = (5.54e-17 * std::pow(t, (-0.397e0)) * std::exp(-473638.0e0 / t) + 4.95e-22 * st * std::exp(-631515.0e0 / t)) * corr;
        RecHp[i] = 2.851e-27 * st * (5.914e0 - 0.5e0 * lt + 1.184e-2 * std::pow(t, (1.0e0 / 3.0e0)));
        RecHep[i] = 1.55e-26 * std::pow(t, 0.3647) + 1.24e-13 / (t * st) * std::exp(-4.7e5 / t) * (1.0e0 + 0.3e0 * std::exp(-9.4e4 / t));
        RecHepp[i] = 1.14e-26 * st * (6.607e0 - 0.5e0 * lt + 7.459e-3 * std::pow(t, (1.0e0 / 3.0e0)));
        Betaff1[i] = (t <= 3.2e5) ? 1.426e-27 * st * (0.79464e0 + 0.1243e0 * l10) : 1.426e-27 * st * (2.13164e0 - 0.1240e0 * l10);
        const double l10q = std::log10(t / 4.0e0);
        Betaff4[i] = (t / 4.0e0 <= 3.2e5) ? 1.426e-27 * st * 4.0e0 * (0.79464e0 + 0.1243e0 * l10q) : 1.426e-27 * st * 4.0e0 * (2.13164e0 - 0.1240e0 * l10q);
    }
    return HC_OK;
}

// interp_to_this_z on the rates image
inline Uvb uvb_at_z(const double* R, double z) {
    Uvb u{0, 0, 0, 0, 0, 0};
    const double* lzr = R + OFF_LZR;
    const double lopz = std::log10(1.0e0 + z);
    if (lopz >= lzr[NCOOLFILE - 1]) return u;
    int j = 0;
    if (!(lopz <= lzr[0])) for (int i = 1; i < NCOOLFILE; ++i) if (lopz < lzr[i]) { j = i - 1; break; }
    const double fact = (lopz - lzr[j]) / (lzr[j + 1] - lzr[j]);
    auto it = [&](int off) { return R[off + j] + (R[off + j + 1] - R[off + j]) * fact; };
    u.ggh0 = it(OFF_RGGH0); u.gghe0 = it(OFF_RGGHE0); u.gghep = it(OFF_RGGHEP);
    u.eh0 = it(OFF_REH0); u.ehe0 = it(OFF_REHE0); u.ehep = it(OFF_REHEP);
    return u;
}

// re-layout of the 15 tables for the kernels (one padding row so row j+1 always exists):
//   ionx [NTAB+1][6]  AlphaHp, AlphaHep, AlphaHepp, Alphad, GammaeH0, GammaeHe0
//   iony [NTAB+1]     GammaeHep
//   cool [NTAB+1][8]  BetaH0, BetaHe0, BetaHep, Betaff1, Betaff4, RecHp, RecHep, RecHepp
inline void interleave_tables(const double* R, std::vector<double>& ionx, std::vector<double>& iony, std::vector<double>& cool) {
    ionx.assign((size_t)(NTAB + 1) * IONX_ROW, 0.0);
    iony.assign((size_t)(NTAB + 1) + 1, 0.0);
    cool.assign((size_t)(NTAB + 1) * COOL_ROW, 0.0);
    for (int j = 0; j < NTAB; ++j) {
        for (int c = 0; c < IONX_ROW; ++c) ionx[(size_t)j * IONX_ROW + c] = R[OFF_TAB0 + c * NTAB + j];
        iony[j] = R[OFF_TAB0 + 6 * NTAB + j];
        for (int c = 0; c < COOL_ROW; ++c) cool[(size_t)j * COOL_ROW + c] = R[OFF_TAB0 + (7 + c) * NTAB + j];
    }
}

// the table of fast_log10 (hc_device.cuh): entry i = {r_i, Lhi_i, Llo_i, 0}, r_i = double(1/c_i), c_i = 1 + (i + 1/2)/128,
// Lhi_i + Llo_i = -log10(r_i) evaluated in 80-bit extended precision, Lhi_i truncated to a multiple of 2^-42
inline void build_log10_table(std::vector<double>& out) {
    out.assign((size_t)LOG_TAB_N * 4, 0.0);
    for (int i = 0; i < LOG_TAB_N; ++i) {
        const double r = (double)(1.0L / (1.0L + ((long double)i + 0.5L) / (long double)LOG_TAB_N));
        const long double L = -log10l((long double)r);
        const double Lhi = (double)(floorl(L * 4398046511104.0L) / 4398046511104.0L);   // 2^42
        out[4 * (size_t)i + 0] = r;
        out[4 * (size_t)i + 1] = Lhi;
        out[4 * (size_t)i + 2] = (double)(L - (long double)Lhi);
    }
}

inline void default_params(HcParams* p) {
    std::memset(p, 0, sizeof *p);
    p->rtol = 1e-4; p->atol_factor = 1e-4; p->h_species = 0.76; p->gamma_minus_1 = 5.0 / 3.0 - 1.0;
    p->uvb_density_A = 1.0; p->uvb_density_B = 0.0; p->zhi_flash = -1.0; p->zheii_flash = -1.0; p->T_zhi = 0.0; p->T_zheii = 0.0;
    p->max_steps = 2000; p->old_max_steps = 3;
}

inline void set_common(Consts& k, const HcParams& p, double dt, double gm1) {
    std::memset(&k, 0, sizeof k);
    k.rtol = p.rtol; k.atol_factor = p.atol_factor; k.tout = dt;
    k.hmax_inv = p.use_typical_steps ? 1.0 / (dt / (p.old_max_steps)) : 0.0;   // CVodeSetMaxStep(dt/old_max_steps): hmax_inv = 1/hmax
    k.max_steps = p.max_steps; k.use_constraint = p.use_constraint;
    k.h_species = p.h_species; k.gm1 = gm1;
    k.yhelium = (1.0 - p.h_species) / (4.0 * p.h_species);
    k.c_mu_num = 1.0e0 + 4.0e0 * k.yhelium;
    k.c_mu_den = 1.0e0 + k.yhelium;
    k.c_T = gm1 * MPROTON / BOLTZMANN;
    k.dt = dt;
}
inline void set_rhs_z(Consts& k, const double* R, double z) {
    k.z = z; k.opz = 1.0e0 + std::fabs(z);
    k.opz4 = k.opz * k.opz * k.opz * k.opz;
    k.tcmb_opz = 2.725e0 * k.opz;
    k.a_rhs = 1.e0 / (1.e0 + std::fabs(z));
    k.uvb_rhs = uvb_at_z(R, z);
    k.mean_rhob = R[OFF_MEAN_RHOB];
}

// Strang path (integrate_state_vec_mfin): rpar[3] = 1/a - 1; finalize uses a' = 1/(z+1) (f_rhs.H:73)
inline Consts make_consts_vec(const double* R, const HcParams& p, double a, double dt) {
    Consts k; set_common(k, p, dt, 2.0 / 3.0);
    const double z = 1 / a - 1;
    set_rhs_z(k, R, z);
    const double a_fin = 1 / (z + 1.0);
    k.a3_eos = a_fin * a_fin * a_fin;
    k.uvb_eos = uvb_at_z(R, 1.e0 / a_fin - 1.e0);
    k.JH0 = 1; k.JHe0 = 1; k.uvb_A = 1.0; k.uvb_B = 0.0; k.a = a;
    return k;
}

// SDC path (integrate_state_struct_mfin)
inline Consts make_consts_struct(const double* R, const HcParams& p, double a, double a_end, double dt, int sdc_iter) {
    Consts k; set_common(k, p, dt, p.gamma_minus_1);
    const double z = 1 / a - 1.0;
    set_rhs_z(k, R, z);
    k.a3_eos = a * a * a;
    k.uvb_eos = uvb_at_z(R, 1.e0 / a - 1.e0);
    k.sdc_has_src = (sdc_iter >= 0);
    k.a = a; k.a_end = a_end; k.asq = a * a; k.aendsq = a_end * a_end; k.ahalf = 0.5 * (a + a_end);
    k.z_end = 1 / (a_end) - 1.0;
    k.uvb_A = p.uvb_density_A; k.uvb_B = p.uvb_density_B;
    k.inhomo = p.inhomo_reion;
    k.flash_h = (p.zhi_flash > 0.0) ? ((p.inhomo_reion > 0) ? 0 : 1) : 0;
    k.flash_he = (p.zheii_flash > 0.0) ? 1 : 0;
    k.JH0 = ((k.flash_h == 1) && (z > p.zhi_flash)) ? 0 : 1;
    k.JHe0 = ((k.flash_he == 1) && (z > p.zheii_flash)) ? 0 : 1;
    k.H_reion_z = p.zhi_flash; k.He_reion_z = p.zheii_flash; k.T_zhi = p.T_zhi; k.T_zheii = p.T_zheii;
    return k;
}

// compute_new_temp (Nyx.cpp:2473-2490): EOS with the caller's a, JH = JHe = 1
inline Consts make_consts_eos(const double* R, const HcParams& p, double a) {
    Consts k; set_common(k, p, 0.0, p.gamma_minus_1);
    set_rhs_z(k, R, 1.e0 / a - 1.e0);
    k.a3_eos = a * a * a;
    k.uvb_eos = uvb_at_z(R, 1.e0 / a - 1.e0);
    k.JH0 = 1; k.JHe0 = 1; k.a = a;
    return k;
}

}  // namespace hc
#endif
