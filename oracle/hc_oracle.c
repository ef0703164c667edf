/*
 * Licence note: the BDF / Newton / diagonal-solver logic below follows SUNDIALS CVODE 6.3.0 (BSD 3-Clause, Copyright (c) 2002-2022 Lawrence
 * Livermore National Security and Southern Methodist University) and the right-hand side / EOS follow Nyx (BSD-style, Copyright (c) 2017 The
 * Regents of the University of California, through Lawrence Berkeley National Laboratory) statement by statement where identical results
 * require it; both notices are reproduced in THIRD_PARTY_NOTICES.md.
 */
/* TEST INFRASTRUCTURE ONLY -- see hc_oracle.h.  Plain C11, compiled with -ffp-contract=off so that
 * every product and sum is rounded separately, as in the reference's x86-64 GCC -O3 build.
 *
 * Path shorthands: EOS/ = Source/EOS/, HC/ = Source/HeatCool/, DRV/ = Source/Driver/,
 * SUN/ = subprojects/sundials/src/.
 */
#include "hc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ constants */
/* EOS/atomic_rates_data.H:9-20 */
#define NCOOLFILE HCO_NCOOLFILE
#define NCOOLTAB HCO_NCOOLTAB
static const double TCOOLMAX = 9.0, TCOOLMIN = 0.0, XACC = 1e-6;
static const double MPROTON = 1.6726230999999999E-024, BOLTZMANN = 1.3806000442045675E-016;
/* DRV/constants_cosmo.H:7-50, same expression order */
#define M_UNIT 1.98848e33
#define L_UNIT 3.0856776e24
#define V_UNIT 1.e5
#define T_UNIT (L_UNIT / V_UNIT)
static const double K_B = 1.38064852e-16 * T_UNIT * T_UNIT / (M_UNIT * L_UNIT * L_UNIT);
static const double M_PROTON_CODE = 1.672621e-24 / M_UNIT;
#define MP_OVER_KB (M_PROTON_CODE / K_B)
static const double DENSITY_TO_CGS = M_UNIT / (L_UNIT * L_UNIT * L_UNIT);
static const double E_TO_CGS = V_UNIT * V_UNIT;
static const double HEAT_FROM_CGS = L_UNIT * (T_UNIT * T_UNIT * T_UNIT / M_UNIT);

/* amrex::max(x, 0.0) == (x < 0.0) ? 0.0 : x  (a NaN passes through) */
static inline double dmax_amrex(double x) { return (x < 0.0) ? 0.0 : x; }

void hco_default_params(hco_params* p) {
    memset(p, 0, sizeof *p);
    p->rtol = 1e-4; p->atol_factor = 1e-4; p->h_species = 0.76; p->gamma_minus_1 = 5.0 / 3.0 - 1.0;
    p->max_steps = 2000; p->old_max_steps = 3; p->uvb_density_A = 1.0; p->uvb_density_B = 0.0;
    p->zhi_flash = -1.0; p->zheii_flash = -1.0; p->T_zhi = 0.0; p->T_zheii = 0.0;
}

/* ------------------------------------------------------------------ A1: tabulate_rates, EOS/atomic_rates.H:10-166 (Katz96 = 0 branch) */
int hco_tabulate_rates(const char* file, double mean_rhob, hco_rates* r) {
    FILE* fp = fopen(file, "r");
    if (!fp) return -1;
    r->mean_rhob = mean_rhob;
    for (int i = 0; i < NCOOLFILE; ++i) {
        if (fscanf(fp, "%lf %lf %lf %lf %lf %lf %lf", &r->lzr[i], &r->rggh0[i], &r->rgghe0[i], &r->rgghep[i], &r->reh0[i],
                   &r->rehe0[i], &r->rehep[i]) != 7) { fclose(fp); return -3; }
    }
    /* :37-51 the reference aborts when more than NCOOLFILE rows are present */
    double extra; int nextra = 0;
    while (fscanf(fp, "%lf", &extra) == 1) ++nextra;
    fclose(fp);
    if (nextra >= 7) return -2;

    const double deltaT = (TCOOLMAX - TCOOLMIN) / NCOOLTAB;
    double t = pow(10, TCOOLMIN);
    for (int i = 0; i <= NCOOLTAB; ++i) {
        const double sqrt_t = sqrt(t);
        r->Alphad[i] = 1.90e-03 / (t * sqrt_t) * exp(-4.7e5 / t) * (1.0e0 + 0.3e0 * exp(-9.4e4 / t));
        r->AlphaHp[i] = 7.982e-11 / (sqrt(t / 3.148e0) * pow((1.0e0 + sqrt(t / 3.148e0)), 0.252) * pow((1.0e0 + sqrt(t / 7.036e5)), 1.748));
        if (t <= 1.0e6)
            r->AlphaHep[i] = 3.294e-11 / (sqrt(t / 15.54e0) * pow((1.0e0 + sqrt(t / 15.54e0)), 0.309) * pow((1.0e0 + sqrt(t / 3.676e7)), 1.691));
        else
            r->AlphaHep[i] = 9.356e-10 / (sqrt(t / 4.266e-2) * pow((1.0e0 + sqrt(t / 4.266e-2)), 0.2108) * pow((1.0e0 + sqrt(t / 4.677e6)), 1.7892));
        r->AlphaHepp[i] = 1.891e-10 / (sqrt(t / 9.37e0) * pow((1.0e0 + sqrt(t / 9.37e0)), 0.2476) * pow((1.0e0 + sqrt(t / 2.774e6)), 1.7524));

        double E = 13.6e0, U = 1.16045e4 * E / t;
        r->GammaeH0[i] = 2.91e-8 * pow(U, 0.39) * exp(-U) / (0.232e0 + U);
        E = 24.6e0; U = 1.16045e4 * E / t;
        r->GammaeHe0[i] = 1.75e-8 * pow(U, 0.35) * exp(-U) / (0.18e0 + U);
        E = 54.4e0; U = 1.16045e4 * E / t;
        r->GammaeHep[i] = 2.05e-9 * (1.0e0 + sqrt(U)) * pow(U, 0.25) * exp(-U) / (0.265e0 + U);

        const double corr_term = 1.e0 / (1.0e0 + sqrt_t / sqrt(5.0e7));
        const double y = log(t);
        if (t <= 1.0e5)
            r->BetaH0[i] = 1.0e-20 * exp(2.137913e2 - 1.139492e2 * y + 2.506062e1 * y * y - 2.762755e0 * y * y * y +
                                          1.515352e-1 * y * y * y * y - 3.290382e-3 * y * y * y * y * y - 1.18415e5 / t);
        else
            r->BetaH0[i] = 1.0e-20 * exp(2.7125446e2 - 9.8019455e1 * y + 1.400728e1 * y * y - 9.780842e-1 * y * y * y +
                                          3.356289e-2 * y * y * y * y - 4.553323e-4 * y * y * y * y * y - 1.18415e5 / t);
        r->BetaHe0[i] = 9.38e-22 * sqrt_t * exp(-285335.4e0 / t) * corr_term;
        r->BetaHep[i] = (5.54e-17 * pow(t, (-0.397e0)) * exp(-473638.0e0 / t) + 4.95e-22 * sqrt_t * exp(-631515.0e0 / t)) * corr_term;

        r->RecHp[i] = 2.851e-27 * sqrt_t * (5.914e0 - 0.5e0 * log(t) + 1.184e-2 * pow(t, (1.0e0 / 3.0e0)));
        r->RecHep[i] = 1.55e-26 * pow(t, 0.3647) + 1.24e-13 / (t * sqrt_t) * exp(-4.7e5 / t) * (1.0e0 + 0.3e0 * exp(-9.4e4 / t));
        r->RecHepp[i] = 1.14e-26 * sqrt_t * (6.607e0 - 0.5e0 * log(t) + 7.459e-3 * pow(t, (1.0e0 / 3.0e0)));

        if (t <= 3.2e5) r->Betaff1[i] = 1.426e-27 * sqrt_t * (0.79464e0 + 0.1243e0 * log10(t));
        else            r->Betaff1[i] = 1.426e-27 * sqrt_t * (2.13164e0 - 0.1240e0 * log10(t));
        if (t / 4.0e0 <= 3.2e5) r->Betaff4[i] = 1.426e-27 * sqrt_t * 4.0e0 * (0.79464e0 + 0.1243e0 * log10(t / 4.0e0));
        else                    r->Betaff4[i] = 1.426e-27 * sqrt_t * 4.0e0 * (2.13164e0 - 0.1240e0 * log10(t / 4.0e0));

        t = t * pow(10, deltaT);   /* :157 multiplicative advance */
    }
    return 0;
}

/* ------------------------------------------------------------------ A4: interp_to_this_z, EOS/eos_hc.H:10-49 */
void hco_interp_to_this_z(const hco_rates* r, double z, double o[6]) {
    const double lopz = log10(1.0e0 + z);
    if (lopz >= r->lzr[NCOOLFILE - 1]) { for (int q = 0; q < 6; ++q) o[q] = 0.0; return; }
    int j = 1;   /* 1-based like the reference */
    if (lopz <= r->lzr[0]) j = 1;
    else for (int i = 2; i <= NCOOLFILE; ++i) if (lopz < r->lzr[i - 1]) { j = i - 1; break; }
    const double fact = (lopz - r->lzr[j - 1]) / (r->lzr[j] - r->lzr[j - 1]);
    o[0] = r->rggh0[j - 1] + (r->rggh0[j] - r->rggh0[j - 1]) * fact;
    o[1] = r->rgghe0[j - 1] + (r->rgghe0[j] - r->rgghe0[j - 1]) * fact;
    o[2] = r->rgghep[j - 1] + (r->rgghep[j] - r->rgghep[j - 1]) * fact;
    o[3] = r->reh0[j - 1] + (r->reh0[j] - r->reh0[j - 1]) * fact;
    o[4] = r->rehe0[j - 1] + (r->rehe0[j] - r->rehe0[j - 1]) * fact;
    o[5] = r->rehep[j - 1] + (r->rehep[j] - r->rehep[j - 1]) * fact;
}

/* ------------------------------------------------------------------ A5: ion_n_device, EOS/eos_hc.H:51-135 */
static void ion_n(const hco_rates* r, int JH, int JHe, double U, double nh, double ne, double* nhp, double* nhep,
                  double* nhepp, double* t, double gamma_minus_1, double h_species, double z) {
    const double smallest_val = DBL_MIN;
    const double deltaT = (TCOOLMAX - TCOOLMIN) / NCOOLTAB;
    const double YHELIUM = (1.0 - h_species) / (4.0 * h_species);
    const double mu = (1.0e0 + 4.0e0 * YHELIUM) / (1.0e0 + YHELIUM + ne);
    *t = gamma_minus_1 * MPROTON / BOLTZMANN * U * mu;
    double logT = log10(*t);
    if (logT >= TCOOLMAX) { *nhp = 1.0e0; *nhep = 0.0e0; *nhepp = YHELIUM; return; }
    if (logT <= TCOOLMIN) logT = TCOOLMIN + 0.5e0 * deltaT;
    const double tmp = (logT - TCOOLMIN) / deltaT;
    int j = (int)floor(tmp);
    const double fhi = tmp - j;
    const double flo = 1.0e0 - fhi;
    /* 0-based j here == reference's (j+1)-1 */
    const double ahp = flo * r->AlphaHp[j] + fhi * r->AlphaHp[j + 1];
    const double ahep = flo * r->AlphaHep[j] + fhi * r->AlphaHep[j + 1];
    const double ahepp = flo * r->AlphaHepp[j] + fhi * r->AlphaHepp[j + 1];
    const double ad = flo * r->Alphad[j] + fhi * r->Alphad[j + 1];
    const double geh0 = flo * r->GammaeH0[j] + fhi * r->GammaeH0[j + 1];
    const double gehe0 = flo * r->GammaeHe0[j] + fhi * r->GammaeHe0[j + 1];
    const double gehep = flo * r->GammaeHep[j] + fhi * r->GammaeHep[j + 1];
    double uvb[6];
    hco_interp_to_this_z(r, z, uvb);
    double ggh0ne, gghe0ne, gghepne;
    if (ne > 0.0) {
        ggh0ne = JH * uvb[0] / (ne * nh);
        gghe0ne = JH * uvb[1] / (ne * nh);
        gghepne = JHe * uvb[2] / (ne * nh);
    } else { ggh0ne = 0.0; gghe0ne = 0.0; gghepne = 0.0; }
    *nhp = 1.0e0 - ahp / (ahp + geh0 + ggh0ne);
    if ((gehe0 + gghe0ne) > smallest_val)
        *nhep = YHELIUM / (1.0e0 + (ahep + ad) / (gehe0 + gghe0ne) + (gehep + gghepne) / ahepp);
    else
        *nhep = 0.0e0;
    if (*nhep > 0.0e0) *nhepp = *nhep * (gehep + gghepne) / ahepp;
    else *nhepp = 0.0e0;
}

void hco_ion_n(const hco_rates* r, int JH, int JHe, double U, double nh, double ne, double gm1, double hsp, double z, double o[4]) {
    ion_n(r, JH, JHe, U, nh, ne, &o[0], &o[1], &o[2], &o[3], gm1, hsp, z);
}

/* ------------------------------------------------------------------ A6: iterate_ne_device, EOS/eos_hc.H:138-188. Returns Newton iterations used. */
static int iterate_ne(const hco_rates* r, int JH, int JHe, double z, double U, double* t, double nh, double* ne, double* nh0,
                      double* nhp, double* nhe0, double* nhep, double* nhepp, double gamma_minus_1, double h_species) {
    const double YHELIUM = (1.0 - h_species) / (4.0 * h_species);
    int i, iters = 0;
    *ne = 1.0e0;
    for (i = 1; i <= 15; ++i) {
        double eps, nhp_plus, nhep_plus, nhepp_plus;
        ++iters;
        ion_n(r, JH, JHe, U, nh, *ne, nhp, nhep, nhepp, t, gamma_minus_1, h_species, z);
        if (*ne > 0.0e0) eps = XACC * (*ne); else eps = 1.0e-24;
        const double ne2 = *ne + eps;
        ion_n(r, JH, JHe, U, nh, ne2, &nhp_plus, &nhep_plus, &nhepp_plus, t, gamma_minus_1, h_species, z);
        const double dnhp_dne = (nhp_plus - *nhp) / eps;
        const double dnhep_dne = (nhep_plus - *nhep) / eps;
        const double dnhepp_dne = (nhepp_plus - *nhepp) / eps;
        const double f = *ne - *nhp - *nhep - 2.0e0 * (*nhepp);
        const double df = 1.0e0 - dnhp_dne - dnhep_dne - 2.0e0 * dnhepp_dne;
        const double dne = f / df;
        *ne = dmax_amrex(*ne - dne);   /* amrex::max((ne-dne), 0.0) */
        if (fabs(dne) < XACC) break;
    }
    ion_n(r, JH, JHe, U, nh, *ne, nhp, nhep, nhepp, t, gamma_minus_1, h_species, z);
    *nh0 = 1.0e0 - *nhp;
    *nhe0 = YHELIUM - (*nhep + *nhepp);
    return iters;
}

int hco_iterate_ne(const hco_rates* r, int JH, int JHe, double z, double U, double nh, double gm1, double hsp, double o[7]) {
    return iterate_ne(r, JH, JHe, z, U, &o[0], nh, &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], gm1, hsp);
}

/* ------------------------------------------------------------------ A7: nyx_eos_T_given_Re_device, EOS/eos_hc.H:190-220.
 * sp = {nh0, nhp, nhe0, nhep, nhepp} in the callee's (declaration) order. */
static int eos_T_given_Re(const hco_rates* r, double gamma_minus_1, double h_species, int JH, int JHe, double* T, double* Ne,
                          double R, double e, double comoving_a, double sp[5]) {
    const double rho = R * DENSITY_TO_CGS / (comoving_a * comoving_a * comoving_a);
    const double U = e * E_TO_CGS;
    const double nh = rho * h_species / MPROTON;
    const double z = 1.e0 / comoving_a - 1.e0;
    return iterate_ne(r, JH, JHe, z, U, T, nh, Ne, &sp[0], &sp[1], &sp[2], &sp[3], &sp[4], gamma_minus_1, h_species);
}

void hco_eos_T_given_Re(const hco_rates* r, double gm1, double hsp, int JH, int JHe, double R, double e, double a, double* T, double* Ne) {
    double sp[5];
    eos_T_given_Re(r, gm1, hsp, JH, JHe, T, Ne, R, e, a, sp);
}

/* ------------------------------------------------------------------ per-cell RHS context */
typedef struct cellctx {
    const hco_rates* r;
    int is_struct;
    double h_species, gamma_minus_1;
    /* Strang path: rpar = {T, ne, rho, z} (HC/integrate_state_vec_3d.cpp:228-232) */
    double rpar[4];
    /* SDC path: RhsData scalars + this cell's slots of its arrays (HC/f_rhs_struct.H:10-43) */
    double a, uvb_density_A, uvb_density_B;
    int JH, JHe, has_src;
    double T, ne, rho, rho_init, rho_src, rhoe_src, e_src;
    long ne_iters;
} cellctx;

/* shared tail of f_rhs_rpar (HC/f_rhs.H:170-248) and f_rhs_struct (HC/f_rhs_struct.H:484-584) */
static double rhs_core(cellctx* c, double* e_in, double rho_vode, double z_vode, int JH, int JHe, double gamma_minus_1,
                       double h_species, double uvbA, double uvbB, double* T_out, double* ne_out, int add_src) {
    const hco_rates* r = c->r;
    const double compt_c = 1.01765467e-37, T_cmb = 2.725e0;
    const double deltaT = (TCOOLMAX - TCOOLMIN) / NCOOLTAB;
    double T_vode, ne_vode, nh0, nhp, nhe0, nhep, nhepp, energy;
    if (*e_in <= 0 || isnan(*e_in)) *e_in = DBL_MIN;

    const double rho = rho_vode * DENSITY_TO_CGS * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode));
    const double U = *e_in * E_TO_CGS;
    const double nh = rho * h_species / MPROTON;
    c->ne_iters += iterate_ne(r, JH, JHe, z_vode, U, &T_vode, nh, &ne_vode, &nh0, &nhp, &nhe0, &nhep, &nhepp, gamma_minus_1, h_species);
    ne_vode = nh * ne_vode; nh0 = nh * nh0; nhp = nh * nhp; nhe0 = nh * nhe0; nhep = nh * nhep; nhepp = nh * nhepp;

    double logT = log10(T_vode);
    if (logT >= TCOOLMAX) {
        const double lambda_ff = 1.42e-27 * sqrt(T_vode) * (1.1e0 + 0.34e0 * exp(-(5.5e0 - logT) * (5.5e0 - logT) / 3.0e0)) * (nhp + 4.0e0 * nhepp) * ne_vode;
        const double lambda_c = compt_c * T_cmb * T_cmb * T_cmb * T_cmb * ne_vode * (T_vode - T_cmb * (1.0e0 + fabs(z_vode))) *
                                (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode));
        energy = (-lambda_ff - lambda_c) * HEAT_FROM_CGS / ((1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)));
        if (add_src) energy = energy / rho_vode * (1.0e0 + fabs(z_vode)) + c->e_src;
        else         energy = energy / rho_vode * (1.0e0 + fabs(z_vode));
        *T_out = T_vode; *ne_out = ne_vode / nh;
        return energy;
    }
    if (logT <= TCOOLMIN) logT = TCOOLMIN + 0.5e0 * deltaT;
    const double tmp = (logT - TCOOLMIN) / deltaT;
    const int j = (int)floor(tmp);
    const double fhi = tmp - j, flo = 1.0e0 - fhi;
    const double bh0 = flo * r->BetaH0[j] + fhi * r->BetaH0[j + 1];
    const double bhe0 = flo * r->BetaHe0[j] + fhi * r->BetaHe0[j + 1];
    const double bhep = flo * r->BetaHep[j] + fhi * r->BetaHep[j + 1];
    const double bff1 = flo * r->Betaff1[j] + fhi * r->Betaff1[j + 1];
    const double bff4 = flo * r->Betaff4[j] + fhi * r->Betaff4[j + 1];
    const double rhp = flo * r->RecHp[j] + fhi * r->RecHp[j + 1];
    const double rhep = flo * r->RecHep[j] + fhi * r->RecHep[j + 1];
    const double rhepp = flo * r->RecHepp[j] + fhi * r->RecHepp[j + 1];

    double lambda = (bh0 * nh0 + bhe0 * nhe0 + bhep * nhep + rhp * nhp + rhep * nhep + rhepp * nhepp + bff1 * (nhp + nhep) + bff4 * nhepp) * ne_vode;
    const double lambda_c = compt_c * T_cmb * T_cmb * T_cmb * T_cmb * ne_vode * (T_vode - T_cmb * (1.0e0 + fabs(z_vode))) *
                            (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode));
    lambda = lambda + lambda_c;
    double uvb[6];
    hco_interp_to_this_z(r, z_vode, uvb);
    double heat = JH * nh0 * uvb[3] + JH * nhe0 * uvb[4] + JHe * nhep * uvb[5];
    const double rho_heat = uvbA * pow((rho_vode / r->mean_rhob), uvbB);
    heat = rho_heat * heat;
    ne_vode = ne_vode / nh;
    energy = (heat - lambda) * HEAT_FROM_CGS / ((1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)) * (1.0e0 + fabs(z_vode)));
    const double a = 1.e0 / (1.e0 + fabs(z_vode));
    if (add_src) energy = energy / rho_vode / a + c->e_src;
    else         energy = energy / rho_vode / a;
    *T_out = T_vode; *ne_out = ne_vode;
    return energy;
}

/* A8 f_rhs_rpar (HC/f_rhs.H:111-249) / A12 f_rhs_struct (HC/f_rhs_struct.H:448-585) */
static double rhs(cellctx* c, double t, double* y) {
    if (!c->is_struct) {
        /* hard-wired JH = JHe = 1, uvb A=1 B=0, gamma-1 = 2/3 (f_rhs.H:131-133,159-160) */
        double T, ne;
        const double en = rhs_core(c, y, c->rpar[2], c->rpar[3], 1, 1, 2.0 / 3.0, c->h_species, 1.0, 0.0, &T, &ne, 0);
        c->rpar[0] = T; c->rpar[1] = ne;
        return en;
    }
    double rho_vode;
    if (c->has_src) rho_vode = c->rho_init + t * c->rho_src;   /* f_rhs_struct.H:476 */
    else rho_vode = c->rho;
    const double z_vode = 1 / (c->a) - 1.0;
    double T, ne;
    const double en = rhs_core(c, y, rho_vode, z_vode, c->JH, c->JHe, c->gamma_minus_1, c->h_species, c->uvb_density_A,
                               c->uvb_density_B, &T, &ne, c->has_src);
    c->T = T; c->ne = ne; c->rho = rho_vode;
    return en;
}

double hco_f_rhs_rpar(const hco_rates* r, double h_species, double* e_in, double rpar[4]) {
    cellctx c; memset(&c, 0, sizeof c);
    c.r = r; c.h_species = h_species; memcpy(c.rpar, rpar, sizeof c.rpar);
    const double en = rhs(&c, 0.0, e_in);
    memcpy(rpar, c.rpar, sizeof c.rpar);
    return en;
}

/* ================================================================== scalar (N = 1) CVODE
 * N_Vector semantics for one component, SUN/nvector/serial/nvector_serial.c. */
static inline double nv_linsum(double a, double x, double b, double y) {   /* :386-470 dispatch */
    if (a == 1.0 && b == 1.0) return x + y;
    if (a == 1.0 && b == -1.0) return x - y;
    if (a == -1.0 && b == 1.0) return y - x;
    if (a == 1.0) return (b * y) + x;
    if (b == 1.0) return (a * x) + y;
    if (a == -1.0) return (b * y) - x;
    if (b == -1.0) return (a * x) - y;
    if (a == b) return a * (x + y);
    if (a == -b) return a * (x - y);
    return (a * x) + (b * y);
}
static inline double nv_scale(double c, double x) { if (c == 1.0) return x; if (c == -1.0) return -x; return c * x; }
static inline double nv_wrms(double x, double w) { const double p = x * w; const double s = p * p; return (s <= 0.0) ? 0.0 : sqrt(s / 1); }
static inline double sun_powr(double b, double e) { return (b <= 0.0) ? 0.0 : pow(b, e); }   /* SUN/sundials/sundials_math.c:62-75 */
static inline double sun_powi(double b, int e) { double p = 1.0; int n = abs(e); for (int i = 1; i <= n; ++i) p *= b; if (e < 0) p = 1.0 / p; return p; }
static inline double dmax(double a, double b) { return (a > b) ? a : b; }   /* SUNMAX: ((A) > (B)) ? (A) : (B) */
static inline double dmin(double a, double b) { return (a < b) ? a : b; }   /* SUNMIN */

/* CVODE return codes (SUN/../include/cvode/cvode.h) and internal control constants (SUN/cvode/cvode_impl.h) */
enum { CV_SUCCESS = 0, CV_TOO_MUCH_WORK = -1, CV_TOO_MUCH_ACC = -2, CV_ERR_FAILURE = -3, CV_CONV_FAILURE = -4,
       CV_LSETUP_FAIL = -6, CV_LSOLVE_FAIL = -7, CV_CONSTR_FAIL = -15, CV_ILL_INPUT = -22, CV_TOO_CLOSE = -27 };
enum { DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, TRY_AGAIN = 5, FIRST_CALL = 6, PREV_CONV_FAIL = 7, PREV_ERR_FAIL = 8,
       CONSTR_RECVR = 10, NLS_CONTINUE = 901, NLS_CONV_RECVR = 902 };
#define QMAX 5
#define ETA_MIN_FX 0.0
#define ETA_MAX_FX 1.5
#define ETA_MAX_FS 10000.0
#define ETA_MAX_ES 10.0
#define ETA_MAX_GS 10.0
#define ETA_MIN 0.1
#define ETA_MAX_EF 0.2
#define ETA_MIN_EF 0.1
#define ETA_CF 0.25
#define SMALL_NST 10
#define SMALL_NEF 2
#define ONEPSM 1.000001
#define ADDON 0.000001
#define BIAS1 6.0
#define BIAS2 6.0
#define BIAS3 10.0
#define LONG_WAIT 10
#define MXNCF 10
#define MXNEF 7
#define MXNEF1 3
#define MSBP 20
#define DGMAX 0.3
#define NLS_MAXCOR 3
#define CRDOWN 0.3
#define RDIV 2.0
#define FRACT 0.1
#define CORTES 0.1

typedef struct cvs {
    cellctx* c;
    double uround, reltol, Vabstol; int atolmin0;
    double zn[QMAX + 1], ewt, y, acor, tempv, ftemp;
    double tn, h, hprime, next_h, eta, hscale, hu, h0u, hmin, hmax_inv, etamax;
    int q, qprime, next_q, qwait, L, qu, indx_acor;
    double tau[QMAX + 2], tq[6], l[QMAX + 2];
    double rl1, gamma, gammap, gamrat, crate, delp, acnrm, saved_tq5, etaqm1, etaq, etaqp1;
    long nst, nfe, ncfn, nnf, netf, nni, nsetups, nstlp, nfeDI, attempts, mxstep;
    int jcur, nls_jcur, use_constraint;
    double M, gammasv;
} cvs;

static int cv_ewt(cvs* m, double ycur, double* w) {   /* cvEwtSetSV, SUN/cvode/cvode.c:4413-4441 */
    double tv = fabs(ycur);
    tv = nv_linsum(m->reltol, tv, 1.0, m->Vabstol);
    if (m->atolmin0) { if (tv <= 0.0) return -1; }
    m->tempv = tv;
    *w = 1.0 / tv;
    return 0;
}

static double cv_f(cvs* m, double t, double* y) { return rhs(m->c, t, y); }

/* SUN/cvode/cvode.c:2054-2090 */
static double cv_upper_bound_h0(cvs* m, double tdist) {
    double temp2 = fabs(m->zn[0]);
    double temp1; cv_ewt(m, m->zn[0], &temp1);
    temp1 = 1.0 / temp1;
    temp1 = nv_linsum(0.1, temp2, 1.0, temp1);
    temp2 = fabs(m->zn[1]);
    temp1 = temp2 / temp1;
    m->tempv = temp1; m->acor = temp2;
    const double hub_inv = fabs(temp1);
    double hub = 0.1 * tdist;
    if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
    return hub;
}

/* SUN/cvode/cvode.c:2099-2115 */
static double cv_ydd_norm(cvs* m, double hg) {
    m->y = nv_linsum(hg, m->zn[1], 1.0, m->zn[0]);
    m->tempv = cv_f(m, m->tn + hg, &m->y); m->nfe++;
    m->tempv = nv_linsum(1.0 / hg, m->tempv, -1.0 / hg, m->zn[1]);
    return nv_wrms(m->tempv, m->ewt);
}

/* SUN/cvode/cvode.c:1945-2045 (the RHS never reports failure, so the recoverable-error retries are dead) */
static int cv_hin(cvs* m, double tout) {
    double tdiff, hg, hs, hnew = 0.0, hrat, h0, yddnrm;
    if ((tdiff = tout - m->tn) == 0.0) return CV_TOO_CLOSE;
    const int sign = (tdiff > 0.0) ? 1 : -1;
    const double tdist = fabs(tdiff);
    const double tround = m->uround * dmax(fabs(m->tn), fabs(tout));
    if (tdist < 2.0 * tround) return CV_TOO_CLOSE;
    const double hlb = 100.0 * tround;
    const double hub = cv_upper_bound_h0(m, tdist);
    hg = sqrt(hlb * hub);
    if (hub < hlb) { m->h = (sign == -1) ? -hg : hg; return CV_SUCCESS; }
    hs = hg;
    for (int count1 = 1; count1 <= 4; ++count1) {
        const double hgs = hg * sign;
        yddnrm = cv_ydd_norm(m, hgs);
        hs = hg;
        hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
        if (count1 == 4) break;
        hrat = hnew / hg;
        if ((hrat > 0.5) && (hrat < 2.0)) break;
        if ((count1 > 1) && (hrat > 2.0)) { hnew = hg; break; }
        hg = hnew;
    }
    (void)hs;
    h0 = 0.5 * hnew;
    if (h0 < hlb) h0 = hlb;
    if (h0 > hub) h0 = hub;
    if (sign == -1) h0 = -h0;
    m->h = h0;
    return CV_SUCCESS;
}

/* SUN/cvode/cvode.c:2457-2473 */
static void cv_rescale(cvs* m) {
    double cvals[QMAX + 1];
    cvals[0] = m->eta;
    for (int j = 1; j <= m->q; ++j) cvals[j] = m->eta * cvals[j - 1];
    for (int j = 0; j < m->q; ++j) m->zn[j + 1] = nv_scale(cvals[j], m->zn[j + 1]);
    m->h = m->hscale * m->eta;
    m->next_h = m->h;
    m->hscale = m->h;
}

/* SUN/cvode/cvode.c:2383-2419 */
static void cv_increase_bdf(cvs* m) {
    double alpha0, alpha1, prod, xi, xiold, hsum, A1;
    for (int i = 0; i <= QMAX; ++i) m->l[i] = 0.0;
    m->l[2] = alpha1 = prod = xiold = 1.0;
    alpha0 = -1.0;
    hsum = m->hscale;
    if (m->q > 1) {
        for (int j = 1; j < m->q; ++j) {
            hsum += m->tau[j + 1];
            xi = hsum / m->hscale;
            prod *= xi;
            alpha0 -= 1.0 / (j + 1);
            alpha1 += 1.0 / xi;
            for (int i = j + 2; i >= 2; --i) m->l[i] = m->l[i] * xiold + m->l[i - 1];
            xiold = xi;
        }
    }
    A1 = (-alpha0 - alpha1) / prod;
    m->zn[m->L] = nv_scale(A1, m->zn[m->indx_acor]);
    if (m->q > 1)
        for (int j = 2; j <= m->q; ++j) m->zn[j] = nv_linsum(m->l[j], m->zn[m->L], 1.0, m->zn[j]);
}

/* SUN/cvode/cvode.c:2431-2454 */
static void cv_decrease_bdf(cvs* m) {
    double hsum, xi;
    for (int i = 0; i <= QMAX; ++i) m->l[i] = 0.0;
    m->l[2] = 1.0;
    hsum = 0.0;
    for (int j = 1; j <= m->q - 2; ++j) {
        hsum += m->tau[j];
        xi = hsum / m->hscale;
        for (int i = j + 2; i >= 2; --i) m->l[i] = m->l[i] * xi + m->l[i - 1];
    }
    if (m->q > 2)
        for (int j = 2; j < m->q; ++j) m->zn[j] = nv_linsum(-m->l[j], m->zn[m->q], 1.0, m->zn[j]);
}

/* cvAdjustOrder, SUN/cvode/cvode.c:2286-2298 */
static void cv_adjust_order(cvs* m, int deltaq) {
    if ((m->q == 2) && (deltaq != 1)) return;
    if (deltaq == 1) cv_increase_bdf(m); else if (deltaq == -1) cv_decrease_bdf(m);
}

/* cvAdjustParams, SUN/cvode/cvode.c:2265-2274 */
static void cv_adjust_params(cvs* m) {
    if (m->qprime != m->q) {
        cv_adjust_order(m, m->qprime - m->q);
        m->q = m->qprime; m->L = m->q + 1; m->qwait = m->L;
    }
    cv_rescale(m);
}

/* SUN/cvode/cvode.c:2485-2505 */
static void cv_predict(cvs* m) {
    m->tn += m->h;
    for (int k = 1; k <= m->q; ++k)
        for (int j = m->q; j >= k; --j) m->zn[j - 1] = m->zn[j - 1] + m->zn[j];
}

/* SUN/cvode/cvode.c:3008-3017 */
static void cv_restore(cvs* m, double saved_t) {
    m->tn = saved_t;
    for (int k = 1; k <= m->q; ++k)
        for (int j = m->q; j >= k; --j) m->zn[j - 1] = m->zn[j - 1] - m->zn[j];
}

/* cvSet + cvSetBDF + cvSetTqBDF, SUN/cvode/cvode.c:2526-2540,2691-2766 */
static void cv_set(cvs* m) {
    double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
    const int q = m->q;
    m->l[0] = m->l[1] = xi_inv = xistar_inv = 1.0;
    for (int i = 2; i <= q; ++i) m->l[i] = 0.0;
    alpha0 = alpha0_hat = -1.0;
    hsum = m->h;
    if (q > 1) {
        for (int j = 2; j < q; ++j) {
            hsum += m->tau[j - 1];
            xi_inv = m->h / hsum;
            alpha0 -= 1.0 / j;
            for (int i = j; i >= 1; --i) m->l[i] += m->l[i - 1] * xi_inv;
        }
        alpha0 -= 1.0 / q;
        xistar_inv = -m->l[1] - alpha0;
        hsum += m->tau[q - 1];
        xi_inv = m->h / hsum;
        alpha0_hat = -m->l[1] - xi_inv;
        for (int i = q; i >= 1; --i) m->l[i] += m->l[i - 1] * xistar_inv;
    }
    {   /* cvSetTqBDF */
        double A1, A2, A3, A4, A5, A6, C, Cpinv, Cppinv;
        A1 = 1.0 - alpha0_hat + alpha0;
        A2 = 1.0 + q * A1;
        m->tq[2] = fabs(A1 / (alpha0 * A2));
        m->tq[5] = fabs(A2 * xistar_inv / (m->l[q] * xi_inv));
        if (m->qwait == 1) {
            if (q > 1) {
                C = xistar_inv / m->l[q];
                A3 = alpha0 + 1.0 / q;
                A4 = alpha0_hat + xi_inv;
                Cpinv = (1.0 - A4 + A3) / A3;
                m->tq[1] = fabs(C * Cpinv);
            } else m->tq[1] = 1.0;
            hsum += m->tau[q];
            xi_inv = m->h / hsum;
            A5 = alpha0 - (1.0 / (q + 1));
            A6 = alpha0_hat - xi_inv;
            Cppinv = (1.0 - A6 + A5) / A2;
            m->tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
        }
        m->tq[4] = CORTES / m->tq[2];
    }
    m->rl1 = 1.0 / m->l[1];
    m->gamma = m->h * m->rl1;
    if (m->nst == 0) m->gammap = m->gamma;
    m->gamrat = (m->nst > 0) ? m->gamma / m->gammap : 1.0;
}

/* CVDiagSetup, SUN/cvode/cvode_diag.c:341-418. ypred = m->y, fpred = m->ftemp. */
static int cv_diag_setup(cvs* m) {
    const double r = FRACT * m->rl1;
    double ft = nv_linsum(m->h, m->ftemp, -1.0, m->zn[1]);
    double yy = nv_linsum(r, ft, 1.0, m->y);
    double M = cv_f(m, m->tn, &yy); m->nfeDI++;
    M = nv_linsum(1.0, M, -1.0, m->ftemp);
    M = nv_linsum(FRACT, ft, -m->h, M);
    yy = ft * m->ewt;
    const double bit = (fabs(yy) >= m->uround) ? 1.0 : 0.0;
    const double bitcomp = bit + (-1.0);
    yy = ft * bit;
    yy = nv_linsum(FRACT, yy, -1.0, bitcomp);
    M = M / yy;
    M = M * bit;
    M = nv_linsum(1.0, M, -1.0, bitcomp);
    if (M == 0.0) { m->M = M; return 1; }   /* N_VInvTest */
    m->M = 1.0 / M;
    m->jcur = 1;
    m->gammasv = m->gamma;
    return 0;
}

/* CVDiagSolve, SUN/cvode/cvode_diag.c:429-468 */
static int cv_diag_solve(cvs* m, double* b) {
    if (m->gammasv != m->gamma) {
        const double r = m->gamma / m->gammasv;
        double M = 1.0 / m->M;
        M = M + (-1.0);
        M = nv_scale(r, M);
        M = M + 1.0;
        if (M == 0.0) { m->M = M; return 1; }
        m->M = 1.0 / M;
        m->gammasv = m->gamma;
    }
    *b = (*b) * m->M;
    return 0;
}

/* cvNlsResidual, SUN/cvode/cvode_nls.c:352-387 */
static void cv_nls_residual(cvs* m, double ycor, double* res) {
    m->y = m->zn[0] + ycor;
    m->ftemp = cv_f(m, m->tn, &m->y); m->nfe++;
    double rr = nv_linsum(m->rl1, m->zn[1], 1.0, ycor);
    rr = nv_linsum(-m->gamma, m->ftemp, 1.0, rr);
    *res = rr;
}

/* cvNlsLSetup, SUN/cvode/cvode_nls.c:251-282 */
static int cv_nls_lsetup(cvs* m) {
    const int retval = cv_diag_setup(m);
    m->nsetups++;
    m->nls_jcur = m->jcur;
    m->gamrat = 1.0; m->gammap = m->gamma; m->crate = 1.0; m->nstlp = m->nst;
    if (retval > 0) return NLS_CONV_RECVR;
    return CV_SUCCESS;
}

/* cvNlsConvTest, SUN/cvode/cvode_nls.c:307-349 */
static int cv_nls_convtest(cvs* m, int curiter, double ycor, double delta, double tol) {
    const double del = nv_wrms(delta, m->ewt);
    if (curiter > 0) m->crate = dmax(CRDOWN * m->crate, del / m->delp);
    const double dcon = del * dmin(1.0, m->crate) / tol;
    if (dcon <= 1.0) { m->acnrm = (curiter == 0) ? del : nv_wrms(ycor, m->ewt); return CV_SUCCESS; }
    if ((curiter >= 1) && (del > RDIV * m->delp)) return NLS_CONV_RECVR;
    m->delp = del;
    return NLS_CONTINUE;
}

/* SUNNonlinSolSolve_Newton, SUN/sunnonlinsol/newton/sunnonlinsol_newton.c:187-337 */
static int cv_newton(cvs* m, int callLSetup, long* niters, long* nconvfails) {
    int retval, curiter;
    double delta;
    *niters = 0; *nconvfails = 0;
    for (;;) {
        cv_nls_residual(m, m->acor, &delta);
        if (callLSetup) {
            retval = cv_nls_lsetup(m);
            if (retval != CV_SUCCESS) break;   /* :268 leaves the setup loop: no retry */
        }
        curiter = 0;
        for (;;) {
            (*niters)++;
            delta = -delta;
            retval = cv_diag_solve(m, &delta);
            if (retval != 0) { retval = NLS_CONV_RECVR; break; }   /* cvNlsLSolve: >0 -> SUN_NLS_CONV_RECVR */
            m->acor = m->acor + delta;
            retval = cv_nls_convtest(m, curiter, m->acor, delta, m->tq[4]);
            if (retval == CV_SUCCESS) { m->nls_jcur = 0; return CV_SUCCESS; }
            if (retval != NLS_CONTINUE) break;
            curiter++;
            if (curiter >= NLS_MAXCOR) { retval = NLS_CONV_RECVR; break; }
            cv_nls_residual(m, m->acor, &delta);
        }
        if ((retval > 0) && !m->nls_jcur) {
            (*nconvfails)++;
            callLSetup = 1;
            m->acor = 0.0;
            continue;
        }
        break;
    }
    (*nconvfails)++;
    return retval;
}

/* cvCheckConstraints with constraints = 2 ("y > 0"), SUN/cvode/cvode.c:2862-2921 */
static int cv_check_constraints(cvs* m) {
    const double cons = 2.0;
    const int violated = (m->y * cons <= 0.0);   /* N_VConstrMask, |c| > 1.5 branch */
    if (!violated) return CV_SUCCESS;
    const double mm = 1.0;
    double tmp = (fabs(cons) >= 1.5) ? 1.0 : 0.0;
    tmp = tmp * cons;
    tmp = tmp / m->ewt;
    tmp = nv_linsum(1.0, m->y, -0.1, tmp);
    tmp = tmp * mm;
    const double vnorm = nv_wrms(tmp, m->ewt);
    if (vnorm <= m->tq[4]) { m->acor = m->acor - tmp; return CV_SUCCESS; }
    if (fabs(m->h) <= m->hmin * ONEPSM) return CV_CONSTR_FAIL;
    tmp = m->zn[0] - m->y;
    tmp = mm * tmp;
    const double minq = (tmp == 0.0) ? DBL_MAX : m->zn[0] / tmp;   /* N_VMinQuotient, BIG_REAL when no denominators */
    m->eta = 0.9 * minq;
    m->eta = dmax(m->eta, 0.1);
    m->eta = dmax(m->eta, m->hmin / fabs(m->h));
    return CONSTR_RECVR;
}

/* cvNls, SUN/cvode/cvode.c:2781-2844 */
static int cv_nls(cvs* m, int nflag) {
    const int callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (m->nst == 0) ||
                          (m->nst >= m->nstlp + MSBP) || (fabs(m->gamrat - 1.0) > DGMAX);
    long ni = 0, nf = 0;
    m->acor = 0.0;
    int flag = cv_newton(m, callSetup, &ni, &nf);
    m->nni += ni;
    m->nnf += nf;   /* CVodeGetNumNonlinSolvConvFails reports this one (SUN/cvode/cvode_io.c) */
    if (flag != CV_SUCCESS) return flag;
    m->y = m->zn[0] + m->acor;
    m->jcur = 0;
    if (m->use_constraint) flag = cv_check_constraints(m);
    return flag;
}

/* cvHandleNFlag, SUN/cvode/cvode.c:2954-2998 */
static int cv_handle_nflag(cvs* m, int* nflag, double saved_t, int* ncf) {
    const int nf = *nflag;
    if (nf == CV_SUCCESS) return DO_ERROR_TEST;
    m->ncfn++;
    cv_restore(m, saved_t);
    if (nf < 0) return nf;
    (*ncf)++;
    m->etamax = 1.0;
    if ((fabs(m->h) <= m->hmin * ONEPSM) || (*ncf == MXNCF)) {
        if (nf == NLS_CONV_RECVR) return CV_CONV_FAILURE;
        if (nf == CONSTR_RECVR) return CV_CONSTR_FAIL;
    }
    if (nf != CONSTR_RECVR) m->eta = dmax(ETA_CF, m->hmin / fabs(m->h));
    *nflag = PREV_CONV_FAIL;
    cv_rescale(m);
    return PREDICT_AGAIN;
}

/* cvDoErrorTest, SUN/cvode/cvode.c:3048-3142 */
static int cv_do_error_test(cvs* m, int* nflag, double saved_t, int* nef, double* dsmp) {
    const double dsm = m->acnrm * m->tq[2];
    *dsmp = dsm;
    if (dsm <= 1.0) return CV_SUCCESS;
    (*nef)++; m->netf++;
    *nflag = PREV_ERR_FAIL;
    cv_restore(m, saved_t);
    if ((fabs(m->h) <= m->hmin * ONEPSM) || (*nef == MXNEF)) return CV_ERR_FAILURE;
    m->etamax = 1.0;
    if (*nef <= MXNEF1) {
        m->eta = 1.0 / (sun_powr(BIAS2 * dsm, 1.0 / m->L) + ADDON);
        m->eta = dmax(ETA_MIN_EF, dmax(m->eta, m->hmin / fabs(m->h)));
        if (*nef >= SMALL_NEF) m->eta = dmin(m->eta, ETA_MAX_EF);
        cv_rescale(m);
        return TRY_AGAIN;
    }
    if (m->q > 1) {
        m->eta = dmax(ETA_MIN_EF, m->hmin / fabs(m->h));
        cv_adjust_order(m, -1);
        m->L = m->q; m->q--; m->qwait = m->L;
        cv_rescale(m);
        return TRY_AGAIN;
    }
    m->eta = dmax(ETA_MIN_EF, m->hmin / fabs(m->h));
    m->h *= m->eta;
    m->next_h = m->h;
    m->hscale = m->h;
    m->qwait = LONG_WAIT;
    m->tempv = cv_f(m, m->tn, &m->zn[0]); m->nfe++;
    m->zn[1] = nv_scale(m->h, m->tempv);
    return TRY_AGAIN;
}

/* cvCompleteStep, SUN/cvode/cvode.c:3162-3207 */
static void cv_complete_step(cvs* m) {
    m->nst++;
    m->hu = m->h; m->qu = m->q;
    for (int i = m->q; i >= 2; --i) m->tau[i] = m->tau[i - 1];
    if ((m->q == 1) && (m->nst > 1)) m->tau[2] = m->tau[1];
    m->tau[1] = m->h;
    for (int j = 0; j <= m->q; ++j) m->zn[j] = nv_linsum(m->l[j], m->acor, 1.0, m->zn[j]);
    m->qwait--;
    if ((m->qwait == 1) && (m->q != QMAX)) {
        m->zn[QMAX] = m->acor;
        m->saved_tq5 = m->tq[5];
        m->indx_acor = QMAX;
    }
}

/* cvSetEta, SUN/cvode/cvode.c:3261-3290 */
static void cv_set_eta(cvs* m) {
    if ((m->eta > ETA_MIN_FX) && (m->eta < ETA_MAX_FX)) {
        m->eta = 1.0;
        m->hprime = m->h;
    } else {
        if (m->eta >= ETA_MAX_FX) {
            m->eta = dmin(m->eta, m->etamax);
            m->eta /= dmax(1.0, fabs(m->h) * m->hmax_inv * m->eta);
        } else {
            m->eta = dmax(m->eta, ETA_MIN);
            m->eta = dmax(m->eta, m->hmin / fabs(m->h));
        }
        m->hprime = m->h * m->eta;
    }
}

/* cvPrepareNextStep (+cvComputeEtaqm1/qp1, cvChooseEta), SUN/cvode/cvode.c:3218-3388 */
static void cv_prepare_next_step(cvs* m, double dsm) {
    if (m->etamax == 1.0) {
        m->qwait = (m->qwait > 2) ? m->qwait : 2;
        m->qprime = m->q;
        m->hprime = m->h;
        m->eta = 1.0;
        return;
    }
    m->etaq = 1.0 / (sun_powr(BIAS2 * dsm, 1.0 / m->L) + ADDON);
    if (m->qwait != 0) {
        m->eta = m->etaq;
        m->qprime = m->q;
        cv_set_eta(m);
        return;
    }
    m->qwait = 2;
    m->etaqm1 = 0.0;
    if (m->q > 1) {
        const double ddn = nv_wrms(m->zn[m->q], m->ewt) * m->tq[1];
        m->etaqm1 = 1.0 / (sun_powr(BIAS1 * ddn, 1.0 / m->q) + ADDON);
    }
    m->etaqp1 = 0.0;
    if (m->q != QMAX) {
        if (m->saved_tq5 != 0.0) {
            const double cquot = (m->tq[5] / m->saved_tq5) * sun_powi(m->h / m->tau[2], m->L);
            m->tempv = nv_linsum(-cquot, m->zn[QMAX], 1.0, m->acor);
            const double dup = nv_wrms(m->tempv, m->ewt) * m->tq[3];
            m->etaqp1 = 1.0 / (sun_powr(BIAS3 * dup, 1.0 / (m->L + 1)) + ADDON);
        }
    }
    {   /* cvChooseEta */
        const double etam = dmax(m->etaqm1, dmax(m->etaq, m->etaqp1));
        if ((etam > ETA_MIN_FX) && (etam < ETA_MAX_FX)) {
            m->eta = 1.0; m->qprime = m->q;
        } else if (etam == m->etaq) {
            m->eta = m->etaq; m->qprime = m->q;
        } else if (etam == m->etaqm1) {
            m->eta = m->etaqm1; m->qprime = m->q - 1;
        } else {
            m->eta = m->etaqp1; m->qprime = m->q + 1;
            m->zn[QMAX] = m->acor;
        }
    }
    cv_set_eta(m);
}

/* cvStep, SUN/cvode/cvode.c:2143-2247 */
static int cv_step(cvs* m) {
    int ncf = 0, nef = 0, nflag, kflag, eflag;
    double dsm = 0.0;
    if ((m->nst > 0) && (m->hprime != m->h)) cv_adjust_params(m);
    const double saved_t = m->tn;
    nflag = FIRST_CALL;
    for (;;) {
        m->attempts++;
        cv_predict(m);
        cv_set(m);
        nflag = cv_nls(m, nflag);
        kflag = cv_handle_nflag(m, &nflag, saved_t, &ncf);
        if (kflag == PREDICT_AGAIN) continue;
        if (kflag != DO_ERROR_TEST) return kflag;
        eflag = cv_do_error_test(m, &nflag, saved_t, &nef, &dsm);
        if (eflag == TRY_AGAIN) continue;
        if (eflag != CV_SUCCESS) return eflag;
        break;
    }
    cv_complete_step(m);
    cv_prepare_next_step(m, dsm);
    m->etamax = (m->nst <= SMALL_NST) ? ETA_MAX_ES : ETA_MAX_GS;
    m->acor = nv_scale(m->tq[2], m->acor);
    return CV_SUCCESS;
}

/* CVodeCreate + CVodeInit + CVodeSVtolerances + CVDiag + CVodeSetMaxNumSteps [+SetMaxStep, SetConstraints] + CVode(CV_NORMAL)
 * for one component: SUN/cvode/cvode.c:255-548,716-775,990-1466,1490-1560; call sequence HC/integrate_state_vec_3d.cpp:251-284. */
static int cvode_scalar(cellctx* c, double y0, double abstol, double reltol, double tout, long mxstep, double hmax_inv,
                        int use_constraint, double* yout, hco_cellstats* st, double* ele) {
    cvs mem; cvs* m = &mem;
    memset(m, 0, sizeof *m);
    m->c = c; m->uround = DBL_EPSILON; m->reltol = reltol; m->Vabstol = abstol; m->atolmin0 = (abstol == 0.0);
    m->mxstep = mxstep; m->hmax_inv = hmax_inv; m->use_constraint = use_constraint;
    m->zn[0] = y0; m->q = 1; m->L = 2; m->qwait = 2; m->etamax = ETA_MAX_FS;
    m->tn = 0.0;
    int istate = CV_SUCCESS;
    *yout = y0;

    /* first-call block :1058-1157 */
    if (use_constraint && (y0 * 2.0 <= 0.0)) { istate = CV_ILL_INPUT; goto done; }   /* cvInitialSetup :1839-1846 */
    if (cv_ewt(m, m->zn[0], &m->ewt) != 0) { istate = CV_ILL_INPUT; goto done; }
    m->zn[1] = cv_f(m, m->tn, &m->zn[0]); m->nfe++;
    {
        const int hflag = cv_hin(m, tout);
        if (hflag != CV_SUCCESS) { istate = hflag; goto done; }
    }
    {
        const double rh = fabs(m->h) * m->hmax_inv;
        if (rh > 1.0) m->h /= rh;
    }
    m->hscale = m->h; m->h0u = m->h; m->hprime = m->h;
    m->zn[1] = nv_scale(m->h, m->zn[1]);

    /* internal step loop :1300-1461 */
    for (long nstloc = 0;;) {
        m->next_h = m->h; m->next_q = m->q;
        if (m->nst > 0) {
            if (cv_ewt(m, m->zn[0], &m->ewt) != 0) { istate = CV_ILL_INPUT; *yout = m->zn[0]; break; }
        }
        if ((m->mxstep > 0) && (nstloc >= m->mxstep)) { istate = CV_TOO_MUCH_WORK; *yout = m->zn[0]; break; }
        const double nrm = nv_wrms(m->zn[0], m->ewt);
        if (m->uround * nrm > 1.0) { istate = CV_TOO_MUCH_ACC; *yout = m->zn[0]; break; }
        const int kflag = cv_step(m);
        if (kflag != CV_SUCCESS) { istate = kflag; *yout = m->zn[0]; break; }
        nstloc++;
        if ((m->tn - tout) * m->h >= 0.0) {
            /* CVodeGetDky(tout, 0): N_VLinearCombination fallback, j descending */
            const double s = (tout - m->tn) / m->h;
            double acc = 0.0;
            for (int j = m->q; j >= 0; --j) {
                double cj = 1.0;
                for (int i = 0; i < j; ++i) cj *= s;
                if (j == m->q) acc = nv_scale(cj, m->zn[j]);
                else acc = nv_linsum(cj, m->zn[j], 1.0, acc);
            }
            *yout = acc;
            istate = CV_SUCCESS;
            break;
        }
    }
done:
    if (ele) *ele = m->acor;   /* CVodeGetEstLocalErrors: N_VScale(ONE, cv_acor, ele), SUN/cvode/cvode_io.c:1346-1360 */
    if (st) {
        st->nst = m->nst; st->netf = m->netf; st->nfe = m->nfe; st->nni = m->nni; st->ncfn = m->nnf; st->nsetups = m->nsetups;
        st->nfeLS = m->nfeDI; st->flag = istate; st->ne_iters = c->ne_iters; st->attempts = m->attempts;
    }
    return istate;
}

/* ------------------------------------------------------------------ drivers */
static inline double* at(const hco_fab* f, int i, int j, int k, int n) {
    return f->p + (i - f->lo[0]) + (long)(j - f->lo[1]) * f->jstride + (long)(k - f->lo[2]) * f->kstride + (long)n * f->nstride;
}
enum { DENS = 0, EDEN = 4, EINT = 5, TEMP = 0, NE = 1, ZHI = 2 };

/* A10: integrate_state_vec_mfin (HC/integrate_state_vec_3d.cpp:72-365) with one CVODE instance per cell,
 * A9: ode_eos_finalize (HC/f_rhs.H:30-109) */
int hco_integrate_state_vec(const hco_rates* r, const hco_params* p, const hco_fab* state, const hco_fab* diag,
                            const int lo[3], const int hi[3], double a, double dt, hco_cellstats* stats) {
    const double hmax_inv = p->use_typical_steps ? 1.0 / (dt / (p->old_max_steps)) : 0.0;   /* CVodeSetMaxStep: hmax_inv = 1/hmax */
    long idx = 0;
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i, ++idx) {
        cellctx c; memset(&c, 0, sizeof c);
        c.r = r; c.is_struct = 0; c.h_species = p->h_species;
        const double rho = *at(state, i, j, k, DENS);
        const double e0 = *at(state, i, j, k, EINT) / rho;
        c.rpar[0] = *at(diag, i, j, k, TEMP); c.rpar[1] = *at(diag, i, j, k, NE); c.rpar[2] = rho; c.rpar[3] = 1 / a - 1;
        const double abstol = nv_scale(p->atol_factor, e0);   /* N_VScale(abstol,u,abstol_vec) :254 */
        double e_out;
        cvode_scalar(&c, e0, abstol, p->rtol, dt, p->max_steps, hmax_inv, p->use_constraint, &e_out, stats ? &stats[idx] : NULL, NULL);
        /* ode_eos_finalize */
        double T_vode = c.rpar[0], ne_vode = c.rpar[1];
        const double rho_vode = c.rpar[2], z_vode = c.rpar[3];
        const double gamma_minus_1 = 2.0 / 3.0;
        const double aa = 1 / (z_vode + 1.0);
        if (e_out < 0.e0) {
            const double YHELIUM = (1.0 - p->h_species) / (4.0 * p->h_species);
            T_vode = 10.0; ne_vode = 0.0;
            const double mu = (1.0e0 + 4.0e0 * YHELIUM) / (1.0e0 + YHELIUM + ne_vode);
            e_out = T_vode / (gamma_minus_1 * MP_OVER_KB * mu);
        }
        double sp[5];
        eos_T_given_Re(r, gamma_minus_1, p->h_species, 1, 1, &T_vode, &ne_vode, rho_vode, e_out, aa, sp);
        *at(diag, i, j, k, TEMP) = T_vode;
        *at(diag, i, j, k, NE) = ne_vode;
        *at(state, i, j, k, EINT) += *at(state, i, j, k, DENS) * (e_out - e0);
        *at(state, i, j, k, EDEN) += *at(state, i, j, k, DENS) * (e_out - e0);
    }
    return 0;
}

/* A14: integrate_state_struct_mfin (HC/integrate_state_with_source_3d.cpp:187-709) per cell;
 * A11: ode_eos_setup / initialize_single / initialize_arrays (HC/f_rhs_struct.H:45-210);
 * A13: ode_eos_finalize_struct (HC/f_rhs_struct.H:273-446) */
int hco_integrate_state_struct(const hco_rates* r, const hco_params* p, const hco_fab* s_old, const hco_fab* s_new,
                               const hco_fab* diag, const hco_fab* hydro_src, const hco_fab* reset_src, const hco_fab* ir,
                               const int lo[3], const int hi[3], double a, double a_end, double dt, int sdc_iter,
                               hco_cellstats* stats) {
    return hco_integrate_state_struct_react(r, p, s_old, s_new, diag, hydro_src, reset_src, ir, NULL, NULL, NULL, lo, hi, a, a_end, dt, sdc_iter, stats);
}

/* the SAVE_REACT build of the same function (HC/integrate_state_with_source_3d.cpp:126-183,602-631): react_in / react_out / react_out_work
 * (7, 7, 9 components; all three NULL: the plain build) as ode_eos_save_react_arrays fills them (HC/f_rhs_struct.H:213-267) */
int hco_integrate_state_struct_react(const hco_rates* r, const hco_params* p, const hco_fab* s_old, const hco_fab* s_new,
                                     const hco_fab* diag, const hco_fab* hydro_src, const hco_fab* reset_src, const hco_fab* ir,
                                     const hco_fab* react_in, const hco_fab* react_out, const hco_fab* react_out_work,
                                     const int lo[3], const int hi[3], double a, double a_end, double dt, int sdc_iter,
                                     hco_cellstats* stats) {
    /* ode_eos_setup :73-93 */
    int flash_h, flash_he;
    if (p->zhi_flash > 0.0) flash_h = (p->inhomo_reion > 0) ? 0 : 1; else flash_h = 0;
    flash_he = (p->zheii_flash > 0.0) ? 1 : 0;
    /* ode_eos_initialize_single :128-142 */
    const double z = 1 / a - 1.0;
    const int JH0 = ((flash_h == 1) && (z > p->zhi_flash)) ? 0 : 1;
    const int JHe0 = ((flash_he == 1) && (z > p->zheii_flash)) ? 0 : 1;
    const double H_reion_z = p->zhi_flash, He_reion_z = p->zheii_flash;
    const double hmax_inv = p->use_typical_steps ? 1.0 / (dt / (p->old_max_steps)) : 0.0;
    const double asq = a * a, aendsq = a_end * a_end;
    const int has_src = (sdc_iter >= 0);
    long idx = 0;
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i, ++idx) {
        cellctx c; memset(&c, 0, sizeof c);
        c.r = r; c.is_struct = 1; c.h_species = p->h_species; c.gamma_minus_1 = p->gamma_minus_1;
        c.a = a; c.uvb_density_A = p->uvb_density_A; c.uvb_density_B = p->uvb_density_B; c.JH = JH0; c.JHe = JHe0; c.has_src = has_src;
        /* ode_eos_initialize_arrays :180-209 */
        const double rho0 = *at(s_old, i, j, k, DENS);
        const double rhoe0 = *at(s_old, i, j, k, EINT);
        const double e0 = rhoe0 / rho0;
        c.T = *at(diag, i, j, k, TEMP); c.ne = *at(diag, i, j, k, NE); c.rho = rho0;
        if (has_src) {
            c.rho_src = *at(hydro_src, i, j, k, DENS) / dt;
            c.rhoe_src = *at(hydro_src, i, j, k, EINT) / dt;
            c.e_src = (((asq * rhoe0 + dt * c.rhoe_src) / aendsq + *at(reset_src, i, j, k, 0)) / (rho0 + dt * c.rho_src) - e0) / dt;
            c.rho_init = rho0;
        }
        double zhi = 0.0;
        if (p->inhomo_reion) { zhi = *at(diag, i, j, k, ZHI); c.JH = (z > zhi) ? 0 : 1; }
        const double abstol = nv_scale(p->atol_factor, e0);
        double e_out, ele = 0.0;
        hco_cellstats st_cell; memset(&st_cell, 0, sizeof st_cell);
        cvode_scalar(&c, e0, abstol, p->rtol, dt, p->max_steps, hmax_inv, p->use_constraint, &e_out, &st_cell, &ele);
        if (stats) stats[idx] = st_cell;
        const double e_cvode = e_out;   /* dptr[idx]: the finalize step works on a local copy (:283) */

        /* ode_eos_finalize_struct */
        const double e_orig = e0;
        const double rho_out = *at(s_new, i, j, k, DENS);
        *at(diag, i, j, k, TEMP) = c.T;    /* :290-291 values of the LAST RHS evaluation */
        *at(diag, i, j, k, NE) = c.ne;
        double ahalf = 0.0, rho_orig = 0.0, IR = 0.0, mu;
        const double rsrc = has_src ? *at(reset_src, i, j, k, 0) : 0.0;
        if (has_src) {
            ahalf = 0.5 * (a + a_end);
            rho_orig = c.rho_init;
            IR = (aendsq * rho_out * e_out - ((asq * rho_orig * e_orig + dt * c.rhoe_src))) / (dt * ahalf) - aendsq * rsrc / (dt * ahalf);
        }
        double Tv = c.T, nev = c.ne;   /* f_rhs_data->T_vode[idx], ne_vode[idx] */
        const double YHELIUM = (1.0 - p->h_species) / (4.0 * p->h_species);
        if (has_src) {
            if ((*at(s_new, i, j, k, EINT) + dt * ahalf * IR / aendsq) / *at(s_new, i, j, k, DENS) < 0.e0) {
                Tv = 10.0; nev = 0.0;
                mu = (1.0e0 + 4.0e0 * YHELIUM) / (1.0e0 + YHELIUM + nev);
                e_out = Tv / (p->gamma_minus_1 * MP_OVER_KB * mu);
                IR = (aendsq * rho_out * e_out - ((asq * rho_orig * e_orig + dt * c.rhoe_src))) / (dt * ahalf) - aendsq * rsrc / (dt * ahalf);
            }
        } else if (e_out < 0.e0) {
            Tv = 10.0; nev = 0.0;
            mu = (1.0e0 + 4.0e0 * YHELIUM) / (1.0e0 + YHELIUM + nev);
            e_out = Tv / (p->gamma_minus_1 * MP_OVER_KB * mu);
        }
        /* :343-348. NOTE the reference passes (nh0, nhep, nhp, nhe0, nhepp) to a callee declared
         * (nh0, nhp, nhe0, nhep, nhepp): its local "nhp" receives the callee's nhe0. Reproduced. */
        double sp[5];
        eos_T_given_Re(r, p->gamma_minus_1, p->h_species, c.JH, c.JHe, &Tv, &nev, c.rho, e_out, a, sp);
        const double caller_nhp = sp[2], caller_nhepp = sp[4];
        const double z_end = 1 / (a_end) - 1.0;
        double T_H = 0.0e0, T_He = 0.0e0;
        if (p->inhomo_reion) {
            if ((zhi < z) && (zhi >= z_end)) T_H = (1.0e0 - caller_nhp) * dmax_amrex(p->T_zhi - Tv);
        } else if (flash_h) {
            if ((H_reion_z < z) && (H_reion_z >= z_end)) T_H = (1.0e0 - caller_nhp) * dmax_amrex(p->T_zhi - Tv);
        }
        if (flash_he) {
            if ((He_reion_z < z) && (He_reion_z >= z_end)) T_He = (1.0e0 - caller_nhepp) * dmax_amrex(p->T_zheii - Tv);
        }
        if ((T_H > 0.0e0) || (T_He > 0.0e0)) {
            Tv = Tv + T_H + T_He;
            nev = 1.0e0 + YHELIUM;
            if (T_He > 0.0e0) nev = nev + YHELIUM;
            mu = (1.0e0 + 4.0e0 * YHELIUM) / (1.0e0 + YHELIUM + nev);
            e_out = Tv / (p->gamma_minus_1 * MP_OVER_KB * mu);
            if (has_src) {
                IR = (aendsq * rho_out * e_out - ((asq * rho_orig * e_orig + dt * c.rhoe_src))) / (dt * ahalf) - aendsq * rsrc / (dt * ahalf);
                if ((*at(s_new, i, j, k, EINT) + dt * ahalf * IR / aendsq) / *at(s_new, i, j, k, DENS) < 0.e0) {
                    Tv = 10.0; nev = 0.0;
                    mu = (1.0e0 + 4.0e0 * YHELIUM) / (1.0e0 + YHELIUM + nev);
                    e_out = Tv / (p->gamma_minus_1 * MP_OVER_KB * mu);
                    IR = (aendsq * rho_out * e_out - ((asq * rho_orig * e_orig + dt * c.rhoe_src))) / (dt * ahalf) - aendsq * rsrc / (dt * ahalf);
                }
            } else if (e_out < 0.e0) {
                Tv = 10.0; nev = 0.0;
                mu = (1.0e0 + 4.0e0 * YHELIUM) / (1.0e0 + YHELIUM + nev);
                e_out = Tv / (p->gamma_minus_1 * MP_OVER_KB * mu);
            }
            eos_T_given_Re(r, p->gamma_minus_1, p->h_species, c.JH, c.JHe, &Tv, &nev, c.rho, e_out, a, sp);
        }
        if (react_in && react_out && react_out_work) {
            /* ode_eos_save_react_arrays :241-266; T_vode / ne_vode are the values the finalize step left (its last EOS solve).  nje and ncfl
             * are never assigned by GetFinalStats (HC/integrate_state_with_source_3d.cpp:792-810): uninitialised in the reference, 0 here.
             * The counters are those of this cell's own CVODE instance (the reference: of the tile-wide instance, in every cell). */
            *at(react_in, i, j, k, 0) = e0;
            *at(react_out, i, j, k, 0) = e_cvode;
            *at(react_in, i, j, k, 1) = c.rho_init;
            *at(react_out, i, j, k, 1) = c.rho_init + dt * c.rho_src;
            *at(react_in, i, j, k, 2) = c.rhoe_src;
            *at(react_in, i, j, k, 3) = c.e_src;
            *at(react_out, i, j, k, 2) = Tv;
            *at(react_out, i, j, k, 3) = nev;
            *at(react_in, i, j, k, 4) = abstol;
            *at(react_out, i, j, k, 4) = ele;
            *at(react_in, i, j, k, 5) = a;
            *at(react_out, i, j, k, 5) = a_end;
            *at(react_in, i, j, k, 6) = 0.0;
            *at(react_out, i, j, k, 6) = dt;
            *at(react_out_work, i, j, k, 0) = (double)st_cell.nst; *at(react_out_work, i, j, k, 1) = (double)st_cell.netf;
            *at(react_out_work, i, j, k, 2) = (double)st_cell.nfe; *at(react_out_work, i, j, k, 3) = (double)st_cell.nni;
            *at(react_out_work, i, j, k, 4) = (double)st_cell.ncfn; *at(react_out_work, i, j, k, 5) = (double)st_cell.nsetups;
            *at(react_out_work, i, j, k, 6) = 0.0; *at(react_out_work, i, j, k, 7) = 0.0;
            *at(react_out_work, i, j, k, 8) = (double)st_cell.nfeLS;
        }
        if (has_src) {
            *at(ir, i, j, k, 0) = IR;
            *at(s_new, i, j, k, EINT) = *at(s_new, i, j, k, EINT) + dt * ahalf * IR / aendsq;
            *at(s_new, i, j, k, EDEN) = *at(s_new, i, j, k, EDEN) + dt * ahalf * IR / aendsq;
        } else {
            *at(s_old, i, j, k, EINT) += *at(s_old, i, j, k, DENS) * (e_out - e_orig);
            *at(s_old, i, j, k, EDEN) += *at(s_old, i, j, k, DENS) * (e_out - e_orig);
        }
    }
    return 0;
}

/* compute_new_temp core (DRV/Nyx.cpp:2473-2490): T, Ne from (rho, e); JH = JHe = 1 */
int hco_eos_box(const hco_rates* r, const hco_params* p, const hco_fab* state, const hco_fab* diag, const int lo[3], const int hi[3], double a) {
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i) {
        const double rho = *at(state, i, j, k, DENS);
        const double e = *at(state, i, j, k, EINT) / rho;
        double sp[5];
        eos_T_given_Re(r, p->gamma_minus_1, p->h_species, 1, 1, at(diag, i, j, k, TEMP), at(diag, i, j, k, NE), rho, e, a, sp);
    }
    return 0;
}

/* nyx_eos_given_RT (EOS/eos_hc.H:222-231): e from (T, Ne) */
static double eos_e_given_T(double gamma_minus_1, double h_species, double T, double Ne) {
    const double YHELIUM = (1.0 - h_species) / (4.0 * h_species);
    const double mu = (1.0 + 4.0 * YHELIUM) / (1.0 + YHELIUM + Ne);
    return T / (gamma_minus_1 * MP_OVER_KB * mu);
}

/* the cell loop of Nyx::compute_new_temp (DRV/Nyx.cpp:2473-2519) */
int hco_compute_new_temp_box(const hco_rates* r, const hco_params* p, const hco_fab* state, const hco_fab* diag, const int lo[3], const int hi[3],
                             double a, double small_temp, double large_temp, int max_temp_dt) {
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i) {
        const double rho = *at(state, i, j, k, DENS);
        const double rhoInv = 1.0 / rho;
        double eint;
        if (*at(state, i, j, k, EINT) > 0.0) {
            double sp[5];
            eos_T_given_Re(r, p->gamma_minus_1, p->h_species, 1, 1, at(diag, i, j, k, TEMP), at(diag, i, j, k, NE), rho,
                           *at(state, i, j, k, EINT) * (1.0 / rho), a, sp);
            if (*at(diag, i, j, k, TEMP) >= large_temp && max_temp_dt == 1) {
                *at(diag, i, j, k, TEMP) = large_temp;
                eint = eos_e_given_T(p->gamma_minus_1, p->h_species, *at(diag, i, j, k, TEMP), *at(diag, i, j, k, NE));
                const double ke = 0.5e0 * (*at(state, i, j, k, 1) * *at(state, i, j, k, 1) + *at(state, i, j, k, 2) * *at(state, i, j, k, 2) +
                                           *at(state, i, j, k, 3) * *at(state, i, j, k, 3)) * rhoInv;
                *at(state, i, j, k, EINT) = rho * eint;
                *at(state, i, j, k, EDEN) = *at(state, i, j, k, EINT) + ke;
            }
        } else {
            eint = eos_e_given_T(p->gamma_minus_1, p->h_species, small_temp, *at(diag, i, j, k, NE));
            const double ke = 0.5e0 * (*at(state, i, j, k, 1) * *at(state, i, j, k, 1) + *at(state, i, j, k, 2) * *at(state, i, j, k, 2) +
                                       *at(state, i, j, k, 3) * *at(state, i, j, k, 3)) * rhoInv;
            *at(diag, i, j, k, TEMP) = small_temp;
            *at(state, i, j, k, EINT) = rho * eint;
            *at(state, i, j, k, EDEN) = *at(state, i, j, k, EINT) + ke;
        }
    }
    return 0;
}

/* reset_internal_e (EOS/reset_internal_e.H:16-68) over a box: the cell loop of Nyx::reset_internal_energy (DRV/Nyx.cpp:2356-2385) */
int hco_reset_internal_e_box(const hco_params* p, const hco_fab* u, const hco_fab* d, const hco_fab* rs, const int lo[3], const int hi[3],
                             double small_temp, int interp) {
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i) {
        const double rhoInv = 1.0 / *at(u, i, j, k, DENS);
        const double Up = *at(u, i, j, k, 1) * rhoInv, Vp = *at(u, i, j, k, 2) * rhoInv, Wp = *at(u, i, j, k, 3) * rhoInv;
        const double ke = 0.5 * *at(u, i, j, k, DENS) * (Up * Up + Vp * Vp + Wp * Wp);
        const double rho_eint = *at(u, i, j, k, EDEN) - ke;
        if (rho_eint > 0.0 && rho_eint / *at(u, i, j, k, EDEN) > 1.0e-6 && interp == 0) {
            *at(rs, i, j, k, 0) = rho_eint - *at(u, i, j, k, EINT);
            *at(u, i, j, k, EINT) = rho_eint;
        } else if (*at(u, i, j, k, EINT) > 0.0) {
            *at(rs, i, j, k, 0) += 0.0;
            *at(u, i, j, k, EDEN) = *at(u, i, j, k, EINT) + ke;
        } else if (*at(u, i, j, k, EINT) <= 0.0) {
            const double eint_new = eos_e_given_T(p->gamma_minus_1, p->h_species, small_temp, *at(d, i, j, k, NE));
            *at(rs, i, j, k, 0) = *at(u, i, j, k, DENS) * eint_new - *at(u, i, j, k, EINT);
            *at(u, i, j, k, EINT) = *at(u, i, j, k, DENS) * eint_new;
            *at(u, i, j, k, EDEN) = *at(u, i, j, k, EINT) + ke;
        }
    }
    return 0;
}

/* ---- SDC source assembly.
 * Nyx::update_state_with_sources (Source/TimeStep/Nyx_update_state_with_sources.cpp:9-121) is three sweeps: (1) the source update of
 * every state component (:33-76), (2) Nyx::enforce_minimum_density over the whole MultiFab (Nyx_enforce_minimum_density.cpp:8-65, floor
 * variant :67-107 with floor_density, Nyx_enforce_minimum_density.H:8-58) -- taken only when the minimum of the NEW density over all boxes
 * is below small_dens, and then it also resets hydro_src(rho) in EVERY cell -- and (3) the gravity update (:92-120).  The minimum is
 * global, so the port has one function per side of it. */

/* sweep (1) over one box; returns the minimum of the new density over the box */
double hco_sources_apply_box(const hco_fab* uin, const hco_fab* uout, const hco_fab* src, const hco_fab* hsrc, const int lo[3], const int hi[3],
                             double dt, double a_old, double a_new) {
    const double a_half = 0.5 * (a_old + a_new);
    const double a_half_inv = 1 / a_half;
    const double a_oldsq = a_old * a_old;
    const double a_newsq = a_new * a_new;
    const double a_new_inv = 1.0 / a_new;
    const double a_newsq_inv = 1.0 / a_newsq;
    double m = DBL_MAX;
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i) {
        for (int n = 0; n < uout->ncomp; ++n) {
            double* o = at(uout, i, j, k, n);
            if (n == DENS) {
                *o = *at(uin, i, j, k, n) + *at(hsrc, i, j, k, n) + dt * *at(src, i, j, k, n) * a_half_inv;
            } else if (n >= 1 && n <= 3) {
                *o = a_old * *at(uin, i, j, k, n) + *at(hsrc, i, j, k, n) + dt * *at(src, i, j, k, n);
                *o = *o * a_new_inv;
            } else if (n == EDEN || n == EINT) {
                *o = a_oldsq * *at(uin, i, j, k, n) + *at(hsrc, i, j, k, n) + a_half * dt * *at(src, i, j, k, n);
                *o = *o * a_newsq_inv;
            } else {
                *o = *at(uin, i, j, k, n) + *at(hsrc, i, j, k, n) + dt * *at(src, i, j, k, n) * a_half_inv;
            }
        }
        if (*at(uout, i, j, k, DENS) < m) m = *at(uout, i, j, k, DENS);
    }
    return m;
}

/* sweeps (2) -- only when enforce != 0, i.e. the caller found min over all boxes < small_dens -- and (3) over one box; sdc != 0: the
 * reference's SDC build, which resets hydro_src(rho) */
int hco_sources_finish_box(const hco_params* p, const hco_fab* uin, const hco_fab* uout, const hco_fab* hsrc, const hco_fab* grav,
                           const int lo[3], const int hi[3], double dt, double a_old, double a_new, double small_dens, double small_temp,
                           int enforce, int sdc) {
    const double a_half = 0.5 * (a_old + a_new);
    const double a_newsq = a_new * a_new;
    const double a_newsq_inv = 1.0 / a_newsq;
    const double dt_a_new = dt / a_new;
    if (enforce) {
        for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i) {
            if (*at(uout, i, j, k, DENS) < small_dens) {   /* floor_density */
                *at(uout, i, j, k, DENS) = small_dens;
                *at(uout, i, j, k, 1) = 0.0; *at(uout, i, j, k, 2) = 0.0; *at(uout, i, j, k, 3) = 0.0;
                const double eint_new = eos_e_given_T(p->gamma_minus_1, p->h_species, small_temp, 0.0);
                *at(uout, i, j, k, EINT) = *at(uout, i, j, k, DENS) * eint_new;
                *at(uout, i, j, k, EDEN) = *at(uout, i, j, k, EINT);
            }
        }
        if (sdc)
            for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
                *at(hsrc, i, j, k, DENS) = *at(uout, i, j, k, DENS) - *at(uin, i, j, k, DENS);
    }
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i) {
        const double rho = *at(uin, i, j, k, DENS);
        const double SrU = rho * *at(grav, i, j, k, 0);
        const double SrV = rho * *at(grav, i, j, k, 1);
        const double SrW = rho * *at(grav, i, j, k, 2);
        *at(uout, i, j, k, 1) += SrU * dt_a_new;
        *at(uout, i, j, k, 2) += SrV * dt_a_new;
        *at(uout, i, j, k, 3) += SrW * dt_a_new;
        const double SrE = *at(uin, i, j, k, 1) * *at(grav, i, j, k, 0) + *at(uin, i, j, k, 2) * *at(grav, i, j, k, 1) +
                           *at(uin, i, j, k, 3) * *at(grav, i, j, k, 2);
        *at(uout, i, j, k, EDEN) = (a_newsq * *at(uout, i, j, k, EDEN) + SrE * (dt * a_half)) * a_newsq_inv;
    }
    return 0;
}

/* the cell loop of Nyx::init_zhi (Source/Initialization/Nyx_initdata.cpp:198-209) over one box */
/* One iteration of Nyx::enforce_minimum_density_cons (TS/Nyx_enforce_minimum_density.cpp:179-236) on one box: sbord = the new state with TWO
 * filled ghost cells (the caller's FillPatch, :183), uout = S_new, rs = reset_e_src.  compute_mu_for_enforce_min on the box grown by one
 * (TS/Nyx_enforce_minimum_density.H:60-157), create_update_for_minimum on the box (:159-200), S_new += update (:226), reset_e_src = update(Eint)
 * (:229, SDC build).  Returns the new minimum density of the box; *bad counts the faces where the reference aborts ("mu < 0"). */
double hco_enforce_min_cons_iter_box(const hco_fab* sbord, const hco_fab* uout, const hco_fab* rs, const int lo[3], const int hi[3],
                                     double small_value, int sdc, long* bad) {
    const int nx = hi[0] - lo[0] + 5, ny = hi[1] - lo[1] + 5, nz = hi[2] - lo[2] + 5;   /* the box grown by two holds every face index touched */
    const size_t npts = (size_t)nx * ny * nz;
    double* mu = (double*)calloc(3 * npts, sizeof(double));   /* mu_x.setVal(0.) ... (:186-188) */
#define MU(d, i, j, k) mu[(size_t)(d) * npts + (size_t)((i) - lo[0] + 2) + (size_t)nx * ((size_t)((j) - lo[1] + 2) + (size_t)ny * (size_t)((k) - lo[2] + 2))]
#define ST(i, j, k) (*at(sbord, i, j, k, DENS))
    const double target = 1.01 * small_value;
    long nbad = 0;
    for (int k = lo[2] - 1; k <= hi[2] + 1; ++k) for (int j = lo[1] - 1; j <= hi[1] + 1; ++j) for (int i = lo[0] - 1; i <= hi[0] + 1; ++i) {
        if (ST(i, j, k) < small_value) {
            const double total_need = target - ST(i, j, k);
            const double avail_from_ihi = dmax((ST(i + 1, j, k) - target) / 6.0, 0.0);
            const double avail_from_ilo = dmax((ST(i - 1, j, k) - target) / 6.0, 0.0);
            const double avail_from_jhi = dmax((ST(i, j + 1, k) - target) / 6.0, 0.0);
            const double avail_from_jlo = dmax((ST(i, j - 1, k) - target) / 6.0, 0.0);
            const double avail_from_khi = dmax((ST(i, j, k + 1) - target) / 6.0, 0.0);
            const double avail_from_klo = dmax((ST(i, j, k - 1) - target) / 6.0, 0.0);
            const double total_avail = avail_from_ihi + avail_from_ilo + avail_from_jhi + avail_from_jlo + avail_from_khi + avail_from_klo;
            double fac;
            if (total_need < total_avail) fac = total_need / total_avail; else fac = 1.0;
            const double from_ihi = fac * avail_from_ihi, from_ilo = fac * avail_from_ilo;
            const double from_jhi = fac * avail_from_jhi, from_jlo = fac * avail_from_jlo;
            const double from_khi = fac * avail_from_khi, from_klo = fac * avail_from_klo;
            if (from_ihi > 0) { MU(0, i + 1, j, k) = from_ihi / (ST(i + 1, j, k) - ST(i, j, k)); if (MU(0, i + 1, j, k) < 0.) nbad++; }
            if (from_ilo > 0) { MU(0, i, j, k) = -from_ilo / (ST(i, j, k) - ST(i - 1, j, k)); if (MU(0, i, j, k) < 0.) nbad++; }
            if (from_jhi > 0) { MU(1, i, j + 1, k) = from_jhi / (ST(i, j + 1, k) - ST(i, j, k)); if (MU(1, i, j + 1, k) < 0.) nbad++; }
            if (from_jlo > 0) { MU(1, i, j, k) = -from_jlo / (ST(i, j, k) - ST(i, j - 1, k)); if (MU(1, i, j, k) < 0.) nbad++; }
            if (from_khi > 0) { MU(2, i, j, k + 1) = from_khi / (ST(i, j, k + 1) - ST(i, j, k)); if (MU(2, i, j, k + 1) < 0.) nbad++; }
            if (from_klo > 0) { MU(2, i, j, k) = -from_klo / (ST(i, j, k) - ST(i, j, k - 1)); if (MU(2, i, j, k) < 0.) nbad++; }
        }
    }
    double m = DBL_MAX;
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i) {
        for (int n = 0; n < 6; ++n) {
#define S(i, j, k) (*at(sbord, i, j, k, n))
            const double update = MU(0, i + 1, j, k) * (S(i + 1, j, k) - S(i, j, k))
                                 - MU(0, i, j, k) * (S(i, j, k) - S(i - 1, j, k))
                                 + MU(1, i, j + 1, k) * (S(i, j + 1, k) - S(i, j, k))
                                 - MU(1, i, j, k) * (S(i, j, k) - S(i, j - 1, k))
                                 + MU(2, i, j, k + 1) * (S(i, j, k + 1) - S(i, j, k))
                                 - MU(2, i, j, k) * (S(i, j, k) - S(i, j, k - 1));
#undef S
            *at(uout, i, j, k, n) += update;
            if (n == EINT && sdc) *at(rs, i, j, k, 0) = update;
        }
        if (*at(uout, i, j, k, DENS) < m) m = *at(uout, i, j, k, DENS);
    }
#undef MU
#undef ST
    free(mu);
    if (bad) *bad = nbad;
    return m;
}

int hco_init_zhi_box(const hco_fab* diag, const hco_fab* zhi, const int lo[3], const int hi[3], int ratio) {
    for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
        *at(diag, i, j, k, ZHI) = *at(zhi, i / ratio, j / ratio, k / ratio, 0);
    return 0;
}
