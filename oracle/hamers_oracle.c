/*
 * hamers_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 * See hamers_oracle.h for scope and parity status.  Citations: path:line under
 * /root/reference.  Arithmetic is written in the reference's association order;
 * compile WITHOUT -ffast-math and without FMA contraction (-ffp-contract=off).
 *
 * Organisation mirrors the reference: full-array passes over freshly allocated
 * temporaries (cell stage -> sensor -> per direction: projection, 6x characteristic
 * transform, WENO, back-projection, bounds flags, fallback, HLLC, HLLC-HLL, sensor
 * select -> face flux -> source), so that timing it is a fair stand-in for the
 * reference's CPU cost structure.
 */
#include "hamers_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPSILON 1.0e-15   /* HAMERS_EPSILON, include/HAMeRS_config.hpp.in:16 */
#define G ORC_GHOSTS

int orc_num_eqn(const orc_desc* d)
{
    if (d->model == ORC_FOUR_EQN_CONSERVATIVE) return d->dim + 1 + d->ns;
    return d->model == ORC_SINGLE_SPECIES ? d->dim + 2 : d->dim + 2 * d->ns;
}
int orc_num_comp(const orc_desc* d)
{
    if (d->model == ORC_FOUR_EQN_CONSERVATIVE) return d->dim + 1 + d->ns;
    return d->model == ORC_SINGLE_SPECIES ? d->dim + 2 : d->dim + 2 * d->ns + 1;
}
static int nz_of(const orc_desc* d) { return d->dim == 3 ? d->n[2] : 1; }
static int gz_of(const orc_desc* d) { return d->dim == 3 ? G : 0; }
long orc_cell_ghost_size(const orc_desc* d)
{
    return (long)(d->n[0] + 2 * G) * (d->n[1] + 2 * G) * (nz_of(d) + 2 * gz_of(d));
}
long orc_cell_size(const orc_desc* d) { return (long)d->n[0] * d->n[1] * nz_of(d); }
long orc_side_size(const orc_desc* d, int dir)
{
    long e[3] = {d->n[0], d->n[1], nz_of(d)};
    e[dir] += 1;
    return e[0] * e[1] * e[2];
}

/* ------------------------------------------------------------------------- */
/* Point kernels                                                              */
/* ------------------------------------------------------------------------- */

/* Ideal-gas equation of state, the expressions of EquationOfStateIdealGas.cpp:29-45 (= :5580 on cell data),
 * :561-577 (= :5909) and :1093-1108 (= :6238); pinned against the reference's scalar members (oracle/_ref). */
static inline double eos_pressure(double gamma, double rho, double epsilon) { return (gamma - 1.0) * rho * epsilon; }
static inline double eos_sound_speed(double gamma, double rho, double p) { return sqrt(gamma * p / rho); }
static inline double eos_internal_energy(double gamma, double rho, double p) { return p / ((gamma - 1.0) * rho); }

/* Point formulas of the sensor chain and of the flux reconstruction, used by the loops below and exported so that they
 * can be pinned against the reference's own statements (oracle/_ref compiles those verbatim):
 *   first derivative, order 2: DerivativeFirstOrder.cpp:229, 382, 601
 *   dilatation / vorticity magnitude: ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:1631, 1653-1657 (2D :713, :729)
 *   sensor value: :2098-2101
 *   6th-order midpoint-and-node flux: :2370-2375 */
static inline double derivative_2nd(double u_R, double u_L, double dx) { return (1.0 / 2.0 * (u_R - u_L)) / dx; }
static inline double dilatation_3d(double dudx, double dvdy, double dwdz) { return dudx + dvdy + dwdz; }
static inline double vorticity_mag_3d(double dudy, double dudz, double dvdx, double dvdz, double dwdx, double dwdy)
{
    const double omega_x = dwdy - dvdz;
    const double omega_y = dudz - dwdx;
    const double omega_z = dvdx - dudy;
    return sqrt(omega_x * omega_x + omega_y * omega_y + omega_z * omega_z);
}
static inline double sensor_value(double theta_L, double theta_R, double Omega_L, double Omega_R)
{
    const double theta_avg = 0.5 * (theta_L + theta_R);
    const double Omega_avg = 0.5 * (Omega_L + Omega_R);
    return -theta_avg / (fabs(theta_avg) + Omega_avg + EPSILON);
}
static inline double face_flux(double dt, double Fm_L, double Fm_0, double Fm_R, double Fn_L, double Fn_R)
{
    return dt * (1.0 / 30.0 * (Fm_R + Fm_L) - 3.0 / 10.0 * (Fn_R + Fn_L) + 23.0 / 15.0 * Fm_0);
}

/* More point formulas of the single-species 3-D x-direction path, used by the loops below and pinned the same way:
 *   velocity, specific internal energy: FlowModelSingleSpecies.cpp:2824-2826, 3049-3051
 *   face averages, characteristic projection and its inverse: FlowModelBasicUtilitiesSingleSpecies.cpp:5000-5001,
 *     6324-6329, 7373-7379
 *   RK update: Euler.cpp:1479 (alpha term), 1544-1548 (beta term) */
static inline double quotient(double a, double rho) { return a / rho; }
static inline double internal_energy_3d(double E, double rho, double u, double v, double w)
{
    return E / rho - 1.0 / 2.0 * (u * u + v * v + w * w);
}
static inline double face_average(double a_L, double a_R) { return 1.0 / 2.0 * (a_L + a_R); }
static inline double char_acoustic_minus(double rho_avg, double c_avg, double un, double p)
{
    return -1.0 / 2.0 * rho_avg * c_avg * un + 1.0 / 2.0 * p;
}
static inline double char_acoustic_plus(double rho_avg, double c_avg, double un, double p)
{
    return 1.0 / 2.0 * rho_avg * c_avg * un + 1.0 / 2.0 * p;
}
static inline double char_entropy(double c_avg, double rho, double p) { return rho - 1.0 / (c_avg * c_avg) * p; }
static inline double back_density(double c_avg, double W0, double W1, double WL)
{
    return 1.0 / (c_avg * c_avg) * W0 + W1 + 1.0 / (c_avg * c_avg) * WL;
}
static inline double back_normal_velocity(double rho_avg, double c_avg, double W0, double WL)
{
    return -1.0 / (rho_avg * c_avg) * W0 + 1.0 / (rho_avg * c_avg) * WL;
}
static inline double back_pressure(double W0, double WL) { return W0 + WL; }
static inline double rk_beta_term_3d(double beta, double FxR, double FxL, double FyT, double FyB, double FzF, double FzB,
                                     double dx0, double dx1, double dx2, double S)
{
    return beta * (-(FxR - FxL) / dx0 - (FyT - FyB) / dx1 - (FzF - FzB) / dx2 + S);
}

/* five-eqn helpers (each is the point formula of one reference statement; pinned through orc_path_points3).
 * EquationOfStateMixingRulesIdealGas.cpp:7544, 7565, 7586 (xi accumulation, gamma_m) */
static inline double fe_xi_accumulate(double xi, double Z, double gamma_species)
{
    const double one_over_denominator = 1.0 / (gamma_species - 1.0);
    return xi + Z * one_over_denominator;
}
static inline double fe_gamma_from_xi(double xi) { return 1.0 / xi + 1.0; }
/* EquationOfStateMixingRulesIdealGas.cpp:7827-7828 (last species from Z_last = 1 - sum Z_i), EquationOfStateIdealGas.cpp:6414 */
static inline double fe_xi_last(double xi, double Z_last, double gamma_last) { return xi + Z_last / (gamma_last - 1.0); }
static inline double fe_internal_energy_from_p(double gamma_m, double rho, double p) { return p / ((gamma_m - 1.0) * rho); }
/* EquationOfStateIdealGas.cpp:5756 */
static inline double fe_pressure(double gamma_m, double rho, double epsilon) { return (gamma_m - 1.0) * rho * epsilon; }
/* EquationOfStateIdealGas.cpp:8157 (Gruneisen), :8308 (Psi), FlowModelFiveEqnAllaire.cpp:4789, 4818 (c^2 terms) */
static inline double fe_gruneisen(double gamma_m) { return gamma_m - 1.0; }
static inline double fe_psi(double p, double rho) { return p / rho; }
static inline double fe_c2_first(double Gamma, double p, double rho) { return Gamma * p / rho; }
static inline double fe_c2_accumulate(double cc, double Y, double Psi) { return cc + Y * Psi; }
/* FlowModelBasicUtilitiesFiveEqnAllaire.cpp:8848-8850, 8913-8914, 8920-8921 (projection, x; y/z are permutations) */
static inline double fe_char_minus(double rho_avg, double c_avg, double un, double p)
{
    return un - 1.0 / (rho_avg * c_avg) * p;
}
static inline double fe_char_plus(double rho_avg, double c_avg, double un, double p)
{
    return un + 1.0 / (rho_avg * c_avg) * p;
}
static inline double fe_char_partial_density(double Zrho_avg, double rho_avg, double c_avg, double Zrho, double p)
{
    return Zrho - Zrho_avg / (rho_avg * c_avg * c_avg) * p;
}
/* :9700-9703, 9751-9752, 9758-9760 (back-projection) */
static inline double fe_back_partial_density(double Zrho_avg, double c_avg, double W0, double Wsi, double Wlast)
{
    return -1.0 / 2.0 * Zrho_avg / c_avg * W0 + Wsi + 1.0 / 2.0 * Zrho_avg / c_avg * Wlast;
}
static inline double fe_back_normal_velocity(double W0, double Wlast) { return 1.0 / 2.0 * W0 + 1.0 / 2.0 * Wlast; }
static inline double fe_back_pressure(double rho_avg, double c_avg, double W0, double Wlast)
{
    return -1.0 / 2.0 * rho_avg * c_avg * W0 + 1.0 / 2.0 * rho_avg * c_avg * Wlast;
}
/* ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2623-2641: one direction's term of the advective source */
static inline double adv_source_term(double um_R, double um_L, double u_R, double u_L, double um_RR, double um_LL, double dx)
{
    return (3.0 / 2.0 * (um_R - um_L) - 3.0 / 10.0 * (u_R - u_L) + 1.0 / 30.0 * (um_RR - um_LL)) / dx;
}

/* ---- four-eqn conservative model: mixture of ideal gases closed by MASS fractions ----
 * EquationOfStateMixingRulesIdealGas.cpp:108-119 (species c_p, c_v), :7000-7120 (array form of
 * computeMixtureThermodynamicPropertiesWithMassFractions: c_p += Y_i c_p_i, c_v += Y_i c_v_i, gamma = c_p/c_v),
 * EquationOfStateIdealGas.cpp:5756 (p), :6414 (epsilon from p), :8157 (Gamma = gamma - 1),
 * EquationOfStateMixingRulesIdealGas.cpp:6406-6409 (Psi_i = ((c_p_i - gamma c_v_i)/c_v + gamma - 1) epsilon),
 * FlowModelFourEqnConservative.cpp:5180-5430 (c = sqrt(Gamma p/rho + sum Y_i Psi_i)).  Pinned through orc_path_points7. */
static inline double fc_species_c_p(double gamma, double R) { return gamma / (gamma - 1.0) * R; }
static inline double fc_species_c_v(double gamma, double R) { return 1.0 / (gamma - 1.0) * R; }
static inline double fc_accumulate(double acc, double Y, double c_species) { return acc + Y * c_species; }
static inline double fc_gamma(double c_p, double c_v) { return c_p / c_v; }
static inline double fc_psi(double c_p_i, double c_v_i, double gamma, double c_v, double epsilon)
{
    return ((c_p_i - gamma * c_v_i) / c_v + gamma - 1.0) * epsilon;
}
/* gamma[0..ns) species gammas, gamma[ns..2ns) species R (see orc_riemann_point) */
static inline void fc_mixture(int ns, const double* gamma, const double* R, const double* Y, double* c_p_o, double* c_v_o)
{
    double c_p = 0.0, c_v = 0.0;
    for (int si = 0; si < ns; si++) {
        c_p = fc_accumulate(c_p, Y[si], fc_species_c_p(gamma[si], R[si]));
        c_v = fc_accumulate(c_v, Y[si], fc_species_c_v(gamma[si], R[si]));
    }
    *c_p_o = c_p;
    *c_v_o = c_v;
}
/* sound speed of a state (rho, Y, p): the reference recomputes epsilon from p for Psi */
static inline double fc_sound_speed(int ns, const double* gamma, const double* R, double rho, const double* Y, double p,
                                    double* eps_o)
{
    double c_p, c_v;
    fc_mixture(ns, gamma, R, Y, &c_p, &c_v);
    const double gamma_m = fc_gamma(c_p, c_v);
    const double eps = fe_internal_energy_from_p(gamma_m, rho, p);
    double cc = fe_c2_first(fe_gruneisen(gamma_m), p, rho);
    for (int si = 0; si < ns; si++)
        cc = fe_c2_accumulate(cc, Y[si], fc_psi(fc_species_c_p(gamma[si], R[si]), fc_species_c_v(gamma[si], R[si]), gamma_m, c_v, eps));
    if (eps_o) *eps_o = eps;
    return sqrt(cc);
}

void orc_path_points3(const double in[56], double out[32])
{
    /* two species, 3-D, x direction.
     * in: Zrho0 Zrho1 rho_u rho_v rho_w E Z0 Z1 gamma0 gamma1 | Zrho0_L Zrho0_R Zrho1_L Zrho1_R rho_L rho_R c_L c_R |
     *     V0..V6 = (Zrho0, Zrho1, u, v, w, p, Z0) of a stencil cell | Wc0..Wc6 |
     *     S dt Q | um_R um_L um_RR um_LL u_R u_L (x) | same (y) | same (z) | dx0 dx1 dx2 */
    double rho = 0.0;
    rho += in[0];
    rho += in[1];
    out[0] = rho;
    out[1] = in[0] / rho;
    out[2] = in[1] / rho;
    out[3] = in[2] / rho;
    out[4] = in[3] / rho;
    out[5] = in[4] / rho;
    out[6] = internal_energy_3d(in[5], rho, out[3], out[4], out[5]);
    double xi = 0.0;
    xi = fe_xi_accumulate(xi, in[6], in[8]);
    xi = fe_xi_accumulate(xi, in[7], in[9]);
    const double gamma_m = fe_gamma_from_xi(xi);
    out[7] = gamma_m;
    const double p = fe_pressure(gamma_m, rho, out[6]);
    out[8] = p;
    double cc = fe_c2_first(fe_gruneisen(gamma_m), p, rho);
    cc = fe_c2_accumulate(cc, out[1], fe_psi(p, rho));
    cc = fe_c2_accumulate(cc, out[2], fe_psi(p, rho));
    out[9] = sqrt(cc);
    const double Zrho_avg[2] = {face_average(in[10], in[11]), face_average(in[12], in[13])};
    const double rho_avg = face_average(in[14], in[15]), c_avg = face_average(in[16], in[17]);
    out[10] = Zrho_avg[0];
    out[11] = Zrho_avg[1];
    out[12] = rho_avg;
    out[13] = c_avg;
    const double* V = in + 18;
    out[14] = fe_char_minus(rho_avg, c_avg, V[2], V[5]);
    out[15] = fe_char_partial_density(Zrho_avg[0], rho_avg, c_avg, V[0], V[5]);
    out[16] = fe_char_partial_density(Zrho_avg[1], rho_avg, c_avg, V[1], V[5]);
    out[17] = V[3];
    out[18] = V[4];
    out[19] = V[6];
    out[20] = fe_char_plus(rho_avg, c_avg, V[2], V[5]);
    const double* Wc = in + 25;
    out[21] = fe_back_partial_density(Zrho_avg[0], c_avg, Wc[0], Wc[1], Wc[6]);
    out[22] = fe_back_partial_density(Zrho_avg[1], c_avg, Wc[0], Wc[2], Wc[6]);
    out[23] = fe_back_normal_velocity(Wc[0], Wc[6]);
    out[24] = Wc[3];
    out[25] = Wc[4];
    out[26] = fe_back_pressure(rho_avg, c_avg, Wc[0], Wc[6]);
    out[27] = Wc[5];
    double S = in[32];
    double acc = adv_source_term(in[35], in[36], in[39], in[40], in[37], in[38], in[53]);
    acc = acc + adv_source_term(in[41], in[42], in[45], in[46], in[43], in[44], in[54]);
    acc = acc + adv_source_term(in[47], in[48], in[51], in[52], in[49], in[50], in[55]);
    S += in[33] * in[34] * acc;
    out[28] = S;
    out[29] = 0.0;
    out[30] = 0.0;
    out[31] = 0.0;
}

static inline void side_thermo(int model, int dim, int ns, const double* gamma, const double* V,
                               double* rho_o, double* c_o, double* eps_o);

void orc_path_points4(const double in[12], double out[4])
{
    /* five-eqn, two species, 3-D: in = V (Zrho0, Zrho1, u, v, w, p, Z0) of one interpolated side, gamma0, gamma1;
     * out = rho, c, epsilon as fed to the Riemann point kernels */
    side_thermo(ORC_FIVE_EQN_ALLAIRE, 3, 2, in + 7, in, &out[0], &out[1], &out[2]);
    out[3] = 0.0;
}

void orc_path_points2(const double in[32], double out[20])
{
    /* in: rho, rho_u, rho_v, rho_w, E | rho_L, rho_R, c_L, c_R | V0..V4 of a stencil cell | Wc0..Wc4 |
     *     Q0, alpha, Qint, beta, FxR, FxL, FyT, FyB, FzF, FzB, dx0, dx1, dx2, S */
    out[0] = quotient(in[1], in[0]);
    out[1] = quotient(in[2], in[0]);
    out[2] = quotient(in[3], in[0]);
    out[3] = internal_energy_3d(in[4], in[0], out[0], out[1], out[2]);
    const double rho_avg = face_average(in[5], in[6]), c_avg = face_average(in[7], in[8]);
    out[4] = rho_avg;
    out[5] = c_avg;
    const double* V = in + 9;
    out[6] = char_acoustic_minus(rho_avg, c_avg, V[1], V[4]);
    out[7] = char_entropy(c_avg, V[0], V[4]);
    out[8] = V[2];
    out[9] = V[3];
    out[10] = char_acoustic_plus(rho_avg, c_avg, V[1], V[4]);
    const double* Wc = in + 14;
    out[11] = back_density(c_avg, Wc[0], Wc[1], Wc[4]);
    out[12] = back_normal_velocity(rho_avg, c_avg, Wc[0], Wc[4]);
    out[13] = Wc[2];
    out[14] = Wc[3];
    out[15] = back_pressure(Wc[0], Wc[4]);
    double Q = in[19];
    Q += in[20] * in[21];
    out[16] = Q;
    Q += rk_beta_term_3d(in[22], in[23], in[24], in[25], in[26], in[27], in[28], in[29], in[30], in[31], in[18]);
    out[17] = Q;
    out[18] = 0.0;
    out[19] = 0.0;
}

void orc_path_points(const double in[16], double out[5])
{
    /* in: u_R, u_L, dx | dudx dudy dudz dvdx dvdy dvdz dwdx dwdy dwdz | (theta, Omega as computed here and scaled copies) |
     *     dt, Fm_L, Fm_0, Fm_R  (Fn_L = in[1], Fn_R = in[0]) */
    out[0] = derivative_2nd(in[0], in[1], in[2]);
    out[1] = dilatation_3d(in[3], in[7], in[11]);
    out[2] = vorticity_mag_3d(in[4], in[5], in[6], in[8], in[9], in[10]);
    out[3] = sensor_value(out[1], 0.75 * out[1] - in[3], out[2], 1.25 * out[2]);
    out[4] = face_flux(in[12], in[13], in[14], in[15], in[1], in[0]);
}

/* hard-switch constants: ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2123 (0.65);
 * FlowModelBasicUtilitiesFiveEqnAllaire.hpp:24-27 (bounds); include/HAMeRS_config.hpp.in:16 (epsilon) */
#define ORC_SENSOR_THRESHOLD 0.65
#define ORC_Y_BOUND_LO (-0.001)
#define ORC_Y_BOUND_UP 1.001
#define ORC_Z_BOUND_LO (-1000.0)
#define ORC_Z_BOUND_UP 1000.0

/* Bounds flag of one interpolated side (1 = bounded).
 * single-species: FlowModelBasicUtilitiesSingleSpecies.cpp:3311-3326 (rho > 0 && p > 0).
 * five-eqn: FlowModelBasicUtilitiesFiveEqnAllaire.cpp:6440-6710 (3-D x), 6750-7025 (y), 7060-7340 (z): volume fractions
 *   (the last one from 1 - sum) and mass fractions strictly inside their bounds, partial densities > 0, c^2 > 0 with the
 *   Gruneisen parameter from ALL ns volume fractions.  REFERENCE QUIRK, mirrored here: the species loop of the c^2 check
 *   accumulates Y_i Psi_i only in the x direction (:5819, 6091, 6678); in the y and z directions it re-assigns
 *   c_sq = Gamma p / rho (:6359, 6992, 7305), so the check there is Gamma p / rho > 0.  The two forms give the same flag
 *   whenever Gamma = 1/xi > 0, i.e. for every physical set of species gammas.  Pinned through orc_path_points5. */
static inline int side_bounded(int model, int dim, int ns, int dir, const double* gamma, const double* Vs)
{
    int ok = 1;
    if (model == ORC_SINGLE_SPECIES) {
        ok &= (Vs[0] > 0.0) ? 1 : 0;
        ok &= (Vs[dim + 1] > 0.0) ? 1 : 0;
        return ok;
    }
    if (model == ORC_FOUR_EQN_CONSERVATIVE) {
        /* FlowModelBasicUtilitiesFourEqnConservative.cpp:3737-4510: rho = sum rho Y_i; every Y_i strictly inside its
         * bounds (FlowModelBasicUtilitiesFourEqnConservative.hpp:24-25, the same numbers as the five-eqn model's); rho > 0; p > 0 */
        double rho = 0.0;
        for (int si = 0; si < ns; si++) rho += Vs[si];
        for (int si = 0; si < ns; si++) {
            const double Y = Vs[si] / rho;
            ok &= (Y > ORC_Y_BOUND_LO && Y < ORC_Y_BOUND_UP) ? 1 : 0;
        }
        ok &= (rho > 0.0) ? 1 : 0;
        ok &= (Vs[ns + dim] > 0.0) ? 1 : 0;
        return ok;
    }
    const double Z_lo = ORC_Z_BOUND_LO, Z_up = ORC_Z_BOUND_UP, Y_lo = ORC_Y_BOUND_LO, Y_up = ORC_Y_BOUND_UP;
    double Z[ORC_MAX_SPECIES];
    Z[ns - 1] = 1.0;
    for (int si = 0; si < ns - 1; si++) {
        Z[si] = Vs[ns + dim + 1 + si];
        Z[ns - 1] -= Z[si];
        ok &= (Z[si] > Z_lo && Z[si] < Z_up) ? 1 : 0;
    }
    ok &= (Z[ns - 1] > Z_lo && Z[ns - 1] < Z_up) ? 1 : 0;
    double rho = 0.0;
    for (int si = 0; si < ns; si++) rho += Vs[si];
    double Y[ORC_MAX_SPECIES];
    for (int si = 0; si < ns; si++) {
        Y[si] = Vs[si] / rho;
        ok &= (Y[si] > Y_lo && Y[si] < Y_up) ? 1 : 0;
    }
    for (int si = 0; si < ns; si++) ok &= (Vs[si] > 0.0) ? 1 : 0;
    const double pp = Vs[ns + dim];
    double xi = 0.0;
    for (int si = 0; si < ns; si++) xi = fe_xi_accumulate(xi, Z[si], gamma[si]);
    const double gamma_m = fe_gamma_from_xi(xi);
    double c_sq = fe_c2_first(fe_gruneisen(gamma_m), pp, rho);
    for (int si = 0; si < ns; si++)
        c_sq = (dir == 0) ? fe_c2_accumulate(c_sq, Y[si], fe_psi(pp, rho)) : fe_c2_first(fe_gruneisen(gamma_m), pp, rho);
    ok &= (c_sq > 0.0) ? 1 : 0;
    return ok;
}

void orc_path_points7(const double in[16], double out[15])
{
    /* four-eqn conservative, two species, 3-D (layout: hamers_oracle.h) */
    const int ns = 2, dim = 3;
    const double* gam = in + 6;     /* gamma0, gamma1, R0, R1 */
    double rho = 0.0;
    for (int si = 0; si < ns; si++) rho += in[si];
    double Y[2];
    for (int si = 0; si < ns; si++) Y[si] = in[si] / rho;
    double vel[3], ke = 0.0;
    for (int a = 0; a < dim; a++) {
        vel[a] = in[ns + a] / rho;
        ke = (a == 0) ? vel[a] * vel[a] : ke + vel[a] * vel[a];
    }
    const double epsilon = in[ns + dim] / rho - 1.0 / 2.0 * ke;
    double c_p, c_v;
    fc_mixture(ns, gam, gam + ns, Y, &c_p, &c_v);
    const double gamma_m = fc_gamma(c_p, c_v);
    const double p = fe_pressure(gamma_m, rho, epsilon);
    const double eps_p = fe_internal_energy_from_p(gamma_m, rho, p);
    out[0] = rho;
    out[1] = Y[0];
    out[2] = Y[1];
    out[3] = epsilon;
    out[4] = c_p;
    out[5] = c_v;
    out[6] = gamma_m;
    out[7] = p;
    for (int si = 0; si < ns; si++)
        out[8 + si] = fc_psi(fc_species_c_p(gam[si], gam[ns + si]), fc_species_c_v(gam[si], gam[ns + si]), gamma_m, c_v, eps_p);
    out[10] = fc_sound_speed(ns, gam, gam + ns, rho, Y, p, 0);
    side_thermo(ORC_FOUR_EQN_CONSERVATIVE, dim, ns, gam, in + 10, &out[11], &out[12], &out[13]);
    out[14] = (double)side_bounded(ORC_FOUR_EQN_CONSERVATIVE, dim, ns, 0, gam, in + 10);
}

void orc_path_points8(const double in[24], double out[16])
{
    /* four-eqn conservative, 3-D x: face averages, projection and back-projection -- the same helper functions as the
     * five-eqn model's (the reference's statements are the same on rho Y_i instead of Z_i rho_i); layout: build_ref.py */
    const double ry_avg[2] = {face_average(in[0], in[1]), face_average(in[2], in[3])};
    const double rho_avg = face_average(in[4], in[5]), c_avg = face_average(in[6], in[7]);
    out[0] = ry_avg[0];
    out[1] = ry_avg[1];
    out[2] = rho_avg;
    out[3] = c_avg;
    const double* V = in + 8;
    out[4] = fe_char_minus(rho_avg, c_avg, V[2], V[5]);
    for (int si = 0; si < 2; si++) out[5 + si] = fe_char_partial_density(ry_avg[si], rho_avg, c_avg, V[si], V[5]);
    out[7] = V[3];
    out[8] = V[4];
    out[9] = fe_char_plus(rho_avg, c_avg, V[2], V[5]);
    const double* W = in + 14;
    for (int si = 0; si < 2; si++) out[10 + si] = fe_back_partial_density(ry_avg[si], c_avg, W[0], W[1 + si], W[5]);
    out[12] = fe_back_normal_velocity(W[0], W[5]);
    out[13] = W[3];
    out[14] = W[4];
    out[15] = fe_back_pressure(rho_avg, c_avg, W[0], W[5]);
}

void orc_path_points5(const double in[16], double out[2])
{
    /* in: five-eqn side V[7] (two species, 3-D), gamma0, gamma1, direction (0 / 1 / 2) | single-species side V[5];
     * out: the two bounds flags */
    out[0] = (double)side_bounded(ORC_FIVE_EQN_ALLAIRE, 3, 2, (int)in[9], in + 7, in);
    out[1] = (double)side_bounded(ORC_SINGLE_SPECIES, 3, 1, 0, in + 7, in + 10);
}

void orc_constants(double out[7])
{
    out[0] = EPSILON;
    out[1] = ORC_SENSOR_THRESHOLD;
    out[2] = ORC_Y_BOUND_LO;
    out[3] = ORC_Y_BOUND_UP;
    out[4] = ORC_Z_BOUND_LO;
    out[5] = ORC_Z_BOUND_UP;
    out[6] = G;
}

void orc_eos_point(double gamma, double rho, double epsilon, double* p, double* c, double* eps_back)
{
    *p = eos_pressure(gamma, rho, epsilon);
    *c = eos_sound_speed(gamma, rho, *p);
    *eps_back = eos_internal_energy(gamma, rho, *p);
}

/* ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:9-18 */
static inline double ipow_(double base, int e)
{
    double r = base;
    for (int i = 1; i < e; i++) r *= base;
    return r;
}

/* One-sided WCNS5-JS midpoint interpolation from (a,b,c,d,e) towards the face between c and d.
 * beta: WCNS5-JS-HLLC-HLL.cpp:31-44 (mirror :58-71); weights :98-106; value :112-117. */
static inline double weno5js_side(double a, double b, double c, double d, double e, int p)
{
    const double beta_0 = 1.0 / 3.0 * (a * (4.0 * a - 19.0 * b + 11.0 * c) + b * (25.0 * b - 31.0 * c) + 10.0 * c * c);
    const double beta_1 = 1.0 / 3.0 * (b * (4.0 * b - 13.0 * c + 5.0 * d) + 13.0 * c * (c - d) + 4.0 * d * d);
    const double beta_2 = 1.0 / 3.0 * (c * (10.0 * c - 31.0 * d + 11.0 * e) + d * (25.0 * d - 19.0 * e) + 4.0 * e * e);

    double omega_0 = 1.0 / 16.0 / ipow_(beta_0 + EPSILON, p);
    double omega_1 = 5.0 / 8.0 / ipow_(beta_1 + EPSILON, p);
    double omega_2 = 5.0 / 16.0 / ipow_(beta_2 + EPSILON, p);

    const double omega_sum = omega_0 + omega_1 + omega_2;
    omega_0 = omega_0 / omega_sum;
    omega_1 = omega_1 / omega_sum;
    omega_2 = omega_2 / omega_sum;

    return 3.0 / 8.0 * omega_0 * a +
           (-10.0 / 8.0 * omega_0 - 1.0 / 8.0 * omega_1) * b +
           (15.0 / 8.0 * omega_0 + 6.0 / 8.0 * omega_1 + 3.0 / 8.0 * omega_2) * c +
           (3.0 / 8.0 * omega_1 + 6.0 / 8.0 * omega_2) * d -
           1.0 / 8.0 * omega_2 * e;
}

void orc_weno5js_point(const double U[6], int p, double* U_minus, double* U_plus)
{
    /* minus: cells 0..4 (:78-118); plus: mirror image, cells 5..1 (:124-164) */
    *U_minus = weno5js_side(U[0], U[1], U[2], U[3], U[4], p);
    *U_plus = weno5js_side(U[5], U[4], U[3], U[2], U[1], p);
}

/* The three smoothness indicators shared by WCNS5-JS / WCNS5-Z / WCNS6-LD (same polynomials in all three files:
 * WCNS5-Z-HLLC-HLL.cpp:31-44, WCNS6-LD-HLLC-HLL.cpp:51-63). */
static inline void beta_012(double a, double b, double c, double d, double e, double* b0, double* b1, double* b2)
{
    *b0 = 1.0 / 3.0 * (a * (4.0 * a - 19.0 * b + 11.0 * c) + b * (25.0 * b - 31.0 * c) + 10.0 * c * c);
    *b1 = 1.0 / 3.0 * (b * (4.0 * b - 13.0 * c + 5.0 * d) + 13.0 * c * (c - d) + 4.0 * d * d);
    *b2 = 1.0 / 3.0 * (c * (10.0 * c - 31.0 * d + 11.0 * e) + d * (25.0 * d - 19.0 * e) + 4.0 * e * e);
}

/* WCNS5-Z one-sided interpolation: weights WCNS5-Z-HLLC-HLL.cpp:96-107 (tau_5 = |beta_0 - beta_2|,
 * omega_k = d_k (1 + (tau_5/(beta_k + eps))^p)), value :113-118; plus side mirrored (:124-165). */
static inline double weno5z_side(double a, double b, double c, double d, double e, int p)
{
    double beta_0, beta_1, beta_2;
    beta_012(a, b, c, d, e, &beta_0, &beta_1, &beta_2);
    const double tau_5 = fabs(beta_0 - beta_2);
    double omega_0 = 1.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_0 + EPSILON), p));
    double omega_1 = 5.0 / 8.0 * (1.0 + ipow_(tau_5 / (beta_1 + EPSILON), p));
    double omega_2 = 5.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_2 + EPSILON), p));
    const double omega_sum = omega_0 + omega_1 + omega_2;
    omega_0 = omega_0 / omega_sum;
    omega_1 = omega_1 / omega_sum;
    omega_2 = omega_2 / omega_sum;
    return 3.0 / 8.0 * omega_0 * a +
           (-10.0 / 8.0 * omega_0 - 1.0 / 8.0 * omega_1) * b +
           (15.0 / 8.0 * omega_0 + 6.0 / 8.0 * omega_1 + 3.0 / 8.0 * omega_2) * c +
           (3.0 / 8.0 * omega_1 + 6.0 / 8.0 * omega_2) * d -
           1.0 / 8.0 * omega_2 * e;
}

void orc_weno5z_point(const double U[6], int p, double* U_minus, double* U_plus)
{
    *U_minus = weno5z_side(U[0], U[1], U[2], U[3], U[4], p);
    *U_plus = weno5z_side(U[5], U[4], U[3], U[2], U[1], p);
}

/* WCNS6-LD one-sided interpolation from the six cells (a..f), upwind cell c, face between c and d.
 * sigma: WCNS6-LD-HLLC-HLL.cpp:23-39 (computed from cells 1..4 = b..e of the UNMIRRORED stencil: the plus side calls
 * the same computeLocalSigma, :275); beta_0..2 :51-63, beta_3 :64-83; upwind weights :167-177; central weights
 * :183-197; blend where R_tau > alpha_tau :203-211; value :217-226.  `sigma` is passed in because it does not mirror. */
static inline double weno6ld_side(double a, double b, double c, double d, double e, double f, double sigma,
                                  int p, int q, double C, double alpha_tau)
{
    double beta_0, beta_1, beta_2;
    beta_012(a, b, c, d, e, &beta_0, &beta_1, &beta_2);
    const double beta_3 = 1.0 / 232243200.0 * (a * (525910327.0 * a - 4562164630.0 * b + 7799501420.0 * c -
        6610694540.0 * d + 2794296070.0 * e - 472758974.0 * f) + 5.0 * b *
        (2146987907.0 * b - 7722406988.0 * c + 6763559276.0 * d - 2926461814.0 * e + 503766638.0 * f) + 20.0 * c *
        (1833221603.0 * c - 3358664662.0 * d + 1495974539.0 * e - 263126407.0 * f) +
        20.0 * d * (1607794163.0 * d - 1486026707.0 * e + 268747951.0 * f) +
        5.0 * e * (1432381427.0 * e - 536951582.0 * f) +
        263126407.0 * f * f);

    double omega_upwind_0, omega_upwind_1, omega_upwind_2;
    const double tau_5 = fabs(beta_0 - beta_2);
    omega_upwind_0 = 1.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_0 + EPSILON), p));
    omega_upwind_1 = 5.0 / 8.0 * (1.0 + ipow_(tau_5 / (beta_1 + EPSILON), p));
    omega_upwind_2 = 5.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_2 + EPSILON), p));
    const double omega_upwind_sum = omega_upwind_0 + omega_upwind_1 + omega_upwind_2;
    omega_upwind_0 = omega_upwind_0 / omega_upwind_sum;
    omega_upwind_1 = omega_upwind_1 / omega_upwind_sum;
    omega_upwind_2 = omega_upwind_2 / omega_upwind_sum;

    double omega_0, omega_1, omega_2, omega_3;
    const double beta_avg = 1.0 / 8.0 * (beta_0 + beta_2 + 6.0 * beta_1);
    const double tau_6 = fabs(beta_3 - beta_avg);
    omega_0 = 1.0 / 32.0 * (C + ipow_(tau_6 / (beta_0 + EPSILON), q));
    omega_1 = 15.0 / 32.0 * (C + ipow_(tau_6 / (beta_1 + EPSILON), q));
    omega_2 = 15.0 / 32.0 * (C + ipow_(tau_6 / (beta_2 + EPSILON), q));
    omega_3 = 1.0 / 32.0 * (C + ipow_(tau_6 / (beta_3 + EPSILON), q));
    const double omega_sum = omega_0 + omega_1 + omega_2 + omega_3;
    omega_0 = omega_0 / omega_sum;
    omega_1 = omega_1 / omega_sum;
    omega_2 = omega_2 / omega_sum;
    omega_3 = omega_3 / omega_sum;

    const double R_tau = tau_6 / (beta_avg + EPSILON);
    if (R_tau > alpha_tau) {
        omega_0 = sigma * omega_upwind_0 + (1.0 - sigma) * omega_0;
        omega_1 = sigma * omega_upwind_1 + (1.0 - sigma) * omega_1;
        omega_2 = sigma * omega_upwind_2 + (1.0 - sigma) * omega_2;
        omega_3 = (1.0 - sigma) * omega_3;
    }

    return 3.0 / 8.0 * omega_0 * a +
           (-10.0 / 8.0 * omega_0 - 1.0 / 8.0 * omega_1) * b +
           (15.0 / 8.0 * omega_0 + 6.0 / 8.0 * omega_1 + 3.0 / 8.0 * omega_2) * c +
           (3.0 / 8.0 * omega_1 + 6.0 / 8.0 * omega_2 + 15.0 / 8.0 * omega_3) * d +
           (-1.0 / 8.0 * omega_2 - 10.0 / 8.0 * omega_3) * e +
           3.0 / 8.0 * omega_3 * f;
}

/* WCNS6-LD-HLLC-HLL.cpp:23-39 */
static inline double weno6ld_sigma(const double U[6])
{
    const double alpha_1 = U[2] - U[1];
    const double alpha_2 = U[3] - U[2];
    const double alpha_3 = U[4] - U[3];
    const double theta_1 = fabs(alpha_1 - alpha_2) / (fabs(alpha_1) + fabs(alpha_2) + EPSILON);
    const double theta_2 = fabs(alpha_2 - alpha_3) / (fabs(alpha_2) + fabs(alpha_3) + EPSILON);
    return fmax(theta_1, theta_2);
}

void orc_weno6ld_point(const double U[6], int p, int q, double C, double alpha_tau, double* U_minus, double* U_plus)
{
    const double sigma = weno6ld_sigma(U);
    *U_minus = weno6ld_side(U[0], U[1], U[2], U[3], U[4], U[5], sigma, p, q, C, alpha_tau);
    *U_plus = weno6ld_side(U[5], U[4], U[3], U[2], U[1], U[0], sigma, p, q, C, alpha_tau);
}

/* Side thermodynamics fed to the Riemann point kernels.
 * single-species: EquationOfStateIdealGas.cpp:5909 (c), :6238 (epsilon), called from
 *   FlowModelRiemannSolverSingleSpeciesHLLC.cpp:2991-3025.
 * five-eqn: FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp:5640-5970: rho = sum Z_rho, Y = Z_rho/rho,
 *   mixture gamma from the ns-1 interpolated volume fractions
 *   (EquationOfStateMixingRulesIdealGas.cpp:7599-7832, Z_last = 1 - sum Z_i, fill :5377-5576),
 *   Gamma = gamma-1 (EquationOfStateIdealGas.cpp:8157), Psi_i = p/rho
 *   (EquationOfStateMixingRulesIdealGas.cpp:6736), epsilon = p/((gamma-1)*rho) (EquationOfStateIdealGas.cpp:6414).
 *   Pinned through orc_path_points4. */
static inline void side_thermo(int model, int dim, int ns, const double* gamma, const double* V,
                               double* rho_o, double* c_o, double* eps_o)
{
    if (model == ORC_SINGLE_SPECIES) {
        const double rho = V[0];
        const double p = V[dim + 1];
        *rho_o = rho;
        *c_o = eos_sound_speed(gamma[0], rho, p);
        *eps_o = eos_internal_energy(gamma[0], rho, p);
    } else if (model == ORC_FOUR_EQN_CONSERVATIVE) {
        /* FlowModelRiemannSolverFourEqnConservativeHLLC-HLL.cpp:5150-5290: rho = sum rho Y_i, Y_i, Gamma, Psi_i, c, epsilon */
        double rho = 0.0;
        for (int si = 0; si < ns; si++) rho += V[si];
        const double p = V[ns + dim];
        double Y[ORC_MAX_SPECIES];
        for (int si = 0; si < ns; si++) Y[si] = V[si] / rho;
        *rho_o = rho;
        *c_o = fc_sound_speed(ns, gamma, gamma + ns, rho, Y, p, eps_o);
    } else {
        double rho = 0.0;
        for (int si = 0; si < ns; si++) rho += V[si];
        const double p = V[ns + dim];
        double Y[ORC_MAX_SPECIES];
        for (int si = 0; si < ns; si++) Y[si] = V[si] / rho;
        double xi = 0.0, Z_last = 1.0;
        for (int si = 0; si < ns - 1; si++) {
            xi = fe_xi_accumulate(xi, V[ns + dim + 1 + si], gamma[si]);
            Z_last -= V[ns + dim + 1 + si];
        }
        xi = fe_xi_last(xi, Z_last, gamma[ns - 1]);
        const double gamma_m = fe_gamma_from_xi(xi);
        double c = fe_c2_first(fe_gruneisen(gamma_m), p, rho);
        for (int si = 0; si < ns; si++) c = fe_c2_accumulate(c, Y[si], fe_psi(p, rho));
        *rho_o = rho;
        *c_o = sqrt(c);
        *eps_o = fe_internal_energy_from_p(gamma_m, rho, p);
    }
}

/* HLLC and HLLC-HLL point kernels, generic in direction `dir` and dimension.
 * single-species: FlowModelRiemannSolverSingleSpeciesHLLC.cpp:604-1074,
 *                 FlowModelRiemannSolverSingleSpeciesHLLC-HLL.cpp:889-1629;
 * five-eqn:       FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp:848-1591,
 *                 FlowModelRiemannSolverFiveEqnAllaireHLLC-HLL.cpp:1283-2440.
 * The HLLC part of the hybrid kernel is the same arithmetic as the pure HLLC kernel, so it
 * is evaluated once.  vel_mid: HLLC.cpp:3381-3388 / FiveEqnAllaireHLLC.cpp:6034-6038. */
static inline void riemann_kernel(int model, int dim, int ns, int dir,
                                  const double* V_L, const double* V_R,
                                  double rho_L, double rho_R, double c_L, double c_R,
                                  double eps_L, double eps_R,
                                  double* F_HLLC, double* F_HLLC_HLL, double* vel_mid)
{
    const int neq = (model == ORC_SINGLE_SPECIES) ? dim + 2 : (model == ORC_FOUR_EQN_CONSERVATIVE ? dim + 1 + ns : dim + 2 * ns);
    const int nm = (model == ORC_SINGLE_SPECIES) ? 1 : ns;     /* number of mass equations */
    const int nz = (model == ORC_FIVE_EQN_ALLAIRE) ? ns - 1 : 0; /* advected volume fractions */
    const int iv = nm;                                         /* first velocity index in V */
    const int ip = nm + dim;                                   /* pressure index in V / energy in Q */

    const double un_L = V_L[iv + dir];
    const double un_R = V_R[iv + dir];
    const double p_L = V_L[ip];
    const double p_R = V_R[ip];

    const double u_average = 1.0 / 2.0 * (un_L + un_R);
    const double c_average = 1.0 / 2.0 * (c_L + c_R);

    const double s_L = fmin(u_average - c_average, un_L - c_L);
    const double s_R = fmax(u_average + c_average, un_R + c_R);

    const double s_minus = fmin(0.0, s_L);
    const double s_plus = fmax(0.0, s_R);

    const double s_star = (p_R - p_L + rho_L * un_L * (s_L - un_L) - rho_R * un_R * (s_R - un_R)) /
                          (rho_L * (s_L - un_L) - rho_R * (s_R - un_R));

    double Q_L[ORC_MAX_EQ], Q_R[ORC_MAX_EQ], F_L[ORC_MAX_EQ], F_R[ORC_MAX_EQ];

    /* conservative states and physical fluxes of both sides */
    for (int side = 0; side < 2; side++) {
        const double* V = side == 0 ? V_L : V_R;
        double* Q = side == 0 ? Q_L : Q_R;
        double* F = side == 0 ? F_L : F_R;
        const double rho = side == 0 ? rho_L : rho_R;
        const double eps = side == 0 ? eps_L : eps_R;
        const double un = V[iv + dir];
        const double p = V[ip];

        double ke = V[iv] * V[iv];
        for (int a = 1; a < dim; a++) ke = ke + V[iv + a] * V[iv + a];

        if (model == ORC_SINGLE_SPECIES) {
            Q[0] = V[0];
            for (int a = 0; a < dim; a++) Q[1 + a] = V[0] * V[1 + a];
            Q[ip] = V[0] * (eps + 1.0 / 2.0 * ke);

            F[0] = Q[1 + dir];
            for (int a = 0; a < dim; a++)
                F[1 + a] = (a == dir) ? Q[1 + dir] * V[1 + a] + p : Q[1 + dir] * V[1 + a];
            F[ip] = un * (Q[ip] + p);
        } else {
            for (int si = 0; si < ns; si++) Q[si] = V[si];
            for (int a = 0; a < dim; a++) Q[iv + a] = rho * V[iv + a];
            Q[ip] = rho * (eps + 1.0 / 2.0 * ke);
            for (int si = 0; si < nz; si++) Q[ip + 1 + si] = V[ip + 1 + si];

            for (int si = 0; si < ns; si++) F[si] = un * V[si];
            for (int a = 0; a < dim; a++)
                F[iv + a] = (a == dir) ? un * Q[iv + a] + p : un * Q[iv + a];
            F[ip] = un * (Q[ip] + p);
            for (int si = 0; si < nz; si++) F[ip + 1 + si] = un * V[ip + 1 + si];
        }
    }

    /* HLLC */
    double Chi_star;
    {
        const int left = (s_star > 0.0);
        const double* V = left ? V_L : V_R;
        const double* Q = left ? Q_L : Q_R;
        const double* F = left ? F_L : F_R;
        const double rho = left ? rho_L : rho_R;
        const double s_K = left ? s_L : s_R;
        const double s_mp = left ? s_minus : s_plus;
        const double un = V[iv + dir];
        const double p = V[ip];

        Chi_star = (s_K - un) / (s_K - s_star);

        double Q_star[ORC_MAX_EQ];
        for (int si = 0; si < nm; si++) Q_star[si] = Chi_star * V[si];
        for (int a = 0; a < dim; a++)
            Q_star[iv + a] = (a == dir) ? Chi_star * rho * s_star : Chi_star * Q[iv + a];
        Q_star[ip] = Chi_star * (Q[ip] + (s_star - un) * (rho * s_star + p / (s_K - un)));
        for (int e = ip + 1; e < neq; e++) Q_star[e] = Chi_star * V[e];

        for (int e = 0; e < neq; e++) F_HLLC[e] = F[e] + s_mp * (Q_star[e] - Q[e]);

        if (vel_mid) *vel_mid = un + s_mp * (Chi_star - 1.0);
    }

    if (!F_HLLC_HLL) return;

    /* HLL for mass, tangential momenta and volume fractions; upwind overrides */
    double F_HLL[ORC_MAX_EQ];
    for (int e = 0; e < neq; e++) {
        const int is_normal_mom = (e == iv + dir);
        const int is_energy = (e == ip);
        if (is_normal_mom || is_energy) continue;
        F_HLL[e] = (s_R * F_L[e] - s_L * F_R[e] + s_R * s_L * (Q_R[e] - Q_L[e])) / (s_R - s_L);
        if (s_L > 0.0) F_HLL[e] = F_L[e];
        if (s_R < 0.0) F_HLL[e] = F_R[e];
    }

    /* weights for hybridisation */
    double diff[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < dim; a++) diff[a] = V_R[iv + a] - V_L[iv + a];
    double mag2 = diff[0] * diff[0];
    for (int a = 1; a < dim; a++) mag2 = mag2 + diff[a] * diff[a];
    const double vel_mag = sqrt(mag2);

    double alpha_1, alpha_2;
    if (vel_mag < EPSILON) {
        alpha_1 = 1.0;
        alpha_2 = 0.0;
    } else {
        alpha_1 = fabs(diff[dir]) / vel_mag;
        alpha_2 = sqrt(1.0 - alpha_1 * alpha_1);
    }
    const double beta_1 = 1.0 / 2.0 * (1.0 + alpha_1 / (alpha_1 + alpha_2));
    const double beta_2 = 1.0 - beta_1;

    for (int e = 0; e < neq; e++) {
        if (e == iv + dir || e == ip)
            F_HLLC_HLL[e] = F_HLLC[e];
        else
            F_HLLC_HLL[e] = beta_1 * F_HLLC[e] + beta_2 * F_HLL[e];
    }
}

void orc_riemann_point(int model, int dim, int ns, const double* gamma, int dir,
                       const double* V_L, const double* V_R,
                       double* F_HLLC, double* F_HLLC_HLL, double* vel_mid)
{
    double rho_L, rho_R, c_L, c_R, e_L, e_R;
    side_thermo(model, dim, ns, gamma, V_L, &rho_L, &c_L, &e_L);
    side_thermo(model, dim, ns, gamma, V_R, &rho_R, &c_R, &e_R);
    riemann_kernel(model, dim, ns, dir, V_L, V_R, rho_L, rho_R, c_L, c_R, e_L, e_R,
                   F_HLLC, F_HLLC_HLL, vel_mid);
}

/* Exposed so that tests can feed identical (rho, c, epsilon) to the reference's point kernels. */
void orc_side_thermo(int model, int dim, int ns, const double* gamma, const double* V,
                     double* rho, double* c, double* eps)
{
    side_thermo(model, dim, ns, gamma, V, rho, c, eps);
}

/* ------------------------------------------------------------------------- */
/* Patch-level passes                                                         */
/* ------------------------------------------------------------------------- */

typedef struct {
    int dim, neq, ncomp, ns, model, nm;
    int n[3], g[3], gd[3];          /* interior dims, ghost widths, ghost-box dims */
    long cs[3];                     /* cell strides in the ghost box */
    long ncell_g;
} geom_t;

static void make_geom(const orc_desc* d, geom_t* q)
{
    q->dim = d->dim;
    q->model = d->model;
    q->ns = d->model == ORC_SINGLE_SPECIES ? 1 : d->ns;
    q->nm = q->ns;
    q->neq = orc_num_eqn(d);
    q->ncomp = orc_num_comp(d);
    for (int a = 0; a < 3; a++) {
        q->n[a] = (a < d->dim) ? d->n[a] : 1;
        q->g[a] = (a < d->dim) ? G : 0;
        q->gd[a] = q->n[a] + 2 * q->g[a];
    }
    q->cs[0] = 1;
    q->cs[1] = q->gd[0];
    q->cs[2] = (long)q->gd[0] * q->gd[1];
    q->ncell_g = (long)q->gd[0] * q->gd[1] * q->gd[2];
}

/* linear index in the ghost box of logical cell (i,j,k), interior origin 0
 * (FlowModelSingleSpecies.cpp:2815-2817) */
static inline long cidx(const geom_t* q, int i, int j, int k)
{
    return (i + q->g[0]) + (long)(j + q->g[1]) * q->cs[1] + (long)(k + q->g[2]) * q->cs[2];
}

static double* dalloc(long n)
{
    double* p = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    return p;
}

/* Derived cell data on the whole ghost box.
 * single-species: FlowModelSingleSpecies.cpp:2824-2826 (velocity), :3049-3051 (epsilon),
 *   EquationOfStateIdealGas.cpp:5580 (p), :5909 (c).
 * five-eqn: EquationOfStateMixingRules.cpp:735-900 (rho), FlowModelFiveEqnAllaire.cpp:3965 (Y),
 *   :4428-4430 (epsilon), EquationOfStateMixingRulesIdealGas.cpp:7520-7587 (gamma_m from ALL ns
 *   stored volume fractions), EquationOfStateIdealGas.cpp:5756 (p), FlowModelFiveEqnAllaire.cpp:4779-4853 (c). */
/* species gammas followed by the species gas constants (used by the four-eqn conservative model only) */
static void model_constants(const orc_desc* d, double* gam)
{
    for (int si = 0; si < ORC_MAX_SPECIES; si++) {
        gam[si] = 0.0;
        gam[ORC_MAX_SPECIES + si] = 0.0;
    }
    const int ns = d->model == ORC_SINGLE_SPECIES ? 1 : d->ns;
    for (int si = 0; si < ns; si++) {
        gam[si] = d->gamma[si];
        gam[ns + si] = (d->model == ORC_FOUR_EQN_CONSERVATIVE) ? d->R[si] : 0.0;
    }
}

static void cell_stage(const geom_t* q, const double* gam, const double* const* Q,
                       double** vel, double* p, double* c, double* rho_m)
{
    const int dim = q->dim, ns = q->ns;
    for (long x = 0; x < q->ncell_g; x++) {
        if (q->model == ORC_SINGLE_SPECIES) {
            const double rho = Q[0][x];
            double ke = 0.0;
            for (int a = 0; a < dim; a++) {
                vel[a][x] = quotient(Q[1 + a][x], rho);
                ke = (a == 0) ? vel[a][x] * vel[a][x] : ke + vel[a][x] * vel[a][x];
            }
            /* = internal_energy_3d for dim == 3 (same operation order: E/rho - 1/2 ((u u + v v) + w w)) */
            const double epsilon = Q[dim + 1][x] / rho - 1.0 / 2.0 * ke;
            p[x] = eos_pressure(gam[0], rho, epsilon);
            c[x] = eos_sound_speed(gam[0], rho, p[x]);
        } else if (q->model == ORC_FOUR_EQN_CONSERVATIVE) {
            /* FlowModelFourEqnConservative.cpp:3877 (rho), :3914 (Y), :4569 (velocity), :4793 (epsilon), :5034 (p), :5087 (c) */
            double rho = 0.0;
            for (int si = 0; si < ns; si++) rho += Q[si][x];
            rho_m[x] = rho;
            double Y[ORC_MAX_SPECIES];
            for (int si = 0; si < ns; si++) Y[si] = Q[si][x] / rho;
            double ke = 0.0;
            for (int a = 0; a < dim; a++) {
                vel[a][x] = Q[ns + a][x] / rho;
                ke = (a == 0) ? vel[a][x] * vel[a][x] : ke + vel[a][x] * vel[a][x];
            }
            const double epsilon = Q[ns + dim][x] / rho - 1.0 / 2.0 * ke;
            double c_p, c_v;
            fc_mixture(ns, gam, gam + ns, Y, &c_p, &c_v);
            p[x] = fe_pressure(fc_gamma(c_p, c_v), rho, epsilon);
            c[x] = fc_sound_speed(ns, gam, gam + ns, rho, Y, p[x], 0);
        } else {
            double rho = 0.0;
            for (int si = 0; si < ns; si++) rho += Q[si][x];
            rho_m[x] = rho;
            double Y[ORC_MAX_SPECIES];
            for (int si = 0; si < ns; si++) Y[si] = Q[si][x] / rho;
            double ke = 0.0;
            for (int a = 0; a < dim; a++) {
                vel[a][x] = Q[ns + a][x] / rho;
                ke = (a == 0) ? vel[a][x] * vel[a][x] : ke + vel[a][x] * vel[a][x];
            }
            const double epsilon = Q[ns + dim][x] / rho - 1.0 / 2.0 * ke;
            double xi = 0.0;
            for (int si = 0; si < ns; si++) xi = fe_xi_accumulate(xi, Q[ns + dim + 1 + si][x], gam[si]);
            const double gamma_m = fe_gamma_from_xi(xi);
            p[x] = fe_pressure(gamma_m, rho, epsilon);
            double cc = fe_c2_first(fe_gruneisen(gamma_m), p[x], rho);
            for (int si = 0; si < ns; si++) cc = fe_c2_accumulate(cc, Y[si], fe_psi(p[x], rho));
            c[x] = sqrt(cc);
        }
    }
}

/* FlowModelSingleSpecies.cpp:4064, 4237, 4365 (max wave speed), Euler.cpp:846-848 (spectral radius); pinned through
 * orc_path_points6 */
static inline double max_wave_speed(double un, double c) { return fabs(un) + c; }
static inline double spectral_radius_of(double lambda_max, double dx) { return lambda_max / dx; }

void orc_path_points6(const double in[8], double out[6])
{
    /* in: u, v, w, c, dx0, dx1, dx2, running maximum of the sum; out: three spectral radii, their sum as accumulated,
     * the updated running maximum and the stable dt from it (Euler.cpp:846-861) */
    double sum = 0.0;
    for (int a = 0; a < 3; a++) {
        out[a] = spectral_radius_of(max_wave_speed(in[a], in[3]), in[4 + a]);
        sum = (a == 0) ? out[a] : sum + out[a];
    }
    out[3] = sum;
    out[4] = fmax(in[7], sum);
    out[5] = 1.0 / out[4];
}

/* Euler.cpp:760-893 (3D; 2D :627-745): spectral radii and stable dt of one patch, see hamers_oracle.h */
int orc_spectral_radii_and_dt(const orc_desc* d, const double* const* Q, int include_ghosts, double* out)
{
    geom_t q;
    make_geom(d, &q);
    const int dim = q.dim;
    double* vel[3] = {dalloc(q.ncell_g), dalloc(q.ncell_g), dalloc(q.ncell_g)};
    double *p = dalloc(q.ncell_g), *c = dalloc(q.ncell_g), *rho_m = dalloc(q.ncell_g);
    double gam[2 * ORC_MAX_SPECIES];
    model_constants(d, gam);
    cell_stage(&q, gam, Q, vel, p, c, rho_m);
    double sr[4] = {0.0, 0.0, 0.0, 0.0};
    const int g = include_ghosts ? G : 0;
    const int gz = (dim == 3) ? g : 0;
    for (int k = -gz; k < nz_of(d) + gz; k++)
        for (int j = -g; j < d->n[1] + g; j++)
            for (int i = -g; i < d->n[0] + g; i++) {
                const long x = cidx(&q, i, j, k);
                double sum = 0.0;
                for (int a = 0; a < dim; a++) {
                    const double lambda_max = max_wave_speed(vel[a][x], c[x]);
                    const double spectral_radius = spectral_radius_of(lambda_max, d->dx[a]);
                    sr[a] = fmax(sr[a], spectral_radius);
                    sum = (a == 0) ? spectral_radius : sum + spectral_radius;
                }
                sr[3] = fmax(sr[3], sum);
            }
    for (int a = 0; a < dim; a++) out[a] = sr[a];
    out[dim] = 1.0 / sr[3];
    for (int a = 0; a < 3; a++) free(vel[a]);
    free(p);
    free(c);
    free(rho_m);
    return 0;
}

/* node flux in direction dir at ghost-box cell x, equation e.
 * single-species: FlowModelSingleSpecies.cpp:3388-3391, 3470-3474, 3611-3614, 3693-3697, 3856-3860;
 * five-eqn: FlowModelFiveEqnAllaire.cpp:5206-5275, 5434-5482, 5559-5628, 5806-5875. */
static inline double node_flux(const geom_t* q, const double* const* Q, double* const* vel,
                               const double* p, int dir, int e, long x)
{
    const int dim = q->dim, ns = q->ns;
    const double un = vel[dir][x];
    if (q->model == ORC_SINGLE_SPECIES) {
        if (e == 0) return Q[1 + dir][x];
        if (e <= dim) return (e - 1 == dir) ? un * Q[e][x] + p[x] : un * Q[e][x];
        return un * (Q[dim + 1][x] + p[x]);
    } else {
        if (e < ns) return un * Q[e][x];
        if (e < ns + dim) return (e - ns == dir) ? un * Q[e][x] + p[x] : un * Q[e][x];
        if (e == ns + dim) return un * (Q[e][x] + p[x]);
        return un * Q[e][x];   /* volume fraction, e = ns+dim+1+si maps to stored Z_si */
    }
}

int orc_compute_flux_and_source(const orc_desc* d, const double* const* Q, double dt,
                                double* const* F, double* const* S,
                                double* const* F_mid_dbg, double* const* sensor_dbg)
{
    geom_t qq;
    make_geom(d, &qq);
    const geom_t* q = &qq;
    const int dim = q->dim, neq = q->neq, ns = q->ns, nm = q->nm;
    const int iv = nm, ip = nm + dim;
    const int p_exp = d->weno_p > 0 ? d->weno_p : 2;
    const int has_adv = (q->model == ORC_FIVE_EQN_ALLAIRE);
    const int multi = (q->model != ORC_SINGLE_SPECIES);       /* partial densities instead of one density */
    const int nz = has_adv ? ns - 1 : 0;                      /* advected volume fractions */
    double gam[2 * ORC_MAX_SPECIES];
    model_constants(d, gam);

    /* ---- step 1: derived cell data (WCNS56-HLLC-HLL.cpp:1392-1410) ---- */
    double* vel[3] = {0, 0, 0};
    for (int a = 0; a < dim; a++) vel[a] = dalloc(q->ncell_g);
    double* p = dalloc(q->ncell_g);
    double* c = dalloc(q->ncell_g);
    double* rho_m = multi ? dalloc(q->ncell_g) : 0;
    cell_stage(q, gam, Q, vel, p, c, rho_m);

    /* primitive variable pointers (FlowModelSingleSpecies / FiveEqnAllaire getCellDataOfPrimitiveVariables) */
    const double* V[ORC_MAX_EQ];
    if (q->model == ORC_SINGLE_SPECIES) {
        V[0] = Q[0];
        for (int a = 0; a < dim; a++) V[1 + a] = vel[a];
        V[dim + 1] = p;
    } else {
        for (int si = 0; si < ns; si++) V[si] = Q[si];
        for (int a = 0; a < dim; a++) V[ns + a] = vel[a];
        V[ns + dim] = p;
        for (int si = 0; si < nz; si++) V[ns + dim + 1 + si] = Q[ns + dim + 1 + si];
    }
    const double* rho_cell = multi ? rho_m : Q[0];

    /* ---- step 2: velocity gradient, dilatation, vorticity (:1523-1659; 2D :663-731) ----
     * arrays with 2 ghosts; derivative (1/2*(uR-uL))/dx, DerivativeFirstOrder.cpp:382,601 */
    int g2[3], d2[3];
    for (int a = 0; a < 3; a++) {
        g2[a] = (a < dim) ? 2 : 0;
        d2[a] = q->n[a] + 2 * g2[a];
    }
    const long n2 = (long)d2[0] * d2[1] * d2[2];
    double* grad[9];
    for (int m = 0; m < dim * dim; m++) grad[m] = dalloc(n2);
    double* theta = dalloc(n2);
    double* Omega = dalloc(n2);
#define IDX2(i, j, k) ((i + g2[0]) + (long)(j + g2[1]) * d2[0] + (long)(k + g2[2]) * d2[0] * d2[1])
    for (int a = 0; a < dim; a++) {
        for (int b = 0; b < dim; b++) {
            double* out = grad[dim * a + b];
            for (int k = -g2[2]; k < q->n[2] + g2[2]; k++)
                for (int j = -g2[1]; j < q->n[1] + g2[1]; j++)
                    for (int i = -g2[0]; i < q->n[0] + g2[0]; i++) {
                        const long x = cidx(q, i, j, k);
                        out[IDX2(i, j, k)] = derivative_2nd(vel[a][x + q->cs[b]], vel[a][x - q->cs[b]], d->dx[b]);
                    }
        }
    }
    for (long x = 0; x < n2; x++) {
        if (dim == 2) {
            theta[x] = grad[0][x] + grad[3][x];
            Omega[x] = fabs(grad[2][x] - grad[1][x]);
        } else {
            /* grad[3a + b] = d u_a / d x_b */
            theta[x] = dilatation_3d(grad[0][x], grad[4][x], grad[8][x]);
            Omega[x] = vorticity_mag_3d(grad[1][x], grad[2][x], grad[3][x], grad[5][x], grad[6][x], grad[7][x]);
        }
    }

    /* HLLC midpoint normal velocities of all directions are needed together by the source */
    double* vel_midpoint[3] = {0, 0, 0};
    double* F_midpoint_all[3][ORC_MAX_EQ];
    long sd_all[3][3];

    for (int dir = 0; dir < dim; dir++) {
        /* side arrays: faces -1..N+1 in the normal direction, interior tangentially */
        long sd[3] = {q->n[0], q->n[1], q->n[2]};
        sd[dir] += 3;
        for (int a = 0; a < 3; a++) sd_all[dir][a] = sd[a];
        const long nside = sd[0] * sd[1] * sd[2];
        const long st = q->cs[dir];
#define SIDX(i, j, k) ((i + (dir == 0)) + sd[0] * ((long)(j + (dir == 1)) + sd[1] * (long)(k + (dir == 2))))
        const int lo[3] = {dir == 0 ? -1 : 0, dir == 1 ? -1 : 0, dir == 2 ? -1 : 0};
        const int hi[3] = {q->n[0] + (dir == 0 ? 2 : 0), q->n[1] + (dir == 1 ? 2 : 0), q->n[2] + (dir == 2 ? 2 : 0)};
#define FOR_FACES                                        \
    for (int k = lo[2]; k < hi[2]; k++)                   \
        for (int j = lo[1]; j < hi[1]; j++)               \
            for (int i = lo[0]; i < hi[0]; i++)

        /* ---- step 3: projection variables (SS :5000-5001; 5eq :7952,7995-7996) ---- */
        double* Zrho_avg[ORC_MAX_SPECIES];
        for (int si = 0; si < ns; si++) Zrho_avg[si] = multi ? dalloc(nside) : 0;
        double* rho_avg = dalloc(nside);
        double* c_avg = dalloc(nside);
        FOR_FACES
        {
            const long xR = cidx(q, i, j, k), xL = xR - st, s = SIDX(i, j, k);
            if (multi)   /* five-eqn :7952; four-eqn FlowModelBasicUtilitiesFourEqnConservative.cpp:4638-5370 (same statements on rho Y_i) */
                for (int si = 0; si < ns; si++) Zrho_avg[si][s] = face_average(Q[si][xL], Q[si][xR]);
            rho_avg[s] = face_average(rho_cell[xL], rho_cell[xR]);
            c_avg[s] = face_average(c[xL], c[xR]);
        }

        /* ---- step 4: characteristic variables of the six stencil cells (:1814-1822) ----
         * SS: BasicUtilitiesSingleSpecies.cpp:6194-6198, 6237-6241, 6324-6329, 6379-6384, 6434-6439
         * 5eq: BasicUtilitiesFiveEqnAllaire.cpp:8835-8921 (x) and the y/z permutations */
        double* W[6][ORC_MAX_EQ];
        for (int m = 0; m < 6; m++)
            for (int e = 0; e < neq; e++) W[m][e] = dalloc(nside);
        for (int m = 0; m < 6; m++) {
            const long off = (long)(m - 3) * st;
            FOR_FACES
            {
                const long x = cidx(q, i, j, k) + off, s = SIDX(i, j, k);
                if (q->model == ORC_SINGLE_SPECIES) {
                    W[m][0][s] = char_acoustic_minus(rho_avg[s], c_avg[s], V[1 + dir][x], V[dim + 1][x]);
                    W[m][1][s] = char_entropy(c_avg[s], V[0][x], V[dim + 1][x]);
                    int w = 2;
                    for (int a = 0; a < dim; a++)
                        if (a != dir) W[m][w++][s] = V[1 + a][x];
                    W[m][dim + 1][s] = char_acoustic_plus(rho_avg[s], c_avg[s], V[1 + dir][x], V[dim + 1][x]);
                } else {
                    W[m][0][s] = fe_char_minus(rho_avg[s], c_avg[s], V[iv + dir][x], V[ip][x]);
                    for (int si = 0; si < ns; si++)
                        W[m][1 + si][s] = fe_char_partial_density(Zrho_avg[si][s], rho_avg[s], c_avg[s], V[si][x], V[ip][x]);
                    int w = ns + 1;
                    for (int a = 0; a < dim; a++)
                        if (a != dir) W[m][w++][s] = V[iv + a][x];
                    for (int si = 0; si < nz; si++) W[m][ns + dim + si][s] = V[ip + 1 + si][x];
                    W[m][neq - 1][s] = fe_char_plus(rho_avg[s], c_avg[s], V[iv + dir][x], V[ip][x]);
                }
            }
        }

        /* ---- step 5: WCNS5-JS interpolation (WCNS5-JS-HLLC-HLL.cpp:492-736) ---- */
        double *W_minus[ORC_MAX_EQ], *W_plus[ORC_MAX_EQ], *V_minus[ORC_MAX_EQ], *V_plus[ORC_MAX_EQ];
        for (int e = 0; e < neq; e++) {
            W_minus[e] = dalloc(nside);
            W_plus[e] = dalloc(nside);
            V_minus[e] = dalloc(nside);
            V_plus[e] = dalloc(nside);
        }
        for (int e = 0; e < neq; e++) {
            FOR_FACES
            {
                const long s = SIDX(i, j, k);
                double U[6];
                for (int m = 0; m < 6; m++) U[m] = W[m][e][s];
                if (d->scheme == ORC_WCNS5_Z)
                    orc_weno5z_point(U, p_exp, &W_minus[e][s], &W_plus[e][s]);
                else if (d->scheme == ORC_WCNS6_LD)
                    orc_weno6ld_point(U, p_exp, d->weno_q > 0 ? d->weno_q : 4, d->weno_C > 0.0 ? d->weno_C : 1.0e9,
                                      d->weno_alpha_tau > 0.0 ? d->weno_alpha_tau : 35.0, &W_minus[e][s], &W_plus[e][s]);
                else
                    orc_weno5js_point(U, p_exp, &W_minus[e][s], &W_plus[e][s]);
            }
        }

        /* ---- step 6: back-projection (SS :7279-7284, 7318-7323, 7373-7379, 7418-7424, 7463-7469;
         *      5eq :9700-9703, 9751-9759 and permutations) ---- */
        for (int side = 0; side < 2; side++) {
            double** Wc = side == 0 ? W_minus : W_plus;
            double** Vs = side == 0 ? V_minus : V_plus;
            FOR_FACES
            {
                const long s = SIDX(i, j, k);
                if (q->model == ORC_SINGLE_SPECIES) {
                    Vs[0][s] = back_density(c_avg[s], Wc[0][s], Wc[1][s], Wc[dim + 1][s]);
                    int w = 2;
                    for (int a = 0; a < dim; a++) {
                        if (a == dir)
                            Vs[1 + a][s] = back_normal_velocity(rho_avg[s], c_avg[s], Wc[0][s], Wc[dim + 1][s]);
                        else
                            Vs[1 + a][s] = Wc[w++][s];
                    }
                    Vs[dim + 1][s] = back_pressure(Wc[0][s], Wc[dim + 1][s]);
                } else {
                    for (int si = 0; si < ns; si++)
                        Vs[si][s] = fe_back_partial_density(Zrho_avg[si][s], c_avg[s], Wc[0][s], Wc[si + 1][s], Wc[neq - 1][s]);
                    int w = ns + 1;
                    for (int a = 0; a < dim; a++) {
                        if (a == dir)
                            Vs[iv + a][s] = fe_back_normal_velocity(Wc[0][s], Wc[neq - 1][s]);
                        else
                            Vs[iv + a][s] = Wc[w++][s];
                    }
                    Vs[ip][s] = fe_back_pressure(rho_avg[s], c_avg[s], Wc[0][s], Wc[neq - 1][s]);
                    for (int si = 0; si < nz; si++) Vs[ip + 1 + si][s] = Wc[ns + dim + si][s];
                }
            }
        }

        /* ---- step 7: bounds flags (SS BasicUtilitiesSingleSpecies.cpp:3013-3433: rho>0 && p>0;
         *      5eq BasicUtilitiesFiveEqnAllaire.cpp:5446-7340) ---- */
        int* flag[2];
        for (int side = 0; side < 2; side++) {
            flag[side] = (int*)malloc(sizeof(int) * (size_t)nside);
            double** Vs = side == 0 ? V_minus : V_plus;
            FOR_FACES
            {
                const long s = SIDX(i, j, k);
                double Vside[ORC_MAX_EQ];
                for (int e = 0; e < neq; e++) Vside[e] = Vs[e][s];
                const int ok = side_bounded(q->model, dim, ns, dir, gam, Vside);
                flag[side][s] = ok;
            }
        }

        /* ---- step 8: first-order fallback (WCNS56-HLLC-HLL.cpp:1884-2039) ---- */
        for (int e = 0; e < neq; e++) {
            FOR_FACES
            {
                const long xR = cidx(q, i, j, k), xL = xR - st, s = SIDX(i, j, k);
                if (flag[0][s] == 0 || flag[1][s] == 0) {
                    V_minus[e][s] = V[e][xL];
                    V_plus[e][s] = V[e][xR];
                }
            }
        }

        /* ---- step 9: HLLC (+ midpoint velocity) and HLLC-HLL (:2045-2070 etc.) ---- */
        double *F_HLLC[ORC_MAX_EQ], *F_HYB[ORC_MAX_EQ], *F_midpoint[ORC_MAX_EQ];
        for (int e = 0; e < neq; e++) {
            F_HLLC[e] = dalloc(nside);
            F_HYB[e] = dalloc(nside);
            F_midpoint[e] = dalloc(nside);
            F_midpoint_all[dir][e] = F_midpoint[e];
        }
        if (has_adv) vel_midpoint[dir] = dalloc(nside);
        double* sensor = dalloc(nside);
        FOR_FACES
        {
            const long s = SIDX(i, j, k);
            double VL[ORC_MAX_EQ], VR[ORC_MAX_EQ], FH[ORC_MAX_EQ], FB[ORC_MAX_EQ], vm;
            for (int e = 0; e < neq; e++) {
                VL[e] = V_minus[e][s];
                VR[e] = V_plus[e][s];
            }
            orc_riemann_point(q->model, dim, ns, gam, dir, VL, VR, FH, FB, &vm);
            for (int e = 0; e < neq; e++) {
                F_HLLC[e][s] = FH[e];
                F_HYB[e][s] = FB[e];
            }
            if (has_adv) vel_midpoint[dir][s] = vm;
        }

        /* ---- step 10: Ducros-like sensor and flux selection (:2072-2134, 2167-2229, 2262-2324) ---- */
        FOR_FACES
        {
            const long s = SIDX(i, j, k);
            const long xR2 = IDX2(i, j, k);
            const long st2 = dir == 0 ? 1 : (dir == 1 ? d2[0] : (long)d2[0] * d2[1]);
            const long xL2 = xR2 - st2;
            sensor[s] = sensor_value(theta[xL2], theta[xR2], Omega[xL2], Omega[xR2]);
        }
        for (int e = 0; e < neq; e++) {
            FOR_FACES
            {
                const long s = SIDX(i, j, k);
                if (sensor[s] > ORC_SENSOR_THRESHOLD)
                    F_midpoint[e][s] = F_HYB[e][s];
                else
                    F_midpoint[e][s] = F_HLLC[e][s];
            }
        }
        if (F_mid_dbg)
            for (int e = 0; e < neq; e++)
                if (F_mid_dbg[dir * neq + e]) memcpy(F_mid_dbg[dir * neq + e], F_midpoint[e], sizeof(double) * (size_t)nside);
        if (sensor_dbg && sensor_dbg[dir]) memcpy(sensor_dbg[dir], sensor, sizeof(double) * (size_t)nside);

        /* ---- step 11: face flux (:2330-2489) ---- */
        {
            long fd[3] = {q->n[0], q->n[1], q->n[2]};
            fd[dir] += 1;
            const long sst = dir == 0 ? 1 : (dir == 1 ? sd[0] : sd[0] * sd[1]);
            const int fhi[3] = {q->n[0] + (dir == 0), q->n[1] + (dir == 1), q->n[2] + (dir == 2)};
            for (int e = 0; e < neq; e++) {
                double* F_face = F[dir * neq + e];
                for (int k = 0; k < fhi[2]; k++)
                    for (int j = 0; j < fhi[1]; j++)
                        for (int i = 0; i < fhi[0]; i++) {
                            const long f = i + fd[0] * ((long)j + fd[1] * (long)k);
                            const long s = SIDX(i, j, k);
                            const long xR = cidx(q, i, j, k), xL = xR - st;
                            F_face[f] = face_flux(dt, F_midpoint[e][s - sst], F_midpoint[e][s], F_midpoint[e][s + sst],
                                                  node_flux(q, Q, vel, p, dir, e, xL), node_flux(q, Q, vel, p, dir, e, xR));
                        }
            }
        }

        /* free per-direction temporaries except the midpoint velocity */
        for (int si = 0; si < ns; si++) free(Zrho_avg[si]);
        free(rho_avg);
        free(c_avg);
        for (int m = 0; m < 6; m++)
            for (int e = 0; e < neq; e++) free(W[m][e]);
        for (int e = 0; e < neq; e++) {
            free(W_minus[e]);
            free(W_plus[e]);
            free(V_minus[e]);
            free(V_plus[e]);
            free(F_HLLC[e]);
            free(F_HYB[e]);
        }
        free(flag[0]);
        free(flag[1]);
        free(sensor);
#undef FOR_FACES
#undef SIDX
    }

    /* ---- step 12: source of the ADVECTIVE equations (:2495-2647; 2D :1240-1369) ---- */
    if (has_adv) {
        for (int si = 0; si < ns - 1; si++) {
            const int e = ns + dim + 1 + si;
            double* Se = S[e];
            for (int k = 0; k < q->n[2]; k++)
                for (int j = 0; j < q->n[1]; j++)
                    for (int i = 0; i < q->n[0]; i++) {
                        const long x = cidx(q, i, j, k);
                        const long xs = i + (long)q->n[0] * ((long)j + (long)q->n[1] * (long)k);
                        double acc = 0.0;
                        for (int dir = 0; dir < dim; dir++) {
                            const long* sd = sd_all[dir];
                            const long s = (i + (dir == 0)) + sd[0] * ((long)(j + (dir == 1)) + sd[1] * (long)(k + (dir == 2)));
                            const long sst = dir == 0 ? 1 : (dir == 1 ? sd[0] : sd[0] * sd[1]);
                            const double* um = vel_midpoint[dir];
                            /* faces: s = face at the low side of the cell ("L"), s+sst = "R" */
                            const double term = adv_source_term(um[s + sst], um[s], vel[dir][x + q->cs[dir]], vel[dir][x - q->cs[dir]],
                                                                um[s + 2 * sst], um[s - sst], d->dx[dir]);
                            acc = (dir == 0) ? term : acc + term;
                        }
                        Se[xs] += dt * Q[e][x] * acc;
                    }
        }
    }

    for (int dir = 0; dir < dim; dir++) {
        for (int e = 0; e < neq; e++) free(F_midpoint_all[dir][e]);
        free(vel_midpoint[dir]);
    }
    for (int m = 0; m < dim * dim; m++) free(grad[m]);
    free(theta);
    free(Omega);
    for (int a = 0; a < dim; a++) free(vel[a]);
    free(p);
    free(c);
    free(rho_m);
#undef IDX2
    return 0;
}

/* Euler::advanceSingleStepOnPatch, Euler.cpp:1003-1679 (3D body :1424-1655);
 * FlowModelFiveEqnAllaire::updateCellDataOfConservativeVariables, FlowModelFiveEqnAllaire.cpp:1739-1886. */
int orc_advance_stage(const orc_desc* d, int ncoef,
                      const double* alpha, const double* beta,
                      const double* const* const* U_int,
                      const double* const* const* F_int,
                      const double* const* const* S_int,
                      double* const* U_out)
{
    geom_t qq;
    make_geom(d, &qq);
    const geom_t* q = &qq;
    const int dim = q->dim, neq = q->neq, ns = q->ns;

    /* fillCellDataOfConservativeVariablesWithZero (all components, whole ghost box) */
    for (int cix = 0; cix < q->ncomp; cix++) memset(U_out[cix], 0, sizeof(double) * (size_t)q->ncell_g);

    for (int n = 0; n < ncoef; n++) {
        if (alpha[n] != 0.0) {
            for (int e = 0; e < neq; e++)
                for (int k = 0; k < q->n[2]; k++)
                    for (int j = 0; j < q->n[1]; j++)
                        for (int i = 0; i < q->n[0]; i++) {
                            const long x = cidx(q, i, j, k);
                            U_out[e][x] += alpha[n] * U_int[n][e][x];
                        }
        }
        if (beta[n] != 0.0) {
            for (int e = 0; e < neq; e++) {
                const double* Fx = F_int[n][0 * neq + e];
                const double* Fy = F_int[n][1 * neq + e];
                const double* Fz = dim == 3 ? F_int[n][2 * neq + e] : 0;
                const double* Sn = S_int[n][e];
                const long nx = q->n[0], ny = q->n[1];
                for (int k = 0; k < q->n[2]; k++)
                    for (int j = 0; j < q->n[1]; j++)
                        for (int i = 0; i < q->n[0]; i++) {
                            const long x = cidx(q, i, j, k);
                            const long xs = i + nx * ((long)j + ny * (long)k);
                            const long fxL = i + (nx + 1) * ((long)j + ny * (long)k);
                            const long fyB = i + nx * ((long)j + (ny + 1) * (long)k);
                            if (dim == 2) {
                                U_out[e][x] += beta[n] * (-(Fx[fxL + 1] - Fx[fxL]) / d->dx[0] -
                                                          (Fy[fyB + nx] - Fy[fyB]) / d->dx[1] + Sn[xs]);
                            } else {
                                const long fzB = xs;
                                U_out[e][x] += rk_beta_term_3d(beta[n], Fx[fxL + 1], Fx[fxL], Fy[fyB + nx], Fy[fyB],
                                                               Fz[fzB + nx * ny], Fz[fzB], d->dx[0], d->dx[1], d->dx[2], Sn[xs]);
                            }
                        }
            }
            if (q->model == ORC_FIVE_EQN_ALLAIRE) {
                const int iz = ns + dim + 1;
                for (int k = 0; k < q->n[2]; k++)
                    for (int j = 0; j < q->n[1]; j++)
                        for (int i = 0; i < q->n[0]; i++) {
                            const long x = cidx(q, i, j, k);
                            U_out[iz + ns - 1][x] = 1.0;
                            for (int si = 0; si < ns - 1; si++) U_out[iz + ns - 1][x] -= U_out[iz + si][x];
                        }
            }
        }
    }
    return 0;
}
