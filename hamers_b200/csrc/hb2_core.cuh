/*
 * hb2_core.cuh -- arithmetic core of the B200 WCNS5-JS / HLLC-HLL path.
 *
 * Everything here is `__host__ __device__` so that the SAME code is (a) inlined into the
 * sm_100a kernels of hb2_sweeps.cu and (b) compiled by g++ into the test-only host emulation
 * harness (tests/host_emu) that checks indexing and arithmetic against the oracle in a
 * container without a GPU.  (b) is never linked into the product library.
 *
 * Reference behaviour restated here (path:line under the reference tree):
 *   cell stage      FlowModelSingleSpecies.cpp:2824-2826, 3049-3051; EquationOfStateIdealGas.cpp:5580, 5909;
 *                   FlowModelFiveEqnAllaire.cpp:3965, 4428-4430, 4779-4853;
 *                   EquationOfStateMixingRulesIdealGas.cpp:7520-7587
 *   projection      FlowModelBasicUtilitiesSingleSpecies.cpp:5000-5001, 6324-6329, 7373-7379;
 *                   FlowModelBasicUtilitiesFiveEqnAllaire.cpp:7952-7996, 8835-8921, 9700-9759
 *   WCNS5-JS        ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:9-164
 *   bounds/fallback FlowModelBasicUtilitiesSingleSpecies.cpp:3013-3433; ...FiveEqnAllaire.cpp:5446-7340;
 *                   ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:1884-2039
 *   Riemann         FlowModelRiemannSolverSingleSpeciesHLLC.cpp:604-1074, ...HLLC-HLL.cpp:889-1629;
 *                   FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp:848-1591, 5640-6040, ...HLLC-HLL.cpp:1283-2440
 *   sensor/select   ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2072-2134
 *   face flux       ...WCNS56-HLLC-HLL.cpp:2330-2489;  source :2495-2647
 *   RK update       Euler.cpp:1424-1655; FlowModelFiveEqnAllaire.cpp:1739-1886
 *
 * This header holds the point arithmetic in the REFERENCE'S OPERATION ORDER (the exact-arithmetic build
 * is bit-identical to the oracle); hb2_fast.cuh holds the re-associated fast variants and hb2_sweep.cuh
 * maps threads of a marching thread block onto them.
 */
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define HB2_HD __host__ __device__ __forceinline__
#else
#define HB2_HD inline
#endif

#define HB2_G 4
#define HB2_EPS 1.0e-15 /* HAMERS_EPSILON, include/HAMeRS_config.hpp.in:16 */
/* hard switches of the path (reported by hb2_constants and compared with the reference's source in the tests) */
#define HB2_SENSOR_THRESHOLD 0.65 /* ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2123, 2218, 2313 */
#define HB2_Y_BOUND_LO (-0.001)   /* FlowModelBasicUtilitiesFiveEqnAllaire.hpp:24-27 */
#define HB2_Y_BOUND_UP 1.001
#define HB2_Z_BOUND_LO (-1000.0)
#define HB2_Z_BOUND_UP 1000.0

namespace hb2 {

enum { SS = 0, FE = 1, FC = 2 };   /* single-species, five-eqn Allaire, four-eqn conservative */
enum { MODE_EMIT = 0, MODE_FUSED = 1 };

template <int MODEL_, int DIM_, int NS_>
struct Traits {
    static constexpr int MODEL = MODEL_;
    static constexpr int DIM = DIM_;
    static constexpr int NS = (MODEL_ == SS) ? 1 : NS_;
    static constexpr int NM = NS;                                  /* mass equations */
    /* FlowModelSingleSpecies.cpp:29, FlowModelFiveEqnAllaire.cpp:29, FlowModelFourEqnConservative.cpp:29 */
    static constexpr int NEQ = (MODEL_ == SS) ? DIM_ + 2 : (MODEL_ == FC ? DIM_ + 1 + NS_ : DIM_ + 2 * NS_);
    static constexpr int NCOMP = (MODEL_ == FE) ? NEQ + 1 : NEQ;   /* + stored Z_last */
    static constexpr int NZ = (MODEL_ == FE) ? NS_ - 1 : 0;        /* advected volume fractions */
    static constexpr int IV = NM;                                  /* first velocity index */
    static constexpr int IP = NM + DIM_;                           /* pressure (V) / energy (Q) index */
    static constexpr bool ADV = (MODEL_ == FE);                    /* has advective equations */
};

struct Geom {
    int dim;
    int n[3];        /* interior cells */
    int g[3];        /* ghost width per direction (0 in the unused 3rd direction of 2D) */
    int gd[3];       /* ghost-box dims */
    long long cs[3]; /* cell strides in the ghost box */
    long long ncell_g;
    double dx[3];
};

struct Consts {
    double gamma[4];
    double inv_gm1[4]; /* 1/(gamma_i - 1), EquationOfStateMixingRulesIdealGas.cpp:7523 */
    double cp[4];      /* four-eqn conservative: species c_p_i = gamma_i/(gamma_i - 1) R_i, c_v_i = 1/(gamma_i - 1) R_i */
    double cv[4];      /* (EquationOfStateMixingRulesIdealGas.cpp:108-119) */
    int weno_p;
    int weno_q;            /* WCNS6-LD: constant_q, constant_C, constant_alpha_tau (WCNS6-LD-HLLC-HLL.cpp:343-361) */
    double weno_C;
    double weno_alpha_tau;
};

/* nonlinear interpolator of the translation unit: the reference's three subclasses of
 * ConvectiveFluxReconstructorWCNS56 differ only in performWENOInterpolation's point kernels */
#define HB2_WCNS5_JS 0
#define HB2_WCNS5_Z 1
#define HB2_WCNS6_LD 2
#ifndef HB2_SCHEME
#define HB2_SCHEME HB2_WCNS5_JS
#endif

#define HB2_MAXC 13
#define HB2_MAXE 12
#define HB2_MAXS 4
#define HB2_MAXT 3 /* states with alpha != 0 in one fused RK stage */

struct DirArgs {
    Geom G;
    Consts K;
    const double* Q[HB2_MAXC]; /* conservative components of the state the flux is evaluated on */
    const unsigned char* hyb;  /* per-cell shock-sensor decisions, ghost-box layout: bit d = face between cells
                                  (c - e_d, c) uses HLLC-HLL (s > 0.65); valid on cells -1..N+1 */
    int mode;                  /* MODE_EMIT | MODE_FUSED */
    double dt;
    double kf[3][3];           /* fast build, fused update: dt/dx_d times {3/2, 1/30, 3/10} (set by dir_args_set_dt) */
    double* F[HB2_MAXE];       /* EMIT: side flux arrays of THIS direction */
    double* S[HB2_MAXE];       /* EMIT, last direction: cell sources (+=), advective equations only */
    double* R[HB2_MAXE];       /* FUSED: running -(dFx/dx) - (dFy/dy) ... per equation, interior layout */
    double* T;                 /* five-eqn: running sum of the velocity-divergence terms, interior layout */
    int ncoef;                 /* FUSED, last direction: RK update */
    double alpha[HB2_MAXS];
    double beta;
    const double* Uint[HB2_MAXS][HB2_MAXC];
    int nterm;                 /* FUSED, last direction: the states with alpha != 0, compacted in stage order */
    double alpha_t[HB2_MAXT];
    const double* Ut[HB2_MAXT][HB2_MAXC];
    double* Uout[HB2_MAXC];
    int seg_len;               /* cells per marching segment along the sweep axis */
    double alpha_q;            /* fast build, FUSED, last direction: coefficient of the flux state (see HB2_NTERM_QREC) */
    /* FUSED, last direction: ghost fill fused into the update.  push[code * NCOMP + comp], code = (ox+1) + 3 (oy+1) +
     * 9 (oz+1), is the address of component `comp` of U_out ON THE PATCH AT OFFSET o (a peer GPU's memory mapped over
     * NVLink, or this patch itself where it is its own periodic neighbour); a cell within the ghost width of a face
     * is also stored into the ghost box of every neighbour that needs it.  NULL table: no push; NULL entry: no
     * neighbour there.  The table lives in device memory. */
    double* const* push;
    int bulk;                  /* 1: the load phase stages whole rows with cp.async.bulk (hb2_sweep.cuh), 0: per-thread cp.async */
};

/* dt and what the update phase derives from it per direction (uniform per launch: kept out of the marching loop, where the
 * division dt/dx cost a call-guarded sequence per iteration) */
inline void dir_args_set_dt(DirArgs* A, double dt)
{
    A->dt = dt;
    for (int d = 0; d < 3; d++) {
        const double k0 = dt / A->G.dx[d];
        A->kf[d][0] = 1.5 * k0;
        A->kf[d][1] = (1.0 / 30.0) * k0;
        A->kf[d][2] = (3.0 / 10.0) * k0;
    }
}

/* Fast build: nterm = HB2_NTERM_QREC + k means "k states in Ut are loaded from HBM and the flux state itself enters
 * the RK combination with alpha_q, rebuilt from the primitive-variable ring" -- 40 B/cell less HBM traffic in the last
 * sweep of every stage (measured: every 5 doubles per cell loaded in a sweep cost it ~1.3 ms at 512^3). */
#define HB2_NTERM_QREC 4
/* NTERM of a launch that materialises the side fluxes (MODE_EMIT): the mode is a compile-time property of the kernel so
 * that neither variant carries the other's code (the sweeps sit at ~30 KB of instructions). */
#define HB2_NTERM_EMIT (-1)

HB2_HD long long cidx(const Geom& G, int i, int j, int k)
{
    return (i + G.g[0]) + (long long)(j + G.g[1]) * G.cs[1] + (long long)(k + G.g[2]) * G.cs[2];
}
/* ghost-0 cell (interior layout) */
HB2_HD long long iidx(const Geom& G, int i, int j, int k)
{
    return i + (long long)G.n[0] * (j + (long long)G.n[1] * k);
}
/* ghost-0 side array of direction dir: extent n+1 in dir */
template <int DIR>
HB2_HD long long sidx(const Geom& G, int i, int j, int k)
{
    const long long e0 = G.n[0] + (DIR == 0 ? 1 : 0);
    const long long e1 = G.n[1] + (DIR == 1 ? 1 : 0);
    return i + e0 * (j + e1 * (long long)k);
}

/* ------------------------------------------------------------------------------------------
 * cell stage: conservative -> primitive + sound speed
 * ---------------------------------------------------------------------------------------- */
/* four-eqn conservative: c = sqrt(Gamma p/rho + sum Y_i Psi_i), Psi_i = ((c_p_i - gamma c_v_i)/c_v + gamma - 1) epsilon with
 * epsilon recomputed from p (FlowModelFourEqnConservative.cpp:5087-5430; EquationOfStateMixingRulesIdealGas.cpp:6406-6409) */
template <int NS>
HB2_HD double fc_sound_speed(const Consts& K, double rho, const double (&Y)[NS], double p, double gamma_m, double c_v)
{
    const double eps = p / ((gamma_m - 1.0) * rho);
    double cc = (gamma_m - 1.0) * p / rho;
#pragma unroll
    for (int si = 0; si < NS; si++) cc += Y[si] * (((K.cp[si] - gamma_m * K.cv[si]) / c_v + gamma_m - 1.0) * eps);
    return sqrt(cc);
}

template <class Tr>
HB2_HD void cons_to_prim(const double (&q)[Tr::NCOMP], const Consts& K, double (&V)[Tr::NEQ], double& c)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS;
    if (Tr::MODEL == SS) {
        const double rho = q[0];
        V[0] = rho;
        double ke = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            V[1 + a] = q[1 + a] / rho;
            ke = (a == 0) ? V[1 + a] * V[1 + a] : ke + V[1 + a] * V[1 + a];
        }
        const double epsilon = q[DIM + 1] / rho - 0.5 * ke;
        const double p = (K.gamma[0] - 1.0) * rho * epsilon;
        V[DIM + 1] = p;
        c = sqrt(K.gamma[0] * p / rho);
    } else if (Tr::MODEL == FC) {
        /* FlowModelFourEqnConservative.cpp:3877-5430; mixture rules with MASS fractions
         * (EquationOfStateMixingRulesIdealGas.cpp:6946-7150, 6367-6564) */
        double rho = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) rho += q[si];
        double Y[NS];
#pragma unroll
        for (int si = 0; si < NS; si++) {
            V[si] = q[si];
            Y[si] = q[si] / rho;
        }
        double ke = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            V[NS + a] = q[NS + a] / rho;
            ke = (a == 0) ? V[NS + a] * V[NS + a] : ke + V[NS + a] * V[NS + a];
        }
        const double epsilon = q[NS + DIM] / rho - 0.5 * ke;
        double c_p = 0.0, c_v = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) {
            c_p += Y[si] * K.cp[si];
            c_v += Y[si] * K.cv[si];
        }
        const double gamma_m = c_p / c_v;
        const double p = (gamma_m - 1.0) * rho * epsilon;
        V[NS + DIM] = p;
        c = fc_sound_speed<NS>(K, rho, Y, p, gamma_m, c_v);
    } else {
        double rho = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) rho += q[si];
        double Y[NS];
#pragma unroll
        for (int si = 0; si < NS; si++) {
            V[si] = q[si];
            Y[si] = q[si] / rho;
        }
        double ke = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            V[NS + a] = q[NS + a] / rho;
            ke = (a == 0) ? V[NS + a] * V[NS + a] : ke + V[NS + a] * V[NS + a];
        }
        const double epsilon = q[NS + DIM] / rho - 0.5 * ke;
        double xi = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) xi += q[NS + DIM + 1 + si] * K.inv_gm1[si];
        const double gamma_m = 1.0 / xi + 1.0;
        const double p = (gamma_m - 1.0) * rho * epsilon;
        V[NS + DIM] = p;
        const double Gamma = gamma_m - 1.0;
        double cc = Gamma * p / rho;
#pragma unroll
        for (int si = 0; si < NS; si++) cc += Y[si] * (p / rho);
        c = sqrt(cc);
#pragma unroll
        for (int si = 0; si < Tr::NZ; si++) V[NS + DIM + 1 + si] = q[NS + DIM + 1 + si];
    }
}

/* node flux of one cell in direction DIR from its conservative (q) and primitive (V) variables */
template <class Tr, int DIR>
HB2_HD void node_flux(const double (&q)[Tr::NCOMP], const double (&V)[Tr::NEQ], double (&Fn)[Tr::NEQ])
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, IV = Tr::IV, IP = Tr::IP;
    const double un = V[IV + DIR];
    const double p = V[IP];
    if (Tr::MODEL == SS) {
        Fn[0] = q[1 + DIR];
#pragma unroll
        for (int a = 0; a < DIM; a++) Fn[1 + a] = (a == DIR) ? un * q[1 + a] + p : un * q[1 + a];
        Fn[DIM + 1] = un * (q[DIM + 1] + p);
    } else {
#pragma unroll
        for (int si = 0; si < NS; si++) Fn[si] = un * q[si];
#pragma unroll
        for (int a = 0; a < DIM; a++) Fn[NS + a] = (a == DIR) ? un * q[NS + a] + p : un * q[NS + a];
        Fn[NS + DIM] = un * (q[NS + DIM] + p);
#pragma unroll
        for (int si = 0; si < Tr::NZ; si++) Fn[NS + DIM + 1 + si] = un * q[NS + DIM + 1 + si];
    }
}

/* ------------------------------------------------------------------------------------------
 * WCNS5-JS one-sided midpoint interpolation
 * ---------------------------------------------------------------------------------------- */
HB2_HD double ipow_(double base, int e)
{
    double r = base;
    for (int i = 1; i < e; i++) r *= base;
    return r;
}

template <int MATH>
HB2_HD double weno5js_side(double a, double b, double c, double d, double e, int p)
{
    const double beta_0 = 1.0 / 3.0 * (a * (4.0 * a - 19.0 * b + 11.0 * c) + b * (25.0 * b - 31.0 * c) + 10.0 * c * c);
    const double beta_1 = 1.0 / 3.0 * (b * (4.0 * b - 13.0 * c + 5.0 * d) + 13.0 * c * (c - d) + 4.0 * d * d);
    const double beta_2 = 1.0 / 3.0 * (c * (10.0 * c - 31.0 * d + 11.0 * e) + d * (25.0 * d - 19.0 * e) + 4.0 * e * e);

    if (MATH == 0) {
        const double b0 = beta_0 + HB2_EPS, b1 = beta_1 + HB2_EPS, b2 = beta_2 + HB2_EPS;
        double omega_0 = 1.0 / 16.0 / (p == 2 ? b0 * b0 : ipow_(b0, p));
        double omega_1 = 5.0 / 8.0 / (p == 2 ? b1 * b1 : ipow_(b1, p));
        double omega_2 = 5.0 / 16.0 / (p == 2 ? b2 * b2 : ipow_(b2, p));
        const double omega_sum = omega_0 + omega_1 + omega_2;
        omega_0 = omega_0 / omega_sum;
        omega_1 = omega_1 / omega_sum;
        omega_2 = omega_2 / omega_sum;
        return 3.0 / 8.0 * omega_0 * a + (-10.0 / 8.0 * omega_0 - 1.0 / 8.0 * omega_1) * b +
               (15.0 / 8.0 * omega_0 + 6.0 / 8.0 * omega_1 + 3.0 / 8.0 * omega_2) * c +
               (3.0 / 8.0 * omega_1 + 6.0 / 8.0 * omega_2) * d - 1.0 / 8.0 * omega_2 * e;
    } else {
        /* One division instead of six: with b_k = (beta_k + eps)^p the normalised weights are
         * omega_k = d_k * prod_{j != k} b_j / sum_k(d_k * prod_{j != k} b_j).  Algebraically
         * identical to the reference; differs by a few ulp of re-association (FP64 range is
         * ample: b_k >= 1e-30 for p = 2). */
        const double e0 = beta_0 + HB2_EPS, e1 = beta_1 + HB2_EPS, e2 = beta_2 + HB2_EPS;
        const double b0 = (p == 2 ? e0 * e0 : ipow_(e0, p));
        const double b1 = (p == 2 ? e1 * e1 : ipow_(e1, p));
        const double b2 = (p == 2 ? e2 * e2 : ipow_(e2, p));
        const double a0 = (1.0 / 16.0) * (b1 * b2);
        const double a1 = (5.0 / 8.0) * (b0 * b2);
        const double a2 = (5.0 / 16.0) * (b0 * b1);
        const double inv = 1.0 / (a0 + a1 + a2);
        /* sub-stencil midpoint values */
        const double P0 = 0.375 * a - 1.25 * b + 1.875 * c;
        const double P1 = -0.125 * b + 0.75 * c + 0.375 * d;
        const double P2 = 0.375 * c + 0.75 * d - 0.125 * e;
        return (a0 * P0 + a1 * P1 + a2 * P2) * inv;
    }
}

/* the three smoothness indicators shared by WCNS5-JS / -Z / WCNS6-LD */
HB2_HD void beta_012(double a, double b, double c, double d, double e, double& b0, double& b1, double& b2)
{
    b0 = 1.0 / 3.0 * (a * (4.0 * a - 19.0 * b + 11.0 * c) + b * (25.0 * b - 31.0 * c) + 10.0 * c * c);
    b1 = 1.0 / 3.0 * (b * (4.0 * b - 13.0 * c + 5.0 * d) + 13.0 * c * (c - d) + 4.0 * d * d);
    b2 = 1.0 / 3.0 * (c * (10.0 * c - 31.0 * d + 11.0 * e) + d * (25.0 * d - 19.0 * e) + 4.0 * e * e);
}

/* WCNS5-Z (ConvectiveFluxReconstructorWCNS5-Z-HLLC-HLL.cpp:78-165), reference operation order */
HB2_HD double weno5z_side(double a, double b, double c, double d, double e, int p)
{
    double beta_0, beta_1, beta_2;
    beta_012(a, b, c, d, e, beta_0, beta_1, beta_2);
    const double tau_5 = fabs(beta_0 - beta_2);
    double omega_0 = 1.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_0 + HB2_EPS), p));
    double omega_1 = 5.0 / 8.0 * (1.0 + ipow_(tau_5 / (beta_1 + HB2_EPS), p));
    double omega_2 = 5.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_2 + HB2_EPS), p));
    const double omega_sum = omega_0 + omega_1 + omega_2;
    omega_0 = omega_0 / omega_sum;
    omega_1 = omega_1 / omega_sum;
    omega_2 = omega_2 / omega_sum;
    return 3.0 / 8.0 * omega_0 * a + (-10.0 / 8.0 * omega_0 - 1.0 / 8.0 * omega_1) * b +
           (15.0 / 8.0 * omega_0 + 6.0 / 8.0 * omega_1 + 3.0 / 8.0 * omega_2) * c +
           (3.0 / 8.0 * omega_1 + 6.0 / 8.0 * omega_2) * d - 1.0 / 8.0 * omega_2 * e;
}

/* WCNS6-LD (ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp:23-39 sigma, :51-83 beta, :143-339 interpolation),
 * reference operation order; sigma comes from the unmirrored stencil for both sides */
HB2_HD double weno6ld_sigma(double w1, double w2, double w3, double w4)
{
    const double alpha_1 = w2 - w1;
    const double alpha_2 = w3 - w2;
    const double alpha_3 = w4 - w3;
    const double theta_1 = fabs(alpha_1 - alpha_2) / (fabs(alpha_1) + fabs(alpha_2) + HB2_EPS);
    const double theta_2 = fabs(alpha_2 - alpha_3) / (fabs(alpha_2) + fabs(alpha_3) + HB2_EPS);
    return fmax(theta_1, theta_2);
}

HB2_HD double weno6ld_side(double a, double b, double c, double d, double e, double f, double sigma, const Consts& K)
{
    const int p = K.weno_p, q = K.weno_q;
    double beta_0, beta_1, beta_2;
    beta_012(a, b, c, d, e, beta_0, beta_1, beta_2);
    const double beta_3 = 1.0 / 232243200.0 * (a * (525910327.0 * a - 4562164630.0 * b + 7799501420.0 * c -
        6610694540.0 * d + 2794296070.0 * e - 472758974.0 * f) + 5.0 * b *
        (2146987907.0 * b - 7722406988.0 * c + 6763559276.0 * d - 2926461814.0 * e + 503766638.0 * f) + 20.0 * c *
        (1833221603.0 * c - 3358664662.0 * d + 1495974539.0 * e - 263126407.0 * f) +
        20.0 * d * (1607794163.0 * d - 1486026707.0 * e + 268747951.0 * f) +
        5.0 * e * (1432381427.0 * e - 536951582.0 * f) +
        263126407.0 * f * f);
    double omega_upwind_0, omega_upwind_1, omega_upwind_2;
    const double tau_5 = fabs(beta_0 - beta_2);
    omega_upwind_0 = 1.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_0 + HB2_EPS), p));
    omega_upwind_1 = 5.0 / 8.0 * (1.0 + ipow_(tau_5 / (beta_1 + HB2_EPS), p));
    omega_upwind_2 = 5.0 / 16.0 * (1.0 + ipow_(tau_5 / (beta_2 + HB2_EPS), p));
    const double omega_upwind_sum = omega_upwind_0 + omega_upwind_1 + omega_upwind_2;
    omega_upwind_0 = omega_upwind_0 / omega_upwind_sum;
    omega_upwind_1 = omega_upwind_1 / omega_upwind_sum;
    omega_upwind_2 = omega_upwind_2 / omega_upwind_sum;
    double omega_0, omega_1, omega_2, omega_3;
    const double beta_avg = 1.0 / 8.0 * (beta_0 + beta_2 + 6.0 * beta_1);
    const double tau_6 = fabs(beta_3 - beta_avg);
    omega_0 = 1.0 / 32.0 * (K.weno_C + ipow_(tau_6 / (beta_0 + HB2_EPS), q));
    omega_1 = 15.0 / 32.0 * (K.weno_C + ipow_(tau_6 / (beta_1 + HB2_EPS), q));
    omega_2 = 15.0 / 32.0 * (K.weno_C + ipow_(tau_6 / (beta_2 + HB2_EPS), q));
    omega_3 = 1.0 / 32.0 * (K.weno_C + ipow_(tau_6 / (beta_3 + HB2_EPS), q));
    const double omega_sum = omega_0 + omega_1 + omega_2 + omega_3;
    omega_0 = omega_0 / omega_sum;
    omega_1 = omega_1 / omega_sum;
    omega_2 = omega_2 / omega_sum;
    omega_3 = omega_3 / omega_sum;
    const double R_tau = tau_6 / (beta_avg + HB2_EPS);
    if (R_tau > K.weno_alpha_tau) {
        omega_0 = sigma * omega_upwind_0 + (1.0 - sigma) * omega_0;
        omega_1 = sigma * omega_upwind_1 + (1.0 - sigma) * omega_1;
        omega_2 = sigma * omega_upwind_2 + (1.0 - sigma) * omega_2;
        omega_3 = (1.0 - sigma) * omega_3;
    }
    return 3.0 / 8.0 * omega_0 * a + (-10.0 / 8.0 * omega_0 - 1.0 / 8.0 * omega_1) * b +
           (15.0 / 8.0 * omega_0 + 6.0 / 8.0 * omega_1 + 3.0 / 8.0 * omega_2) * c +
           (3.0 / 8.0 * omega_1 + 6.0 / 8.0 * omega_2 + 15.0 / 8.0 * omega_3) * d +
           (-1.0 / 8.0 * omega_2 - 10.0 / 8.0 * omega_3) * e + 3.0 / 8.0 * omega_3 * f;
}

/* minus / plus midpoint values of one characteristic field from its six stencil values, interpolator of this
 * translation unit (HB2_SCHEME) */
template <int MATH>
HB2_HD void weno_pair(double w0, double w1, double w2, double w3, double w4, double w5, const Consts& K, double& wm,
                      double& wp)
{
#if HB2_SCHEME == HB2_WCNS5_Z
    wm = weno5z_side(w0, w1, w2, w3, w4, K.weno_p);
    wp = weno5z_side(w5, w4, w3, w2, w1, K.weno_p);
#elif HB2_SCHEME == HB2_WCNS6_LD
    const double sigma = weno6ld_sigma(w1, w2, w3, w4);
    wm = weno6ld_side(w0, w1, w2, w3, w4, w5, sigma, K);
    wp = weno6ld_side(w5, w4, w3, w2, w1, w0, sigma, K);
#else
    wm = weno5js_side<MATH>(w0, w1, w2, w3, w4, K.weno_p);
    wp = weno5js_side<MATH>(w5, w4, w3, w2, w1, K.weno_p);
#endif
}

/* ------------------------------------------------------------------------------------------
 * side thermodynamics for the Riemann solvers
 * ---------------------------------------------------------------------------------------- */
template <class Tr>
HB2_HD void side_thermo(const double (&V)[Tr::NEQ], const Consts& K, double& rho, double& c, double& eps)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS;
    if (Tr::MODEL == SS) {
        rho = V[0];
        const double p = V[DIM + 1];
        c = sqrt(K.gamma[0] * p / rho);
        eps = p / ((K.gamma[0] - 1.0) * rho);
    } else if (Tr::MODEL == FC) {
        /* FlowModelRiemannSolverFourEqnConservativeHLLC-HLL.cpp:5150-5290 */
        double r = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) r += V[si];
        const double p = V[NS + DIM];
        double Y[NS];
        double c_p = 0.0, c_v = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) Y[si] = V[si] / r;
#pragma unroll
        for (int si = 0; si < NS; si++) {
            c_p += Y[si] * K.cp[si];
            c_v += Y[si] * K.cv[si];
        }
        const double gamma_m = c_p / c_v;
        rho = r;
        eps = p / ((gamma_m - 1.0) * r);
        c = fc_sound_speed<NS>(K, r, Y, p, gamma_m, c_v);
    } else {
        double r = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) r += V[si];
        const double p = V[NS + DIM];
        double xi = 0.0, Z_last = 1.0;
#pragma unroll
        for (int si = 0; si < NS - 1; si++) {
            xi += V[NS + DIM + 1 + si] * K.inv_gm1[si];
            Z_last -= V[NS + DIM + 1 + si];
        }
        xi += Z_last / (K.gamma[NS - 1] - 1.0);
        const double gamma_m = 1.0 / xi + 1.0;
        const double Gamma = gamma_m - 1.0;
        double cc = Gamma * p / r;
#pragma unroll
        for (int si = 0; si < NS; si++) cc += (V[si] / r) * (p / r);
        rho = r;
        c = sqrt(cc);
        eps = p / ((gamma_m - 1.0) * r);
    }
}

/* bounded flag of one interpolated side.  Five-eqn c^2 check: the reference accumulates Y_i Psi_i onto Gamma p / rho only in
 * its x-direction block (FlowModelBasicUtilitiesFiveEqnAllaire.cpp:6654-6678); its y and z blocks re-assign
 * c_sq = Gamma p / rho inside the species loop (:6968-6992, 7281-7305), so the test there is Gamma p / rho > 0.  Mirrored. */
template <class Tr, int DIR>
HB2_HD int side_bounded(const double (&V)[Tr::NEQ], const Consts& K)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, NEQ = Tr::NEQ;
    int ok = 1;
    if (Tr::MODEL == SS) {
        ok &= (V[0] > 0.0) ? 1 : 0;
        ok &= (V[NEQ - 1] > 0.0) ? 1 : 0;
    } else if (Tr::MODEL == FC) {
        /* FlowModelBasicUtilitiesFourEqnConservative.cpp:3737-4510 */
        double rho = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) rho += V[si];
#pragma unroll
        for (int si = 0; si < NS; si++) {
            const double Y = V[si] / rho;
            ok &= (Y > HB2_Y_BOUND_LO && Y < HB2_Y_BOUND_UP) ? 1 : 0;
        }
        ok &= (rho > 0.0) ? 1 : 0;
        ok &= (V[NS + DIM] > 0.0) ? 1 : 0;
    } else {
        const double Z_lo = HB2_Z_BOUND_LO, Z_up = HB2_Z_BOUND_UP, Y_lo = HB2_Y_BOUND_LO, Y_up = HB2_Y_BOUND_UP;
        double Z[NS];
        Z[NS - 1] = 1.0;
#pragma unroll
        for (int si = 0; si < NS - 1; si++) {
            Z[si] = V[NS + DIM + 1 + si];
            Z[NS - 1] -= Z[si];
            ok &= (Z[si] > Z_lo && Z[si] < Z_up) ? 1 : 0;
        }
        ok &= (Z[NS - 1] > Z_lo && Z[NS - 1] < Z_up) ? 1 : 0;
        double rho = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) rho += V[si];
        double Y[NS];
#pragma unroll
        for (int si = 0; si < NS; si++) {
            Y[si] = V[si] / rho;
            ok &= (Y[si] > Y_lo && Y[si] < Y_up) ? 1 : 0;
        }
#pragma unroll
        for (int si = 0; si < NS; si++) ok &= (V[si] > 0.0) ? 1 : 0;
        const double p = V[NS + DIM];
        double xi = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) xi += Z[si] * K.inv_gm1[si];
        const double gamma_m = 1.0 / xi + 1.0;
        const double Gamma = gamma_m - 1.0;
        double c_sq = Gamma * p / rho;
        if (DIR == 0) {
#pragma unroll
            for (int si = 0; si < NS; si++) c_sq += Y[si] * (p / rho);
        }
        ok &= (c_sq > 0.0) ? 1 : 0;
    }
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * HLLC (always) and HLLC-HLL (when `hybrid`) midpoint flux + HLLC midpoint normal velocity
 * ---------------------------------------------------------------------------------------- */
template <class Tr, int DIR>
HB2_HD void riemann(const double (&V_L)[Tr::NEQ], const double (&V_R)[Tr::NEQ], const Consts& K, bool hybrid,
                    double (&Fm)[Tr::NEQ], double& vel_mid)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, NEQ = Tr::NEQ, NM = Tr::NM, IV = Tr::IV, IP = Tr::IP;

    double rho_L, rho_R, c_L, c_R, eps_L, eps_R;
    side_thermo<Tr>(V_L, K, rho_L, c_L, eps_L);
    side_thermo<Tr>(V_R, K, rho_R, c_R, eps_R);

    const double un_L = V_L[IV + DIR], un_R = V_R[IV + DIR];
    const double p_L = V_L[IP], p_R = V_R[IP];

    const double u_average = 0.5 * (un_L + un_R);
    const double c_average = 0.5 * (c_L + c_R);
    const double s_L = fmin(u_average - c_average, un_L - c_L);
    const double s_R = fmax(u_average + c_average, un_R + c_R);
    const double s_minus = fmin(0.0, s_L);
    const double s_plus = fmax(0.0, s_R);
    const double s_star = (p_R - p_L + rho_L * un_L * (s_L - un_L) - rho_R * un_R * (s_R - un_R)) /
                          (rho_L * (s_L - un_L) - rho_R * (s_R - un_R));

    /* conservative state and physical flux of one side */
    auto side_QF = [&](const double (&V)[NEQ], double rho, double eps, double (&Q)[NEQ], double (&F)[NEQ]) {
        const double un = V[IV + DIR];
        const double p = V[IP];
        double ke = V[IV] * V[IV];
#pragma unroll
        for (int a = 1; a < DIM; a++) ke = ke + V[IV + a] * V[IV + a];
        if (Tr::MODEL == SS) {
            Q[0] = V[0];
#pragma unroll
            for (int a = 0; a < DIM; a++) Q[1 + a] = V[0] * V[1 + a];
            Q[IP] = V[0] * (eps + 0.5 * ke);
            F[0] = Q[1 + DIR];
#pragma unroll
            for (int a = 0; a < DIM; a++) F[1 + a] = (a == DIR) ? Q[1 + DIR] * V[1 + a] + p : Q[1 + DIR] * V[1 + a];
            F[IP] = un * (Q[IP] + p);
        } else {
#pragma unroll
            for (int si = 0; si < NS; si++) Q[si] = V[si];
#pragma unroll
            for (int a = 0; a < DIM; a++) Q[IV + a] = rho * V[IV + a];
            Q[IP] = rho * (eps + 0.5 * ke);
#pragma unroll
            for (int si = 0; si < Tr::NZ; si++) Q[IP + 1 + si] = V[IP + 1 + si];
#pragma unroll
            for (int si = 0; si < NS; si++) F[si] = un * V[si];
#pragma unroll
            for (int a = 0; a < DIM; a++) F[IV + a] = (a == DIR) ? un * Q[IV + a] + p : un * Q[IV + a];
            F[IP] = un * (Q[IP] + p);
#pragma unroll
            for (int si = 0; si < Tr::NZ; si++) F[IP + 1 + si] = un * V[IP + 1 + si];
        }
    };

    /* HLLC star state on the upwind side of the contact */
    auto hllc = [&](const double (&V)[NEQ], const double (&Q)[NEQ], const double (&F)[NEQ], double rho, double s_K,
                    double s_mp) {
        const double un = V[IV + DIR];
        const double p = V[IP];
        const double Chi = (s_K - un) / (s_K - s_star);
        double Qs[NEQ];
#pragma unroll
        for (int si = 0; si < NM; si++) Qs[si] = Chi * V[si];
#pragma unroll
        for (int a = 0; a < DIM; a++) Qs[IV + a] = (a == DIR) ? Chi * rho * s_star : Chi * Q[IV + a];
        Qs[IP] = Chi * (Q[IP] + (s_star - un) * (rho * s_star + p / (s_K - un)));
#pragma unroll
        for (int e = IP + 1; e < NEQ; e++) Qs[e] = Chi * V[e];
#pragma unroll
        for (int e = 0; e < NEQ; e++) Fm[e] = F[e] + s_mp * (Qs[e] - Q[e]);
        vel_mid = un + s_mp * (Chi - 1.0);
    };

    if (!hybrid) {
        double Q[NEQ], F[NEQ];
        if (s_star > 0.0) {
            side_QF(V_L, rho_L, eps_L, Q, F);
            hllc(V_L, Q, F, rho_L, s_L, s_minus);
        } else {
            side_QF(V_R, rho_R, eps_R, Q, F);
            hllc(V_R, Q, F, rho_R, s_R, s_plus);
        }
        return;
    }

    double Q_L[NEQ], Q_R[NEQ], F_L[NEQ], F_R[NEQ];
    side_QF(V_L, rho_L, eps_L, Q_L, F_L);
    side_QF(V_R, rho_R, eps_R, Q_R, F_R);
    if (s_star > 0.0)
        hllc(V_L, Q_L, F_L, rho_L, s_L, s_minus);
    else
        hllc(V_R, Q_R, F_R, rho_R, s_R, s_plus);

    double diff[DIM];
#pragma unroll
    for (int a = 0; a < DIM; a++) diff[a] = V_R[IV + a] - V_L[IV + a];
    double mag2 = diff[0] * diff[0];
#pragma unroll
    for (int a = 1; a < DIM; a++) mag2 = mag2 + diff[a] * diff[a];
    const double vel_mag = sqrt(mag2);
    double alpha_1, alpha_2;
    if (vel_mag < HB2_EPS) {
        alpha_1 = 1.0;
        alpha_2 = 0.0;
    } else {
        alpha_1 = fabs(diff[DIR]) / vel_mag;
        alpha_2 = sqrt(1.0 - alpha_1 * alpha_1);
    }
    const double beta_1 = 0.5 * (1.0 + alpha_1 / (alpha_1 + alpha_2));
    const double beta_2 = 1.0 - beta_1;

#pragma unroll
    for (int e = 0; e < NEQ; e++) {
        if (e == IV + DIR || e == IP) continue;
        double F_HLL = (s_R * F_L[e] - s_L * F_R[e] + s_R * s_L * (Q_R[e] - Q_L[e])) / (s_R - s_L);
        if (s_L > 0.0) F_HLL = F_L[e];
        if (s_R < 0.0) F_HLL = F_R[e];
        Fm[e] = beta_1 * Fm[e] + beta_2 * F_HLL;
    }
}

/* ------------------------------------------------------------------------------------------
 * One midpoint flux from the six stencil cells f-3..f+2 (primitive variables + sound speed of
 * cells L = 2 and R = 3; `hybrid` = the face's shock-sensor decision, see face_sensor below).
 * ---------------------------------------------------------------------------------------- */
template <class Tr, int DIR, int MATH>
HB2_HD void face_midpoint(const double (&V)[6][Tr::NEQ], double c_cellL, double c_cellR, bool hybrid, const Consts& K,
                          double (&Fm)[Tr::NEQ], double& vel_mid)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, NEQ = Tr::NEQ, IV = Tr::IV, IP = Tr::IP;

    double V_minus[NEQ], V_plus[NEQ];
    const double c_avg = 0.5 * (c_cellL + c_cellR);

    if (Tr::MODEL == SS) {
        const double rho_avg = 0.5 * (V[2][0] + V[3][0]);
        const double kn = -0.5 * rho_avg * c_avg; /* (-1/2*rho_avg)*c_avg */
        const double kp = 0.5 * rho_avg * c_avg;
        const double r_cc = 1.0 / (c_avg * c_avg);
        const double r_rc = 1.0 / (rho_avg * c_avg);
        double W[6], W0m, W0p, W1m, W1p, WLm, WLp;
        /* field 0 */
#pragma unroll
        for (int m = 0; m < 6; m++) W[m] = kn * V[m][1 + DIR] + 0.5 * V[m][IP];
        weno_pair<MATH>(W[0], W[1], W[2], W[3], W[4], W[5], K, W0m, W0p);
        /* field 1 */
#pragma unroll
        for (int m = 0; m < 6; m++) W[m] = V[m][0] - r_cc * V[m][IP];
        weno_pair<MATH>(W[0], W[1], W[2], W[3], W[4], W[5], K, W1m, W1p);
        /* last field */
#pragma unroll
        for (int m = 0; m < 6; m++) W[m] = kp * V[m][1 + DIR] + 0.5 * V[m][IP];
        weno_pair<MATH>(W[0], W[1], W[2], W[3], W[4], W[5], K, WLm, WLp);
        /* tangential velocities are their own characteristic fields */
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            if (a == DIR) continue;
            weno_pair<MATH>(V[0][1 + a], V[1][1 + a], V[2][1 + a], V[3][1 + a], V[4][1 + a], V[5][1 + a], K, V_minus[1 + a], V_plus[1 + a]);
        }
        V_minus[0] = r_cc * W0m + W1m + r_cc * WLm;
        V_plus[0] = r_cc * W0p + W1p + r_cc * WLp;
        V_minus[1 + DIR] = -r_rc * W0m + r_rc * WLm;
        V_plus[1 + DIR] = -r_rc * W0p + r_rc * WLp;
        V_minus[IP] = W0m + WLm;
        V_plus[IP] = W0p + WLp;
    } else {
        double rho_L = 0.0, rho_R = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) {
            rho_L += V[2][si];
            rho_R += V[3][si];
        }
        const double rho_avg = 0.5 * (rho_L + rho_R);
        const double rc = rho_avg * c_avg;
        const double r_rc = 1.0 / rc;
        double W[6], W0m, W0p, WLm, WLp;
#pragma unroll
        for (int m = 0; m < 6; m++) W[m] = V[m][IV + DIR] - r_rc * V[m][IP];
        weno_pair<MATH>(W[0], W[1], W[2], W[3], W[4], W[5], K, W0m, W0p);
#pragma unroll
        for (int m = 0; m < 6; m++) W[m] = V[m][IV + DIR] + r_rc * V[m][IP];
        weno_pair<MATH>(W[0], W[1], W[2], W[3], W[4], W[5], K, WLm, WLp);
#pragma unroll
        for (int si = 0; si < NS; si++) {
            const double Zrho_avg = 0.5 * (V[2][si] + V[3][si]);
            const double zc = Zrho_avg / (rc * c_avg);
#pragma unroll
            for (int m = 0; m < 6; m++) W[m] = V[m][si] - zc * V[m][IP];
            double Wm, Wp;
            weno_pair<MATH>(W[0], W[1], W[2], W[3], W[4], W[5], K, Wm, Wp);
            const double yn = -0.5 * Zrho_avg / c_avg;
            const double yp = 0.5 * Zrho_avg / c_avg;
            V_minus[si] = yn * W0m + Wm + yp * WLm;
            V_plus[si] = yn * W0p + Wp + yp * WLp;
        }
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            if (a == DIR) continue;
            weno_pair<MATH>(V[0][IV + a], V[1][IV + a], V[2][IV + a], V[3][IV + a], V[4][IV + a], V[5][IV + a], K, V_minus[IV + a], V_plus[IV + a]);
        }
#pragma unroll
        for (int si = 0; si < Tr::NZ; si++) {
            const int e = IP + 1 + si;
            weno_pair<MATH>(V[0][e], V[1][e], V[2][e], V[3][e], V[4][e], V[5][e], K, V_minus[e], V_plus[e]);
        }
        V_minus[IV + DIR] = 0.5 * W0m + 0.5 * WLm;
        V_plus[IV + DIR] = 0.5 * W0p + 0.5 * WLp;
        const double kn = -0.5 * rho_avg * c_avg;
        const double kp = 0.5 * rho_avg * c_avg;
        V_minus[IP] = kn * W0m + kp * WLm;
        V_plus[IP] = kn * W0p + kp * WLp;
    }

    /* bounds check and first-order fallback */
    const int ok = side_bounded<Tr, DIR>(V_minus, K) & side_bounded<Tr, DIR>(V_plus, K);
    if (!ok) {
#pragma unroll
        for (int e = 0; e < NEQ; e++) {
            V_minus[e] = V[2][e];
            V_plus[e] = V[3][e];
        }
    }

    riemann<Tr, DIR>(V_minus, V_plus, K, hybrid, Fm, vel_mid);
}

/* ------------------------------------------------------------------------------------------
 * helpers shared by the sweeps
 * ---------------------------------------------------------------------------------------- */
/* streaming accesses of data touched once per sweep (running right-hand side, RK states): HB2_STREAM_HINTS=1 uses the
 * evict-first / no-allocate cache hints */
#ifndef HB2_STREAM_HINTS
#define HB2_STREAM_HINTS 0
#endif
HB2_HD double load_stream(const double* p)
{
#if defined(__CUDA_ARCH__) && HB2_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}
HB2_HD void store_stream(double* p, double v)
{
#if defined(__CUDA_ARCH__) && HB2_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}

template <class Tr>
HB2_HD void load_cons(const DirArgs& A, long long x, double (&q)[Tr::NCOMP])
{
#pragma unroll
    for (int cix = 0; cix < Tr::NCOMP; cix++) q[cix] = A.Q[cix][x];
}

/* RK update of one interior cell from the complete right-hand side (FUSED, last direction).
 * Order of Euler.cpp:1479, 1544-1548: zero; += alpha_n*U_n for alpha_n != 0; += beta*(...). */
/* Same-level ghost fill (xfer::RefineSchedule::fillData, RungeKuttaLevelIntegrator.cpp:1568/1701) fused into the
 * producer: the new values of a cell within the ghost width of a patch face go straight into the ghost cells of the
 * neighbouring patches (all equal-sized boxes of a regular decomposition; neighbour at offset o sees my cell c at
 * c - o * n). */
/* table: the push addresses (DirArgs::push), staged in shared memory by the kernel (a cell near an x face is 4
 * lanes of a warp: the dependent global loads of the table made the whole warp wait, +17 % on the last sweep at 256^3) */
template <class Tr>
HB2_HD void push_cell(const Geom& G, double* const* table, int ci, int cj, int ck, const double (&v)[Tr::NCOMP])
{
    constexpr int DIM = Tr::DIM, NCOMP = Tr::NCOMP;
    const int c[3] = {ci, cj, ck};
    bool lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        lo[a] = (a < DIM) && c[a] < HB2_G;
        hi[a] = (a < DIM) && c[a] >= G.n[a] - HB2_G;
    }
    for (int oz = (lo[2] ? -1 : 0); oz <= (hi[2] ? 1 : 0); oz++)
        for (int oy = (lo[1] ? -1 : 0); oy <= (hi[1] ? 1 : 0); oy++)
            for (int ox = (lo[0] ? -1 : 0); ox <= (hi[0] ? 1 : 0); ox++) {
                if (!(ox | oy | oz)) continue;
                double* const* T = table + ((ox + 1) + 3 * (oy + 1) + 9 * (oz + 1)) * NCOMP;
                if (!T[0]) continue;
                const long long x = cidx(G, ci - ox * G.n[0], cj - oy * G.n[1], ck - oz * G.n[2]);
#pragma unroll
                for (int q = 0; q < NCOMP; q++) T[q][x] = v[q];
            }
}

/* true for cells within the ghost width of a patch face */
template <class Tr>
HB2_HD bool near_face(const Geom& G, int ci, int cj, int ck)
{
    /* signed comparisons: boxes narrower than two ghost widths (4 < n < 8) have cells near both faces */
    bool near = ci < HB2_G || ci >= G.n[0] - HB2_G || cj < HB2_G || cj >= G.n[1] - HB2_G;
    if (Tr::DIM == 3) near = near || ck < HB2_G || ck >= G.n[2] - HB2_G;
    return near;
}

template <class Tr>
HB2_HD void rk_update_cell(const DirArgs& A, double* const* push_table, long long x, int ci, int cj, int ck,
                           const double (&ua)[Tr::NEQ], const double (&rhs)[Tr::NEQ])
{
    constexpr int NEQ = Tr::NEQ, NS = Tr::NS, DIM = Tr::DIM;
    double Unew[Tr::NCOMP];
#pragma unroll
    for (int e = 0; e < NEQ; e++) {
        double u = ua[e]; /* sum_n alpha_n U_n, accumulated in the reference's order by the caller */
        u += A.beta * rhs[e];
        Unew[e] = u;
        store_stream(A.Uout[e] + x, u);
    }
    if (Tr::MODEL == FE) {
        double zl = 1.0;
#pragma unroll
        for (int si = 0; si < NS - 1; si++) zl -= Unew[NS + DIM + 1 + si];
        A.Uout[NEQ][x] = zl;
        Unew[Tr::NCOMP - 1] = zl;
    }
    if (A.push && near_face<Tr>(A.G, ci, cj, ck)) push_cell<Tr>(A.G, push_table, ci, cj, ck, Unew);
}

/* Shock-sensor decision of the face between cells L and R (ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2088-2123):
 * s = -theta_avg/(|theta_avg| + Omega_avg + eps); HLLC-HLL iff s > 0.65. */
HB2_HD bool face_sensor(double th_L, double th_R, double Om_L, double Om_R)
{
    const double theta_avg = 0.5 * (th_L + th_R);
    const double Omega_avg = 0.5 * (Om_L + Om_R);
    const double s = -theta_avg / (fabs(theta_avg) + Omega_avg + HB2_EPS);
    return s > HB2_SENSOR_THRESHOLD;
}

}  // namespace hb2
