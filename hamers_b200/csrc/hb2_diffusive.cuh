/*
 * hb2_diffusive.cuh -- SURVEY.md row f4: node-based sixth-order diffusive (viscous) flux of the single-species
 * Navier-Stokes application, per-thread arithmetic.
 *
 * `__host__ __device__` like hb2_core.cuh: inlined into the sm_100a kernels of hb2_diffusive.cu and compiled by g++
 * into the test-only host emulation (tests/host_emu/emu_diffusive.cpp).  Reference operation order throughout (the
 * translation unit is built with -fmad=false): results are bit-identical to oracle/oracle_diffusive.c.
 *
 * Reference behaviour restated here (path:line under the reference tree):
 *   driver            src/flow/diffusive_flux_reconstructors/node/DiffusiveFluxReconstructorNode.cpp:31-1736
 *   derivative        .../node/DiffusiveFluxReconstructorNodeSixthOrder.cpp:65-505   (a_n, b_n, c_n; times 1/dx)
 *   reconstruction    .../node/DiffusiveFluxReconstructorNodeSixthOrder.cpp:507-939  (a_r, b_r, c_r; times dt, "+=" on 0)
 *   terms             src/flow/flow_models/single-species/FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:654-2363
 *   diffusivities     ...FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:4180-4193 (2-D), 4262-4280 (3-D)
 *   T, kappa          EquationOfStateIdealGas.cpp:6897; EquationOfThermalConductivityPrandtl.cpp:309
 *   RK update         src/apps/Navier-Stokes/NavierStokes.cpp:1715-1751 (2-D), 2085-2092 (3-D)
 *
 * Data layout: cell data on the ghost box of width 6 (d_num_diff_ghosts), x fastest; the primitive scratch (velocity
 * components, temperature) and the node-flux scratch use the same box; side fluxes on ghost 0 like the convective ones.
 */
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define HB2D_HD __host__ __device__ __forceinline__
#else
#define HB2D_HD inline
#endif

#define HB2_DIFF_G 6 /* DiffusiveFluxReconstructorNodeSixthOrder.cpp:24 */
#ifndef HB2_MAXS
#define HB2_MAXS 4   /* RK coefficients per stage row, as in hb2_core.cuh */
#endif

namespace hb2 {

struct DiffGeom {
    int dim;
    int n[3];          /* interior cells (n[2] = 1 in 2-D) */
    int g[3];          /* 6, 6, 6 (0 in the unused direction of 2-D) */
    int gd[3];         /* ghost-box dims */
    long long cs[3];   /* cell strides in the ghost box */
    long long ncell_g;
    double dx_inv[3];  /* double(1)/dx[d], DiffusiveFluxReconstructorNode.cpp:1762 */
    double dx[3];
};

struct DiffConsts {
    double gamma, c_v;     /* ideal gas */
    double mu, mu_v;       /* CONSTANT shear / bulk viscosity */
    double kappa;          /* c_p*mu/Pr */
};

/* one term of a node flux: derivative of variable `var` (velocity component, or DIM for the temperature) in one
 * direction times the diffusivity D[diff] */
struct DiffTerm {
    signed char var, diff;
};
struct DiffTermList {
    int n;
    DiffTerm t[4];
};

/* terms[flux direction][derivative direction][equation], in the reference's order of accumulation
 * (FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:978-1503 and 1838-2363) */
template <int DIM>
struct DiffTerms;

template <>
struct DiffTerms<3> {
    static constexpr int NEQ = 5, ND = 13, T = 3;
    HB2D_HD static constexpr DiffTermList get(int f, int d, int e)
    {
        constexpr DiffTermList none = {0, {{0, 0}, {0, 0}, {0, 0}, {0, 0}}};
        if (e == 0) return none;
        if (f == 0) {
            if (d == 0) return e == 1 ? DiffTermList{1, {{0, 0}}} : e == 2 ? DiffTermList{1, {{1, 2}}} : e == 3 ? DiffTermList{1, {{2, 2}}}
                                                                  : DiffTermList{4, {{0, 3}, {1, 10}, {2, 11}, {T, 12}}};
            if (d == 1) return e == 1 ? DiffTermList{1, {{1, 1}}} : e == 2 ? DiffTermList{1, {{0, 2}}} : e == 3 ? none
                                                                  : DiffTermList{2, {{0, 10}, {1, 6}}};
            return e == 1 ? DiffTermList{1, {{2, 1}}} : e == 2 ? none : e == 3 ? DiffTermList{1, {{0, 2}}}
                                                                  : DiffTermList{2, {{0, 11}, {2, 6}}};
        }
        if (f == 1) {
            if (d == 0) return e == 1 ? DiffTermList{1, {{1, 2}}} : e == 2 ? DiffTermList{1, {{0, 1}}} : e == 3 ? none
                                                                  : DiffTermList{2, {{0, 7}, {1, 9}}};
            if (d == 1) return e == 1 ? DiffTermList{1, {{0, 2}}} : e == 2 ? DiffTermList{1, {{1, 0}}} : e == 3 ? DiffTermList{1, {{2, 2}}}
                                                                  : DiffTermList{4, {{0, 9}, {1, 4}, {2, 11}, {T, 12}}};
            return e == 1 ? none : e == 2 ? DiffTermList{1, {{2, 1}}} : e == 3 ? DiffTermList{1, {{1, 2}}}
                                                                  : DiffTermList{2, {{1, 11}, {2, 7}}};
        }
        if (d == 0) return e == 1 ? DiffTermList{1, {{2, 2}}} : e == 2 ? none : e == 3 ? DiffTermList{1, {{0, 1}}}
                                                              : DiffTermList{2, {{0, 8}, {2, 9}}};
        if (d == 1) return e == 1 ? none : e == 2 ? DiffTermList{1, {{2, 2}}} : e == 3 ? DiffTermList{1, {{1, 1}}}
                                                              : DiffTermList{2, {{1, 8}, {2, 10}}};
        return e == 1 ? DiffTermList{1, {{0, 2}}} : e == 2 ? DiffTermList{1, {{1, 2}}} : e == 3 ? DiffTermList{1, {{2, 0}}}
                                                              : DiffTermList{4, {{0, 9}, {1, 10}, {2, 5}, {T, 12}}};
    }
};

template <>
struct DiffTerms<2> {
    static constexpr int NEQ = 4, ND = 10, T = 2;
    HB2D_HD static constexpr DiffTermList get(int f, int d, int e)
    {
        constexpr DiffTermList none = {0, {{0, 0}, {0, 0}, {0, 0}, {0, 0}}};
        if (e == 0) return none;
        if (f == 0) {
            if (d == 0) return e == 1 ? DiffTermList{1, {{0, 0}}} : e == 2 ? DiffTermList{1, {{1, 2}}}
                                                                  : DiffTermList{3, {{0, 3}, {1, 8}, {T, 9}}};
            return e == 1 ? DiffTermList{1, {{1, 1}}} : e == 2 ? DiffTermList{1, {{0, 2}}} : DiffTermList{2, {{0, 8}, {1, 5}}};
        }
        if (d == 0) return e == 1 ? DiffTermList{1, {{1, 2}}} : e == 2 ? DiffTermList{1, {{0, 1}}} : DiffTermList{2, {{0, 6}, {1, 7}}};
        return e == 1 ? DiffTermList{1, {{0, 2}}} : e == 2 ? DiffTermList{1, {{1, 0}}} : DiffTermList{3, {{0, 7}, {1, 4}, {T, 9}}};
    }
};

/* velocity and temperature of one cell from its conservative variables (FlowModelSingleSpecies.cpp:2824-2826, 3049-3051;
 * EquationOfStateIdealGas.cpp:5580, 6897) */
template <int DIM>
HB2D_HD void diff_primitives(const double (&Q)[DIM + 2], const DiffConsts& K, double (&P)[DIM + 1])
{
    const double rho = Q[0];
    double ke = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
        P[a] = Q[1 + a] / rho;
        ke = (a == 0) ? P[a] * P[a] : ke + P[a] * P[a];
    }
    const double epsilon = Q[DIM + 1] / rho - 1.0 / 2.0 * ke;
    const double p = (K.gamma - 1.0) * rho * epsilon;
    P[DIM] = p / ((K.gamma - 1.0) * K.c_v * rho);
}

/* the diffusivities of one cell (FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:4180-4193, 4262-4280) */
template <int DIM>
HB2D_HD void diff_diffusivities(const double* vel, const DiffConsts& K, double (&D)[DiffTerms<DIM>::ND])
{
    const double mu = K.mu, mu_v = K.mu_v;
    const double u = vel[0], v = vel[1];
    D[0] = -(4.0 / 3.0 * mu + mu_v);
    D[1] = 2.0 / 3.0 * mu - mu_v;
    D[2] = -mu;
    if constexpr (DIM == 2) {
        D[3] = -u * (4.0 / 3.0 * mu + mu_v);
        D[4] = -v * (4.0 / 3.0 * mu + mu_v);
        D[5] = u * (2.0 / 3.0 * mu - mu_v);
        D[6] = v * (2.0 / 3.0 * mu - mu_v);
        D[7] = -u * mu;
        D[8] = -v * mu;
        D[9] = -K.kappa;
    } else {
        const double w = vel[DIM - 1];
        D[3] = -u * (4.0 / 3.0 * mu + mu_v);
        D[4] = -v * (4.0 / 3.0 * mu + mu_v);
        D[5] = -w * (4.0 / 3.0 * mu + mu_v);
        D[6] = u * (2.0 / 3.0 * mu - mu_v);
        D[7] = v * (2.0 / 3.0 * mu - mu_v);
        D[8] = w * (2.0 / 3.0 * mu - mu_v);
        D[9] = -u * mu;
        D[10] = -v * mu;
        D[11] = -w * mu;
        D[12] = -K.kappa;
    }
}

/* DiffusiveFluxReconstructorNodeSixthOrder.cpp:236-239: sixth-order first derivative at a node, p points at the node */
HB2D_HD double diff_first_derivative(const double* p, long long stride, double dx_inv)
{
    const double a_n = 3.0 / 4.0;
    const double b_n = -(3.0 / 20.0);
    const double c_n = 1.0 / 60.0;
    return (a_n * (p[stride] - p[-stride]) + b_n * (p[2 * stride] - p[-2 * stride]) + c_n * (p[3 * stride] - p[-3 * stride])) *
           dx_inv;
}

/* :679-683: face flux from the six nodes LLL..RRR around the face; p points at node R (the cell on the high side) */
HB2D_HD double diff_reconstruct(const double* p, long long stride, double dt)
{
    const double a_n = 3.0 / 4.0;
    const double b_n = -(3.0 / 20.0);
    const double c_n = 1.0 / 60.0;
    const double a_r = a_n + b_n + c_n;
    const double b_r = b_n + c_n;
    const double c_r = c_n;
    double F = 0.0;                              /* diffusive_flux->fillAll(0), then "+=" */
    F += dt * (a_r * (p[-stride] + p[0]) + b_r * (p[-2 * stride] + p[stride]) + c_r * (p[-3 * stride] + p[2 * stride]));
    return F;
}

/* The node fluxes of ALL flux directions at ghost-box cell x in one pass (DiffusiveFluxReconstructorNode.cpp:888-1055 and
 * the y / z copies: per direction and equation zero, then "+=" term by term, x-, y-, then z-derivative terms).  The three
 * directions draw on the same twelve (2-D: six) first derivatives, so each is evaluated once -- like the reference's
 * derivatives_*_computed maps do -- and used up to three times.  P[v]: primitive scratch arrays on the ghost box. */
/* node fluxes of all directions from the velocity and the (DIM + 1) x DIM first derivatives of the node */
template <int DIM>
HB2D_HD void diff_node_flux_from_derivatives(const DiffConsts& K, const double (&vel)[3], const double (&der)[DIM + 1][DIM],
                                             double (&Fn)[DIM][DIM + 2])
{
    using TT = DiffTerms<DIM>;
    double D[TT::ND];
    diff_diffusivities<DIM>(vel, K, D);
#pragma unroll
    for (int f = 0; f < DIM; f++) {
#pragma unroll
        for (int e = 0; e < DIM + 2; e++) {
            double acc = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                const DiffTermList tl = TT::get(f, d, e);
#pragma unroll
                for (int ti = 0; ti < 4; ti++)
                    if (ti < tl.n) acc += D[tl.t[ti].diff] * der[tl.t[ti].var][d];
            }
            Fn[f][e] = acc;
        }
    }
}

template <int DIM>
HB2D_HD void diff_node_flux_all(const DiffGeom& G, const DiffConsts& K, const double* const* P, long long x,
                                double (&Fn)[DIM][DIM + 2])
{
    double vel[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < DIM; a++) vel[a] = P[a][x];
    double der[DIM + 1][DIM];
#pragma unroll
    for (int v = 0; v < DIM + 1; v++) {
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            /* dT/dx_d enters the energy flux of direction d only; every velocity derivative is used by some direction */
            der[v][d] = diff_first_derivative(P[v] + x, G.cs[d], G.dx_inv[d]);
        }
    }
    diff_node_flux_from_derivatives<DIM>(K, vel, der, Fn);
}

/* the same derivative from six values (tiled kernels: register ring along the marching direction) */
HB2D_HD double diff_first_derivative6(double m3, double m2, double m1, double p1, double p2, double p3, double dx_inv)
{
    const double a_n = 3.0 / 4.0;
    const double b_n = -(3.0 / 20.0);
    const double c_n = 1.0 / 60.0;
    return (a_n * (p1 - m1) + b_n * (p2 - m2) + c_n * (p3 - m3)) * dx_inv;
}

/* face flux from six node values LLL, LL, L, R, RR, RRR (diff_reconstruct with p[-3..2]) */
HB2D_HD double diff_reconstruct6(double lll, double ll, double l, double r, double rr, double rrr, double dt)
{
    const double a_n = 3.0 / 4.0;
    const double b_n = -(3.0 / 20.0);
    const double c_n = 1.0 / 60.0;
    const double a_r = a_n + b_n + c_n;
    const double b_r = b_n + c_n;
    const double c_r = c_n;
    double F = 0.0;
    F += dt * (a_r * (l + r) + b_r * (ll + rr) + c_r * (lll + rrr));
    return F;
}

/* ---- midpoint family: DiffusiveFluxReconstructorMidpointSixthOrder ("MIDPOINT_SIXTH_ORDER"; no shipped deck selects it) ----
 *   driver   src/flow/diffusive_flux_reconstructors/midpoint/DiffusiveFluxReconstructorMidpoint.cpp:38-2330
 *   kernels  .../midpoint/DiffusiveFluxReconstructorMidpointSixthOrder.cpp:68-1799 (reads the cells within 5 of the interior)
 *   side diffusivities  FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:2365-2797, 2799-3657
 * The flux is formed AT THE MIDPOINTS (faces) of the flux direction f: derivatives along f by a staggered sixth-order
 * difference of the node values, derivatives along the other directions by sixth-order node derivatives interpolated along f,
 * diffusivities from mu, mu_v, kappa and the velocity interpolated along f; a face flux is a five-midpoint combination,
 * times dt.  The reference stages every intermediate in its own patch-sized array; here one thread forms the midpoint flux
 * of all equations from the primitive scratch (every value is the same function of the same inputs: bit-identical). */

/* DiffusiveFluxReconstructorMidpointSixthOrder.cpp:80-82, 219-222: p points at node R (the cell on the high side) */
HB2D_HD double diff_mid_derivative(const double* p, long long s, double dx_inv)
{
    const double a_n = 75.0 / 64.0;
    const double b_n = -(25.0 / 384.0);
    const double c_n = 3.0 / 640.0;
    return (a_n * (p[0] - p[-s]) + b_n * (p[s] - p[-2 * s]) + c_n * (p[2 * s] - p[-3 * s])) * dx_inv;
}

/* :962-964, 1101-1104: six values R, L, RR, LL, RRR, LLL in the order the reference adds them */
HB2D_HD double diff_mid_interpolate6(double r, double l, double rr, double ll, double rrr, double lll)
{
    const double a_n = 75.0 / 128.0;
    const double b_n = -(25.0 / 256.0);
    const double c_n = 3.0 / 256.0;
    return (a_n * (r + l) + b_n * (rr + ll) + c_n * (rrr + lll));
}
HB2D_HD double diff_mid_interpolate(const double* p, long long s)
{
    return diff_mid_interpolate6(p[0], p[-s], p[s], p[-2 * s], p[2 * s], p[-3 * s]);
}

/* :1400-1406, 1552-1556: face value from the midpoint fluxes LL, L, own, R, RR; p points at the face's own midpoint */
HB2D_HD double diff_mid_reconstruct(const double* p, long long s, double dt)
{
    const double a_m = 75.0 / 64.0;
    const double b_m = -(25.0 / 384.0);
    const double c_m = 3.0 / 640.0;
    const double a_r = a_m + b_m + c_m;
    const double b_r = b_m + c_m;
    const double c_r = c_m;
    double F = 0.0;                              /* diffusive_flux->fillAll(0), then "+=" */
    F += dt * (a_r * (p[0]) + b_r * (p[-s] + p[s]) + c_r * (p[-2 * s] + p[2 * s]));
    return F;
}

/* side diffusivities of direction FDIR (FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:2682-2691, 2725-2734, 2768-2777;
 * 2-D :2581-2589, 2615-2623) from the interpolated coefficients and velocity */
template <int DIM, int FDIR>
HB2D_HD void diff_side_diffusivities(double mu, double mu_v, double kappa, const double (&vel)[3], double (&D)[8])
{
    const double un = vel[FDIR];
    D[0] = -(4.0 / 3.0 * mu + mu_v);
    D[1] = 2.0 / 3.0 * mu - mu_v;
    D[2] = -mu;
    D[3] = -un * (4.0 / 3.0 * mu + mu_v);
    D[4] = un * (2.0 / 3.0 * mu - mu_v);
    int m = 5;
#pragma unroll
    for (int a = 0; a < DIM; a++)
        if (a != FDIR) D[m++] = -vel[a] * mu;
    D[m] = -kappa;
    if (DIM == 2) D[7] = 0.0;
}

/* index of the side diffusivity of term ti of (flux direction, derivative direction, equation); the variables are those of
 * DiffTerms (getCellDataOfDiffusiveFluxVariablesForDerivative serves both families); :2799-3657 */
template <int DIM>
HB2D_HD constexpr int diff_side_index(int f, int d, int e, int ti)
{
    if (DIM == 3) {
        constexpr signed char T[3][3][5][4] = {
            {{{-1, -1, -1, -1}, {0, -1, -1, -1}, {2, -1, -1, -1}, {2, -1, -1, -1}, {3, 5, 6, 7}},
             {{-1, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {-1, -1, -1, -1}, {5, 4, -1, -1}},
             {{-1, -1, -1, -1}, {1, -1, -1, -1}, {-1, -1, -1, -1}, {2, -1, -1, -1}, {6, 4, -1, -1}}},
            {{{-1, -1, -1, -1}, {2, -1, -1, -1}, {1, -1, -1, -1}, {-1, -1, -1, -1}, {4, 5, -1, -1}},
             {{-1, -1, -1, -1}, {2, -1, -1, -1}, {0, -1, -1, -1}, {2, -1, -1, -1}, {5, 3, 6, 7}},
             {{-1, -1, -1, -1}, {-1, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {6, 4, -1, -1}}},
            {{{-1, -1, -1, -1}, {2, -1, -1, -1}, {-1, -1, -1, -1}, {1, -1, -1, -1}, {4, 5, -1, -1}},
             {{-1, -1, -1, -1}, {-1, -1, -1, -1}, {2, -1, -1, -1}, {1, -1, -1, -1}, {4, 6, -1, -1}},
             {{-1, -1, -1, -1}, {2, -1, -1, -1}, {2, -1, -1, -1}, {0, -1, -1, -1}, {5, 6, 3, 7}}}};
        return T[f][d][e][ti];
    } else {
        constexpr signed char T[2][2][4][4] = {
            {{{-1, -1, -1, -1}, {0, -1, -1, -1}, {2, -1, -1, -1}, {3, 5, 6, -1}},
             {{-1, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {5, 4, -1, -1}}},
            {{{-1, -1, -1, -1}, {2, -1, -1, -1}, {1, -1, -1, -1}, {4, 5, -1, -1}},
             {{-1, -1, -1, -1}, {2, -1, -1, -1}, {0, -1, -1, -1}, {5, 3, 6, -1}}}};
        return T[f][d][e][ti];
    }
}

/* does the flux of direction f use the derivative of variable v in direction d? */
template <int DIM>
HB2D_HD constexpr bool diff_mid_uses(int f, int v, int d)
{
    for (int e = 1; e < DIM + 2; e++) {
        const DiffTermList tl = DiffTerms<DIM>::get(f, d, e);
        for (int ti = 0; ti < tl.n; ti++)
            if (tl.t[ti].var == v) return true;
    }
    return false;
}

struct DiffMidPtrs {
    const double* P[4];      /* velocity components, temperature on the ghost box */
    double* Fm[5];           /* midpoint flux of the current direction: ghost-box layout, midpoint i (the face between cells
                                i - 1 and i) at cell index i; equation 0 unused */
    double* F[5];            /* side flux of the current direction (ghost 0) */
};

template <int DIM, int FDIR>
HB2D_HD long long diff_mid_count(const DiffGeom& G)
{
    return (long long)(G.n[0] + (FDIR == 0 ? 5 : 0)) * (G.n[1] + (FDIR == 1 ? 5 : 0)) * (G.n[2] + (FDIR == 2 ? 5 : 0));
}

/* midpoint t of direction FDIR: midpoints -2 .. n + 2 along FDIR, interior cells of the other directions
 * (DiffusiveFluxReconstructorMidpoint.cpp:1466-1470) */
template <int DIM, int FDIR>
HB2D_HD void diff_mid_flux_thread(const DiffGeom& G, const DiffConsts& K, const DiffMidPtrs& A, long long t)
{
    using TT = DiffTerms<DIM>;
    const int e0 = G.n[0] + (FDIR == 0 ? 5 : 0), e1 = G.n[1] + (FDIR == 1 ? 5 : 0);
    const int i = (int)(t % e0) - (FDIR == 0 ? 2 : 0), j = (int)((t / e0) % e1) - (FDIR == 1 ? 2 : 0);
    const int k = (int)(t / ((long long)e0 * e1)) - (FDIR == 2 ? 2 : 0);
    const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
    const long long s = G.cs[FDIR];
    /* the transport coefficients are cell data like any other: interpolated, not copied (the weights do not sum to one in
     * floating point) */
    const double mu_m = diff_mid_interpolate6(K.mu, K.mu, K.mu, K.mu, K.mu, K.mu);
    const double mu_v_m = diff_mid_interpolate6(K.mu_v, K.mu_v, K.mu_v, K.mu_v, K.mu_v, K.mu_v);
    const double kappa_m = diff_mid_interpolate6(K.kappa, K.kappa, K.kappa, K.kappa, K.kappa, K.kappa);
    double vel[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < DIM; a++) vel[a] = diff_mid_interpolate(A.P[a] + x, s);
    double D[8];
    diff_side_diffusivities<DIM, FDIR>(mu_m, mu_v_m, kappa_m, vel, D);
    double der[DIM + 1][DIM];
#pragma unroll
    for (int v = 0; v < DIM + 1; v++) {
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            der[v][d] = 0.0;
            if (!diff_mid_uses<DIM>(FDIR, v, d)) continue;
            const double* p = A.P[v] + x;
            if (d == FDIR) {
                der[v][d] = diff_mid_derivative(p, s, G.dx_inv[d]);
            } else {
                /* node derivatives in direction d at the six nodes along FDIR, then the interpolation to the midpoint */
                const long long sd = G.cs[d];
                const double inv = G.dx_inv[d];
                der[v][d] = diff_mid_interpolate6(diff_first_derivative(p, sd, inv), diff_first_derivative(p - s, sd, inv),
                                                  diff_first_derivative(p + s, sd, inv), diff_first_derivative(p - 2 * s, sd, inv),
                                                  diff_first_derivative(p + 2 * s, sd, inv), diff_first_derivative(p - 3 * s, sd, inv));
            }
        }
    }
#pragma unroll
    for (int e = 1; e < DIM + 2; e++) {
        double acc = 0.0;                        /* fillAll(0), then "+=" term by term: x-, y-, z-derivative terms */
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const DiffTermList tl = TT::get(FDIR, d, e);
#pragma unroll
            for (int ti = 0; ti < 4; ti++)
                if (ti < tl.n) acc += D[diff_side_index<DIM>(FDIR, d, e, ti)] * der[tl.t[ti].var][d];
        }
        A.Fm[e][x] = acc;
    }
}

/* face t of direction FDIR (side-data order) from the five midpoint fluxes around it */
template <int DIM, int FDIR>
HB2D_HD void diff_mid_face_thread(const DiffGeom& G, const DiffMidPtrs& A, double dt, long long t)
{
    const int f0 = G.n[0] + (FDIR == 0), f1 = G.n[1] + (FDIR == 1);
    const int i = (int)(t % f0), j = (int)((t / f0) % f1), k = (int)(t / ((long long)f0 * f1));
    const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
    A.F[0][t] = 0.0;                             /* the midpoint flux of the continuity equation is the fillAll(0) */
#pragma unroll
    for (int e = 1; e < DIM + 2; e++) A.F[e][t] = diff_mid_reconstruct(A.Fm[e] + x, G.cs[FDIR], dt);
}

/* ---- re-associated arithmetic of the flux-free route (HB2_MATH_FAST; hb2_diffusive_plan_set_math): the same terms with
 * fewer FP64 instructions -- one reciprocal per cell instead of five divisions (T = epsilon/c_v: the (gamma - 1) rho of the
 * pressure cancels), derivative and reconstruction coefficients pre-multiplied by 1/dx and by -beta dt/dx, explicit FMAs,
 * the energy flux assembled from the momentum fluxes of the same direction (u . tau_f + kappa dT/dx_f; the tables of the
 * reference expand it into eight products per direction), and the two faces of a cell differenced analytically:
 *   F_R - F_L = dt (a_n (p[1] - p[-1]) + b_n (p[2] - p[-2]) + c_n (p[3] - p[-3]))
 * (a_r - b_r = a_n, b_r - c_r = b_n, c_r = c_n: the face difference of the reconstruction IS the sixth-order derivative).
 * 3-D: ~130 FP64 instructions per node instead of ~300, ~75 per cell update instead of ~335; <= 1e-12 relative. ---- */
struct DiffFast {
    double inv_cv;         /* 1/c_v */
    double cd[3][3];       /* {a_n, b_n, c_n} / dx_d */
    double kd[3][3];       /* -beta dt {a_n, b_n, c_n} / dx_d */
    double D0, D1, D2;     /* -(4/3 mu + mu_v), 2/3 mu - mu_v, -mu */
    double mkappa;         /* -kappa */
};

inline void make_diff_fast(const DiffGeom& G, const DiffConsts& K, double dt, double beta, DiffFast* F)
{
    const double a_n = 3.0 / 4.0, b_n = -(3.0 / 20.0), c_n = 1.0 / 60.0;
    F->inv_cv = 1.0 / K.c_v;
    for (int d = 0; d < 3; d++) {
        F->cd[d][0] = a_n * G.dx_inv[d];
        F->cd[d][1] = b_n * G.dx_inv[d];
        F->cd[d][2] = c_n * G.dx_inv[d];
        const double s = -beta * dt / G.dx[d];
        F->kd[d][0] = s * a_n;
        F->kd[d][1] = s * b_n;
        F->kd[d][2] = s * c_n;
    }
    F->D0 = -(4.0 / 3.0 * K.mu + K.mu_v);
    F->D1 = 2.0 / 3.0 * K.mu - K.mu_v;
    F->D2 = -K.mu;
    F->mkappa = -K.kappa;
}

HB2D_HD double diff_rcp_fast(double x)
{
#if defined(__CUDA_ARCH__)
    /* MUFU.RCP64H seed (relative error e0 ~ 2^-22) + one third-order step r0 (1 + e + e^2): ~1 ulp, 3 FP64 instructions */
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#else
    return 1.0 / x;
#endif
}

HB2D_HD void diff_primitives_fast(const double (&Q)[5], const DiffFast& F, double (&P)[4])
{
    const double r = diff_rcp_fast(Q[0]);
    P[0] = Q[1] * r;
    P[1] = Q[2] * r;
    P[2] = Q[3] * r;
    const double ke = fma(P[2], P[2], fma(P[1], P[1], P[0] * P[0]));
    P[3] = fma(Q[4], r, -0.5 * ke) * F.inv_cv;
}

HB2D_HD double diff_first_derivative6_fast(double m3, double m2, double m1, double p1, double p2, double p3, const double (&c)[3])
{
    return fma(c[2], p3 - m3, fma(c[1], p2 - m2, c[0] * (p1 - m1)));
}

/* node fluxes of the three directions (momentum and energy rows) from the velocity and the twelve derivatives
 * der[variable][direction]: F_f = -(tau e_f, u . tau e_f + kappa dT/dx_f) */
HB2D_HD void diff_node_flux_fast(const DiffFast& F, const double (&vel)[3], const double (&der)[4][3], double (&Fn)[3][5])
{
    const double div = der[0][0] + der[1][1] + der[2][2];
    const double dm = F.D0 - F.D1;                      /* D0 d_f u_f + D1 (div - d_f u_f) */
#pragma unroll
    for (int f = 0; f < 3; f++) {
        Fn[f][0] = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++)
            Fn[f][1 + a] = (a == f) ? fma(dm, der[f][f], F.D1 * div) : F.D2 * (der[a][f] + der[f][a]);
        Fn[f][4] = fma(vel[0], Fn[f][1], fma(vel[1], Fn[f][2], fma(vel[2], Fn[f][3], F.mkappa * der[3][f])));
    }
}

/* contribution of direction d to the update of one component: sum_m kd[m] (p[m] - p[-m]) added to acc */
HB2D_HD double diff_divergence_fast(double acc, double m3, double m2, double m1, double p1, double p2, double p3, const double (&k)[3])
{
    return fma(k[2], p3 - m3, fma(k[1], p2 - m2, fma(k[0], p1 - m1, acc)));
}

/* ---- one thread of each kernel (the kernels of hb2_diffusive.cu are grid-stride loops over these; the host emulation
 * calls them from plain loops) ---- */
struct DiffPtrs {
    const double* Q[5];
    double* P[4];
    double* Fn[5];
    double* F[5];
};

template <int DIM>
HB2D_HD void diff_primitives_thread(const DiffConsts& K, const DiffPtrs& A, long long x)
{
    double Q[DIM + 2], P[DIM + 1];
#pragma unroll
    for (int c = 0; c < DIM + 2; c++) Q[c] = A.Q[c][x];
    diff_primitives<DIM>(Q, K, P);
#pragma unroll
    for (int v = 0; v < DIM + 1; v++) A.P[v][x] = P[v];
}

struct DiffAllPtrs {
    const double* P[4];
    double* Fn[3][5];
};

template <int DIM>
HB2D_HD long long diff_node_all_count(const DiffGeom& G)
{
    return (long long)(G.n[0] + 6) * (G.n[1] + 6) * (G.n[2] + (DIM == 3 ? 6 : 0));
}

/* node t of the interior extended by 3 cells in EVERY direction (the union of the three per-direction node sets plus
 * the edge / corner nodes between them, which no face reads: (n + 6)^3 nodes instead of 3 n^2 (n + 6)) */
template <int DIM>
HB2D_HD void diff_node_all_thread(const DiffGeom& G, const DiffConsts& K, const DiffAllPtrs& A, long long t)
{
    const int e0 = G.n[0] + 6, e1 = G.n[1] + 6;
    const int i = (int)(t % e0) - 3, j = (int)((t / e0) % e1) - 3;
    const int k = (int)(t / ((long long)e0 * e1)) - (DIM == 3 ? 3 : 0);
    const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
    const double* P[DIM + 1];
#pragma unroll
    for (int v = 0; v < DIM + 1; v++) P[v] = A.P[v];
    double Fn[DIM][DIM + 2];
    diff_node_flux_all<DIM>(G, K, P, x, Fn);
#pragma unroll
    for (int f = 0; f < DIM; f++) {
#pragma unroll
        for (int e = 1; e < DIM + 2; e++) A.Fn[f][e][x] = Fn[f][e];
    }
}

template <int DIM, int FDIR>
HB2D_HD long long diff_face_count(const DiffGeom& G)
{
    return (long long)(G.n[0] + (FDIR == 0)) * (G.n[1] + (FDIR == 1)) * (G.n[2] + (FDIR == 2));
}

/* face t of direction FDIR (side-data order) */
template <int DIM, int FDIR>
HB2D_HD void diff_face_thread(const DiffGeom& G, const DiffPtrs& A, double dt, long long t)
{
    const int f0 = G.n[0] + (FDIR == 0), f1 = G.n[1] + (FDIR == 1);
    const int i = (int)(t % f0), j = (int)((t / f0) % f1), k = (int)(t / ((long long)f0 * f1));
    const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
    A.F[0][t] = 0.0;                             /* no diffusive mass flux: the fillAll(0) of the reference */
#pragma unroll
    for (int e = 1; e < DIM + 2; e++) A.F[e][t] = diff_reconstruct(A.Fn[e] + x, G.cs[FDIR], dt);
}

struct NsArgs {
    DiffGeom G;       /* geometry of the STATE arrays: their ghost width need not be 6 */
    int neq, ncoef;
    double alpha[HB2_MAXS], beta[HB2_MAXS];
    const double* U[HB2_MAXS][5];
    const double* Fc[HB2_MAXS][15];
    const double* Fd[HB2_MAXS][15];
    const double* S[HB2_MAXS][5];
    double* Uout[5];
};

/* NavierStokes::advanceSingleStepOnPatch, conservative diffusive form (NavierStokes.cpp:1715-1751, 2085-2092), interior
 * cell t */
template <int DIM>
HB2D_HD void advance_ns_thread(const NsArgs& A, long long t)
{
    const DiffGeom& G = A.G;
    const int i = (int)(t % G.n[0]), j = (int)((t / G.n[0]) % G.n[1]), k = (int)(t / ((long long)G.n[0] * G.n[1]));
    const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
    const long long fxL = i + (long long)(G.n[0] + 1) * (j + (long long)G.n[1] * k), fxR = fxL + 1;
    const long long fyB = i + (long long)G.n[0] * (j + (long long)(G.n[1] + 1) * k), fyT = fyB + G.n[0];
    const long long fzB = t, fzF = t + (long long)G.n[0] * G.n[1];
    for (int e = 0; e < A.neq; e++) {
        double Q = 0.0;
        for (int n = 0; n < A.ncoef; n++) {
            if (A.alpha[n] != 0.0) Q += A.alpha[n] * A.U[n][e][x];
            if (A.beta[n] != 0.0) {
                const double *Fcx = A.Fc[n][e], *Fcy = A.Fc[n][A.neq + e], *Fdx = A.Fd[n][e], *Fdy = A.Fd[n][A.neq + e];
                if (DIM == 2) {
                    Q += A.beta[n] * (-(Fcx[fxR] - Fcx[fxL] + Fdx[fxR] - Fdx[fxL]) / G.dx[0] -
                                      (Fcy[fyT] - Fcy[fyB] + Fdy[fyT] - Fdy[fyB]) / G.dx[1] + A.S[n][e][t]);
                } else {
                    const double *Fcz = A.Fc[n][2 * A.neq + e], *Fdz = A.Fd[n][2 * A.neq + e];
                    Q += A.beta[n] * (-(Fcx[fxR] - Fcx[fxL] + Fdx[fxR] - Fdx[fxL]) / G.dx[0] -
                                      (Fcy[fyT] - Fcy[fyB] + Fdy[fyT] - Fdy[fyB]) / G.dx[1] -
                                      (Fcz[fzF] - Fcz[fzB] + Fdz[fzF] - Fdz[fzB]) / G.dx[2] + A.S[n][e][t]);
                }
            }
        }
        A.Uout[e][x] = Q;
    }
}

/* Diffusive part of the stage update on top of a state that already holds sum alpha U + beta (-div F_c + S) (the fused
 * convective stage): U += beta (-(Fd_R - Fd_L)/dx_0 - ...).  Same terms as NavierStokes.cpp:2085-2092, associated
 * differently (the reference adds the two flux differences before dividing): <= a few ulp, used by the
 * re-associated (HB2_MATH_FAST) path only. */
struct NsAccArgs {
    DiffGeom G;       /* geometry of U */
    int neq;
    double beta;
    const double* Fd[15];
    double* U[5];
};

template <int DIM>
HB2D_HD void diff_accumulate_thread(const NsAccArgs& A, long long t)
{
    const DiffGeom& G = A.G;
    const int i = (int)(t % G.n[0]), j = (int)((t / G.n[0]) % G.n[1]), k = (int)(t / ((long long)G.n[0] * G.n[1]));
    const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
    const long long fxL = i + (long long)(G.n[0] + 1) * (j + (long long)G.n[1] * k), fxR = fxL + 1;
    const long long fyB = i + (long long)G.n[0] * (j + (long long)(G.n[1] + 1) * k), fyT = fyB + G.n[0];
    const long long fzB = t, fzF = t + (long long)G.n[0] * G.n[1];
    for (int e = 1; e < A.neq; e++) {            /* the continuity equation has no diffusive flux */
        const double *Fdx = A.Fd[e], *Fdy = A.Fd[A.neq + e];
        double div = -(Fdx[fxR] - Fdx[fxL]) / G.dx[0] - (Fdy[fyT] - Fdy[fyB]) / G.dx[1];
        if (DIM == 3) {
            const double* Fdz = A.Fd[2 * A.neq + e];
            div -= (Fdz[fzF] - Fdz[fzB]) / G.dx[2];
        }
        A.U[e][x] += A.beta * div;
    }
}

/* The same update without materialising the diffusive side flux: both faces of the cell are reconstructed from the node
 * fluxes of every direction (kept in three scratch sets) and differenced on the spot.  Every operation and its order are
 * those of diff_face_thread followed by diff_accumulate_thread: bit-identical to that route. */
struct NsDivArgs {
    DiffGeom G6;      /* geometry of the node-flux scratch (six ghosts) */
    DiffGeom GU;      /* geometry of U */
    int neq;
    double beta, dt;
    const double* Fn[3][5];
    double* U[5];
};

template <int DIM>
HB2D_HD void diff_divergence_accumulate_thread(const NsDivArgs& A, long long t)
{
    const DiffGeom &G = A.G6, &GU = A.GU;
    const int i = (int)(t % G.n[0]), j = (int)((t / G.n[0]) % G.n[1]), k = (int)(t / ((long long)G.n[0] * G.n[1]));
    const long long x6 = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
    const long long xu = (i + GU.g[0]) + GU.cs[1] * (j + GU.g[1]) + GU.cs[2] * (k + GU.g[2]);
    for (int e = 1; e < A.neq; e++) {
        /* face "L" of the cell is the face whose high-side cell is the cell itself, face "R" the next one */
        const double FxL = diff_reconstruct(A.Fn[0][e] + x6, G.cs[0], A.dt), FxR = diff_reconstruct(A.Fn[0][e] + x6 + G.cs[0], G.cs[0], A.dt);
        const double FyB = diff_reconstruct(A.Fn[1][e] + x6, G.cs[1], A.dt), FyT = diff_reconstruct(A.Fn[1][e] + x6 + G.cs[1], G.cs[1], A.dt);
        double div = -(FxR - FxL) / G.dx[0] - (FyT - FyB) / G.dx[1];
        if (DIM == 3) {
            const double FzB = diff_reconstruct(A.Fn[2][e] + x6, G.cs[2], A.dt), FzF = diff_reconstruct(A.Fn[2][e] + x6 + G.cs[2], G.cs[2], A.dt);
            div -= (FzF - FzB) / G.dx[2];
        }
        A.U[e][xu] += A.beta * div;
    }
}

/* Diffusive spectral radius of one cell (FlowModelSingleSpecies.cpp:4661-4665 MAX_DIFFUSIVITY; NavierStokes.cpp:884-886,
 * 1083-1086): 2 max_d D_max / dx_d^2 with D_max = max(mu/rho, mu_v/rho, kappa/(rho c_p)) */
template <int DIM>
HB2D_HD double diff_spectral_radius_cell(const DiffGeom& G, const DiffConsts& K, double c_p_eos, double rho)
{
    double D_max = fmax(K.mu / rho, K.mu_v / rho);
    D_max = fmax(D_max, K.kappa / (rho * c_p_eos));
    if (DIM == 2) return 2.0 * fmax(D_max / (G.dx[0] * G.dx[0]), D_max / (G.dx[1] * G.dx[1]));
    return 2.0 * fmax(D_max / (G.dx[0] * G.dx[0]), fmax(D_max / (G.dx[1] * G.dx[1]), D_max / (G.dx[2] * G.dx[2])));
}

/* ---- state management of a six-ghost level (what xfer::RefineSchedule::fillData does for a periodic single-patch level,
 * and the four-ghost view the convective reconstructor reads) ---- */
struct DiffStatePtrs {
    double* U[5];
    const double* src[5];
};

HB2D_HD int diff_wrap(int i, int n)
{
    int r = i % n;
    return r < 0 ? r + n : r;
}

/* ghost-box cell t of G: ghost cells of the directions in `mask` take the value of their periodic image. */
HB2D_HD void diff_fill_periodic_thread(const DiffGeom& G, const DiffStatePtrs& A, int ncomp, int mask, long long t)
{
    const int i = (int)(t % G.gd[0]) - G.g[0];
    const int j = (int)((t / G.gd[0]) % G.gd[1]) - G.g[1];
    const int k = (int)(t / ((long long)G.gd[0] * G.gd[1])) - G.g[2];
    const bool in0 = i >= 0 && i < G.n[0], in1 = j >= 0 && j < G.n[1], in2 = k >= 0 && k < G.n[2];
    /* a cell is filled when it is a ghost in at least one MASKED direction; only the masked coordinates are wrapped, so
     * the fill is ghost-inclusive in the other directions (after an exchange across ranks in those directions the edge
     * and corner ghosts come from the already exchanged slabs, like k_fill_periodic of the convective plan) */
    if (!((!in0 && (mask & 1)) || (!in1 && (mask & 2)) || (!in2 && (mask & 4)))) return;
    const int si = (mask & 1) ? diff_wrap(i, G.n[0]) : i, sj = (mask & 2) ? diff_wrap(j, G.n[1]) : j,
              sk = (mask & 4) ? diff_wrap(k, G.n[2]) : k;
    const long long s = (si + G.g[0]) + G.cs[1] * (sj + G.g[1]) + G.cs[2] * (sk + G.g[2]);
    for (int c = 0; c < ncomp; c++) A.U[c][t] = A.U[c][s];
}

/* cell t of the ghost box of Gd (fewer ghosts) copied from the same cell of the ghost box of Gs */
HB2D_HD void diff_extract_view_thread(const DiffGeom& Gs, const DiffGeom& Gd, const DiffStatePtrs& A, int ncomp, long long t)
{
    const int i = (int)(t % Gd.gd[0]) - Gd.g[0];
    const int j = (int)((t / Gd.gd[0]) % Gd.gd[1]) - Gd.g[1];
    const int k = (int)(t / ((long long)Gd.gd[0] * Gd.gd[1])) - Gd.g[2];
    const long long s = (i + Gs.g[0]) + Gs.cs[1] * (j + Gs.g[1]) + Gs.cs[2] * (k + Gs.g[2]);
    for (int c = 0; c < ncomp; c++) A.U[c][t] = A.src[c][s];
}

inline void make_diff_geom(int dim, const int* n, const double* dx, int g, DiffGeom* G)
{
    G->dim = dim;
    for (int a = 0; a < 3; a++) {
        G->n[a] = (a < dim) ? n[a] : 1;
        G->g[a] = (a < dim) ? g : 0;
        G->gd[a] = G->n[a] + 2 * G->g[a];
        G->dx[a] = (a < dim) ? dx[a] : 1.0;
        G->dx_inv[a] = (a < dim) ? 1.0 / dx[a] : 0.0;
    }
    G->cs[0] = 1;
    G->cs[1] = G->gd[0];
    G->cs[2] = (long long)G->gd[0] * G->gd[1];
    G->ncell_g = (long long)G->gd[0] * G->gd[1] * G->gd[2];
}

}  // namespace hb2
