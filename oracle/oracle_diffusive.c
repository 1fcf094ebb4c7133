/*
 * oracle_diffusive.c -- CPU ORACLE of SURVEY.md row f4 (test infrastructure, NOT a product path).
 *
 * Restates, in the reference's operation order, the conservative-form viscous flux of the single-species
 * Navier-Stokes application with the node-based sixth-order reconstructor (the one all ten shipped viscous decks use,
 * diffusive_flux_reconstructor = "SIXTH_ORDER"):
 *
 *   DiffusiveFluxReconstructorNode::computeDiffusiveFluxOnPatch
 *       src/flow/diffusive_flux_reconstructors/node/DiffusiveFluxReconstructorNode.cpp:31-1736
 *   DiffusiveFluxReconstructorNodeSixthOrder::computeFirstDerivativesIn{X,Y,Z}, reconstructFlux{X,Y,Z}
 *       .../node/DiffusiveFluxReconstructorNodeSixthOrder.cpp:65-939   (6 ghost cells, :24-27)
 *   FlowModelDiffusiveFluxUtilitiesSingleSpecies: variables to differentiate (:654-1503), diffusivity of each term
 *       (:1505-2363), the diffusivities D_00..D_12 (:4031-4296)
 *       src/flow/flow_models/single-species/FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp
 *   temperature  EquationOfStateIdealGas.cpp:6897   T = p/((gamma - 1) c_v rho)
 *   viscosities  EquationOfShearViscosityConstant.cpp:273, EquationOfBulkViscosityConstant (same form)
 *   conductivity EquationOfThermalConductivityPrandtl.cpp:309   kappa = c_p mu / Pr
 *   RK update    NavierStokes.cpp:2085-2092   (conservative form of the diffusive flux)
 *
 * Parity status: the derivative and reconstruction kernels, the point formulas and the term tables (which derivative
 * carries which diffusivity in which equation, in the order of accumulation) are pinned against the reference's own code,
 * compiled verbatim (oracle/build_ref.py: diffusive_kernels, diffusive_term_tables; tests/test_oracle_diffusive.py); the
 * loop ranges and the order of the x / y / z derivative groups are restated from the driver cited above.
 */
#include "oracle_diffusive.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GD ORC_DIFF_GHOSTS

/* DiffusiveFluxReconstructorNodeSixthOrder.cpp:76-78 */
static const double a_n = 3.0 / 4.0;
static const double b_n = -(3.0 / 20.0);
static const double c_n = 1.0 / 60.0;

/* :236-239 (x), :391-394 (y), :488-491 (z): note the multiplication by the inverse mesh width */
double orc_diff_first_derivative(const double u[7], double dx_inv)
{
    return (a_n * (u[4] - u[2]) + b_n * (u[5] - u[1]) + c_n * (u[6] - u[0])) * dx_inv;
}

/* :679-683: face value from the six nodes around the face (LLL, LL, L, R, RR, RRR), already times dt */
double orc_diff_reconstruct(const double F[6], double dt)
{
    const double a_r = a_n + b_n + c_n;
    const double b_r = b_n + c_n;
    const double c_r = c_n;
    return dt * (a_r * (F[2] + F[3]) + b_r * (F[1] + F[4]) + c_r * (F[0] + F[5]));
}

/* EquationOfStateIdealGas.cpp:6897; EquationOfThermalConductivityPrandtl.cpp:309 */
double orc_diff_temperature(double gamma, double c_v, double rho, double p) { return p / ((gamma - 1.0) * c_v * rho); }
double orc_diff_conductivity(double c_p, double mu, double Pr) { return c_p * mu / Pr; }

/* FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:4262-4280 (3-D: D_00..D_12), :4180-4193 (2-D: D_00..D_09) */
void orc_diff_diffusivities(int dim, double mu, double mu_v, double kappa, const double* vel, double* D)
{
    const double u = vel[0], v = vel[1];
    if (dim == 2) {
        D[0] = -(4.0 / 3.0 * mu + mu_v);
        D[1] = 2.0 / 3.0 * mu - mu_v;
        D[2] = -mu;
        D[3] = -u * (4.0 / 3.0 * mu + mu_v);
        D[4] = -v * (4.0 / 3.0 * mu + mu_v);
        D[5] = u * (2.0 / 3.0 * mu - mu_v);
        D[6] = v * (2.0 / 3.0 * mu - mu_v);
        D[7] = -u * mu;
        D[8] = -v * mu;
        D[9] = -kappa;
    } else {
        const double w = vel[2];
        D[0] = -(4.0 / 3.0 * mu + mu_v);
        D[1] = 2.0 / 3.0 * mu - mu_v;
        D[2] = -mu;
        D[3] = -u * (4.0 / 3.0 * mu + mu_v);
        D[4] = -v * (4.0 / 3.0 * mu + mu_v);
        D[5] = -w * (4.0 / 3.0 * mu + mu_v);
        D[6] = u * (2.0 / 3.0 * mu - mu_v);
        D[7] = v * (2.0 / 3.0 * mu - mu_v);
        D[8] = w * (2.0 / 3.0 * mu - mu_v);
        D[9] = -u * mu;
        D[10] = -v * mu;
        D[11] = -w * mu;
        D[12] = -kappa;
    }
}

/* One term of a node flux: derivative of variable `var` (0..dim-1 velocity component, dim = temperature) times
 * diffusivity D[diff].  terms[flux dir][derivative dir][equation] lists them in the reference's order
 * (FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:978-1503 variables, :1838-2363 diffusivities for 3-D;
 * :760-976, :1640-1836 for 2-D).  The continuity equation has no term. */
typedef struct {
    int n;
    struct {
        int var, diff;
    } t[4];
} term_list;

#define T3 3 /* temperature slot, 3-D */
static const term_list terms3[3][3][5] = {
    /* flux x */
    {{{0, {{0, 0}}}, {1, {{0, 0}}}, {1, {{1, 2}}}, {1, {{2, 2}}}, {4, {{0, 3}, {1, 10}, {2, 11}, {T3, 12}}}},      /* d/dx */
     {{0, {{0, 0}}}, {1, {{1, 1}}}, {1, {{0, 2}}}, {0, {{0, 0}}}, {2, {{0, 10}, {1, 6}}}},                         /* d/dy */
     {{0, {{0, 0}}}, {1, {{2, 1}}}, {0, {{0, 0}}}, {1, {{0, 2}}}, {2, {{0, 11}, {2, 6}}}}},                        /* d/dz */
    /* flux y */
    {{{0, {{0, 0}}}, {1, {{1, 2}}}, {1, {{0, 1}}}, {0, {{0, 0}}}, {2, {{0, 7}, {1, 9}}}},
     {{0, {{0, 0}}}, {1, {{0, 2}}}, {1, {{1, 0}}}, {1, {{2, 2}}}, {4, {{0, 9}, {1, 4}, {2, 11}, {T3, 12}}}},
     {{0, {{0, 0}}}, {0, {{0, 0}}}, {1, {{2, 1}}}, {1, {{1, 2}}}, {2, {{1, 11}, {2, 7}}}}},
    /* flux z */
    {{{0, {{0, 0}}}, {1, {{2, 2}}}, {0, {{0, 0}}}, {1, {{0, 1}}}, {2, {{0, 8}, {2, 9}}}},
     {{0, {{0, 0}}}, {0, {{0, 0}}}, {1, {{2, 2}}}, {1, {{1, 1}}}, {2, {{1, 8}, {2, 10}}}},
     {{0, {{0, 0}}}, {1, {{0, 2}}}, {1, {{1, 2}}}, {1, {{2, 0}}}, {4, {{0, 9}, {1, 10}, {2, 5}, {T3, 12}}}}},
};
#define T2 2 /* temperature slot, 2-D */
static const term_list terms2[2][2][4] = {
    {{{0, {{0, 0}}}, {1, {{0, 0}}}, {1, {{1, 2}}}, {3, {{0, 3}, {1, 8}, {T2, 9}}}},
     {{0, {{0, 0}}}, {1, {{1, 1}}}, {1, {{0, 2}}}, {2, {{0, 8}, {1, 5}}}}},
    {{{0, {{0, 0}}}, {1, {{1, 2}}}, {1, {{0, 1}}}, {2, {{0, 6}, {1, 7}}}},
     {{0, {{0, 0}}}, {1, {{0, 2}}}, {1, {{1, 0}}}, {3, {{0, 7}, {1, 4}, {T2, 9}}}}},
};

static const term_list* terms_of(int dim, int fdir, int ddir, int e)
{
    return dim == 3 ? &terms3[fdir][ddir][e] : &terms2[fdir][ddir][e];
}

void orc_diff_terms(int dim, int fdir, int ddir, int e, int* n, int var[4], int diff[4])
{
    const term_list* tl = terms_of(dim, fdir, ddir, e);
    *n = tl->n;
    for (int i = 0; i < tl->n; i++) {
        var[i] = tl->t[i].var;
        diff[i] = tl->t[i].diff;
    }
}

long orc_diff_ghost_size(const orc_desc* d)
{
    long s = 1;
    for (int a = 0; a < d->dim; a++) s *= d->n[a] + 2 * GD;
    return s;
}

/* computeFirstDerivativesIn{X,Y,Z} over the range the base class passes (DiffusiveFluxReconstructorNode.cpp:1795-1812):
 * the whole ghost box, shrunk by 3 cells on both sides of the derivative direction.  u, out: ghost-box layout. */
void orc_diff_derivative_array(int dim, int ddir, const double* u, const int* n, double dx_inv, double* out)
{
    const int g2 = dim == 3 ? GD : 0;
    const int n2 = dim == 3 ? n[2] : 1;
    const long e0 = n[0] + 2 * GD, e1 = n[1] + 2 * GD;
    const long cs[3] = {1, e0, e0 * e1};
    int lo[3] = {-GD, -GD, -g2}, hi[3] = {n[0] + GD, n[1] + GD, n2 + g2};
    lo[ddir] += 3;
    hi[ddir] -= 3;
    for (int k = lo[2]; k < hi[2]; k++)
        for (int j = lo[1]; j < hi[1]; j++)
            for (int i = lo[0]; i < hi[0]; i++) {
                const long x = (long)(i + GD) + e0 * ((long)(j + GD) + e1 * (long)(k + g2));
                double s[7];
                for (int m = 0; m < 7; m++) s[m] = u[x + (m - 3) * cs[ddir]];
                out[x] = orc_diff_first_derivative(s, dx_inv);
            }
}

/* reconstructFlux{X,Y,Z} over the faces of the interior (DiffusiveFluxReconstructorNode.cpp:2276-2297): "+=" into the
 * side data.  F_node: ghost-box layout; F_face: side layout of direction fdir, ghost 0. */
void orc_diff_reconstruct_array(int dim, int fdir, const double* F_node, const int* n, double dt, double* F_face)
{
    const int g2 = dim == 3 ? GD : 0;
    const int n2 = dim == 3 ? n[2] : 1;
    const long e0 = n[0] + 2 * GD, e1 = n[1] + 2 * GD;
    const long cs[3] = {1, e0, e0 * e1};
    const int fn[3] = {n[0] + (fdir == 0), n[1] + (fdir == 1), n2 + (fdir == 2)};
    for (int k = 0; k < fn[2]; k++)
        for (int j = 0; j < fn[1]; j++)
            for (int i = 0; i < fn[0]; i++) {
                const long x = (long)(i + GD) + e0 * ((long)(j + GD) + e1 * (long)(k + g2));   /* cell on the high side */
                double s[6];
                for (int m = 0; m < 6; m++) s[m] = F_node[x + (m - 3) * cs[fdir]];
                F_face[i + (long)fn[0] * ((long)j + (long)fn[1] * (long)k)] += orc_diff_reconstruct(s, dt);
            }
}

int orc_compute_diffusive_flux(const orc_desc* d, const orc_transport* tr, const double* const* Q, double dt,
                               double* const* F)
{
    if (d->model != ORC_SINGLE_SPECIES || (d->dim != 2 && d->dim != 3)) return 1;
    const int dim = d->dim, neq = dim + 2;
    const int n0 = d->n[0], n1 = d->n[1], n2 = dim == 3 ? d->n[2] : 1;
    const int g2 = dim == 3 ? GD : 0;
    const long e0 = n0 + 2 * GD, e1 = n1 + 2 * GD;
    const long ncell = orc_diff_ghost_size(d);
#define CIDX(i, j, k) ((long)((i) + GD) + e0 * ((long)((j) + GD) + e1 * (long)((k) + g2)))

    /* derived cell data on the whole ghost box: velocity, pressure, temperature, diffusivities
     * (FlowModelDiffusiveFluxUtilitiesSingleSpecies::computeDerivedCellData, :490-567) */
    const int nD = dim == 3 ? 13 : 10;
    double* var[4];      /* velocity components, then temperature */
    for (int a = 0; a <= dim; a++) var[a] = (double*)malloc(sizeof(double) * (size_t)ncell);
    double* D[13];
    for (int m = 0; m < nD; m++) D[m] = (double*)malloc(sizeof(double) * (size_t)ncell);
    const double kappa = orc_diff_conductivity(tr->c_p, tr->mu, tr->Pr);
    for (long x = 0; x < ncell; x++) {
        const double rho = Q[0][x];
        double vel[3] = {0.0, 0.0, 0.0}, ke = 0.0, Dx[13];
        for (int a = 0; a < dim; a++) {
            vel[a] = Q[1 + a][x] / rho;                                   /* FlowModelSingleSpecies.cpp:2824-2826 */
            ke = (a == 0) ? vel[a] * vel[a] : ke + vel[a] * vel[a];
            var[a][x] = vel[a];
        }
        const double epsilon = Q[dim + 1][x] / rho - 1.0 / 2.0 * ke;       /* :3049-3051 */
        const double p = (d->gamma[0] - 1.0) * rho * epsilon;              /* EquationOfStateIdealGas.cpp:5580 */
        var[dim][x] = orc_diff_temperature(d->gamma[0], tr->c_v, rho, p);
        orc_diff_diffusivities(dim, tr->mu, tr->mu_v, kappa, vel, Dx);
        for (int m = 0; m < nD; m++) D[m][x] = Dx[m];
    }

    /* first derivatives of every variable in every direction, each computed once like the reference's
     * derivatives_{x,y,z}_computed maps do (DiffusiveFluxReconstructorNode.cpp:1764-1838) */
    double* der[4][3];
    for (int a = 0; a <= dim; a++)
        for (int ddir = 0; ddir < dim; ddir++) {
            der[a][ddir] = (double*)malloc(sizeof(double) * (size_t)ncell);
            orc_diff_derivative_array(dim, ddir, var[a], d->n, 1.0 / d->dx[ddir], der[a][ddir]);
        }

    double* Fn = (double*)malloc(sizeof(double) * (size_t)ncell);
    for (int fdir = 0; fdir < dim; fdir++) {
        /* node flux on the interior cells extended by 3 in the flux direction (DiffusiveFluxReconstructorNode.cpp:
         * 3-D x :888-1055, y :1161-1328, z :1434-1601): fillAll(0), then "+=" term by term, x- then y- then
         * z-derivatives */
        const int lo[3] = {fdir == 0 ? -3 : 0, fdir == 1 ? -3 : 0, fdir == 2 ? -3 : 0};
        const int hi[3] = {n0 + (fdir == 0 ? 3 : 0), n1 + (fdir == 1 ? 3 : 0), n2 + (fdir == 2 ? 3 : 0)};
        for (int e = 0; e < neq; e++) {
            memset(Fn, 0, sizeof(double) * (size_t)ncell);
            for (int ddir = 0; ddir < dim; ddir++) {
                const term_list* tl = terms_of(dim, fdir, ddir, e);
                for (int ti = 0; ti < tl->n; ti++) {
                    const double* mu = D[tl->t[ti].diff];
                    const double* dudx = der[tl->t[ti].var][ddir];
                    for (int k = lo[2]; k < hi[2]; k++)
                        for (int j = lo[1]; j < hi[1]; j++)
                            for (int i = lo[0]; i < hi[0]; i++) {
                                const long x = CIDX(i, j, k);
                                Fn[x] += mu[x] * dudx[x];
                            }
                }
            }
            /* face flux: the side data is zero-filled first (DiffusiveFluxReconstructorNode.cpp:84) */
            double* Fs = F[fdir * neq + e];
            memset(Fs, 0, sizeof(double) * (size_t)orc_side_size(d, fdir));
            orc_diff_reconstruct_array(dim, fdir, Fn, d->n, dt, Fs);
        }
    }
    free(Fn);
    for (int a = 0; a <= dim; a++)
        for (int ddir = 0; ddir < dim; ddir++) free(der[a][ddir]);
    for (int m = 0; m < nD; m++) free(D[m]);
    for (int a = 0; a <= dim; a++) free(var[a]);
#undef CIDX
    return 0;
}


/* ====================================================================================================================
 * Midpoint family: DiffusiveFluxReconstructorMidpointSixthOrder ("MIDPOINT_SIXTH_ORDER"; no shipped deck selects it).
 *   driver   src/flow/diffusive_flux_reconstructors/midpoint/DiffusiveFluxReconstructorMidpoint.cpp:38-2330
 *   kernels  .../midpoint/DiffusiveFluxReconstructorMidpointSixthOrder.cpp:68-1799 (5 ghost cells, :24-30)
 *   side diffusivities  FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:2365-2797 (what is interpolated, the D of a
 *            direction), :2799-3657 (which D multiplies which derivative)
 * The flux is formed AT THE MIDPOINTS (faces) of the flux direction f: derivatives along f by a staggered sixth-order
 * difference of the node values, derivatives along the other directions by sixth-order node derivatives interpolated along
 * f, diffusivities from mu, mu_v, kappa and the velocity interpolated along f; the face flux is a five-midpoint
 * combination, times dt.  Arrays staged like the reference stages them (every intermediate is its own array).
 * Parity status: the twelve kernels, the side-diffusivity statements and the side term table are pinned against the
 * reference's own code compiled verbatim (oracle/build_ref.py: midpoint_kernels; tests/test_oracle_diffusive_midpoint.py);
 * loop ranges and the order of the derivative groups are restated from the driver.
 * Midpoint arrays here use the ghost-box layout: midpoint i of direction f (the face between cells i - 1 and i) is stored
 * at cell index i. */

/* DiffusiveFluxReconstructorMidpointSixthOrder.cpp:80-82 / :962-964 / :1400-1406 */
static const double a_dm = 75.0 / 64.0, b_dm = -(25.0 / 384.0), c_dm = 3.0 / 640.0;
static const double a_im = 75.0 / 128.0, b_im = -(25.0 / 256.0), c_im = 3.0 / 256.0;

/* u[0..5] = nodes LLL, LL, L, R, RR, RRR around the midpoint */
double orc_mid_derivative(const double u[6], double dx_inv)
{
    return (a_dm * (u[3] - u[2]) + b_dm * (u[4] - u[1]) + c_dm * (u[5] - u[0])) * dx_inv;
}
double orc_mid_interpolate(const double u[6])
{
    return (a_im * (u[3] + u[2]) + b_im * (u[4] + u[1]) + c_im * (u[5] + u[0]));
}
/* F[0..4] = midpoints LL, L, the face's own, R, RR */
double orc_mid_reconstruct(const double F[5], double dt)
{
    const double a_r = a_dm + b_dm + c_dm;
    const double b_r = b_dm + c_dm;
    const double c_r = c_dm;
    return dt * (a_r * (F[2]) + b_r * (F[1] + F[3]) + c_r * (F[0] + F[4]));
}

/* FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:2682-2691 (x), 2725-2734 (y), 2768-2777 (z); 2-D :2581-2589, 2615-2623 */
void orc_mid_side_diffusivities(int dim, int dir, double mu, double mu_v, double kappa, const double* vel, double* D)
{
    const double un = vel[dir];
    D[0] = -(4.0 / 3.0 * mu + mu_v);
    D[1] = 2.0 / 3.0 * mu - mu_v;
    D[2] = -mu;
    D[3] = -un * (4.0 / 3.0 * mu + mu_v);
    D[4] = un * (2.0 / 3.0 * mu - mu_v);
    int m = 5;
    for (int a = 0; a < dim; a++)
        if (a != dir) D[m++] = -vel[a] * mu;
    D[m] = -kappa;
}

/* index of the side diffusivity of term ti of (flux direction, derivative direction, equation); the variables are those of
 * terms3 / terms2 (one function of the reference serves both reconstructor families) */
static const signed char side3[3][3][5][4] = {
    {{{-1, -1, -1, -1}, {0, -1, -1, -1}, {2, -1, -1, -1}, {2, -1, -1, -1}, {3, 5, 6, 7}},
     {{-1, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {-1, -1, -1, -1}, {5, 4, -1, -1}},
     {{-1, -1, -1, -1}, {1, -1, -1, -1}, {-1, -1, -1, -1}, {2, -1, -1, -1}, {6, 4, -1, -1}}},
    {{{-1, -1, -1, -1}, {2, -1, -1, -1}, {1, -1, -1, -1}, {-1, -1, -1, -1}, {4, 5, -1, -1}},
     {{-1, -1, -1, -1}, {2, -1, -1, -1}, {0, -1, -1, -1}, {2, -1, -1, -1}, {5, 3, 6, 7}},
     {{-1, -1, -1, -1}, {-1, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {6, 4, -1, -1}}},
    {{{-1, -1, -1, -1}, {2, -1, -1, -1}, {-1, -1, -1, -1}, {1, -1, -1, -1}, {4, 5, -1, -1}},
     {{-1, -1, -1, -1}, {-1, -1, -1, -1}, {2, -1, -1, -1}, {1, -1, -1, -1}, {4, 6, -1, -1}},
     {{-1, -1, -1, -1}, {2, -1, -1, -1}, {2, -1, -1, -1}, {0, -1, -1, -1}, {5, 6, 3, 7}}}};
static const signed char side2[2][2][4][4] = {
    {{{-1, -1, -1, -1}, {0, -1, -1, -1}, {2, -1, -1, -1}, {3, 5, 6, -1}},
     {{-1, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {5, 4, -1, -1}}},
    {{{-1, -1, -1, -1}, {2, -1, -1, -1}, {1, -1, -1, -1}, {4, 5, -1, -1}},
     {{-1, -1, -1, -1}, {2, -1, -1, -1}, {0, -1, -1, -1}, {5, 3, 6, -1}}}};

void orc_mid_side_terms(int dim, int fdir, int ddir, int e, int* n, int var[4], int diff[4])
{
    const term_list* tl = terms_of(dim, fdir, ddir, e);
    *n = tl->n;
    for (int i = 0; i < tl->n; i++) {
        var[i] = tl->t[i].var;
        diff[i] = dim == 3 ? side3[fdir][ddir][e][i] : side2[fdir][ddir][e][i];
    }
}

int orc_compute_diffusive_flux_midpoint(const orc_desc* d, const orc_transport* tr, const double* const* Q, double dt,
                                        double* const* F)
{
    if (d->model != ORC_SINGLE_SPECIES || (d->dim != 2 && d->dim != 3)) return 1;
    const int dim = d->dim, neq = dim + 2;
    const int n[3] = {d->n[0], d->n[1], dim == 3 ? d->n[2] : 1};
    const int g2 = dim == 3 ? GD : 0;
    const long e0 = n[0] + 2 * GD, e1 = n[1] + 2 * GD;
    const long cs[3] = {1, e0, e0 * e1};
    const long ncell = orc_diff_ghost_size(d);
#define CIDX(i, j, k) ((long)((i) + GD) + e0 * ((long)((j) + GD) + e1 * (long)((k) + g2)))
    /* cell data on the ghost box: velocity, temperature, and the three transport coefficients (constant, but interpolated
     * like any other cell data: the interpolation weights do not sum to one in floating point) */
    double* var[4];
    for (int a = 0; a <= dim; a++) var[a] = (double*)malloc(sizeof(double) * (size_t)ncell);
    const double kappa = orc_diff_conductivity(tr->c_p, tr->mu, tr->Pr);
    for (long x = 0; x < ncell; x++) {
        const double rho = Q[0][x];
        double ke = 0.0;
        for (int a = 0; a < dim; a++) {
            const double v = Q[1 + a][x] / rho;
            ke = (a == 0) ? v * v : ke + v * v;
            var[a][x] = v;
        }
        const double epsilon = Q[dim + 1][x] / rho - 1.0 / 2.0 * ke;
        const double p = (d->gamma[0] - 1.0) * rho * epsilon;
        var[dim][x] = orc_diff_temperature(d->gamma[0], tr->c_v, rho, p);
    }
    /* node derivatives of every variable in every direction (DiffusiveFluxReconstructorMidpoint.cpp:3013-3136: the whole
     * ghost box shrunk by three along the derivative direction) */
    double* dnode[4][3];
    for (int a = 0; a <= dim; a++)
        for (int dd = 0; dd < dim; dd++) {
            dnode[a][dd] = (double*)calloc((size_t)ncell, sizeof(double));
            orc_diff_derivative_array(dim, dd, var[a], d->n, 1.0 / d->dx[dd], dnode[a][dd]);
        }
    double* dmid = (double*)malloc(sizeof(double) * (size_t)ncell);
    double* Fm = (double*)malloc(sizeof(double) * (size_t)ncell);
    double* Dm[8];
    for (int m = 0; m < 8; m++) Dm[m] = (double*)malloc(sizeof(double) * (size_t)ncell);
    const double cmu[6] = {tr->mu, tr->mu, tr->mu, tr->mu, tr->mu, tr->mu};
    const double cmv[6] = {tr->mu_v, tr->mu_v, tr->mu_v, tr->mu_v, tr->mu_v, tr->mu_v};
    const double cka[6] = {kappa, kappa, kappa, kappa, kappa, kappa};
    const double mu_m = orc_mid_interpolate(cmu), mu_v_m = orc_mid_interpolate(cmv), kappa_m = orc_mid_interpolate(cka);
    for (int f = 0; f < dim; f++) {
        /* midpoints -2 .. n_f + 2 of the flux direction, interior cells of the others (:1466-1470) */
        int lo[3] = {0, 0, 0}, hi[3] = {n[0], n[1], n[2]};
        lo[f] = -2;
        hi[f] = n[f] + 3;
#define FOR_MIDPOINTS for (int k = lo[2]; k < hi[2]; k++) for (int j = lo[1]; j < hi[1]; j++) for (int i = lo[0]; i < hi[0]; i++)
        /* side diffusivities of direction f from the interpolated transport coefficients and velocity */
        FOR_MIDPOINTS {
            const long x = CIDX(i, j, k);
            double vel[3] = {0.0, 0.0, 0.0}, Dx[8], s[6];
            for (int a = 0; a < dim; a++) {
                for (int m = 0; m < 6; m++) s[m] = var[a][x + (m - 3) * cs[f]];
                vel[a] = orc_mid_interpolate(s);
            }
            orc_mid_side_diffusivities(dim, f, mu_m, mu_v_m, kappa_m, vel, Dx);
            for (int m = 0; m < (dim == 3 ? 8 : 7); m++) Dm[m][x] = Dx[m];
        }
        for (int e = 0; e < neq; e++) {
            memset(Fm, 0, sizeof(double) * (size_t)ncell);
            for (int dd = 0; dd < dim; dd++) {
                int nt, tv[4], td[4];
                orc_mid_side_terms(dim, f, dd, e, &nt, tv, td);
                for (int ti = 0; ti < nt; ti++) {
                    const double* src = (dd == f) ? var[tv[ti]] : dnode[tv[ti]][dd];
                    FOR_MIDPOINTS {
                        const long x = CIDX(i, j, k);
                        double s[6];
                        for (int m = 0; m < 6; m++) s[m] = src[x + (m - 3) * cs[f]];
                        dmid[x] = (dd == f) ? orc_mid_derivative(s, 1.0 / d->dx[f]) : orc_mid_interpolate(s);
                    }
                    const double* mu = Dm[td[ti]];
                    FOR_MIDPOINTS {
                        const long x = CIDX(i, j, k);
                        Fm[x] += mu[x] * dmid[x];
                    }
                }
            }
            double* Fs = F[f * neq + e];
            const int fn[3] = {n[0] + (f == 0), n[1] + (f == 1), n[2] + (f == 2)};
            for (int k = 0; k < fn[2]; k++)
                for (int j = 0; j < fn[1]; j++)
                    for (int i = 0; i < fn[0]; i++) {
                        const long x = CIDX(i, j, k);
                        double s[5];
                        for (int m = 0; m < 5; m++) s[m] = Fm[x + (m - 2) * cs[f]];
                        double v = 0.0;                                   /* fillAll(0), then "+=" */
                        v += orc_mid_reconstruct(s, dt);
                        Fs[i + (long)fn[0] * ((long)j + (long)fn[1] * (long)k)] = v;
                    }
        }
#undef FOR_MIDPOINTS
    }
    for (int m = 0; m < 8; m++) free(Dm[m]);
    free(Fm);
    free(dmid);
    for (int a = 0; a <= dim; a++)
        for (int dd = 0; dd < dim; dd++) free(dnode[a][dd]);
    for (int a = 0; a <= dim; a++) free(var[a]);
#undef CIDX
    return 0;
}

/* FlowModelSingleSpecies.cpp:4661-4665: MAX_DIFFUSIVITY = max(mu/rho, mu_v/rho, kappa/(rho c_p)), c_p the isobaric specific
 * heat of the equation of state (gamma/(gamma - 1) R, EquationOfStateMixingRulesIdealGas.cpp:113) */
double orc_diff_max_diffusivity(double mu, double mu_v, double kappa, double c_p_eos, double rho)
{
    double D_max = fmax(mu / rho, mu_v / rho);
    D_max = fmax(D_max, kappa / (rho * c_p_eos));
    return D_max;
}

/* NavierStokes.cpp:1083-1086 (3-D), 884-886 (2-D): diffusive spectral radius of one cell */
double orc_diff_spectral_radius(int dim, double D_max, const double* dx)
{
    if (dim == 2) return 2.0 * fmax(D_max / (dx[0] * dx[0]), D_max / (dx[1] * dx[1]));
    return 2.0 * fmax(D_max / (dx[0] * dx[0]), fmax(D_max / (dx[1] * dx[1]), D_max / (dx[2] * dx[2])));
}

/* NavierStokes::computeSpectralRadiusesAndStableDtOnPatch (NavierStokes.cpp:585-1102; 3-D :1006-1089), no source terms:
 * over the whole ghost box (six ghosts)
 *   out[a]   = max (|u_a| + c)/dx_a,  a < dim
 *   out[dim] = 1 / (max(max sum_a (|u_a| + c)/dx_a, max diffusive spectral radius) + HAMERS_EPSILON)
 *   out[dim + 1] = max diffusive spectral radius (not an output of the reference; exported for the tests) */
int orc_ns_spectral_radii_and_dt(const orc_desc* d, const orc_transport* tr, double c_p_eos, const double* const* Q, double* out)
{
    if (d->model != ORC_SINGLE_SPECIES || (d->dim != 2 && d->dim != 3)) return 1;
    const int dim = d->dim;
    const long ncell = orc_diff_ghost_size(d);
    const double kappa = orc_diff_conductivity(tr->c_p, tr->mu, tr->Pr);
    double sr[3] = {0.0, 0.0, 0.0}, sr_sum = 0.0, sr_diff = 0.0;
    for (long x = 0; x < ncell; x++) {
        const double rho = Q[0][x];
        double vel[3], ke = 0.0;
        for (int a = 0; a < dim; a++) {
            vel[a] = Q[1 + a][x] / rho;
            ke = (a == 0) ? vel[a] * vel[a] : ke + vel[a] * vel[a];
        }
        const double epsilon = Q[dim + 1][x] / rho - 1.0 / 2.0 * ke;
        double p, c, eb;
        orc_eos_point(d->gamma[0], rho, epsilon, &p, &c, &eb);
        double sum = 0.0;
        for (int a = 0; a < dim; a++) {
            const double s = (fabs(vel[a]) + c) / d->dx[a];
            sr[a] = fmax(sr[a], s);
            sum = (a == 0) ? s : sum + s;
        }
        sr_sum = fmax(sr_sum, sum);
        sr_diff = fmax(sr_diff, orc_diff_spectral_radius(dim, orc_diff_max_diffusivity(tr->mu, tr->mu_v, kappa, c_p_eos, rho), d->dx));
    }
    for (int a = 0; a < dim; a++) out[a] = sr[a];
    out[dim] = 1.0 / (fmax(sr_diff, sr_sum) + 1.0e-15);
    out[dim + 1] = sr_diff;
    return 0;
}

/* NavierStokes::advanceSingleStepOnPatch, conservative diffusive form (NavierStokes.cpp:1947-2097, 3-D :2085-2092;
 * 2-D :1715-1751).  U on ghost boxes of width g (same for U_int and U_out), fluxes and sources on ghost 0. */
int orc_advance_stage_ns(const orc_desc* d, int g, int ncoef, const double* alpha, const double* beta,
                         const double* const* const* U_int, const double* const* const* Fc_int,
                         const double* const* const* Fd_int, const double* const* const* S_int, double* const* U_out)
{
    const int dim = d->dim, neq = orc_num_eqn(d);
    const int n0 = d->n[0], n1 = d->n[1], n2 = dim == 3 ? d->n[2] : 1;
    const int g2 = dim == 3 ? g : 0;
    const long e0 = n0 + 2 * g, e1 = n1 + 2 * g, e2 = n2 + 2 * g2;
    const long ncell = e0 * e1 * e2;
    for (int e = 0; e < neq; e++) memset(U_out[e], 0, sizeof(double) * (size_t)ncell);
    for (int n = 0; n < ncoef; n++) {
        if (alpha[n] != 0.0)
            for (int e = 0; e < neq; e++)
                for (int k = 0; k < n2; k++)
                    for (int j = 0; j < n1; j++)
                        for (int i = 0; i < n0; i++) {
                            const long x = (i + g) + e0 * ((long)(j + g) + e1 * (long)(k + g2));
                            U_out[e][x] += alpha[n] * U_int[n][e][x];
                        }
        if (beta[n] != 0.0)
            for (int e = 0; e < neq; e++) {
                const double *Fcx = Fc_int[n][0 * neq + e], *Fcy = Fc_int[n][1 * neq + e];
                const double *Fdx = Fd_int[n][0 * neq + e], *Fdy = Fd_int[n][1 * neq + e];
                const double* Fcz = dim == 3 ? Fc_int[n][2 * neq + e] : NULL;
                const double* Fdz = dim == 3 ? Fd_int[n][2 * neq + e] : NULL;
                const double* S = S_int[n][e];
                for (int k = 0; k < n2; k++)
                    for (int j = 0; j < n1; j++)
                        for (int i = 0; i < n0; i++) {
                            const long x = (i + g) + e0 * ((long)(j + g) + e1 * (long)(k + g2));
                            const long xs = i + (long)n0 * ((long)j + (long)n1 * (long)k);
                            const long fxL = i + (long)(n0 + 1) * ((long)j + (long)n1 * (long)k), fxR = fxL + 1;
                            const long fyB = i + (long)n0 * ((long)j + (long)(n1 + 1) * (long)k), fyT = fyB + n0;
                            if (dim == 2) {
                                U_out[e][x] += beta[n] * (-(Fcx[fxR] - Fcx[fxL] + Fdx[fxR] - Fdx[fxL]) / d->dx[0] -
                                                          (Fcy[fyT] - Fcy[fyB] + Fdy[fyT] - Fdy[fyB]) / d->dx[1] + S[xs]);
                            } else {
                                const long fzB = xs, fzF = fzB + (long)n0 * n1;
                                U_out[e][x] += beta[n] * (-(Fcx[fxR] - Fcx[fxL] + Fdx[fxR] - Fdx[fxL]) / d->dx[0] -
                                                          (Fcy[fyT] - Fcy[fyB] + Fdy[fyT] - Fdy[fyB]) / d->dx[1] -
                                                          (Fcz[fzF] - Fcz[fzB] + Fdz[fzF] - Fdz[fzB]) / d->dx[2] + S[xs]);
                            }
                        }
            }
    }
    return 0;
}
