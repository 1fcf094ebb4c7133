/*
 * oracle_diffusive.h -- CPU ORACLE of SURVEY.md row f4: node-based sixth-order diffusive (viscous) flux of the
 * single-species Navier-Stokes application.  TEST INFRASTRUCTURE ONLY (see hamers_oracle.h).
 */
#ifndef ORACLE_DIFFUSIVE_H
#define ORACLE_DIFFUSIVE_H
#include "hamers_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_DIFF_GHOSTS 6      /* DiffusiveFluxReconstructorNodeSixthOrder.cpp:24 */

/* CONSTANT shear / bulk viscosity, PRANDTL thermal conductivity, ideal gas: what the shipped viscous decks use
 * (problems/support_files/3D_Couette_flow/x-direction/input_3D_Couette_flow.txt:25-55) */
typedef struct {
    double mu;      /* species_mu */
    double mu_v;    /* species_mu_v */
    double c_p;     /* species_c_p */
    double c_v;     /* R/(gamma - 1) */
    double Pr;      /* species_Pr */
} orc_transport;

long orc_diff_ghost_size(const orc_desc* d);

/*
 * DiffusiveFluxReconstructorNode::computeDiffusiveFluxOnPatch with the SIXTH_ORDER kernels.
 *   Q[c]         : conservative components on the ghost box of width 6 (all of it filled), SAMRAI CellData layout
 *   F[dir*neq+e] : output side flux (ghost 0), already multiplied by dt, fully overwritten
 */
int orc_compute_diffusive_flux(const orc_desc* d, const orc_transport* tr, const double* const* Q, double dt,
                               double* const* F);

/* NavierStokes::advanceSingleStepOnPatch with the conservative diffusive flux (NavierStokes.cpp:2085-2092); all cell
 * data on ghost boxes of width g */
int orc_advance_stage_ns(const orc_desc* d, int g, int ncoef, const double* alpha, const double* beta,
                         const double* const* const* U_int, const double* const* const* Fc_int,
                         const double* const* const* Fd_int, const double* const* const* S_int, double* const* U_out);

/* NavierStokes::computeSpectralRadiusesAndStableDtOnPatch without source terms, over the six-ghost box: out[0..dim-1] the
 * acoustic spectral radii, out[dim] the stable dt, out[dim+1] the maximum diffusive spectral radius.  c_p_eos: isobaric
 * specific heat of the equation of state, gamma/(gamma - 1) R. */
int orc_ns_spectral_radii_and_dt(const orc_desc* d, const orc_transport* tr, double c_p_eos, const double* const* Q, double* out);
double orc_diff_max_diffusivity(double mu, double mu_v, double kappa, double c_p_eos, double rho);
double orc_diff_spectral_radius(int dim, double D_max, const double* dx);

/* point formulas exported for pinning against oracle/_ref */
double orc_diff_first_derivative(const double u[7], double dx_inv);
double orc_diff_reconstruct(const double F[6], double dt);
void orc_diff_derivative_array(int dim, int ddir, const double* u, const int* n, double dx_inv, double* out);
void orc_diff_reconstruct_array(int dim, int fdir, const double* F_node, const int* n, double dt, double* F_face);
double orc_diff_temperature(double gamma, double c_v, double rho, double p);
double orc_diff_conductivity(double c_p, double mu, double Pr);
void orc_diff_diffusivities(int dim, double mu, double mu_v, double kappa, const double* vel, double* D);
/* the terms of equation e of the node flux in direction fdir that carry a derivative in direction ddir */
void orc_diff_terms(int dim, int fdir, int ddir, int e, int* n, int var[4], int diff[4]);

/* ---- midpoint family: DiffusiveFluxReconstructorMidpointSixthOrder ("MIDPOINT_SIXTH_ORDER") ----
 * same argument meaning as orc_compute_diffusive_flux; reads the cells within 5 of the interior */
int orc_compute_diffusive_flux_midpoint(const orc_desc* d, const orc_transport* tr, const double* const* Q, double dt,
                                        double* const* F);
double orc_mid_derivative(const double u[6], double dx_inv);
double orc_mid_interpolate(const double u[6]);
double orc_mid_reconstruct(const double F[5], double dt);
void orc_mid_side_diffusivities(int dim, int dir, double mu, double mu_v, double kappa, const double* vel, double* D);
void orc_mid_side_terms(int dim, int fdir, int ddir, int e, int* n, int var[4], int diff[4]);

#ifdef __cplusplus
}
#endif
#endif
