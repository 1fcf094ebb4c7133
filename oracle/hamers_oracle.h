/*
 * hamers_oracle.h -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A plain-C, scalar, FP64 restatement of HAMeRS's per-patch WCNS5-JS / HLLC-HLL
 * convective-flux hot path in the reference's own operation order.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (hamers_b200/) never links or calls it.
 *
 * Parity status: the point kernels for the WCNS5-JS / WCNS5-Z / WCNS6-LD interpolations,
 * the HLLC / HLLC-HLL Riemann solvers, the ideal-gas equation of state, the sensor chain
 * and the face-flux formula are PINNED against the reference's own compiled code (oracle/_ref, built by
 * oracle/build_ref.py from /root/reference).  Everything else (derived cell data, characteristic
 * projection, bounds check / fallback, sensor, flux differencing, RK update) is a
 * restatement of formulas cited below ("parity unpinned" for those pieces: the
 * reference holds no golden vectors and cannot be built here without SAMRAI).
 *
 * All citations are path:line under /root/reference.
 */
#ifndef HAMERS_ORACLE_H
#define HAMERS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SPECIES 4
#define ORC_MAX_EQ 12          /* dim + 2*ns <= 3 + 8 */
#define ORC_GHOSTS 4           /* ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:22 */

#define ORC_SINGLE_SPECIES 0
#define ORC_FIVE_EQN_ALLAIRE 1
/* FlowModelFourEqnConservative (FlowModelManager.cpp "FOUR_EQN_CONSERVATIVE"): partial densities rho Y_i,
 * momentum, total energy; all equations conservative; mixture of ideal gases closed by mass fractions */
#define ORC_FOUR_EQN_CONSERVATIVE 2

typedef struct {
    int dim;                         /* 2 or 3 */
    int n[3];                        /* interior cells of the patch (n[2] ignored in 2D) */
    int model;                       /* ORC_SINGLE_SPECIES | ORC_FIVE_EQN_ALLAIRE */
    int ns;                          /* number of species (1 for single-species) */
    double gamma[ORC_MAX_SPECIES];   /* species ratio of specific heats */
    double dx[3];
    int weno_p;                      /* constant_p, default 2 (WCNS5-JS-HLLC-HLL.cpp:188-191) */
    /* the other nonlinear interpolators of the same base class.  0 = WCNS5-JS, 1 = WCNS5-Z
     * (ConvectiveFluxReconstructorWCNS5-Z-HLLC-HLL.cpp), 2 = WCNS6-LD (ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp)
     * with constant_q (default 4), constant_C (1.0e9), constant_alpha_tau (35): :343-361. */
    int scheme;
    int weno_q;
    double weno_C;
    double weno_alpha_tau;
    /* four-eqn conservative: species gas constants, Equation_of_state_mixing_rules { species_R }
     * (EquationOfStateMixingRulesIdealGas.cpp:60-125: c_p_i = gamma_i/(gamma_i - 1) R_i, c_v_i = 1/(gamma_i - 1) R_i) */
    double R[ORC_MAX_SPECIES];
} orc_desc;

#define ORC_WCNS5_JS 0
#define ORC_WCNS5_Z 1
#define ORC_WCNS6_LD 2

/* d + 2 (single-species), d + 2*ns (five-eqn: FlowModelFiveEqnAllaire.cpp:29) or d + 1 + ns (four-eqn conservative:
 * FlowModelFourEqnConservative.cpp:29) */
int orc_num_eqn(const orc_desc* d);
/* number of stored conservative components: num_eqn (+1 for the five-eqn Z_last) */
int orc_num_comp(const orc_desc* d);
/* number of doubles of one ghost-box cell component / one ghost-0 cell component / one side component */
long orc_cell_ghost_size(const orc_desc* d);
long orc_cell_size(const orc_desc* d);
long orc_side_size(const orc_desc* d, int dir);

/*
 * ConvectiveFluxReconstructorWCNS56::computeConvectiveFluxAndSourceOnPatch
 * (ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:39-2656).
 *   Q[c]        : conservative components on the ghost box (g = 4), SAMRAI CellData layout
 *   F[dir*neq+e]: output side flux (ghost 0), already multiplied by dt, fully overwritten
 *   S[e]        : output cell source (ghost 0), "+=" for ADVECTIVE equations only
 * Optional debug outputs (may be NULL): F_mid[dir*neq+e] on faces -1..N+1 (normal) x interior,
 * sensor[dir] same shape.
 */
int orc_compute_flux_and_source(const orc_desc* d, const double* const* Q, double dt,
                                double* const* F, double* const* S,
                                double* const* F_mid_dbg, double* const* sensor_dbg);

/*
 * Euler::advanceSingleStepOnPatch (Euler.cpp:1003-1679) for one patch.
 *   ncoef        : number of coefficients (= stage number + 1)
 *   U_int[m][c]  : intermediate states (ghost box layout), m < ncoef
 *   F_int[m][..] : intermediate fluxes (side layout as above), S_int[m][e] sources
 *   U_out[c]     : SCRATCH state, ghost box layout; whole array zero-filled first
 *                  (FlowModelSingleSpecies.cpp:1602-1620), interior then updated.
 */
int orc_advance_stage(const orc_desc* d, int ncoef,
                      const double* alpha, const double* beta,
                      const double* const* const* U_int,
                      const double* const* const* F_int,
                      const double* const* const* S_int,
                      double* const* U_out);

/*
 * Euler::computeSpectralRadiusesAndStableDtOnPatch (Euler.cpp:489-900; 3D body :760-893) with the derived data
 * MAX_WAVE_SPEED_d = |u_d| + c (FlowModelSingleSpecies.cpp:3959, 4003, 4064; five-eqn FlowModelFiveEqnAllaire.cpp
 * computeCellDataOfMaxWaveSpeedWithVelocityAndSoundSpeed) for one patch, WITHOUT source terms:
 *   out[a]   = max over the cells of (|u_a| + c)/dx_a,  a < dim
 *   out[dim] = 1 / max over the cells of sum_a (|u_a| + c)/dx_a          (the patch's stable dt for CFL = 1)
 * include_ghosts != 0 loops over the ghost box like the reference (its ghost cells hold neighbour data by then),
 * 0 over the interior only.
 */
int orc_spectral_radii_and_dt(const orc_desc* d, const double* const* Q, int include_ghosts, double* out);

/* Point kernels exported for pinning against oracle/_ref (the reference's own functions). */
/* first derivative, dilatation, vorticity magnitude, sensor value and the face-flux formula as the loops evaluate them
 * (layout of in/out: hamers_oracle.c) */
void orc_path_points(const double in[16], double out[5]);
/* velocity, internal energy, face averages, characteristic projection and its inverse (single-species, 3-D, x), RK update */
void orc_path_points2(const double in[32], double out[20]);
/* five-eqn (two species, 3-D, x): mixture density, mass fractions, velocity, internal energy, mixture gamma, pressure,
 * sound speed, face averages, characteristic projection and its inverse, advective source */
void orc_path_points3(const double in[56], double out[32]);
/* five-eqn: (rho, c, epsilon) rebuilt from one interpolated side in front of the Riemann point kernels */
void orc_path_points4(const double in[12], double out[4]);
/* bounds flags of one interpolated side: five-eqn (direction-dependent, see side_bounded) and single-species */
void orc_path_points5(const double in[16], double out[2]);
/* max wave speeds, spectral radii, their sum, running maximum and stable dt of one cell */
void orc_path_points6(const double in[8], double out[6]);
/* four-eqn conservative (two species, 3-D): mixture properties from mass fractions, pressure, sound speed of a cell; the
 * (rho, c, epsilon) of an interpolated side; its bounds flag.  in: rhoY0, rhoY1, mx, my, mz, E, gamma0, gamma1, R0, R1 |
 * side V[6]; out: rho, Y0, Y1, epsilon, c_p, c_v, gamma, p, Psi0, Psi1, c | rho, c, eps | flag */
void orc_path_points7(const double in[16], double out[15]);
/* four-eqn conservative (3-D, x): face averages, characteristic projection and its inverse */
void orc_path_points8(const double in[24], double out[16]);
void orc_constants(double out[7]); /* eps, sensor threshold, Y lo/up, Z lo/up, ghost width */
void orc_eos_point(double gamma, double rho, double epsilon, double* p, double* c, double* eps_back);
void orc_weno5js_point(const double U[6], int p, double* U_minus, double* U_plus);
void orc_weno5z_point(const double U[6], int p, double* U_minus, double* U_plus);
void orc_weno6ld_point(const double U[6], int p, int q, double C, double alpha_tau, double* U_minus, double* U_plus);
/* V layout: single-species [rho, vel(d), p]; five-eqn [Zrho(ns), vel(d), p, Z(ns-1)]; four-eqn [rhoY(ns), vel(d), p].
 * gamma: ns species gammas, followed by the ns species gas constants R for the four-eqn model */
void orc_riemann_point(int model, int dim, int ns, const double* gamma, int dir,
                       const double* V_L, const double* V_R,
                       double* F_HLLC, double* F_HLLC_HLL, double* vel_mid);

#ifdef __cplusplus
}
#endif
#endif
