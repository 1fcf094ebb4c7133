/*
 * hamers_b200.h -- C ABI of the B200-native WCNS5-JS / HLLC-HLL convective-flux path.
 *
 * This is the drop-in boundary for ONE hot path of HAMeRS (all citations are path:line under
 * the reference tree):
 *
 *   ConvectiveFluxReconstructor::computeConvectiveFluxAndSourceOnPatch
 *       include/flow/convective_flux_reconstructors/ConvectiveFluxReconstructor.hpp:73-81
 *       (implementation replaced: src/flow/convective_flux_reconstructors/WCNS56/
 *        ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:39-2656 with the WCNS5-JS interpolator
 *        ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:234-737)
 *   RungeKuttaPatchStrategy::{computeFluxesAndSourcesOnPatch, advanceSingleStepOnPatch}
 *       include/algs/patch_strategy/RungeKuttaPatchStrategy.hpp:149-190
 *       (Euler.cpp:904-999 and :1003-1679)
 *   the same-level ghost fill that RungeKuttaLevelIntegrator::advanceLevel triggers per stage
 *       src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1568, :1701
 *
 * Conventions
 *   - plain C, no C++/torch types; every entry point returns 0 on success and a negative
 *     code on failure, with a message available from hb2_last_error() (the C++ wrapper turns
 *     non-zero into TBOX_ERROR, the reference's only error convention).
 *   - arrays are SAMRAI pdat layouts: cell data with ghost width g = 4,
 *     x fastest, one pointer per depth component; side data ghost 0, one pointer per
 *     (direction, component): flux[dir*num_eqn + e].
 *   - conservative components, in order:
 *       single-species : rho, rho*u, rho*v, (rho*w), E                      (num_comp = dim+2)
 *       five-eqn       : Zrho_1..Zrho_ns, rho*u, rho*v, (rho*w), E, Z_1..Z_ns (num_comp = dim+2ns+1;
 *                        the last volume fraction is stored but is not an equation)
 *       four-eqn cons. : rhoY_1..rhoY_ns, rho*u, rho*v, (rho*w), E           (num_comp = dim+1+ns)
 *   - pointers in the *_dev calls are DEVICE pointers on the plan's device and must stay valid
 *     until the plan's stream has been synchronised; the *_host calls take HOST pointers and do
 *     the H2D / D2H copies themselves (the reference-facing, host-memory drop-in).
 *   - a plan is not re-entrant (the reference's FlowModel is a stateful singleton too,
 *     FlowModelSingleSpecies.cpp:706-712); use one plan per patch shape per calling thread.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef HAMERS_B200_H
#define HAMERS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB2_MAX_SPECIES 4
#define HB2_MAX_EQ 12
#define HB2_MAX_COMP 13
#define HB2_MAX_STAGES 4
#define HB2_GHOSTS 4            /* ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:22 */

#define HB2_SINGLE_SPECIES 0    /* FlowModelManager.cpp:20 "SINGLE_SPECIES" */
#define HB2_FIVE_EQN_ALLAIRE 1  /* FlowModelManager.cpp:44 "FIVE_EQN_ALLAIRE" */
/* FlowModelManager.cpp "FOUR_EQN_CONSERVATIVE" (src/flow/flow_models/four-eqn_conservative/): partial densities
 * rho Y_1..rho Y_ns, momentum, total energy (num_comp = num_eqn = dim + 1 + ns), all equations conservative; mixture of
 * ideal gases closed by mass fractions (species_gamma AND species_R are needed).  Built for num_species = 2 and 3; runs the
 * reference-order kernels whatever `math` says. */
#define HB2_FOUR_EQN_CONSERVATIVE 2

/* arithmetic variants */
#define HB2_MATH_EXACT 0        /* reference operation order, no FMA contraction: bit-identical to the oracle */
#define HB2_MATH_FAST 1         /* FMA contraction + reciprocal sharing: <= 1e-12 relative of the oracle */

typedef struct hb2_patch_desc {
    int32_t dim;                              /* 2 or 3 */
    int32_t n[3];                             /* interior cells of the patch */
    int32_t flow_model;                       /* HB2_SINGLE_SPECIES | HB2_FIVE_EQN_ALLAIRE */
    int32_t num_species;                      /* 1; 2 or 3 for the five-eqn and four-eqn models (the reference is generic in
                                                 d_num_species, FlowModelFiveEqnAllaire.cpp:29; three species run the
                                                 reference-order kernels whatever `math` says) */
    double species_gamma[HB2_MAX_SPECIES];    /* Equation_of_state_mixing_rules{species_gamma} */
    double dx[3];                             /* CartesianPatchGeometry::getDx() */
    int32_t weno_p;                           /* Convective_flux_reconstructor{constant_p}, default 2 */
    int32_t math;                             /* HB2_MATH_EXACT | HB2_MATH_FAST */
    int32_t device;                           /* CUDA device ordinal, -1 = current */
    /* Which subclass of ConvectiveFluxReconstructorWCNS56 the plan stands for (ConvectiveFluxReconstructorManager.cpp:
     * "WCNS5_JS_HLLC_HLL" / "WCNS5_Z_HLLC_HLL" / "WCNS6_LD_HLLC_HLL") and the WCNS6-LD constants
     * Convective_flux_reconstructor{constant_q, constant_C, constant_alpha_tau} (defaults 4, 1.0e9, 35:
     * ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp:343-361; 0 selects the default).  HB2_MATH_FAST exists for
     * constant_p = 2 (and constant_q = 4 for WCNS6-LD); every other combination runs the reference-order kernels. */
    int32_t scheme;                           /* HB2_WCNS5_JS | HB2_WCNS5_Z | HB2_WCNS6_LD */
    int32_t weno_q;
    double weno_C;
    double weno_alpha_tau;
    /* Ghost width of the cell-data LAYOUT the plan's state arrays use: 0 = HB2_GHOSTS.  The kernels read four layers
     * whatever it is; a Navier-Stokes application allocates the state with six (the diffusive reconstructor's width,
     * HB2_DIFF_GHOSTS) and hands the same arrays to both reconstructors, like SAMRAI does.  4 <= num_ghosts <= 8. */
    int32_t num_ghosts;
    /* Equation_of_state_mixing_rules{species_R}: gas constants of the species, used by HB2_FOUR_EQN_CONSERVATIVE only
     * (EquationOfStateMixingRulesIdealGas.cpp:60-119) */
    double species_R[HB2_MAX_SPECIES];
} hb2_patch_desc;

#define HB2_WCNS5_JS 0
#define HB2_WCNS5_Z 1
#define HB2_WCNS6_LD 2

typedef struct hb2_plan_s* hb2_plan_t;

/* ---- introspection ------------------------------------------------------------------ */
const char* hb2_last_error(void);
const char* hb2_version(void);
int hb2_device_count(int32_t* count);
/* The constants of the path's hard switches as compiled into the kernels: out = { HAMERS_EPSILON
 * (include/HAMeRS_config.hpp.in:16), the sensor threshold 0.65 (ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2123),
 * Y / Z bounds lo, up (FlowModelBasicUtilitiesFiveEqnAllaire.hpp:24-27), ghost width (…WCNS56-HLLC-HLL.cpp:22) }. */
int hb2_constants(double out[7]);
int hb2_num_eqn(const hb2_patch_desc* d, int32_t* num_eqn);      /* FlowModel::getNumberOfEquations() */
int hb2_num_comp(const hb2_patch_desc* d, int32_t* num_comp);
/* ConvectiveFluxReconstructor::getConvectiveFluxNumberOfGhostCells(): 4 in every direction */
int hb2_num_ghosts(const hb2_patch_desc* d, int32_t ghosts[3]);
/* element counts of one component: ghost-box cell data, ghost-0 cell data, ghost-0 side data */
int64_t hb2_cell_ghost_size(const hb2_patch_desc* d);
int64_t hb2_cell_size(const hb2_patch_desc* d);
int64_t hb2_side_size(const hb2_patch_desc* d, int32_t dir);

/* ---- plan --------------------------------------------------------------------------- */
int hb2_plan_create(const hb2_patch_desc* d, hb2_plan_t* plan);
int hb2_plan_destroy(hb2_plan_t plan);
/* cudaStream_t to launch on; 0 is the legacy default stream.  A new plan launches on a private
 * non-blocking stream; hb2_plan_use_own_stream() returns to it. */
int hb2_plan_set_stream(hb2_plan_t plan, void* cuda_stream);
int hb2_plan_use_own_stream(hb2_plan_t plan);
int hb2_plan_synchronize(hb2_plan_t plan);
/* number of kernels this plan has launched since creation (bench.py's gpu_launches) */
int64_t hb2_plan_launch_count(hb2_plan_t plan);
/* bytes of device workspace the plan owns */
int64_t hb2_plan_workspace_bytes(hb2_plan_t plan);
/* Per-kernel device timing with CUDA events on the launching stream (the analogue of the reference's
 * tbox::TimerManager timers t_compute_fluxes_sources / t_advance_step, Euler.cpp:44-52).
 * kinds: 0 sensor, 1 x sweep, 2 y sweep, 3 z sweep, 4 advance, 5 periodic fill, 6 pack, 7 unpack */
#define HB2_NUM_KERNEL_KINDS 8
int hb2_plan_set_profiling(hb2_plan_t plan, int32_t on);
int hb2_plan_get_profile(hb2_plan_t plan, double* ms_total, int64_t* launches, int32_t reset);

/* ---- the hot path, device-resident data ------------------------------------------------ */

/* computeConvectiveFluxAndSourceOnPatch.  Q: num_comp ghost-box components (ghosts filled by
 * the caller).  flux: dim*num_eqn side arrays, fully overwritten on faces 0..N of each
 * direction, ALREADY multiplied by dt.  source: num_eqn ghost-0 cell arrays, "+=" and only for
 * ADVECTIVE equations (the five-eqn volume fractions); entries of non-advective equations may
 * be NULL. */
int hb2_compute_flux_and_source_dev(hb2_plan_t plan, const double* const* Q, double dt,
                                    double* const* flux, double* const* source);

/* advanceSingleStepOnPatch.  ncoef = RK stage number + 1.  U_int[m*num_comp + c] are the
 * intermediate states (ghost-box layout), F_int[m*dim*num_eqn + dir*num_eqn + e] and
 * S_int[m*num_eqn + e] their fluxes / sources (rows with beta[m] == 0 and gamma[m] == 0 may be
 * NULL).  U_out: the SCRATCH state; its interior is overwritten (ghosts are left untouched: they
 * are refilled before they are next read).  F_acc / S_acc (may be NULL): the gamma-weighted
 * accumulation used by AMR flux correction, "+=". */
int hb2_advance_stage_dev(hb2_plan_t plan, int32_t ncoef,
                          const double* alpha, const double* beta, const double* gamma,
                          const double* const* U_int, const double* const* F_int,
                          const double* const* S_int, double* const* U_out,
                          double* const* F_acc, double* const* S_acc);

/* computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch in one pass that never writes the
 * side fluxes to HBM (uniform level, no AMR flux sums).  Requires beta[m] == 0 for m < ncoef-1
 * (true for SSP-RK3 and every "newest flux only" table); the flux is evaluated on
 * U_int[ncoef-1].  U_out must not alias U_int[ncoef-1] (older states may be overwritten in place). */
int hb2_fused_stage_dev(hb2_plan_t plan, int32_t ncoef, const double* alpha, const double* beta,
                        const double* const* U_int, double dt, double* const* U_out);

/* hb2_fused_stage_dev with the same-level ghost fill of the NEW state (what the next stage's
 * xfer::RefineSchedule::fillData would do, RungeKuttaLevelIntegrator.cpp:1568/1701) fused into the update: every new
 * cell value within the ghost width of a patch face is also stored into the ghost box of each neighbouring patch.
 * push: 27 * ncomp device addresses, push[code * ncomp + c] with code = (ox+1) + 3 (oy+1) + 9 (oz+1) = component c of
 * U_out of the equal-sized patch at offset (ox,oy,oz) -- memory of a peer GPU opened with hb2_ipc_open (stores
 * travel over NVLink) or this patch's own U_out where the level is periodic over a single patch; NULL entries: no
 * neighbour.  The caller orders stages across GPUs (one barrier between a stage and the next). */
int hb2_fused_stage_push_dev(hb2_plan_t plan, int32_t ncoef, const double* alpha, const double* beta,
                             const double* const* U_int, double dt, double* const* U_out,
                             double* const* push);

/* Device memory that can be shared between the one-process-per-GPU ranks of a box (CUDA IPC): allocate, export
 * a 64-byte handle, open a peer's handle, close it. */
#define HB2_IPC_HANDLE_BYTES 64
int hb2_device_malloc(int64_t bytes, void** ptr);
int hb2_device_free(void* ptr);
int hb2_ipc_export(const void* ptr, uint8_t handle[HB2_IPC_HANDLE_BYTES]);
int hb2_ipc_open(const uint8_t handle[HB2_IPC_HANDLE_BYTES], void** ptr);
int hb2_ipc_close(void* ptr);

/* Same-level periodic ghost fill of one patch that covers the whole periodic level in the
 * directions flagged in periodic_mask (bit d).  All 4-cell ghosts incl. edges and corners. */
int hb2_fill_ghosts_periodic_dev(hb2_plan_t plan, double* const* U, int32_t periodic_mask);

/* RungeKuttaLevelIntegrator::advanceLevel stage loop (RungeKuttaLevelIntegrator.cpp:1672-1745) for one
 * patch covering a periodic level: per stage a periodic ghost fill of the newest state, then the fused
 * stage.  alpha/beta: row-major [nstages][nstages] (SSP-RK3 default: :3894-3929).  U: U^n in, U^{n+1}
 * (interior) out; nstages <= 3. */
int hb2_advance_level_dev(hb2_plan_t plan, int32_t nstages, const double* alpha, const double* beta,
                          double dt, int32_t periodic_mask, double* const* U);

/* Halo exchange building blocks for GPU-resident neighbouring patches: copy the box
 * [lo, hi) (cell indices relative to the interior origin, may extend into ghosts) of every
 * component to / from a contiguous buffer (component-major, x fastest). */
int hb2_pack_box_dev(hb2_plan_t plan, const double* const* U, const int32_t lo[3],
                     const int32_t hi[3], double* buffer);
int hb2_unpack_box_dev(hb2_plan_t plan, double* const* U, const int32_t lo[3],
                       const int32_t hi[3], const double* buffer);

/* The same for up to HB2_MAX_BOXES boxes in ONE launch (what one same-level ghost fill of
 * xfer::RefineSchedule::fillData, RungeKuttaLevelIntegrator.cpp:1568/1701, moves between a patch and its up to 26
 * neighbours): box b = [lo + 3b, hi + 3b) is copied to / from buffer + offsets[b] (offsets in doubles). */
#define HB2_MAX_BOXES 32
int hb2_pack_boxes_dev(hb2_plan_t plan, const double* const* U, int32_t nbox, const int32_t* lo,
                       const int32_t* hi, const int64_t* offsets, double* buffer);
int hb2_unpack_boxes_dev(hb2_plan_t plan, double* const* U, int32_t nbox, const int32_t* lo,
                         const int32_t* hi, const int64_t* offsets, const double* buffer);

/* Same-level ghost fill between equally sized patches WITHOUT message buffers: box b of this patch's state is stored
 * straight into the state array of a neighbouring patch -- dst[b] is component 0 of that array (components comp_stride
 * doubles apart; the same ghost-box geometry as the plan's), on this GPU or on a peer GPU whose allocation was opened
 * with hb2_ipc_open -- at cell (i, j, k) - shift[3b..3b+2].  For a periodic neighbour at offset o, lo / hi is the
 * interior slab next to face o and shift = o * n: what xfer::RefineSchedule::fillData
 * (RungeKuttaLevelIntegrator.cpp:1568/1701) copies, travelling as NVLink stores instead of pack -> message -> unpack.
 * The caller orders "all ranks' stores have landed" before the ghosts are read (one stream-ordered all-reduce). */
int hb2_push_boxes_dev(hb2_plan_t plan, const double* const* U, int32_t nbox, const int32_t* lo,
                       const int32_t* hi, double* const* dst, const int32_t* shift, int64_t comp_stride);

/* Euler::computeSpectralRadiusesAndStableDtOnPatch (Euler.cpp:489-900) without source terms, with
 * MAX_WAVE_SPEED_d = |u_d| + c (FlowModelSingleSpecies.cpp:3884-4388) evaluated on the fly:
 *   out_dev[a]   = max over the interior of (|u_a| + c)/dx_a, a < dim   (the reference also visits the ghost cells,
 *                  which hold neighbour interiors: the level-wide maximum is the same)
 *   out_dev[3]   = max over the interior of sum_a (|u_a| + c)/dx_a     (stable dt of the patch = CFL / out_dev[3])
 * out_dev: 4 doubles of DEVICE memory; non-negative doubles, so a MAX all-reduce over ranks
 * (RungeKuttaLevelIntegrator.cpp:1864, 1911) can be applied to them directly. */
int hb2_max_wave_speed_dev(hb2_plan_t plan, const double* const* Q, double* out_dev);

/* ---- the hot path, HOST buffers (H2D + kernels + D2H inside the call) ------------------- */
int hb2_compute_flux_and_source_host(hb2_plan_t plan, const double* const* Q_host, double dt,
                                     double* const* flux_host, double* const* source_host);
/* One RK stage on host memory: uploads U_int rows (ghost-filled by the caller), runs the fused
 * stage, downloads the INTERIOR of U_out (the ghost cells of the host arrays are left untouched). */
int hb2_fused_stage_host(hb2_plan_t plan, int32_t ncoef, const double* alpha, const double* beta,
                         const double* const* U_int_host, double dt, double* const* U_out_host);

/* advanceLevel on host memory: H2D of U^n, all stages on the device, D2H of U^{n+1}. */
int hb2_advance_level_host(hb2_plan_t plan, int32_t nstages, const double* alpha, const double* beta,
                           double dt, int32_t periodic_mask, double* const* U_host);

/* ---- measurement helpers --------------------------------------------------------------- */
/* Dependent-free DFMA loop on every SM; returns achieved FP64 FLOP/s (2 per FMA). */
int hb2_probe_fp64_peak(int32_t device, double seconds_hint, double* flops_per_s);
/* device-to-device copy bandwidth, bytes read + written per second */
int hb2_probe_hbm_bandwidth(int32_t device, int64_t bytes, double* bytes_per_s);

/* ---- diffusive (viscous) flux of the single-species Navier-Stokes application -------------------
 * Replaces DiffusiveFluxReconstructorNodeSixthOrder ("SIXTH_ORDER", the reconstructor of every shipped viscous deck):
 *   src/flow/diffusive_flux_reconstructors/node/DiffusiveFluxReconstructorNode.cpp:31-1736 (computeDiffusiveFluxOnPatch)
 *   src/flow/diffusive_flux_reconstructors/node/DiffusiveFluxReconstructorNodeSixthOrder.cpp:65-939 (kernels)
 *   src/flow/flow_models/single-species/FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:654-2363, 4031-4296
 * with equation_of_shear_viscosity / equation_of_bulk_viscosity = "CONSTANT", equation_of_thermal_conductivity =
 * "PRANDTL" and the ideal-gas temperature (what those decks select).  Cell data carries SIX ghost cells here
 * (d_num_diff_ghosts, DiffusiveFluxReconstructorNodeSixthOrder.cpp:24), all of them filled by the caller. */
#define HB2_DIFF_GHOSTS 6
typedef struct {
    int32_t dim;            /* 2 or 3 */
    int32_t n[3];           /* interior cells of the patch */
    double dx[3];
    double species_gamma;   /* Equation_of_state_mixing_rules { species_gamma } */
    double species_c_v;     /* R/(gamma - 1) with species_R of the same block */
    double species_mu;      /* Equation_of_shear_viscosity_mixing_rules { species_mu } */
    double species_mu_v;    /* Equation_of_bulk_viscosity_mixing_rules { species_mu_v } */
    double species_c_p;     /* Equation_of_thermal_conductivity_mixing_rules { species_c_p, species_Pr } */
    double species_Pr;
    int32_t device;         /* CUDA device index; -1 = the calling thread's current device */
} hb2_diffusive_desc;
typedef struct hb2_diff_plan_s* hb2_diff_plan_t;

int hb2_diffusive_plan_create(const hb2_diffusive_desc* desc, hb2_diff_plan_t* plan);
int hb2_diffusive_plan_destroy(hb2_diff_plan_t plan);
int hb2_diffusive_plan_set_stream(hb2_diff_plan_t plan, void* cuda_stream);
int hb2_diffusive_plan_launches(hb2_diff_plan_t plan, int64_t* launches);
/* Arithmetic of the flux-free route (hb2_diffusive_divergence_accumulate_dev), like hb2_patch_desc::math of the convective
 * plan: HB2_MATH_EXACT (default) keeps the reference's operations in the reference's order (bit-identical to the oracle);
 * HB2_MATH_FAST re-associates them (one reciprocal per cell, coefficients pre-multiplied by 1/dx, explicit FMAs, the energy
 * flux assembled from the momentum fluxes, the two faces of a cell differenced analytically): <= 1e-12 relative, 3-D only
 * (2-D patches keep the exact arithmetic).  The materialised side flux (hb2_compute_diffusive_flux_*) is always exact. */
int hb2_diffusive_plan_set_math(hb2_diff_plan_t plan, int32_t math);
/* Which reconstructor of the reference's DiffusiveFluxReconstructorManager (DiffusiveFluxReconstructorManager.cpp:13-92)
 * hb2_compute_diffusive_flux_* stands for:
 *   HB2_DIFF_NODE_SIXTH_ORDER      "SIXTH_ORDER"          DiffusiveFluxReconstructorNodeSixthOrder (default; every shipped deck)
 *   HB2_DIFF_MIDPOINT_SIXTH_ORDER  "MIDPOINT_SIXTH_ORDER" DiffusiveFluxReconstructorMidpointSixthOrder
 *       (src/flow/diffusive_flux_reconstructors/midpoint/DiffusiveFluxReconstructorMidpoint.cpp:38-2330,
 *       DiffusiveFluxReconstructorMidpointSixthOrder.cpp:68-1799): the flux is formed at the midpoints of the flux direction
 *       from staggered derivatives, interpolated node derivatives and interpolated diffusivities
 *       (FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:2365-3657).  Same six-ghost layout of Q (the cells within 5 of
 *       the interior are read); the flux-free update exists for the node reconstructor only. */
#define HB2_DIFF_NODE_SIXTH_ORDER 0
#define HB2_DIFF_MIDPOINT_SIXTH_ORDER 1
int hb2_diffusive_plan_set_reconstructor(hb2_diff_plan_t plan, int32_t reconstructor);

/* DiffusiveFluxReconstructor::computeDiffusiveFluxOnPatch
 * (include/flow/diffusive_flux_reconstructors/DiffusiveFluxReconstructor.hpp; called from NavierStokes.cpp:1153-1160).
 * Q: dim+2 conservative components on the ghost box of width 6 (device pointers).  flux[dir*num_eqn + e]: ghost-0 side
 * arrays, fully overwritten, ALREADY multiplied by dt (the continuity rows are +0.0). */
int hb2_compute_diffusive_flux_dev(hb2_diff_plan_t plan, const double* const* Q, double dt, double* const* flux);
/* same with host pointers (H2D, kernels and D2H inside the call): the API-preserving seam */
int hb2_compute_diffusive_flux_host(hb2_diff_plan_t plan, const double* const* Q_host, double dt, double* const* flux_host);

/* The diffusive part of NavierStokes::computeSpectralRadiusesAndStableDtOnPatch (NavierStokes.cpp:1064-1089; 2-D :868-891):
 * out_dev[0] = max over the six-ghost box of 2 max_d MAX_DIFFUSIVITY / dx_d^2, MAX_DIFFUSIVITY = max(mu/rho, mu_v/rho,
 * kappa/(rho c_p)) (FlowModelSingleSpecies.cpp:4661-4665) with c_p = species_c_p_eos, the isobaric specific heat of the
 * equation of state (gamma/(gamma - 1) R).  The patch's stable dt is then
 *   1 / (max(out_dev[0], <max sum of the acoustic radii: hb2_max_wave_speed_dev's 4th output>) + HAMERS_EPSILON). */
int hb2_diffusive_max_spectral_radius_dev(hb2_diff_plan_t plan, const double* const* Q, double species_c_p_eos, double* out_dev);

/* Same-level periodic ghost fill of a six-ghost state (what xfer::RefineSchedule::fillData does for one patch covering a
 * periodic level, RungeKuttaLevelIntegrator.cpp:1568, 1701); periodic_mask bit d = direction d is periodic. */
int hb2_diffusive_fill_ghosts_periodic_dev(hb2_diff_plan_t plan, double* const* U, int32_t periodic_mask);
/* The num_ghosts-wide view of a six-ghost state, as a separate array: the convective reconstructor reads four ghost
 * cells (hb2_compute_flux_and_source_dev), SAMRAI hands it the same allocation with a larger ghost box instead. */
int hb2_diffusive_extract_view_dev(hb2_diff_plan_t plan, const double* const* U, int32_t num_ghosts, double* const* U_view);

/* NavierStokes::advanceSingleStepOnPatch with the conservative diffusive flux (NavierStokes.cpp:1715-1751, 2085-2092):
 *   U_out = sum_m alpha[m] U_int[m] + beta[m] ( -(Fc_R - Fc_L + Fd_R - Fd_L)/dx_0 - ... + S[m] )
 * U_int[m*num_eqn + e], U_out[e]: cell data with num_ghosts ghost cells; Fc_int / Fd_int[m*dim*num_eqn + dir*num_eqn + e],
 * S_int[m*num_eqn + e]: ghost 0.  Rows with a zero coefficient may be NULL.  The ghosts of U_out are left untouched. */
int hb2_advance_stage_ns_dev(hb2_diff_plan_t plan, int32_t num_ghosts, int32_t ncoef, const double* alpha, const double* beta,
                             const double* const* U_int, const double* const* Fc_int, const double* const* Fd_int,
                             const double* const* S_int, double* const* U_out);

/* The diffusive part of that update on top of a state that already holds sum_m alpha[m] U_int[m] + beta (-div F_c + S),
 * i.e. the output of hb2_fused_stage_dev on the same arrays (plan created with num_ghosts = 6):
 *   U += beta ( -(Fd_R - Fd_L)/dx_0 - ... ).
 * Same terms as NavierStokes.cpp:2085-2092 associated differently (a few ulp): the HB2_MATH_FAST route of the stage,
 * which never materialises the convective flux.  U[e]: cell data with num_ghosts ghost cells. */
int hb2_diffusive_accumulate_dev(hb2_diff_plan_t plan, int32_t num_ghosts, double beta, const double* const* Fd, double* const* U);
/* hb2_compute_diffusive_flux_dev(Q) + hb2_diffusive_accumulate_dev in one call that never writes the diffusive side flux
 * either: the node fluxes of all directions stay in the plan's scratch and both faces of a cell are reconstructed where
 * they are differenced.  Bit-identical to the two-call route.  Q: six-ghost state the flux is taken of; U: the state to
 * update (num_ghosts ghost cells; may not alias Q). */
int hb2_diffusive_divergence_accumulate_dev(hb2_diff_plan_t plan, const double* const* Q, double dt, int32_t num_ghosts,
                                            double beta, double* const* U);

/* ---- patch-data operators of a two-level AMR step (Berger-Colella flux correction) -----------------------
 * What RungeKuttaLevelIntegrator does around the hot path when a finer level exists (src/algs/integrator/
 * RungeKuttaLevelIntegrator.cpp): coarse-fine ghost fill with time interpolation (:1568, xfer::RefineSchedule), flux
 * integrals on the outer sides of fine patches (:2968-3230, src/algs/integrator/fortran/algs_upfluxsum{2,3}d.f), their
 * coarsening onto the coarse flux (:2147-2158), the repeated conservative difference Euler::synchronizeFluxes
 * (Euler.cpp:1682-1949 = hb2_advance_stage_dev with alpha = beta = 1 on the accumulated flux) and the conservative
 * coarsening of the solution (:2189-2203).  One coarse patch and one fine patch per call; all pointers are DEVICE
 * pointers; cell data in the SAMRAI ghost-box layout, side data ghost 0.  The three SAMRAI geom operators are restated
 * from SAMRAI 4.1.0's published algorithm (parity unpinned: SAMRAI is not part of the reference tree). */
typedef struct hb2_amr_pair {
    int32_t dim;            /* 2 or 3 */
    int32_t nc[3];          /* interior cells of the coarse patch */
    int32_t nf[3];          /* interior cells of the fine patch = ratio * (coarse cells it covers) */
    int32_t ratio[3];       /* PatchHierarchy{ratio_to_coarser} */
    int32_t origin[3];      /* index, in the coarse patch, of the coarse cell that holds fine cell 0 */
    int32_t ghosts_c;       /* ghost width of the coarse / fine cell-data layouts (HB2_GHOSTS for the convective state) */
    int32_t ghosts_f;
    int32_t ncomp;          /* conservative components moved (hb2_num_comp) */
    int32_t neq;            /* equations with a flux (hb2_num_eqn) */
    double dxc[3], dxf[3];
} hb2_amr_pair;

/* Coarse-to-fine fill of the fine cells [lo, hi) (fine patch indices, ghost cells included): linear time interpolation
 * between Uc_old and Uc_new (tfrac = (t - t_old)/(t_new - t_old); Uc_new = NULL: Uc_old alone), then
 * "CONSERVATIVE_LINEAR_REFINE".  The coarse ghost cells the stencil reaches must be filled by the caller. */
int hb2_amr_refine_dev(const hb2_amr_pair* pair, const double* const* Uc_old, const double* const* Uc_new, double tfrac,
                       const int32_t lo[3], const int32_t hi[3], double* const* Uf, void* cuda_stream);
/* "CONSERVATIVE_COARSEN": the coarse cells [lo, hi) (coarse patch indices, covered by the fine patch) become the
 * volume-weighted averages of their fine cells. */
int hb2_amr_coarsen_dev(const hb2_amr_pair* pair, const double* const* Uf, const int32_t lo[3], const int32_t hi[3],
                        double* const* Uc, void* cuda_stream);
/* upfluxsumside{2,3}d: fluxsum[(2 dir + side) * neq + e][tangential cell] += F_fine[dir * neq + e][face on that patch side].
 * fluxsum arrays are dense over the tangential cells of the fine patch, lower direction fastest; zeroed by the caller on
 * the first fine step of a coarse step (preprocessFluxAndSourceData, :2880-2960). */
int hb2_amr_fluxsum_update_dev(const hb2_amr_pair* pair, const double* const* F_fine, double* const* fluxsum, void* cuda_stream);
/* The coarse side flux on every face of the fine patch's boundary := area-weighted average of the fine flux integrals. */
int hb2_amr_coarsen_fluxsum_dev(const hb2_amr_pair* pair, const double* const* fluxsum, double* const* F_coarse, void* cuda_stream);
/* BDRY_COND::BASIC::FLOW (BasicCartesianBoundaryUtilities2.cpp:310-345): the ghost cells beyond face (dir, side) of the
 * plan's patch copy the adjacent interior cell; the other directions run over the interior (fill periodic / other
 * boundaries afterwards). */
int hb2_fill_ghosts_extrapolate_dev(hb2_plan_t plan, double* const* U, int32_t dir, int32_t side);

/* ---- device-resident patch level: the loop of RungeKuttaLevelIntegrator::advanceLevel over ALL patches a rank owns -----
 * (RungeKuttaLevelIntegrator.cpp:1672-1745; RungeKuttaPatchStrategy.hpp:149-190).  The conservative variables of every
 * patch are registered once and stay in HBM (three state buffers per patch); boxes may have any sizes; the same-level
 * ghost fill between the rank's patches (periodic images included) is one kernel launch; every patch advances with the
 * fused stage.  `model`: flow model, species, dx, math, scheme, num_ghosts, device (n is ignored).  lo / hi: npatch x 3
 * box corners in LEVEL index space, hi exclusive.  level_n: cells of the level (periodic wrap).  A level spread over
 * several ranks exchanges the ghosts that belong to other ranks through hb2_pack_boxes_dev / NCCL on the states returned
 * by hb2_level_patch_state_dev, like hamers_b200/level.py does for one box per rank. */
typedef struct hb2_level_s* hb2_level_t;
int hb2_level_create(const hb2_patch_desc* model, int32_t npatch, const int32_t* lo, const int32_t* hi, const int32_t level_n[3],
                     int32_t periodic_mask, hb2_level_t* level);
int hb2_level_destroy(hb2_level_t level);
int hb2_level_num_patches(hb2_level_t level);
int64_t hb2_level_launch_count(hb2_level_t level);
int hb2_level_synchronize(hb2_level_t level);
/* HOST ghost-box arrays (SAMRAI CellData::getPointer(c)) <-> the patch's current state in HBM */
int hb2_level_upload_patch(hb2_level_t level, int32_t patch, const double* const* U_host);
int hb2_level_download_patch(hb2_level_t level, int32_t patch, double* const* U_host);
/* device pointers (num_comp of them) of intermediate state `state` (0 = current) of a patch */
int hb2_level_patch_state_dev(hb2_level_t level, int32_t patch, int32_t state, double** U_dev);
/* xfer::RefineSchedule::fillData, same level: every ghost cell of state `state` covered by a patch of this level (or a
 * periodic image of one) is copied from it */
int hb2_level_fill_ghosts(hb2_level_t level, int32_t state);
/* computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch of every patch for RK stage ncoef - 1 (coefficient rows of
 * length ncoef); last_stage != 0: the result becomes the current state */
int hb2_level_advance_stage(hb2_level_t level, int32_t ncoef, const double* alpha, const double* beta, double dt, int32_t last_stage);
/* the same patch by patch (the reference calls the patch strategy once per patch and stage, RungeKuttaLevelIntegrator.cpp:
 * 1709-1741): advance one patch through stage ncoef - 1; then, once per stage, hb2_level_end_stage */
int hb2_level_advance_stage_patch(hb2_level_t level, int32_t patch, int32_t ncoef, const double* alpha, const double* beta, double dt);
int hb2_level_end_stage(hb2_level_t level, int32_t ncoef, const double* alpha, int32_t last_stage);
/* the stage loop of advanceLevel: alpha / beta row-major [nstages][nstages] */
int hb2_level_advance(hb2_level_t level, int32_t nstages, const double* alpha, const double* beta, double dt);
/* advanceLevel on HOST memory, pipelined over the patches: U_host[patch * num_comp + c] are ghost-box host arrays (pinned
 * memory for asynchronous copies), new state on return.  The upload of later patches overlaps the first stage of the
 * patches that have arrived, the download of finished patches overlaps the last stage of the others. */
int hb2_level_advance_host(hb2_level_t level, int32_t nstages, const double* alpha, const double* beta, double dt,
                           double* const* U_host);
/* max over the rank's patches of the spectral radii (hb2_max_wave_speed_dev per patch); out_host: 4 doubles */
int hb2_level_max_wave_speed(hb2_level_t level, double out_host[4]);

#ifdef __cplusplus
}
#endif
#endif /* HAMERS_B200_H */
