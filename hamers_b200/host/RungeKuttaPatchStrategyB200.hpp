/*
 * RungeKuttaPatchStrategyB200.hpp -- device-resident "seam 2": the patch strategy the reference's level integrator calls,
 * on patch data that stays in HBM.
 *
 * Mirrors (path:line under the reference tree):
 *   RungeKuttaPatchStrategy::computeFluxesAndSourcesOnPatch / advanceSingleStepOnPatch /
 *       computeSpectralRadiusesAndStableDtOnPatch       include/algs/patch_strategy/RungeKuttaPatchStrategy.hpp:121-190
 *   their caller, the stage loop of advanceLevel          src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1672-1745
 *   the Euler implementation                              src/apps/Euler/Euler.cpp:489-900, 904-999, 1003-1679
 * Same method names, argument order and meaning.  What is different is where the data lives: registerPatchLevel() uploads
 * the conservative variables of every local patch ONCE (hb2_level_create / hb2_level_upload_patch); after that the stage
 * loop runs on the device copies -- computeFluxesAndSourcesOnPatch only records the stage (the fused kernels evaluate the
 * flux inside the update and never write it), advanceSingleStepOnPatch launches the fused stage of the patch,
 * fillGhostCellsOnLevel() stands where fill_schedule->fillData() stands in advanceLevel (same-level copies between the
 * registered patches, periodic images included, one kernel launch), and synchronizePatchToHost() brings a patch back when
 * the host needs it (output, regridding).  No arithmetic happens here; errors go through TBOX_ERROR.
 */
#ifndef HAMERS_B200_RUNGE_KUTTA_PATCH_STRATEGY_B200_HPP
#define HAMERS_B200_RUNGE_KUTTA_PATCH_STRATEGY_B200_HPP

#include "ConvectiveFluxReconstructorB200.hpp"

#include <map>
#include <string>
#include <vector>

class RungeKuttaPatchStrategyB200 {
public:
    /* convective_flux_reconstructor: "WCNS5_JS_HLLC_HLL" | "WCNS5_Z_HLLC_HLL" | "WCNS6_LD_HLLC_HLL" (the input key of
     * ConvectiveFluxReconstructorManager.cpp:37-48); math: HB2_MATH_EXACT | HB2_MATH_FAST */
    RungeKuttaPatchStrategyB200(const std::string& object_name, const tbox::Dimension& dim, const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                                const std::string& convective_flux_reconstructor, int math);
    ~RungeKuttaPatchStrategyB200();

    /* Register the local patches of one level (boxes in level index space; all with the same dx) and upload their
     * conservative variables from `data_context`.  domain_cells / periodic: the level's index box and periodic directions
     * (CartesianGeometry{domain_boxes, periodic_dimension}). */
    void registerPatchLevel(const std::vector<HAMERS_SHARED_PTR<hier::Patch> >& patches, const hier::IntVector& domain_cells,
                            const hier::IntVector& periodic, const HAMERS_SHARED_PTR<hier::VariableContext>& data_context);

    int getNumberOfSpectralRadiuses() const { return d_dim.getValue(); }
    /* level-wide form of computeSpectralRadiusesAndStableDtOnPatch: spectral radii per direction, then the stable dt for
     * CFL = 1 (1 / (max sum of radii + HAMERS_EPSILON), Euler.cpp:846-861) over the registered patches */
    std::vector<double> computeSpectralRadiusesAndStableDtOnLevel();

    /* stands where fill_schedule(_intermediate)->fillData() stands: ghosts of the intermediate state RK_step_number */
    void fillGhostCellsOnLevel(const int RK_step_number);

    void computeFluxesAndSourcesOnPatch(hier::Patch& patch, const double time, const double dt, const int RK_step_number,
                                        const HAMERS_SHARED_PTR<hier::VariableContext>& data_context = HAMERS_SHARED_PTR<hier::VariableContext>());

    void advanceSingleStepOnPatch(hier::Patch& patch, const double time, const double dt, const std::vector<double>& alpha,
                                  const std::vector<double>& beta, const std::vector<double>& gamma,
                                  const std::vector<HAMERS_SHARED_PTR<hier::VariableContext> >& intermediate_context);

    /* once per stage, after the patch loop (the scratch -> intermediate hand-over of copyTimeDependentData, :1681) */
    void finishStageOnLevel(const std::vector<double>& alpha, const bool last_stage);

    /* device -> host: the patch's current state into the patch data of `data_context` */
    void synchronizePatchToHost(hier::Patch& patch, const HAMERS_SHARED_PTR<hier::VariableContext>& data_context);

    long long getNumberOfKernelLaunches() const;

private:
    int patchIndex(const hier::Patch& patch) const;
    void gather(hier::Patch& patch, const HAMERS_SHARED_PTR<hier::VariableContext>& ctx, std::vector<double*>& ptrs) const;

    std::string d_object_name;
    tbox::Dimension d_dim;
    HAMERS_SHARED_PTR<FlowModel> d_flow_model;
    int d_scheme, d_math;
    hb2_level_t d_level;
    std::map<std::vector<int>, int> d_patch_of_box;        /* lower corner -> patch index */
    int d_stage_recorded;
    double d_dt_recorded;
};

#endif
