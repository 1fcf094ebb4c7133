/* RungeKuttaPatchStrategyB200.cpp -- see the header.  Marshalling only: every number is computed by the kernels behind
 * the hb2_level_* entry points of include/hamers_b200.h. */
#include "RungeKuttaPatchStrategyB200.hpp"

#include <cstring>

#define HB2_CHECK(call)                                                                              \
    do {                                                                                             \
        if ((call) != 0) TBOX_ERROR(d_object_name << ": " << #call << " failed: " << hb2_last_error() << std::endl); \
    } while (0)

RungeKuttaPatchStrategyB200::RungeKuttaPatchStrategyB200(const std::string& object_name, const tbox::Dimension& dim,
                                                         const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                                                         const std::string& convective_flux_reconstructor, int math)
    : d_object_name(object_name), d_dim(dim), d_flow_model(flow_model), d_scheme(HB2_WCNS5_JS), d_math(math), d_level(0),
      d_stage_recorded(-1), d_dt_recorded(0.0)
{
    if (convective_flux_reconstructor == "WCNS5_JS_HLLC_HLL")
        d_scheme = HB2_WCNS5_JS;
    else if (convective_flux_reconstructor == "WCNS5_Z_HLLC_HLL")
        d_scheme = HB2_WCNS5_Z;
    else if (convective_flux_reconstructor == "WCNS6_LD_HLLC_HLL")
        d_scheme = HB2_WCNS6_LD;
    else
        TBOX_ERROR(d_object_name << ": unknown convective_flux_reconstructor '" << convective_flux_reconstructor << "'" << std::endl);
}

RungeKuttaPatchStrategyB200::~RungeKuttaPatchStrategyB200()
{
    if (d_level) hb2_level_destroy(d_level);
}

void RungeKuttaPatchStrategyB200::gather(hier::Patch& patch, const HAMERS_SHARED_PTR<hier::VariableContext>& ctx,
                                         std::vector<double*>& ptrs) const
{
    const std::vector<HAMERS_SHARED_PTR<pdat::CellVariable<double> > >& cons = d_flow_model->getConservativeVariables();
    for (size_t v = 0; v < cons.size(); v++) {
        HAMERS_SHARED_PTR<pdat::CellData<double> > data(
            HAMERS_SHARED_PTR_CAST<pdat::CellData<double>, hier::PatchData>(patch.getPatchData(cons[v], ctx)));
        if (!data) TBOX_ERROR(d_object_name << ": conservative variable '" << cons[v]->getName() << "' is not cell data" << std::endl);
        const hier::IntVector g = data->getGhostCellWidth();
        for (int a = 0; a < d_dim.getValue(); a++)
            if (g[a] != HB2_GHOSTS) TBOX_ERROR(d_object_name << ": the conservative variables must carry " << HB2_GHOSTS << " ghost cells" << std::endl);
        for (int c = 0; c < data->getDepth(); c++) ptrs.push_back(data->getPointer(c));
    }
}

void RungeKuttaPatchStrategyB200::registerPatchLevel(const std::vector<HAMERS_SHARED_PTR<hier::Patch> >& patches,
                                                     const hier::IntVector& domain_cells, const hier::IntVector& periodic,
                                                     const HAMERS_SHARED_PTR<hier::VariableContext>& data_context)
{
    if (patches.empty()) TBOX_ERROR(d_object_name << ": registerPatchLevel() needs at least one patch" << std::endl);
    const int dim = d_dim.getValue();
    hb2_patch_desc model;
    memset(&model, 0, sizeof model);
    model.dim = dim;
    const FLOW_MODEL::TYPE type = d_flow_model->getType();
    model.flow_model = type == FLOW_MODEL::SINGLE_SPECIES ? HB2_SINGLE_SPECIES
                       : (type == FLOW_MODEL::FIVE_EQN_ALLAIRE ? HB2_FIVE_EQN_ALLAIRE : HB2_FOUR_EQN_CONSERVATIVE);
    model.num_species = d_flow_model->getNumberOfSpecies();
    for (int s = 0; s < model.num_species; s++) model.species_gamma[s] = d_flow_model->getSpeciesGamma()[s];
    if (model.flow_model == HB2_FOUR_EQN_CONSERVATIVE) {
        const std::vector<double> R = d_flow_model->getFlowModelDatabase()->getDoubleVector("species_R");
        for (int s = 0; s < model.num_species; s++) model.species_R[s] = R[s];
    }
    HAMERS_SHARED_PTR<geom::CartesianPatchGeometry> geom0(
        HAMERS_SHARED_PTR_CAST<geom::CartesianPatchGeometry, hier::PatchGeometry>(patches[0]->getPatchGeometry()));
    if (!geom0) TBOX_ERROR(d_object_name << ": patch without Cartesian geometry" << std::endl);
    for (int a = 0; a < dim; a++) model.dx[a] = geom0->getDx()[a];
    model.weno_p = 2;
    model.math = d_math;
    model.device = -1;
    model.scheme = d_scheme;
    std::vector<int32_t> lo(3 * patches.size(), 0), hi(3 * patches.size(), 1);
    for (size_t p = 0; p < patches.size(); p++) {
        const hier::Box& box = patches[p]->getBox();
        std::vector<int> key(3, 0);
        for (int a = 0; a < dim; a++) {
            lo[3 * p + a] = box.lower()[a];
            hi[3 * p + a] = box.upper()[a] + 1;
            key[a] = box.lower()[a];
        }
        d_patch_of_box[key] = (int)p;
    }
    int32_t level_n[3] = {1, 1, 1}, mask = 0;
    for (int a = 0; a < dim; a++) {
        level_n[a] = domain_cells[a];
        if (periodic[a]) mask |= 1 << a;
    }
    if (d_level) hb2_level_destroy(d_level);
    d_level = 0;
    HB2_CHECK(hb2_level_create(&model, (int32_t)patches.size(), lo.data(), hi.data(), level_n, mask, &d_level));
    for (size_t p = 0; p < patches.size(); p++) {
        std::vector<double*> ptrs;
        gather(*patches[p], data_context, ptrs);
        HB2_CHECK(hb2_level_upload_patch(d_level, (int32_t)p, ptrs.data()));
    }
}

int RungeKuttaPatchStrategyB200::patchIndex(const hier::Patch& patch) const
{
    std::vector<int> key(3, 0);
    for (int a = 0; a < d_dim.getValue(); a++) key[a] = patch.getBox().lower()[a];
    std::map<std::vector<int>, int>::const_iterator it = d_patch_of_box.find(key);
    if (it == d_patch_of_box.end()) TBOX_ERROR(d_object_name << ": the patch was not registered with registerPatchLevel()" << std::endl);
    return it->second;
}

std::vector<double> RungeKuttaPatchStrategyB200::computeSpectralRadiusesAndStableDtOnLevel()
{
    if (!d_level) TBOX_ERROR(d_object_name << ": no level registered" << std::endl);
    double sr[4];
    HB2_CHECK(hb2_level_max_wave_speed(d_level, sr));
    std::vector<double> out(sr, sr + d_dim.getValue());
    out.push_back(1.0 / (sr[3] + 1.0e-15));
    return out;
}

void RungeKuttaPatchStrategyB200::fillGhostCellsOnLevel(const int RK_step_number)
{
    if (!d_level) TBOX_ERROR(d_object_name << ": no level registered" << std::endl);
    HB2_CHECK(hb2_level_fill_ghosts(d_level, RK_step_number));
}

void RungeKuttaPatchStrategyB200::computeFluxesAndSourcesOnPatch(hier::Patch& patch, const double time, const double dt,
                                                                 const int RK_step_number,
                                                                 const HAMERS_SHARED_PTR<hier::VariableContext>& data_context)
{
    (void)time;
    (void)data_context;
    (void)patchIndex(patch);        /* must be a registered patch */
    /* fused path: the flux of U^(RK_step_number) is evaluated inside advanceSingleStepOnPatch and never written */
    d_stage_recorded = RK_step_number;
    d_dt_recorded = dt;
}

void RungeKuttaPatchStrategyB200::advanceSingleStepOnPatch(hier::Patch& patch, const double time, const double dt,
                                                           const std::vector<double>& alpha, const std::vector<double>& beta,
                                                           const std::vector<double>& gamma,
                                                           const std::vector<HAMERS_SHARED_PTR<hier::VariableContext> >& intermediate_context)
{
    (void)time;
    (void)gamma;                    /* flux sums for the AMR synchronisation: hamers_b200/amr.py route (materialised fluxes) */
    (void)intermediate_context;     /* the intermediate states are the level's device buffers */
    const int ncoef = (int)alpha.size();
    if ((int)beta.size() != ncoef) TBOX_ERROR(d_object_name << ": alpha and beta must have the same length" << std::endl);
    if (d_stage_recorded != ncoef - 1 || d_dt_recorded != dt)
        TBOX_ERROR(d_object_name << ": advanceSingleStepOnPatch() must follow computeFluxesAndSourcesOnPatch() of the same stage" << std::endl);
    HB2_CHECK(hb2_level_advance_stage_patch(d_level, patchIndex(patch), ncoef, alpha.data(), beta.data(), dt));
}

void RungeKuttaPatchStrategyB200::finishStageOnLevel(const std::vector<double>& alpha, const bool last_stage)
{
    HB2_CHECK(hb2_level_end_stage(d_level, (int32_t)alpha.size(), alpha.data(), last_stage ? 1 : 0));
}

void RungeKuttaPatchStrategyB200::synchronizePatchToHost(hier::Patch& patch, const HAMERS_SHARED_PTR<hier::VariableContext>& data_context)
{
    std::vector<double*> ptrs;
    gather(patch, data_context, ptrs);
    HB2_CHECK(hb2_level_download_patch(d_level, patchIndex(patch), ptrs.data()));
}

long long RungeKuttaPatchStrategyB200::getNumberOfKernelLaunches() const { return d_level ? hb2_level_launch_count(d_level) : 0; }
