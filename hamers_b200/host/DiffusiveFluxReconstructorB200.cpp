/* DiffusiveFluxReconstructorB200.cpp -- see the header.  Marshalling only. */
#include "DiffusiveFluxReconstructorB200.hpp"

#include <cstring>

namespace {
/* `key` as in an input file, `d_key` as putToRestart of the reference's mixing-rule classes writes it */
double first_of(const HAMERS_SHARED_PTR<tbox::Database>& db, const std::string& key, const std::string& who)
{
    const std::string k = db->keyExists(key) ? key : "d_" + key;
    if (!db->keyExists(k)) TBOX_ERROR(who << ": key '" << key << "' not found in the flow model database." << std::endl);
    const std::vector<double> v = db->getDoubleVector(k);
    if (v.size() != 1) TBOX_ERROR(who << ": '" << key << "' must have one entry (single-species flow model)." << std::endl);
    return v[0];
}
}  // namespace

DiffusiveFluxReconstructorNodeSixthOrder_B200::DiffusiveFluxReconstructorNodeSixthOrder_B200(
    const std::string& object_name, const tbox::Dimension& dim, const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry,
    const int& num_eqn, const HAMERS_SHARED_PTR<FlowModel>& flow_model,
    const HAMERS_SHARED_PTR<tbox::Database>& diffusive_flux_reconstructor_db)
    : DiffusiveFluxReconstructor(object_name, dim, grid_geometry, num_eqn, flow_model, diffusive_flux_reconstructor_db)
{
    /* DiffusiveFluxReconstructorNodeSixthOrder.cpp:24 */
    d_num_diff_ghosts = hier::IntVector::getOne(d_dim) * HB2_DIFF_GHOSTS;
    if (d_flow_model->getType() != FLOW_MODEL::SINGLE_SPECIES)
        TBOX_ERROR(d_object_name << ": the B200 diffusive-flux path is built for the SINGLE_SPECIES flow model." << std::endl);
    const HAMERS_SHARED_PTR<tbox::Database>& db = d_flow_model->getFlowModelDatabase();
    d_species_gamma = d_flow_model->getSpeciesGamma()[0];
    /* EquationOfStateMixingRulesIdealGas.cpp:119: c_v = 1/(gamma - 1)*R */
    d_species_c_v = 1.0 / (d_species_gamma - 1.0) * first_of(db, "species_R", d_object_name);
    d_species_mu = first_of(db, "species_mu", d_object_name);
    d_species_mu_v = first_of(db, "species_mu_v", d_object_name);
    d_species_c_p = first_of(db, "species_c_p", d_object_name);
    d_species_Pr = first_of(db, "species_Pr", d_object_name);
}

DiffusiveFluxReconstructorNodeSixthOrder_B200::~DiffusiveFluxReconstructorNodeSixthOrder_B200()
{
    for (std::map<std::vector<double>, hb2_diff_plan_t>::iterator it = d_plans.begin(); it != d_plans.end(); ++it)
        hb2_diffusive_plan_destroy(it->second);
}

void DiffusiveFluxReconstructorNodeSixthOrder_B200::printClassData(std::ostream& os) const
{
    os << "\nPrint DiffusiveFluxReconstructorNodeSixthOrder_B200 object..." << std::endl;
    os << std::endl;
    os << "DiffusiveFluxReconstructorNodeSixthOrder_B200: this = " << (const void*)this << std::endl;
    os << "d_object_name = " << d_object_name << std::endl;
    os << "backend = " << hb2_version() << std::endl;
}

void DiffusiveFluxReconstructorNodeSixthOrder_B200::putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const
{
    /* DiffusiveFluxReconstructorNodeSixthOrder.cpp:58 */
    restart_db->putString("d_diffusive_flux_reconstructor", "SIXTH_ORDER");
}

hb2_diff_plan_t DiffusiveFluxReconstructorNodeSixthOrder_B200::getPlan(const hier::Patch& patch)
{
    const int dim = d_dim.getValue();
    const hier::IntVector interior_dims = patch.getBox().numberCells();
    const HAMERS_SHARED_PTR<geom::CartesianPatchGeometry> patch_geom(
        HAMERS_SHARED_PTR_CAST<geom::CartesianPatchGeometry, hier::PatchGeometry>(patch.getPatchGeometry()));
    if (!patch_geom) TBOX_ERROR(d_object_name << ": patch has no Cartesian patch geometry." << std::endl);
    const double* const dx = patch_geom->getDx();
    std::vector<double> key;
    for (int a = 0; a < dim; a++) {
        key.push_back(interior_dims[a]);
        key.push_back(dx[a]);
    }
    std::map<std::vector<double>, hb2_diff_plan_t>::iterator it = d_plans.find(key);
    if (it != d_plans.end()) return it->second;

    hb2_diffusive_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    desc.dim = dim;
    for (int a = 0; a < 3; a++) {
        desc.n[a] = a < dim ? interior_dims[a] : 1;
        desc.dx[a] = a < dim ? dx[a] : 1.0;
    }
    desc.species_gamma = d_species_gamma;
    desc.species_c_v = d_species_c_v;
    desc.species_mu = d_species_mu;
    desc.species_mu_v = d_species_mu_v;
    desc.species_c_p = d_species_c_p;
    desc.species_Pr = d_species_Pr;
    desc.device = -1;
    hb2_diff_plan_t plan = 0;
    if (hb2_diffusive_plan_create(&desc, &plan) != 0) TBOX_ERROR(d_object_name << ": " << hb2_last_error() << std::endl);
    d_plans[key] = plan;
    return plan;
}

void DiffusiveFluxReconstructorNodeSixthOrder_B200::computeDiffusiveFluxOnPatch(
    hier::Patch& patch, const HAMERS_SHARED_PTR<pdat::SideVariable<double> >& variable_diffusive_flux,
    const HAMERS_SHARED_PTR<hier::VariableContext>& data_context, const double time, const double dt, const int RK_step_number)
{
    NULL_USE(time);
    NULL_USE(RK_step_number);
    const int dim = d_dim.getValue();
    hb2_diff_plan_t plan = getPlan(patch);

    HAMERS_SHARED_PTR<pdat::SideData<double> > diffusive_flux(
        HAMERS_SHARED_PTR_CAST<pdat::SideData<double>, hier::PatchData>(patch.getPatchData(variable_diffusive_flux, data_context)));
    TBOX_ASSERT(diffusive_flux);
    TBOX_ASSERT(diffusive_flux->getGhostCellWidth() == hier::IntVector::getZero(d_dim));
    TBOX_ASSERT(diffusive_flux->getDepth() == d_num_eqn);

    std::vector<double*> Q;
    const std::vector<HAMERS_SHARED_PTR<pdat::CellVariable<double> > >& vars = d_flow_model->getConservativeVariables();
    for (size_t v = 0; v < vars.size(); v++) {
        HAMERS_SHARED_PTR<pdat::CellData<double> > data(
            HAMERS_SHARED_PTR_CAST<pdat::CellData<double>, hier::PatchData>(patch.getPatchData(vars[v], data_context)));
        if (!data) TBOX_ERROR(d_object_name << ": conservative variable '" << vars[v]->getName() << "' is not cell data." << std::endl);
        if (!(data->getGhostCellWidth() == d_num_diff_ghosts))
            TBOX_ERROR(d_object_name << ": conservative variables need " << HB2_DIFF_GHOSTS << " ghost cells." << std::endl);
        for (int d = 0; d < data->getDepth(); d++) Q.push_back(data->getPointer(d));
    }
    std::vector<double*> F;
    for (int n = 0; n < dim; n++)
        for (int e = 0; e < d_num_eqn; e++) F.push_back(diffusive_flux->getPointer(n, e));
    if (hb2_compute_diffusive_flux_host(plan, (const double* const*)Q.data(), dt, F.data()) != 0)
        TBOX_ERROR(d_object_name << ": " << hb2_last_error() << std::endl);
}
