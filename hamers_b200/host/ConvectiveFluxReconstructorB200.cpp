/*
 * ConvectiveFluxReconstructorB200.cpp -- see the header.  Marshalling only: every number is produced by
 * libhamers_b200.so on the GPU (there is no CPU fallback; a missing device surfaces as TBOX_ERROR).
 */
#include "ConvectiveFluxReconstructorB200.hpp"

#include <cstring>

FlowModel::FlowModel(const std::string& object_name, const tbox::Dimension& dim, const FLOW_MODEL::TYPE& type, int num_species,
                     const HAMERS_SHARED_PTR<tbox::Database>& flow_model_db)
    : d_object_name(object_name), d_dim(dim), d_type(type), d_num_species(num_species), d_num_eqn(0), d_flow_model_db(flow_model_db)
{
    const int d = dim.getValue();
    /* Flow_model { Equation_of_state_mixing_rules { species_gamma = ... } } */
    d_species_gamma = flow_model_db->getDoubleVector("species_gamma");
    if ((int)d_species_gamma.size() != num_species)
        TBOX_ERROR(d_object_name << ": number of 'species_gamma' entries is not equal to the number of species." << std::endl);
    if (type == FLOW_MODEL::SINGLE_SPECIES) {
        if (num_species != 1) TBOX_ERROR(d_object_name << ": single-species flow model needs num_species = 1." << std::endl);
        d_num_eqn = d + 2; /* FlowModelSingleSpecies.cpp:29 */
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "density", 1)));
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "momentum", d)));
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "total energy", 1)));
    } else if (type == FLOW_MODEL::FIVE_EQN_ALLAIRE) {
        d_num_eqn = d + 2 * num_species; /* FlowModelFiveEqnAllaire.cpp:29 */
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "partial densities", num_species)));
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "momentum", d)));
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "total energy", 1)));
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "volume fractions", num_species)));
    } else if (type == FLOW_MODEL::FOUR_EQN_CONSERVATIVE) {
        /* FlowModelFourEqnConservative.cpp:29, registerConservativeVariables :608-645; species_R is read where
         * the plans are created */
        d_num_eqn = d + 1 + num_species;
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "partial densities", num_species)));
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "momentum", d)));
        d_cons.push_back(HAMERS_SHARED_PTR<pdat::CellVariable<double> >(new pdat::CellVariable<double>(dim, "total energy", 1)));
    } else {
        TBOX_ERROR(d_object_name << ": unknown flow model." << std::endl);
    }
}

int FlowModel::getNumberOfStoredComponents() const
{
    int n = 0;
    for (size_t v = 0; v < d_cons.size(); v++) n += d_cons[v]->getDepth();
    return n;
}

ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200(
    const std::string& object_name, const tbox::Dimension& dim, const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry,
    const int& num_eqn, const FLOW_MODEL::TYPE& flow_model_type, const HAMERS_SHARED_PTR<FlowModel>& flow_model,
    const HAMERS_SHARED_PTR<tbox::Database>& convective_flux_reconstructor_db)
    : ConvectiveFluxReconstructor(object_name, dim, grid_geometry, num_eqn, flow_model_type, flow_model, convective_flux_reconstructor_db),
      d_scheme(HB2_WCNS5_JS), d_constant_q(4), d_constant_C(1.0e9), d_constant_alpha_tau(35.0), d_math(HB2_MATH_EXACT)
{
    /* ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:22: four ghost cells in every direction */
    d_num_conv_ghosts = hier::IntVector::getOne(d_dim) * HB2_GHOSTS;
    /* ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:188-191 */
    d_constant_p = d_convective_flux_reconstructor_db->getIntegerWithDefault("constant_p", 2);
    d_constant_p = d_convective_flux_reconstructor_db->getIntegerWithDefault("d_constant_p", d_constant_p);
    if (num_eqn != flow_model->getNumberOfEquations())
        TBOX_ERROR(d_object_name << ": num_eqn does not match the flow model." << std::endl);
    if (dim.getValue() < 2)
        TBOX_ERROR(d_object_name << ": the 1D branch of WCNS56 has no shock sensor and is not on the B200 path." << std::endl);
}

ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::~ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200()
{
    for (std::map<std::vector<double>, hb2_plan_t>::iterator it = d_plans.begin(); it != d_plans.end(); ++it) hb2_plan_destroy(it->second);
}

void ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::printClassData(std::ostream& os) const
{
    os << "\nPrint ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200 object..." << std::endl;
    os << std::endl;
    os << "ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200: this = " << (const void*)this << std::endl;
    os << "d_object_name = " << d_object_name << std::endl;
    os << "d_constant_p = " << d_constant_p << std::endl;
    os << "backend = " << hb2_version() << std::endl;
}

void ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const
{
    restart_db->putInteger("d_constant_p", d_constant_p);
}

ConvectiveFluxReconstructorWCNS6_LD_HLLC_HLL_B200::ConvectiveFluxReconstructorWCNS6_LD_HLLC_HLL_B200(
    const std::string& object_name, const tbox::Dimension& dim, const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry,
    const int& num_eqn, const FLOW_MODEL::TYPE& flow_model_type, const HAMERS_SHARED_PTR<FlowModel>& flow_model,
    const HAMERS_SHARED_PTR<tbox::Database>& convective_flux_reconstructor_db)
    : ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200(object_name, dim, grid_geometry, num_eqn, flow_model_type, flow_model,
                                                        convective_flux_reconstructor_db)
{
    d_scheme = HB2_WCNS6_LD;
    /* ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp:343-361 */
    d_constant_q = d_convective_flux_reconstructor_db->getIntegerWithDefault("constant_q", 4);
    d_constant_q = d_convective_flux_reconstructor_db->getIntegerWithDefault("d_constant_q", d_constant_q);
    d_constant_C = d_convective_flux_reconstructor_db->getDoubleWithDefault("constant_C", 1.0e9);
    d_constant_C = d_convective_flux_reconstructor_db->getDoubleWithDefault("d_constant_C", d_constant_C);
    d_constant_alpha_tau = d_convective_flux_reconstructor_db->getDoubleWithDefault("constant_alpha_tau", 35.0);
    d_constant_alpha_tau = d_convective_flux_reconstructor_db->getDoubleWithDefault("d_constant_alpha_tau", d_constant_alpha_tau);
}

void ConvectiveFluxReconstructorWCNS6_LD_HLLC_HLL_B200::printClassData(std::ostream& os) const
{
    ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::printClassData(os);
    os << "d_constant_q = " << d_constant_q << std::endl;
    os << "d_constant_C = " << d_constant_C << std::endl;
    os << "d_constant_alpha_tau = " << d_constant_alpha_tau << std::endl;
}

void ConvectiveFluxReconstructorWCNS6_LD_HLLC_HLL_B200::putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const
{
    /* ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp:406-409 */
    restart_db->putInteger("d_constant_p", d_constant_p);
    restart_db->putInteger("d_constant_q", d_constant_q);
    restart_db->putDouble("d_constant_C", d_constant_C);
    restart_db->putDouble("d_constant_alpha_tau", d_constant_alpha_tau);
}

hb2_plan_t ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::getPlan(const hier::Patch& patch, int num_ghosts)
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
    key.push_back(d_math);
    key.push_back(d_scheme);
    key.push_back(num_ghosts);
    std::map<std::vector<double>, hb2_plan_t>::iterator it = d_plans.find(key);
    if (it != d_plans.end()) return it->second;

    hb2_patch_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    desc.dim = dim;
    for (int a = 0; a < 3; a++) {
        desc.n[a] = a < dim ? interior_dims[a] : 1;
        desc.dx[a] = a < dim ? dx[a] : 1.0;
    }
    desc.flow_model = d_flow_model_type == FLOW_MODEL::SINGLE_SPECIES
                          ? HB2_SINGLE_SPECIES
                          : (d_flow_model_type == FLOW_MODEL::FIVE_EQN_ALLAIRE ? HB2_FIVE_EQN_ALLAIRE : HB2_FOUR_EQN_CONSERVATIVE);
    desc.num_species = d_flow_model->getNumberOfSpecies();
    for (int s = 0; s < desc.num_species && s < HB2_MAX_SPECIES; s++) desc.species_gamma[s] = d_flow_model->getSpeciesGamma()[s];
    if (desc.flow_model == HB2_FOUR_EQN_CONSERVATIVE) {
        /* Equation_of_state_mixing_rules { species_R } (EquationOfStateMixingRulesIdealGas.cpp:60-100) */
        const std::vector<double> R = d_flow_model->getFlowModelDatabase()->getDoubleVector("species_R");
        for (int s = 0; s < desc.num_species && s < HB2_MAX_SPECIES && s < (int)R.size(); s++) desc.species_R[s] = R[s];
    }
    desc.weno_p = d_constant_p;
    desc.scheme = d_scheme;
    desc.weno_q = d_constant_q;
    desc.weno_C = d_constant_C;
    desc.weno_alpha_tau = d_constant_alpha_tau;
    desc.math = d_math;
    desc.device = -1;
    desc.num_ghosts = num_ghosts;
    hb2_plan_t plan = 0;
    if (hb2_plan_create(&desc, &plan) != 0) TBOX_ERROR(d_object_name << ": " << hb2_last_error() << std::endl);
    d_plans[key] = plan;
    return plan;
}

int ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::gatherConservative(hier::Patch& patch,
                                                                        const HAMERS_SHARED_PTR<hier::VariableContext>& ctx,
                                                                        std::vector<double*>& ptrs) const
{
    const std::vector<HAMERS_SHARED_PTR<pdat::CellVariable<double> > >& vars = d_flow_model->getConservativeVariables();
    int ghosts = -1;
    for (size_t v = 0; v < vars.size(); v++) {
        HAMERS_SHARED_PTR<pdat::CellData<double> > data(
            HAMERS_SHARED_PTR_CAST<pdat::CellData<double>, hier::PatchData>(patch.getPatchData(vars[v], ctx)));
        if (!data) TBOX_ERROR(d_object_name << ": conservative variable '" << vars[v]->getName() << "' is not cell data." << std::endl);
        /* at least the four layers this reconstructor reads, the same width in every direction and variable */
        const hier::IntVector gw = data->getGhostCellWidth();
        const int g = gw[0];
        bool uniform = true;
        for (int a = 0; a < d_dim.getValue(); a++) uniform = uniform && gw[a] == g;
        if (!uniform || g < HB2_GHOSTS || g > 8 || (ghosts >= 0 && g != ghosts))
            TBOX_ERROR(d_object_name << ": conservative variables need the same ghost width, at least " << HB2_GHOSTS
                                     << " (at most 8), in every direction." << std::endl);
        ghosts = g;
        for (int d = 0; d < data->getDepth(); d++) ptrs.push_back(data->getPointer(d));
    }
    return ghosts;
}

void ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::computeConvectiveFluxAndSourceOnPatch(
    hier::Patch& patch, const HAMERS_SHARED_PTR<pdat::SideVariable<double> >& variable_convective_flux,
    const HAMERS_SHARED_PTR<pdat::CellVariable<double> >& variable_source, const HAMERS_SHARED_PTR<hier::VariableContext>& data_context,
    const double time, const double dt, const int RK_step_number)
{
    NULL_USE(time);
    NULL_USE(RK_step_number);
    const int dim = d_dim.getValue();

    HAMERS_SHARED_PTR<pdat::SideData<double> > convective_flux(
        HAMERS_SHARED_PTR_CAST<pdat::SideData<double>, hier::PatchData>(patch.getPatchData(variable_convective_flux, data_context)));
    HAMERS_SHARED_PTR<pdat::CellData<double> > source(
        HAMERS_SHARED_PTR_CAST<pdat::CellData<double>, hier::PatchData>(patch.getPatchData(variable_source, data_context)));
    TBOX_ASSERT(convective_flux);
    TBOX_ASSERT(convective_flux->getGhostCellWidth() == hier::IntVector::getZero(d_dim));
    TBOX_ASSERT(convective_flux->getDepth() == d_num_eqn);
    TBOX_ASSERT(source);
    TBOX_ASSERT(source->getGhostCellWidth() == hier::IntVector::getZero(d_dim));
    TBOX_ASSERT(source->getDepth() == d_num_eqn);

    std::vector<double*> Q;
    hb2_plan_t plan = getPlan(patch, gatherConservative(patch, data_context, Q));
    std::vector<double*> F, S;
    for (int n = 0; n < dim; n++)
        for (int e = 0; e < d_num_eqn; e++) F.push_back(convective_flux->getPointer(n, e));
    for (int e = 0; e < d_num_eqn; e++) S.push_back(source->getPointer(e));

    if (hb2_compute_flux_and_source_host(plan, (const double* const*)Q.data(), dt, F.data(), S.data()) != 0)
        TBOX_ERROR(d_object_name << ": " << hb2_last_error() << std::endl);
}

void ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200::advanceFusedStageOnPatch(
    hier::Patch& patch, const double dt, const std::vector<double>& alpha, const std::vector<double>& beta,
    const std::vector<HAMERS_SHARED_PTR<hier::VariableContext> >& intermediate_contexts,
    const HAMERS_SHARED_PTR<hier::VariableContext>& output_context)
{
    const size_t ncoef = alpha.size();
    if (beta.size() != ncoef || intermediate_contexts.size() != ncoef || ncoef == 0)
        TBOX_ERROR(d_object_name << ": alpha, beta and the intermediate contexts must have the same, non-zero length." << std::endl);
    std::vector<double*> U_int, U_out;
    int ghosts = -1;
    for (size_t m = 0; m < ncoef; m++) {
        const int g = gatherConservative(patch, intermediate_contexts[m], U_int);
        if (ghosts >= 0 && g != ghosts) TBOX_ERROR(d_object_name << ": the intermediate states differ in ghost width." << std::endl);
        ghosts = g;
    }
    if (gatherConservative(patch, output_context, U_out) != ghosts)
        TBOX_ERROR(d_object_name << ": the output state differs in ghost width from the intermediate states." << std::endl);
    hb2_plan_t plan = getPlan(patch, ghosts);
    if (hb2_fused_stage_host(plan, (int32_t)ncoef, alpha.data(), beta.data(), (const double* const*)U_int.data(), dt, U_out.data()) != 0)
        TBOX_ERROR(d_object_name << ": " << hb2_last_error() << std::endl);
}
