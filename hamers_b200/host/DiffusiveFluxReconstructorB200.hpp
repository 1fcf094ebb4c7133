/*
 * DiffusiveFluxReconstructorB200.hpp -- host-side mirror of the reference's diffusive-flux reconstructor interface
 * (SURVEY.md row f4), marshalling SAMRAI patch data into the C ABI of include/hamers_b200.h.
 *
 * Mirrors (path:line under the reference tree):
 *   DiffusiveFluxReconstructor                include/flow/diffusive_flux_reconstructors/DiffusiveFluxReconstructor.hpp
 *   DiffusiveFluxReconstructorNodeSixthOrder  src/flow/diffusive_flux_reconstructors/node/
 *                                             DiffusiveFluxReconstructorNodeSixthOrder.cpp:7-63 (ctor, print, restart)
 *   computeDiffusiveFluxOnPatch               .../node/DiffusiveFluxReconstructorNode.cpp:31-1736
 *   input keys                                EquationOfStateMixingRulesIdealGas.cpp:67-119 (species_R -> c_v),
 *                                             EquationOfShearViscosityMixingRulesConstant.cpp:24-48 (species_mu),
 *                                             EquationOfBulkViscosityMixingRulesConstant.cpp:24-48 (species_mu_v),
 *                                             EquationOfThermalConductivityMixingRulesPrandtl.cpp (species_c_p, species_Pr)
 *
 * Same class / method names and argument meaning as the reference ("SIXTH_ORDER" in DiffusiveFluxReconstructorManager.cpp:
 * 73-80); errors go through TBOX_ERROR.  No arithmetic happens here.
 */
#ifndef HAMERS_B200_DIFFUSIVE_FLUX_RECONSTRUCTOR_B200_HPP
#define HAMERS_B200_DIFFUSIVE_FLUX_RECONSTRUCTOR_B200_HPP

#include "ConvectiveFluxReconstructorB200.hpp"

class DiffusiveFluxReconstructor {
public:
    DiffusiveFluxReconstructor(const std::string& object_name, const tbox::Dimension& dim,
                               const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry, const int& num_eqn,
                               const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                               const HAMERS_SHARED_PTR<tbox::Database>& diffusive_flux_reconstructor_db)
        : d_object_name(object_name), d_dim(dim), d_grid_geometry(grid_geometry), d_num_diff_ghosts(hier::IntVector::getZero(d_dim)),
          d_num_eqn(num_eqn), d_flow_model(flow_model), d_diffusive_flux_reconstructor_db(diffusive_flux_reconstructor_db)
    {
    }
    virtual ~DiffusiveFluxReconstructor() {}
    hier::IntVector getDiffusiveFluxNumberOfGhostCells(void) const { return d_num_diff_ghosts; }
    virtual void printClassData(std::ostream& os) const = 0;
    virtual void putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const = 0;
    virtual void computeDiffusiveFluxOnPatch(hier::Patch& patch,
                                             const HAMERS_SHARED_PTR<pdat::SideVariable<double> >& variable_diffusive_flux,
                                             const HAMERS_SHARED_PTR<hier::VariableContext>& data_context, const double time,
                                             const double dt, const int RK_step_number) = 0;

protected:
    const std::string d_object_name;
    const tbox::Dimension d_dim;
    const HAMERS_SHARED_PTR<geom::CartesianGridGeometry> d_grid_geometry;
    hier::IntVector d_num_diff_ghosts;
    const int d_num_eqn;
    const HAMERS_SHARED_PTR<FlowModel> d_flow_model;
    const HAMERS_SHARED_PTR<tbox::Database> d_diffusive_flux_reconstructor_db;
};

class DiffusiveFluxReconstructorNodeSixthOrder_B200 : public DiffusiveFluxReconstructor {
public:
    DiffusiveFluxReconstructorNodeSixthOrder_B200(const std::string& object_name, const tbox::Dimension& dim,
                                                  const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry,
                                                  const int& num_eqn, const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                                                  const HAMERS_SHARED_PTR<tbox::Database>& diffusive_flux_reconstructor_db);
    ~DiffusiveFluxReconstructorNodeSixthOrder_B200();

    virtual void printClassData(std::ostream& os) const;
    virtual void putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const;

    /* Fully overwrites diffusive_flux (ghost 0, already multiplied by dt) on faces 0..N of every direction; the
     * conservative variables (SIX ghost cells, all filled by the caller) are read-only.  Host pointers: the call uploads,
     * computes on the GPU and downloads. */
    void computeDiffusiveFluxOnPatch(hier::Patch& patch, const HAMERS_SHARED_PTR<pdat::SideVariable<double> >& variable_diffusive_flux,
                                     const HAMERS_SHARED_PTR<hier::VariableContext>& data_context, const double time, const double dt,
                                     const int RK_step_number);

private:
    hb2_diff_plan_t getPlan(const hier::Patch& patch);

    double d_species_gamma, d_species_c_v, d_species_mu, d_species_mu_v, d_species_c_p, d_species_Pr;
    std::map<std::vector<double>, hb2_diff_plan_t> d_plans; /* keyed by (n, dx) */
};

#endif
