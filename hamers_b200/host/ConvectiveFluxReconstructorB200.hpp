/*
 * ConvectiveFluxReconstructorB200.hpp -- host-side mirror of the reference's reconstructor interface for the
 * WCNS5-JS / HLLC-HLL path, marshalling SAMRAI patch data into the C ABI of include/hamers_b200.h.
 *
 * Mirrors (path:line under the reference tree):
 *   ConvectiveFluxReconstructor                 include/flow/convective_flux_reconstructors/ConvectiveFluxReconstructor.hpp:23-127
 *   ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL  src/flow/convective_flux_reconstructors/WCNS56/
 *                                               ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:166-227 (ctor, print, restart)
 *   computeConvectiveFluxAndSourceOnPatch       .../ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:39-2656
 *   FlowModel (the part the reconstructor uses)  src/flow/flow_models/single-species/FlowModelSingleSpecies.cpp:90-97
 *                                               (conservative variables: density, momentum, total energy),
 *                                               five-eqn_Allaire/FlowModelFiveEqnAllaire.cpp ctor
 *                                               (partial densities, momentum, total energy, volume fractions)
 *   Euler::advanceSingleStepOnPatch             src/apps/Euler/Euler.cpp:1003-1679 (fused variant: advanceFusedStageOnPatch)
 *
 * Same class / method names and argument meaning as the reference; errors go through TBOX_ERROR.  The class owns
 * one hb2 plan per patch shape (the reference's FlowModel is a per-app singleton that is not re-entrant either,
 * FlowModelSingleSpecies.cpp:706-712).  No arithmetic happens here.
 */
#ifndef HAMERS_B200_CONVECTIVE_FLUX_RECONSTRUCTOR_B200_HPP
#define HAMERS_B200_CONVECTIVE_FLUX_RECONSTRUCTOR_B200_HPP

#include "samrai_shim.hpp"

#include "../../include/hamers_b200.h"

#include <map>
#include <ostream>
#include <string>
#include <vector>

using namespace SAMRAI;

namespace FLOW_MODEL {
enum TYPE { SINGLE_SPECIES, FOUR_EQN_CONSERVATIVE, FIVE_EQN_ALLAIRE }; /* include/flow/flow_models/FlowModels.hpp */
}

/* The part of FlowModel the convective-flux path needs: equation count, species gammas and the registered
 * conservative cell variables, in the reference's order. */
class FlowModel {
public:
    FlowModel(const std::string& object_name, const tbox::Dimension& dim, const FLOW_MODEL::TYPE& type, int num_species,
              const HAMERS_SHARED_PTR<tbox::Database>& flow_model_db);

    int getNumberOfEquations() const { return d_num_eqn; }
    int getNumberOfSpecies() const { return d_num_species; }
    FLOW_MODEL::TYPE getType() const { return d_type; }
    const std::vector<double>& getSpeciesGamma() const { return d_species_gamma; }
    /* density | momentum | total energy   or   partial densities | momentum | total energy | volume fractions */
    const std::vector<HAMERS_SHARED_PTR<pdat::CellVariable<double> > >& getConservativeVariables() const { return d_cons; }
    /* number of stored components (the five-eqn model stores all ns volume fractions) */
    int getNumberOfStoredComponents() const;
    /* the Flow_model input block (equation of state and transport mixing-rule keys) */
    const HAMERS_SHARED_PTR<tbox::Database>& getFlowModelDatabase() const { return d_flow_model_db; }

private:
    std::string d_object_name;
    tbox::Dimension d_dim;
    FLOW_MODEL::TYPE d_type;
    int d_num_species, d_num_eqn;
    std::vector<double> d_species_gamma;
    std::vector<HAMERS_SHARED_PTR<pdat::CellVariable<double> > > d_cons;
    HAMERS_SHARED_PTR<tbox::Database> d_flow_model_db;
};

class ConvectiveFluxReconstructor {
public:
    ConvectiveFluxReconstructor(const std::string& object_name, const tbox::Dimension& dim,
                                const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry, const int& num_eqn,
                                const FLOW_MODEL::TYPE& flow_model_type, const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                                const HAMERS_SHARED_PTR<tbox::Database>& convective_flux_reconstructor_db)
        : d_object_name(object_name), d_dim(dim), d_grid_geometry(grid_geometry), d_num_conv_ghosts(hier::IntVector::getZero(d_dim)),
          d_num_eqn(num_eqn), d_flow_model_type(flow_model_type), d_flow_model(flow_model),
          d_convective_flux_reconstructor_db(convective_flux_reconstructor_db)
    {
    }
    virtual ~ConvectiveFluxReconstructor() {}
    hier::IntVector getConvectiveFluxNumberOfGhostCells(void) const { return d_num_conv_ghosts; }
    virtual void printClassData(std::ostream& os) const = 0;
    virtual void putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const = 0;
    virtual void computeConvectiveFluxAndSourceOnPatch(hier::Patch& patch,
                                                       const HAMERS_SHARED_PTR<pdat::SideVariable<double> >& variable_convective_flux,
                                                       const HAMERS_SHARED_PTR<pdat::CellVariable<double> >& variable_source,
                                                       const HAMERS_SHARED_PTR<hier::VariableContext>& data_context, const double time,
                                                       const double dt, const int RK_step_number) = 0;

protected:
    const std::string d_object_name;
    const tbox::Dimension d_dim;
    const HAMERS_SHARED_PTR<geom::CartesianGridGeometry> d_grid_geometry;
    hier::IntVector d_num_conv_ghosts;
    const int d_num_eqn;
    const FLOW_MODEL::TYPE d_flow_model_type;
    const HAMERS_SHARED_PTR<FlowModel> d_flow_model;
    const HAMERS_SHARED_PTR<tbox::Database> d_convective_flux_reconstructor_db;
};

class ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200 : public ConvectiveFluxReconstructor {
public:
    ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200(const std::string& object_name, const tbox::Dimension& dim,
                                                      const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry,
                                                      const int& num_eqn, const FLOW_MODEL::TYPE& flow_model_type,
                                                      const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                                                      const HAMERS_SHARED_PTR<tbox::Database>& convective_flux_reconstructor_db);
    ~ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200();

    virtual void printClassData(std::ostream& os) const;
    virtual void putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const;

    /* Fully overwrites convective_flux (ghost 0, already multiplied by dt) on faces 0..N of every direction and "+="s
     * the advective-equation entries of source (ghost 0); the conservative variables (ghost 4, filled by the caller)
     * are read-only.  Host pointers: the call uploads, computes on the GPU and downloads. */
    void computeConvectiveFluxAndSourceOnPatch(hier::Patch& patch,
                                               const HAMERS_SHARED_PTR<pdat::SideVariable<double> >& variable_convective_flux,
                                               const HAMERS_SHARED_PTR<pdat::CellVariable<double> >& variable_source,
                                               const HAMERS_SHARED_PTR<hier::VariableContext>& data_context, const double time,
                                               const double dt, const int RK_step_number);

    /* RungeKuttaPatchStrategy::computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch (Euler.cpp:904-1679) in one
     * pass for the "newest flux only" RK tables: U_out = sum_m alpha[m] U^(m) + beta[n] L(U^(n)), n = num_coeffs - 1.
     * intermediate_contexts[m] hold the ghost-filled U^(m); the result goes to the interior of the data in
     * output_context. */
    void advanceFusedStageOnPatch(hier::Patch& patch, const double dt, const std::vector<double>& alpha,
                                  const std::vector<double>& beta,
                                  const std::vector<HAMERS_SHARED_PTR<hier::VariableContext> >& intermediate_contexts,
                                  const HAMERS_SHARED_PTR<hier::VariableContext>& output_context);

    /* arithmetic variant of the plans created from now on: HB2_MATH_EXACT (default, bit-identical to the reference's
     * operation order) or HB2_MATH_FAST */
    void setMathMode(int math) { d_math = math; }

protected:
    /* which nonlinear interpolator the plans use (HB2_WCNS5_JS here; the subclasses below set the others) */
    int d_scheme;
    int d_constant_p;
    int d_constant_q;
    double d_constant_C;
    double d_constant_alpha_tau;

private:
    /* one plan per (patch shape, ghost width of the state arrays): an application that allocates the state with more
     * ghost cells than the four this reconstructor reads -- Navier-Stokes: six -- passes the same arrays */
    hb2_plan_t getPlan(const hier::Patch& patch, int num_ghosts);
    /* appends the component pointers; returns the ghost width of the data (uniform over variables and directions) */
    int gatherConservative(hier::Patch& patch, const HAMERS_SHARED_PTR<hier::VariableContext>& ctx, std::vector<double*>& ptrs) const;

    int d_math;
    std::map<std::vector<double>, hb2_plan_t> d_plans; /* keyed by (n, dx) */
};

/* "WCNS5_Z_HLLC_HLL" (ConvectiveFluxReconstructorWCNS5-Z-HLLC-HLL.cpp): same data flow, WCNS5-Z weights. */
class ConvectiveFluxReconstructorWCNS5_Z_HLLC_HLL_B200 : public ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200 {
public:
    ConvectiveFluxReconstructorWCNS5_Z_HLLC_HLL_B200(const std::string& object_name, const tbox::Dimension& dim,
                                                     const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry,
                                                     const int& num_eqn, const FLOW_MODEL::TYPE& flow_model_type,
                                                     const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                                                     const HAMERS_SHARED_PTR<tbox::Database>& convective_flux_reconstructor_db)
        : ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200(object_name, dim, grid_geometry, num_eqn, flow_model_type, flow_model,
                                                            convective_flux_reconstructor_db)
    {
        d_scheme = HB2_WCNS5_Z; /* constant_p: ConvectiveFluxReconstructorWCNS5-Z-HLLC-HLL.cpp:192-195, read by the base */
    }
};

/* "WCNS6_LD_HLLC_HLL" (ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp): localized-dissipation six-point weights. */
class ConvectiveFluxReconstructorWCNS6_LD_HLLC_HLL_B200 : public ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200 {
public:
    ConvectiveFluxReconstructorWCNS6_LD_HLLC_HLL_B200(const std::string& object_name, const tbox::Dimension& dim,
                                                      const HAMERS_SHARED_PTR<geom::CartesianGridGeometry>& grid_geometry,
                                                      const int& num_eqn, const FLOW_MODEL::TYPE& flow_model_type,
                                                      const HAMERS_SHARED_PTR<FlowModel>& flow_model,
                                                      const HAMERS_SHARED_PTR<tbox::Database>& convective_flux_reconstructor_db);
    void printClassData(std::ostream& os) const;
    void putToRestart(const HAMERS_SHARED_PTR<tbox::Database>& restart_db) const;
};

#endif
