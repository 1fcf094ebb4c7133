/*
 * test_diffusive.cpp -- drives DiffusiveFluxReconstructorNodeSixthOrder_B200 the way NavierStokes::
 * computeFluxesAndSourcesOnPatch does (NavierStokes.cpp:1153-1160): databases like main() reads them from a viscous
 * input deck, a hier::Patch with pdat::CellData (six ghost cells) in a named context, one call, side data written out.
 *   test_diffusive <input> <output>
 * input : int32 dim, n[3]; double gamma, R, mu, mu_v, c_p, Pr, dx[3], dt; then the dim+2 conservative components on the
 *         ghost box (6);  output: the side fluxes, direction by direction, equation by equation.
 */
#include "../../hamers_b200/host/DiffusiveFluxReconstructorB200.hpp"

#include <cstdio>
#include <iostream>
#include <stdexcept>

static void must(bool ok, const char* what)
{
    if (!ok) {
        std::fprintf(stderr, "test_diffusive: %s\n", what);
        std::exit(2);
    }
}

int main(int argc, char** argv)
{
    must(argc == 3, "usage: test_diffusive <input> <output>");
    FILE* fi = std::fopen(argv[1], "rb");
    must(fi != 0, "cannot open input");
    int32_t hdr[4];
    double par[10];
    must(std::fread(hdr, 4, 4, fi) == 4 && std::fread(par, 8, 10, fi) == 10, "short header");
    const int d = hdr[0];
    const double dt = par[9];
    const tbox::Dimension dim((unsigned short)d);
    try {
        /* Flow_model { Equation_of_state_mixing_rules {...} Equation_of_shear_viscosity_mixing_rules {...} ... } flattened */
        HAMERS_SHARED_PTR<tbox::Database> flow_model_db(new tbox::Database("Flow_model"));
        flow_model_db->putDoubleVector("species_gamma", std::vector<double>(1, par[0]));
        flow_model_db->putDoubleVector("species_R", std::vector<double>(1, par[1]));
        flow_model_db->putDoubleVector("species_mu", std::vector<double>(1, par[2]));
        flow_model_db->putDoubleVector("species_mu_v", std::vector<double>(1, par[3]));
        flow_model_db->putDoubleVector("species_c_p", std::vector<double>(1, par[4]));
        flow_model_db->putDoubleVector("species_Pr", std::vector<double>(1, par[5]));
        HAMERS_SHARED_PTR<tbox::Database> reconstructor_db(new tbox::Database("Diffusive_flux_reconstructor"));
        HAMERS_SHARED_PTR<FlowModel> flow_model(new FlowModel("flow model", dim, FLOW_MODEL::SINGLE_SPECIES, 1, flow_model_db));
        HAMERS_SHARED_PTR<geom::CartesianGridGeometry> grid_geometry(new geom::CartesianGridGeometry(dim));
        /* what DiffusiveFluxReconstructorManager does with "SIXTH_ORDER" (DiffusiveFluxReconstructorManager.cpp:73-80) */
        DiffusiveFluxReconstructorNodeSixthOrder_B200 reconstructor("SIXTH_ORDER", dim, grid_geometry, flow_model->getNumberOfEquations(),
                                                                    flow_model, reconstructor_db);
        reconstructor.printClassData(std::cout);
        HAMERS_SHARED_PTR<tbox::Database> restart_db(new tbox::Database("restart"));
        reconstructor.putToRestart(restart_db);
        must(restart_db->getStringWithDefault("d_diffusive_flux_reconstructor", "") == "SIXTH_ORDER", "putToRestart");
        const hier::IntVector ghosts = reconstructor.getDiffusiveFluxNumberOfGhostCells();
        for (int a = 0; a < d; a++) must(ghosts[a] == 6, "ghost width");

        hier::IntVector lo(dim, 0), hi(dim, 0);
        for (int a = 0; a < d; a++) hi[a] = hdr[1 + a] - 1;
        hier::Box box(lo, hi);
        hier::Patch patch(box);
        patch.setPatchGeometry(HAMERS_SHARED_PTR<hier::PatchGeometry>(new geom::CartesianPatchGeometry(par + 6, 0, d)));
        HAMERS_SHARED_PTR<hier::VariableContext> ctx(new hier::VariableContext("INTERMEDIATE_0"));
        const int neq = flow_model->getNumberOfEquations();
        const std::vector<HAMERS_SHARED_PTR<pdat::CellVariable<double> > >& cons = flow_model->getConservativeVariables();
        for (size_t v = 0; v < cons.size(); v++) {
            HAMERS_SHARED_PTR<pdat::CellData<double> > data(new pdat::CellData<double>(box, cons[v]->getDepth(), ghosts));
            for (int c = 0; c < cons[v]->getDepth(); c++) {
                size_t n = 1;
                const hier::IntVector gd = data->getGhostBox().numberCells();
                for (int a = 0; a < d; a++) n *= (size_t)gd[a];
                must(std::fread(data->getPointer(c), 8, n, fi) == n, "short state");
            }
            patch.setPatchData(cons[v], ctx, data);
        }
        std::fclose(fi);
        HAMERS_SHARED_PTR<pdat::SideVariable<double> > var_flux(new pdat::SideVariable<double>(dim, "diffusive flux", neq));
        HAMERS_SHARED_PTR<pdat::SideData<double> > flux(new pdat::SideData<double>(box, neq, hier::IntVector::getZero(dim)));
        patch.setPatchData(var_flux, ctx, flux);

        reconstructor.computeDiffusiveFluxOnPatch(patch, var_flux, ctx, 0.0, dt, 0);

        FILE* fo = std::fopen(argv[2], "wb");
        must(fo != 0, "cannot open output");
        const hier::IntVector nc = box.numberCells();
        for (int nd = 0; nd < d; nd++) {
            size_t n = 1;
            for (int a = 0; a < d; a++) n *= (size_t)(nc[a] + (a == nd ? 1 : 0));
            for (int e = 0; e < neq; e++) std::fwrite(flux->getPointer(nd, e), 8, n, fo);
        }
        std::fclose(fo);

        /* error convention: state registered with the convective ghost width (4) must surface as TBOX_ERROR */
        bool threw = false;
        try {
            HAMERS_SHARED_PTR<hier::VariableContext> ctx4(new hier::VariableContext("FOUR_GHOSTS"));
            for (size_t v = 0; v < cons.size(); v++)
                patch.setPatchData(cons[v], ctx4, HAMERS_SHARED_PTR<pdat::CellData<double> >(new pdat::CellData<double>(
                                                      box, cons[v]->getDepth(), hier::IntVector::getOne(dim) * 4)));
            patch.setPatchData(var_flux, ctx4, flux);
            reconstructor.computeDiffusiveFluxOnPatch(patch, var_flux, ctx4, 0.0, dt, 0);
        } catch (const std::runtime_error&) {
            threw = true;
        }
        must(threw, "a state with four ghost cells did not raise TBOX_ERROR");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "TBOX_ERROR: %s\n", e.what());
        return 3;
    }
    return 0;
}
