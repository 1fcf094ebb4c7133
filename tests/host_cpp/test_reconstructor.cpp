/*
 * test_reconstructor.cpp -- drives the host-side C++ mirror of the reference interface
 * (hamers_b200/host/ConvectiveFluxReconstructorB200) the way HAMeRS drives its reconstructor: build a hier::Patch
 * with pdat::CellData / SideData in named variable contexts, call computeConvectiveFluxAndSourceOnPatch and the
 * fused RK stage, dump the results.  tests/test_host_cpp.py writes the input, runs this program on the GPU box and
 * compares the dump with the oracle.
 *
 *   input  (binary): int32 dim, n[3], model, ns, math (+ 10 x scheme, + 100: the state is allocated with SIX ghost cells,
 *                    like the Navier-Stokes application does); double gamma[4], dx[3], dt; then ncomp ghost-box components
 *   output (binary): per direction num_eqn side components, num_eqn source components, ncomp ghost-box components of
 *                    the fused first RK stage
 */
#include "../../hamers_b200/host/ConvectiveFluxReconstructorB200.hpp"

#include <cstdio>
#include <cstdlib>
#include <iostream>

static void must(bool ok, const char* what)
{
    if (!ok) {
        std::fprintf(stderr, "test_reconstructor: %s\n", what);
        std::exit(2);
    }
}

int main(int argc, char** argv)
{
    must(argc == 3, "usage: test_reconstructor <input> <output>");
    FILE* fi = std::fopen(argv[1], "rb");
    must(fi != 0, "cannot open input");
    int32_t hdr[7];
    double gam[4], dx[3], dt;
    must(std::fread(hdr, 4, 7, fi) == 7 && std::fread(gam, 8, 4, fi) == 4 && std::fread(dx, 8, 3, fi) == 3 && std::fread(&dt, 8, 1, fi) == 1, "short header");
    const int d = hdr[0], model = hdr[4], ns = hdr[5], math = hdr[6] % 10, scheme = (hdr[6] % 100) / 10;
    const int state_ghosts = hdr[6] >= 100 ? 6 : 4;
    const tbox::Dimension dim((unsigned short)d);

    try {
        /* input databases, as main.cpp would read them from the input file */
        HAMERS_SHARED_PTR<tbox::Database> flow_model_db(new tbox::Database("Flow_model"));
        flow_model_db->putDoubleVector("species_gamma", std::vector<double>(gam, gam + ns));
        HAMERS_SHARED_PTR<tbox::Database> reconstructor_db(new tbox::Database("Convective_flux_reconstructor"));
        reconstructor_db->putInteger("constant_p", 2);

        const FLOW_MODEL::TYPE type = model == 0 ? FLOW_MODEL::SINGLE_SPECIES : FLOW_MODEL::FIVE_EQN_ALLAIRE;
        HAMERS_SHARED_PTR<FlowModel> flow_model(new FlowModel("flow model", dim, type, ns, flow_model_db));
        HAMERS_SHARED_PTR<geom::CartesianGridGeometry> grid_geometry(new geom::CartesianGridGeometry(dim));
        /* what ConvectiveFluxReconstructorManager does with the input string (ConvectiveFluxReconstructorManager.cpp:37-48) */
        const char* names[3] = {"WCNS5_JS_HLLC_HLL", "WCNS5_Z_HLLC_HLL", "WCNS6_LD_HLLC_HLL"};
        must(scheme >= 0 && scheme < 3, "unknown scheme");
        HAMERS_SHARED_PTR<ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200> holder;
        if (scheme == 0)
            holder.reset(new ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200(names[0], dim, grid_geometry, flow_model->getNumberOfEquations(),
                                                                               type, flow_model, reconstructor_db));
        else if (scheme == 1)
            holder.reset(new ConvectiveFluxReconstructorWCNS5_Z_HLLC_HLL_B200(names[1], dim, grid_geometry, flow_model->getNumberOfEquations(),
                                                                              type, flow_model, reconstructor_db));
        else
            holder.reset(new ConvectiveFluxReconstructorWCNS6_LD_HLLC_HLL_B200(names[2], dim, grid_geometry, flow_model->getNumberOfEquations(),
                                                                               type, flow_model, reconstructor_db));
        ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200& reconstructor = *holder;
        reconstructor.setMathMode(math);
        reconstructor.printClassData(std::cout);
        HAMERS_SHARED_PTR<tbox::Database> restart_db(new tbox::Database("restart"));
        reconstructor.putToRestart(restart_db);
        must(restart_db->getInteger("d_constant_p") == 2, "putToRestart");
        const hier::IntVector conv_ghosts = reconstructor.getConvectiveFluxNumberOfGhostCells();
        for (int a = 0; a < d; a++) must(conv_ghosts[a] == 4, "ghost width");
        /* the application allocates max(convective, diffusive) ghost cells: 4 for Euler, 6 for Navier-Stokes */
        const hier::IntVector ghosts = hier::IntVector::getOne(dim) * state_ghosts;

        /* one patch, interior box [0, n-1] */
        hier::IntVector lo(dim, 0), hi(dim, 0);
        for (int a = 0; a < d; a++) hi[a] = hdr[1 + a] - 1;
        hier::Box box(lo, hi);
        hier::Patch patch(box);
        patch.setPatchGeometry(HAMERS_SHARED_PTR<hier::PatchGeometry>(new geom::CartesianPatchGeometry(dx, 0, d)));

        HAMERS_SHARED_PTR<hier::VariableContext> ctx_intermediate(new hier::VariableContext("INTERMEDIATE_0"));
        HAMERS_SHARED_PTR<hier::VariableContext> ctx_scratch(new hier::VariableContext("SCRATCH"));
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
            patch.setPatchData(cons[v], ctx_intermediate, data);
            HAMERS_SHARED_PTR<pdat::CellData<double> > out(new pdat::CellData<double>(box, cons[v]->getDepth(), ghosts));
            patch.setPatchData(cons[v], ctx_scratch, out);
        }
        std::fclose(fi);

        HAMERS_SHARED_PTR<pdat::SideVariable<double> > var_flux(new pdat::SideVariable<double>(dim, "convective flux", neq));
        HAMERS_SHARED_PTR<pdat::CellVariable<double> > var_source(new pdat::CellVariable<double>(dim, "source", neq));
        HAMERS_SHARED_PTR<pdat::SideData<double> > flux(new pdat::SideData<double>(box, neq, hier::IntVector::getZero(dim)));
        HAMERS_SHARED_PTR<pdat::CellData<double> > source(new pdat::CellData<double>(box, neq, hier::IntVector::getZero(dim)));
        source->fillAll(0.0); /* Euler::computeFluxesAndSourcesOnPatch zero-fills the source, Euler.cpp:917-932 */
        patch.setPatchData(var_flux, ctx_intermediate, flux);
        patch.setPatchData(var_source, ctx_intermediate, source);

        reconstructor.computeConvectiveFluxAndSourceOnPatch(patch, var_flux, var_source, ctx_intermediate, 0.0, dt, 0);

        std::vector<HAMERS_SHARED_PTR<hier::VariableContext> > ctxs(1, ctx_intermediate);
        reconstructor.advanceFusedStageOnPatch(patch, dt, std::vector<double>(1, 1.0), std::vector<double>(1, 1.0), ctxs, ctx_scratch);

        FILE* fo = std::fopen(argv[2], "wb");
        must(fo != 0, "cannot open output");
        const hier::IntVector nc = box.numberCells();
        for (int nd = 0; nd < d; nd++) {
            size_t n = 1;
            for (int a = 0; a < d; a++) n *= (size_t)(nc[a] + (a == nd ? 1 : 0));
            for (int e = 0; e < neq; e++) std::fwrite(flux->getPointer(nd, e), 8, n, fo);
        }
        size_t ncell = 1, nghost = 1;
        for (int a = 0; a < d; a++) {
            ncell *= (size_t)nc[a];
            nghost *= (size_t)(nc[a] + 2 * state_ghosts);
        }
        for (int e = 0; e < neq; e++) std::fwrite(source->getPointer(e), 8, ncell, fo);
        for (size_t v = 0; v < cons.size(); v++) {
            HAMERS_SHARED_PTR<pdat::CellData<double> > out(
                HAMERS_SHARED_PTR_CAST<pdat::CellData<double>, hier::PatchData>(patch.getPatchData(cons[v], ctx_scratch)));
            for (int c = 0; c < cons[v]->getDepth(); c++) std::fwrite(out->getPointer(c), 8, nghost, fo);
        }
        std::fclose(fo);

        /* error convention: a patch without the registered data must surface as TBOX_ERROR */
        bool threw = false;
        try {
            HAMERS_SHARED_PTR<hier::VariableContext> ctx_missing(new hier::VariableContext("MISSING"));
            reconstructor.computeConvectiveFluxAndSourceOnPatch(patch, var_flux, var_source, ctx_missing, 0.0, dt, 0);
        } catch (const std::runtime_error&) {
            threw = true;
        }
        must(threw, "missing patch data did not raise TBOX_ERROR");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "TBOX_ERROR: %s\n", e.what());
        return 3;
    }
    std::printf("test_reconstructor OK\n");
    return 0;
}
