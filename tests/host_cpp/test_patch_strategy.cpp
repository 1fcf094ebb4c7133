/*
 * test_patch_strategy.cpp -- drives RungeKuttaPatchStrategyB200 (hamers_b200/host) the way
 * RungeKuttaLevelIntegrator::advanceLevel drives its patch strategy (RungeKuttaLevelIntegrator.cpp:1672-1745): a periodic
 * level cut into SEVERAL patches of unequal sizes, registered once, then per time step and RK stage: ghost fill, per patch
 * computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch, stage hand-over; the patches come back to host memory only at
 * the end.  tests/test_host_cpp.py writes the input, runs this program on the GPU box and compares the dump with the
 * oracle's orc_level_advance on the same level.
 *
 *   input  (binary): int32 dim, N[3], model, ns, math, nsteps, cuts[3] (level cut at these indices: 2 patches per direction);
 *                    double gamma[4], R[4], dx[3], dt; then ncomp LEVEL-interior components (x fastest)
 *   output (binary): ncomp level-interior components after nsteps SSP-RK3 steps; then dim + 1 doubles: spectral radii, stable dt
 */
#include "../../hamers_b200/host/RungeKuttaPatchStrategyB200.hpp"

#include <cstdio>
#include <cstdlib>
#include <iostream>

static void must(bool ok, const char* what)
{
    if (!ok) {
        std::fprintf(stderr, "test_patch_strategy: %s\n", what);
        std::exit(2);
    }
}

int main(int argc, char** argv)
{
    must(argc == 3, "usage: test_patch_strategy <input> <output>");
    FILE* fi = std::fopen(argv[1], "rb");
    must(fi != 0, "cannot open input");
    int32_t hdr[11];
    double gam[4], R[4], dx[3], dt;
    must(std::fread(hdr, 4, 11, fi) == 11 && std::fread(gam, 8, 4, fi) == 4 && std::fread(R, 8, 4, fi) == 4 && std::fread(dx, 8, 3, fi) == 3 &&
             std::fread(&dt, 8, 1, fi) == 1, "short header");
    const int d = hdr[0], model = hdr[4], ns = hdr[5], math = hdr[6], nsteps = hdr[7];
    const int N[3] = {hdr[1], hdr[2], d == 3 ? hdr[3] : 1}, cut[3] = {hdr[8], hdr[9], hdr[10]};
    const tbox::Dimension dim((unsigned short)d);
    try {
        HAMERS_SHARED_PTR<tbox::Database> flow_model_db(new tbox::Database("Flow_model"));
        flow_model_db->putDoubleVector("species_gamma", std::vector<double>(gam, gam + ns));
        flow_model_db->putDoubleVector("species_R", std::vector<double>(R, R + ns));
        const FLOW_MODEL::TYPE type = model == 0 ? FLOW_MODEL::SINGLE_SPECIES : (model == 1 ? FLOW_MODEL::FIVE_EQN_ALLAIRE : FLOW_MODEL::FOUR_EQN_CONSERVATIVE);
        HAMERS_SHARED_PTR<FlowModel> flow_model(new FlowModel("flow model", dim, type, ns, flow_model_db));
        const std::vector<HAMERS_SHARED_PTR<pdat::CellVariable<double> > >& cons = flow_model->getConservativeVariables();
        const int ncomp = flow_model->getNumberOfStoredComponents();
        const size_t nlevel = (size_t)N[0] * N[1] * N[2];
        std::vector<std::vector<double> > U(ncomp, std::vector<double>(nlevel));
        for (int c = 0; c < ncomp; c++) must(std::fread(U[c].data(), 8, nlevel, fi) == nlevel, "short state");
        std::fclose(fi);

        /* the level: 2 x 2 (x 2) patches of unequal sizes */
        HAMERS_SHARED_PTR<hier::VariableContext> ctx_current(new hier::VariableContext("CURRENT"));
        std::vector<HAMERS_SHARED_PTR<hier::Patch> > patches;
        const hier::IntVector ghosts = hier::IntVector::getOne(dim) * 4;
        for (int pz = 0; pz < (d == 3 ? 2 : 1); pz++)
            for (int py = 0; py < 2; py++)
                for (int px = 0; px < 2; px++) {
                    const int pc[3] = {px, py, pz};
                    hier::IntVector lo(dim, 0), hi(dim, 0);
                    for (int a = 0; a < d; a++) {
                        lo[a] = pc[a] == 0 ? 0 : cut[a];
                        hi[a] = (pc[a] == 0 ? cut[a] : N[a]) - 1;
                    }
                    hier::Box box(lo, hi);
                    HAMERS_SHARED_PTR<hier::Patch> patch(new hier::Patch(box));
                    patch->setPatchGeometry(HAMERS_SHARED_PTR<hier::PatchGeometry>(new geom::CartesianPatchGeometry(dx, 0, d)));
                    int comp = 0;
                    const hier::IntVector nc = box.numberCells();
                    for (size_t v = 0; v < cons.size(); v++) {
                        HAMERS_SHARED_PTR<pdat::CellData<double> > data(new pdat::CellData<double>(box, cons[v]->getDepth(), ghosts));
                        for (int c = 0; c < cons[v]->getDepth(); c++, comp++) {
                            double* q = data->getPointer(c);
                            const long long gd0 = nc[0] + 8, gd1 = nc[1] + 8;
                            for (int k = 0; k < (d == 3 ? nc[2] : 1); k++)
                                for (int j = 0; j < nc[1]; j++)
                                    for (int i = 0; i < nc[0]; i++) {
                                        const size_t src = (size_t)(lo[0] + i) + (size_t)N[0] * ((size_t)(lo[1] + j) + (size_t)N[1] * (size_t)((d == 3 ? lo[2] : 0) + k));
                                        const size_t dst = (size_t)(i + 4) + (size_t)gd0 * ((size_t)(j + 4) + (size_t)gd1 * (size_t)(d == 3 ? k + 4 : 0));
                                        q[dst] = U[comp][src];
                                    }
                        }
                        patch->setPatchData(cons[v], ctx_current, data);
                    }
                    patches.push_back(patch);
                }

        RungeKuttaPatchStrategyB200 strategy("Euler on B200", dim, flow_model, "WCNS5_JS_HLLC_HLL", math);
        hier::IntVector domain(dim, 0), periodic(dim, 1);
        for (int a = 0; a < d; a++) domain[a] = N[a];
        strategy.registerPatchLevel(patches, domain, periodic, ctx_current);

        /* SSPRK(3,3), RungeKuttaLevelIntegrator.cpp:3894-3929 */
        const double al[3][3] = {{1.0, 0.0, 0.0}, {3.0 / 4.0, 1.0 / 4.0, 0.0}, {1.0 / 3.0, 0.0, 2.0 / 3.0}};
        const double be[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0 / 4.0, 0.0}, {0.0, 0.0, 2.0 / 3.0}};
        const double ga[3][3] = {{1.0 / 6.0, 0.0, 0.0}, {0.0, 1.0 / 6.0, 0.0}, {0.0, 0.0, 2.0 / 3.0}};
        std::vector<HAMERS_SHARED_PTR<hier::VariableContext> > intermediate;
        for (int sn = 0; sn < 3; sn++) intermediate.push_back(HAMERS_SHARED_PTR<hier::VariableContext>(new hier::VariableContext("INTERMEDIATE")));
        double time = 0.0;
        for (int step = 0; step < nsteps; step++) {
            for (int sn = 0; sn < 3; sn++) {
                strategy.fillGhostCellsOnLevel(sn);
                const std::vector<double> alpha(al[sn], al[sn] + sn + 1), beta(be[sn], be[sn] + sn + 1), gamma(ga[sn], ga[sn] + sn + 1);
                for (size_t p = 0; p < patches.size(); p++) {
                    strategy.computeFluxesAndSourcesOnPatch(*patches[p], time, dt, sn, intermediate[sn]);
                    strategy.advanceSingleStepOnPatch(*patches[p], time, dt, alpha, beta, gamma, intermediate);
                }
                strategy.finishStageOnLevel(alpha, sn == 2);
            }
            time += dt;
        }
        const std::vector<double> sr = strategy.computeSpectralRadiusesAndStableDtOnLevel();
        for (size_t p = 0; p < patches.size(); p++) {
            strategy.synchronizePatchToHost(*patches[p], ctx_current);
            const hier::Box& box = patches[p]->getBox();
            const hier::IntVector nc = box.numberCells(), lo = box.lower();
            int comp = 0;
            for (size_t v = 0; v < cons.size(); v++) {
                HAMERS_SHARED_PTR<pdat::CellData<double> > data(
                    HAMERS_SHARED_PTR_CAST<pdat::CellData<double>, hier::PatchData>(patches[p]->getPatchData(cons[v], ctx_current)));
                for (int c = 0; c < cons[v]->getDepth(); c++, comp++) {
                    const double* q = data->getPointer(c);
                    const long long gd0 = nc[0] + 8, gd1 = nc[1] + 8;
                    for (int k = 0; k < (d == 3 ? nc[2] : 1); k++)
                        for (int j = 0; j < nc[1]; j++)
                            for (int i = 0; i < nc[0]; i++) {
                                const size_t dst = (size_t)(lo[0] + i) + (size_t)N[0] * ((size_t)(lo[1] + j) + (size_t)N[1] * (size_t)((d == 3 ? lo[2] : 0) + k));
                                const size_t src = (size_t)(i + 4) + (size_t)gd0 * ((size_t)(j + 4) + (size_t)gd1 * (size_t)(d == 3 ? k + 4 : 0));
                                U[comp][dst] = q[src];
                            }
                }
            }
        }
        FILE* fo = std::fopen(argv[2], "wb");
        must(fo != 0, "cannot open output");
        for (int c = 0; c < ncomp; c++) std::fwrite(U[c].data(), 8, nlevel, fo);
        std::fwrite(sr.data(), 8, sr.size(), fo);
        std::fclose(fo);
        std::printf("test_patch_strategy OK: %d patches, %lld kernel launches\n", (int)patches.size(), strategy.getNumberOfKernelLaunches());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "TBOX_ERROR: %s\n", e.what());
        return 3;
    }
    return 0;
}
