/*
 * oracle_level.c -- CPU ORACLE level driver (test infrastructure, NOT a product path).
 *
 * Restates the stage loop of RungeKuttaLevelIntegrator::advanceLevel
 * (src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1672-1745) for ONE uniform, fully
 * periodic level tiled into patches: per stage, fill the 4-cell ghosts of every patch
 * from the level (what xfer::RefineSchedule::fillData does at :1568/:1701 for a periodic
 * single level: plain copies), then per patch computeFluxesAndSourcesOnPatch
 * (Euler.cpp:904-999: zero the source, call the reconstructor) and advanceSingleStepOnPatch
 * (Euler.cpp:1003-1679).  Patches are distributed over OpenMP threads, one thread standing
 * in for one MPI rank of the reference.
 */
#include "hamers_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define G ORC_GHOSTS
#define MAX_STAGES 8

static inline int wrap(int i, int n)
{
    int r = i % n;
    return r < 0 ? r + n : r;
}

/* Copy the ghost box of patch with lower corner lo[] out of the periodic level array. */
static void gather_patch(const orc_desc* lvl, const orc_desc* pd, const int lo[3],
                         const double* src, double* dst)
{
    const int dim = lvl->dim;
    const int N0 = lvl->n[0], N1 = lvl->n[1], N2 = dim == 3 ? lvl->n[2] : 1;
    const int g2 = dim == 3 ? G : 0;
    const int e0 = pd->n[0] + 2 * G, e1 = pd->n[1] + 2 * G;
    const int n2 = dim == 3 ? pd->n[2] : 1;
    for (int k = -g2; k < n2 + g2; k++) {
        const int K = wrap(lo[2] + k, N2);
        for (int j = -G; j < pd->n[1] + G; j++) {
            const int J = wrap(lo[1] + j, N1);
            double* drow = dst + (long)e0 * ((j + G) + (long)e1 * (k + g2));
            const double* srow = src + (long)N0 * (J + (long)N1 * K);
            for (int i = -G; i < pd->n[0] + G; i++) drow[i + G] = srow[wrap(lo[0] + i, N0)];
        }
    }
}

static void scatter_interior(const orc_desc* lvl, const orc_desc* pd, const int lo[3],
                             const double* src, double* dst)
{
    const int dim = lvl->dim;
    const int N0 = lvl->n[0], N1 = lvl->n[1];
    const int g2 = dim == 3 ? G : 0;
    const int e0 = pd->n[0] + 2 * G, e1 = pd->n[1] + 2 * G;
    const int n2 = dim == 3 ? pd->n[2] : 1;
    for (int k = 0; k < n2; k++)
        for (int j = 0; j < pd->n[1]; j++) {
            const double* srow = src + (long)e0 * ((j + G) + (long)e1 * (k + g2)) + G;
            double* drow = dst + (long)N0 * ((lo[1] + j) + (long)N1 * (lo[2] + k)) + lo[0];
            memcpy(drow, srow, sizeof(double) * (size_t)pd->n[0]);
        }
}

/*
 * Advance a periodic uniform level by `nsteps` Runge-Kutta steps of size dt, in place.
 *   lvl      : descriptor whose n[] are the LEVEL dims
 *   patch    : patch dims (must divide the level dims)
 *   U        : ncomp arrays of the level interior (no ghosts), x fastest
 *   nstages, alpha, beta : RK table, row-major [stage][m] with row length nstages
 *                          (RungeKuttaLevelIntegrator.cpp:3894-3929 for the default)
 * Returns 0, or a negative error code.
 */
int orc_level_advance(const orc_desc* lvl, const int patch[3], double* const* U,
                      double dt, int nsteps, int nstages,
                      const double* alpha, const double* beta, int nthreads)
{
    const int dim = lvl->dim;
    const int ncomp = orc_num_comp(lvl), neq = orc_num_eqn(lvl);
    if (nstages > MAX_STAGES) return -1;
    int np[3] = {1, 1, 1};
    orc_desc pd = *lvl;
    for (int a = 0; a < dim; a++) {
        if (lvl->n[a] % patch[a] != 0) return -2;
        np[a] = lvl->n[a] / patch[a];
        pd.n[a] = patch[a];
    }
    const long npatch = (long)np[0] * np[1] * np[2];
    const long nlevel = orc_cell_size(lvl);
    const long ncg = orc_cell_ghost_size(&pd), ncs = orc_cell_size(&pd);

    /* a flux of stage m must be kept if a later stage uses it */
    int keep[MAX_STAGES];
    for (int m = 0; m < nstages; m++) {
        keep[m] = 0;
        for (int s = m + 1; s < nstages; s++)
            if (beta[s * nstages + m] != 0.0) keep[m] = 1;
    }

    /* level storage of the intermediate states U^(m) */
    double* Uint[MAX_STAGES][ORC_MAX_EQ + 1];
    for (int m = 0; m < nstages; m++)
        for (int c = 0; c < ncomp; c++) Uint[m][c] = (double*)malloc(sizeof(double) * (size_t)nlevel);
    double* Unew[ORC_MAX_EQ + 1];
    for (int c = 0; c < ncomp; c++) Unew[c] = (double*)malloc(sizeof(double) * (size_t)nlevel);

    /* per (stage, patch) flux/source storage when needed by later stages */
    double** Fkeep[MAX_STAGES];
    for (int m = 0; m < nstages; m++) Fkeep[m] = keep[m] ? (double**)calloc((size_t)npatch, sizeof(double*)) : 0;
    long fsz[3], foff[3 * ORC_MAX_EQ + ORC_MAX_EQ + 1];
    long ftot = 0;
    for (int a = 0; a < dim; a++) fsz[a] = orc_side_size(&pd, a);
    for (int a = 0; a < dim; a++)
        for (int e = 0; e < neq; e++) {
            foff[a * neq + e] = ftot;
            ftot += fsz[a];
        }
    for (int e = 0; e < neq; e++) {
        foff[dim * neq + e] = ftot;
        ftot += ncs;
    }

#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif

    int err = 0;
    for (int step = 0; step < nsteps; step++) {
        for (int sn = 0; sn < nstages; sn++) {
            /* copyTimeDependentData(SCRATCH -> INTERMEDIATE[sn]) (:1677) */
            for (int c = 0; c < ncomp; c++) memcpy(Uint[sn][c], U[c], sizeof(double) * (size_t)nlevel);

#pragma omp parallel for schedule(dynamic, 1)
            for (long ip = 0; ip < npatch; ip++) {
                int lo[3] = {(int)(ip % np[0]) * patch[0], (int)((ip / np[0]) % np[1]) * patch[1],
                             dim == 3 ? (int)(ip / ((long)np[0] * np[1])) * patch[2] : 0};
                /* ghost-filled copies of every intermediate state used by this stage */
                double* Ug[MAX_STAGES][ORC_MAX_EQ + 1];
                for (int m = 0; m <= sn; m++)
                    for (int c = 0; c < ncomp; c++) {
                        Ug[m][c] = (double*)malloc(sizeof(double) * (size_t)ncg);
                        gather_patch(lvl, &pd, lo, Uint[m][c], Ug[m][c]);
                    }
                /* flux + source of this stage: Euler.cpp:917-932 zero-fills the source first */
                double* Fbuf = (double*)calloc((size_t)ftot, sizeof(double));
                double* Fp[3 * ORC_MAX_EQ];
                double* Sp[ORC_MAX_EQ];
                for (int a = 0; a < dim; a++)
                    for (int e = 0; e < neq; e++) Fp[a * neq + e] = Fbuf + foff[a * neq + e];
                for (int e = 0; e < neq; e++) Sp[e] = Fbuf + foff[dim * neq + e];
                const double* Qp[ORC_MAX_EQ + 1];
                for (int c = 0; c < ncomp; c++) Qp[c] = Ug[sn][c];
                if (orc_compute_flux_and_source(&pd, Qp, dt, Fp, Sp, 0, 0) != 0) err = -3;
                if (keep[sn]) Fkeep[sn][ip] = Fbuf;

                /* advanceSingleStepOnPatch */
                const double* const* Uptr[MAX_STAGES];
                const double* const* Fptr[MAX_STAGES];
                const double* const* Sptr[MAX_STAGES];
                const double* Ftab[MAX_STAGES][3 * ORC_MAX_EQ];
                const double* Stab[MAX_STAGES][ORC_MAX_EQ];
                for (int m = 0; m <= sn; m++) {
                    Uptr[m] = (const double* const*)Ug[m];
                    const double* base = (m == sn) ? Fbuf : (keep[m] ? Fkeep[m][ip] : 0);
                    for (int q = 0; q < dim * neq; q++) Ftab[m][q] = base ? base + foff[q] : 0;
                    for (int e = 0; e < neq; e++) Stab[m][e] = base ? base + foff[dim * neq + e] : 0;
                    Fptr[m] = Ftab[m];
                    Sptr[m] = Stab[m];
                }
                double* Uo[ORC_MAX_EQ + 1];
                for (int c = 0; c < ncomp; c++) Uo[c] = (double*)malloc(sizeof(double) * (size_t)ncg);
                orc_advance_stage(&pd, sn + 1, alpha + sn * nstages, beta + sn * nstages, Uptr, Fptr, Sptr, Uo);
                for (int c = 0; c < ncomp; c++) {
                    scatter_interior(lvl, &pd, lo, Uo[c], Unew[c]);
                    free(Uo[c]);
                }
                for (int m = 0; m <= sn; m++)
                    for (int c = 0; c < ncomp; c++) free(Ug[m][c]);
                if (!keep[sn]) free(Fbuf);
            }
            for (int c = 0; c < ncomp; c++) memcpy(U[c], Unew[c], sizeof(double) * (size_t)nlevel);
        }
        for (int m = 0; m < nstages; m++)
            if (keep[m])
                for (long ip = 0; ip < npatch; ip++) {
                    free(Fkeep[m][ip]);
                    Fkeep[m][ip] = 0;
                }
    }

    for (int m = 0; m < nstages; m++) {
        for (int c = 0; c < ncomp; c++) free(Uint[m][c]);
        free(Fkeep[m]);
    }
    for (int c = 0; c < ncomp; c++) free(Unew[c]);
    return err;
}
