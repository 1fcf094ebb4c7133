/* TEST-ONLY host emulation of the AMR operator kernels: the `__host__ __device__` thread functions of
 * hamers_b200/csrc/hb2_amr.cuh driven by plain loops (forward and reverse thread order give the same result: every
 * thread writes its own cell).  Never linked into the product library. */
#include "../../hamers_b200/csrc/hb2_amr.cuh"

#include <cstring>

using namespace hb2;

struct EmuPair {
    int dim, nc[3], nf[3], ratio[3], origin[3], ghosts_c, ghosts_f, ncomp, neq;
    double dxc[3], dxf[3];
};

static void fix3(const EmuPair* p, int* ratio, int* origin, double* dxc, double* dxf)
{
    for (int a = 0; a < 3; a++) {
        const bool on = a < p->dim;
        ratio[a] = on ? p->ratio[a] : 1;
        origin[a] = on ? p->origin[a] : 0;
        dxc[a] = on ? p->dxc[a] : 1.0;
        dxf[a] = on ? p->dxf[a] : 1.0;
    }
}

extern "C" int emu_amr_refine(const EmuPair* p, const double* const* Uold, const double* const* Unew, double tfrac,
                              const int* lo, const int* hi, double* const* Uf, int reverse)
{
    AmrRefineArgs A;
    memset(&A, 0, sizeof A);
    amr_make_layout(p->dim, p->nc, p->ghosts_c, &A.C);
    amr_make_layout(p->dim, p->nf, p->ghosts_f, &A.F);
    fix3(p, A.ratio, A.origin, A.dxc, A.dxf);
    long long total = 1;
    for (int a = 0; a < 3; a++) {
        A.lo[a] = a < p->dim ? lo[a] : 0;
        A.hi[a] = a < p->dim ? hi[a] : 1;
        total *= A.hi[a] - A.lo[a];
    }
    A.ncomp = p->ncomp;
    A.has_new = Unew ? 1 : 0;
    A.tfrac = tfrac;
    for (int c = 0; c < p->ncomp; c++) {
        A.Uold[c] = Uold[c];
        A.Unew[c] = Unew ? Unew[c] : nullptr;
        A.Uf[c] = Uf[c];
    }
    for (long long t = 0; t < total; t++) amr_refine_thread(A, reverse ? total - 1 - t : t);
    return 0;
}

extern "C" int emu_amr_coarsen(const EmuPair* p, const double* const* Uf, const int* lo, const int* hi, double* const* Uc)
{
    AmrCoarsenArgs A;
    memset(&A, 0, sizeof A);
    amr_make_layout(p->dim, p->nc, p->ghosts_c, &A.C);
    amr_make_layout(p->dim, p->nf, p->ghosts_f, &A.F);
    fix3(p, A.ratio, A.origin, A.dxc, A.dxf);
    long long total = 1;
    for (int a = 0; a < 3; a++) {
        A.lo[a] = a < p->dim ? lo[a] : 0;
        A.hi[a] = a < p->dim ? hi[a] : 1;
        total *= A.hi[a] - A.lo[a];
    }
    A.ncomp = p->ncomp;
    for (int c = 0; c < p->ncomp; c++) {
        A.Uf[c] = Uf[c];
        A.Uc[c] = Uc[c];
    }
    for (long long t = 0; t < total; t++) amr_coarsen_thread(A, t);
    return 0;
}

extern "C" int emu_amr_fluxsum(const EmuPair* p, const double* const* F, double* const* fsum)
{
    AmrFluxsumArgs A;
    memset(&A, 0, sizeof A);
    A.dim = p->dim;
    A.neq = p->neq;
    for (int a = 0; a < 3; a++) A.n[a] = a < p->dim ? p->nf[a] : 1;
    for (int d = 0; d < p->dim; d++)
        for (int e = 0; e < p->neq; e++) {
            A.F[d * p->neq + e] = F[d * p->neq + e];
            for (int s = 0; s < 2; s++) A.fsum[(2 * d + s) * p->neq + e] = fsum[(2 * d + s) * p->neq + e];
        }
    for (int d = 0; d < p->dim; d++)
        for (int s = 0; s < 2; s++) {
            long long total = 1;
            for (int a = 0; a < p->dim; a++)
                if (a != d) total *= A.n[a];
            for (long long t = 0; t < total; t++) amr_fluxsum_thread(A, d, s, t);
        }
    return 0;
}

extern "C" int emu_amr_coarsen_fluxsum(const EmuPair* p, const double* const* fsum, double* const* Fc)
{
    AmrCoarsenFluxsumArgs A;
    memset(&A, 0, sizeof A);
    A.dim = p->dim;
    A.neq = p->neq;
    fix3(p, A.ratio, A.origin, A.dxc, A.dxf);
    for (int a = 0; a < 3; a++) {
        A.nf[a] = a < p->dim ? p->nf[a] : 1;
        A.nc[a] = a < p->dim ? p->nc[a] : 1;
    }
    for (int d = 0; d < p->dim; d++)
        for (int e = 0; e < p->neq; e++) {
            A.Fc[d * p->neq + e] = Fc[d * p->neq + e];
            for (int s = 0; s < 2; s++) A.fsum[(2 * d + s) * p->neq + e] = fsum[(2 * d + s) * p->neq + e];
        }
    for (int d = 0; d < p->dim; d++)
        for (int s = 0; s < 2; s++) {
            long long total = 1;
            for (int a = 0; a < p->dim; a++)
                if (a != d) total *= A.nf[a] / A.ratio[a];
            for (long long t = 0; t < total; t++) amr_coarsen_fluxsum_thread(A, d, s, t);
        }
    return 0;
}

extern "C" int emu_amr_extrapolate(int dim, const int* n, int g, int ncomp, double* const* U, int dir, int side)
{
    AmrExtrapArgs A;
    memset(&A, 0, sizeof A);
    amr_make_layout(dim, n, g, &A.L);
    A.dir = dir;
    A.side = side;
    A.ncomp = ncomp;
    for (int c = 0; c < ncomp; c++) A.U[c] = U[c];
    long long total = 1;
    for (int a = 0; a < 3; a++) total *= (a == dir) ? A.L.g[a] : A.L.n[a];
    for (long long t = 0; t < total; t++) amr_extrapolate_thread(A, t);
    return 0;
}
