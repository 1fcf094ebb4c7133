/*
 * hb2_amr.cuh -- SURVEY row f3: the patch-data operators of a two-level AMR step around the convective hot path.
 *
 * `__host__ __device__` thread functions (one thread = one fine cell / coarse cell / face), shared by the sm_100a kernels
 * of hb2_amr.cu and by the test-only host emulation (tests/host_emu/emu_amr.cpp).  All HBM-bound copies with a handful
 * of FP64 operations; compiled with -fmad=false so that the operation order below is what runs.
 *
 * What each one replaces in the reference's time step (RungeKuttaLevelIntegrator.cpp, path:line under the reference tree):
 *   amr_refine_thread         the coarse-to-fine part of xfer::RefineSchedule::fillData (:1568) for the ghost cells of a
 *                             fine patch at a coarse-fine boundary: linear time interpolation of the coarser level's OLD
 *                             and NEW data (SAMRAI pdat CellDoubleLinearTimeInterpolateOp, `arrayold*oldfrac +
 *                             arraynew*tfrac`), then "CONSERVATIVE_LINEAR_REFINE" (registered at FlowModelSingleSpecies.cpp:578,
 *                             FlowModelFourEqnConservative.cpp registerConservativeVariables): SAMRAI geom
 *                             CartesianCellDoubleConservativeLinearRefine -- per direction a central slope limited to twice
 *                             the smaller one-sided difference, zero at extrema.
 *   amr_coarsen_thread        "CONSERVATIVE_COARSEN" of d_coarsen_sync_data (:2189-2203): SAMRAI geom
 *                             CartesianCellDoubleWeightedAverage, sum of fine values times dV_f, divided by dV_c.
 *   amr_fluxsum_thread        postprocessFluxAndSourceData (:2968-3230) = algs_upfluxsum{2,3}d.f (upfluxsumside*): the
 *                             flux on the outer sides of a fine patch is added to the patch's flux integrals.
 *   amr_coarsen_fluxsum_thread  d_coarsen_fluxsum (:2147-2158): "CONSERVATIVE_COARSEN" of OutersideData onto the coarse SideData
 *                             flux (SAMRAI geom CartesianOutersideDoubleWeightedAverage): sum over the fine faces of a
 *                             coarse face times their area, divided by the coarse area.
 *   amr_extrapolate_thread    BDRY_COND::BASIC::FLOW (BasicCartesianBoundaryUtilities2.cpp:310-345, ...3.cpp): ghost cells
 *                             of a physical boundary copy the adjacent interior cell.
 * SAMRAI itself (v4.1.0, circleci/install-SAMRAI.sh:4) is not part of /root/reference: the three SAMRAI operators are
 * restated from its published algorithm -- PARITY UNPINNED for them; the oracle (oracle/amr.py) restates
 * them independently in numpy and the tests check their defining properties (conservation, exactness on linear data,
 * monotonicity).
 */
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HB2A_HD __host__ __device__ __forceinline__
#else
#define HB2A_HD inline
#endif

#define HB2A_MAXC 13

namespace hb2 {

/* ghost-box cell-data layout of one patch (x fastest, SAMRAI pdat::CellData) */
struct AmrLayout {
    int dim;
    int n[3];
    int g[3];
    long long cs[3];
    long long ncell_g;
};

HB2A_HD long long amr_cidx(const AmrLayout& L, int i, int j, int k)
{
    return (i + L.g[0]) + (long long)(j + L.g[1]) * L.cs[1] + (long long)(k + L.g[2]) * L.cs[2];
}

HB2A_HD int amr_floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

struct AmrRefineArgs {
    AmrLayout C, F;
    int ratio[3];
    int origin[3];          /* index, in the coarse patch, of the coarse cell that holds fine cell 0 */
    int lo[3], hi[3];       /* fine cells [lo, hi) to fill (may lie in the fine ghost box) */
    int ncomp;
    int has_new;            /* 0: refine Uold alone (initial data) */
    double tfrac;           /* (t - t_old)/(t_new - t_old) */
    double dxc[3], dxf[3];
    const double* Uold[HB2A_MAXC];
    const double* Unew[HB2A_MAXC];
    double* Uf[HB2A_MAXC];
};

/* time-interpolated coarse value */
HB2A_HD double amr_coarse_value(const AmrRefineArgs& A, int c, long long x)
{
    if (!A.has_new) return A.Uold[c][x];
    const double oldfrac = 1.0 - A.tfrac;
    return A.Uold[c][x] * oldfrac + A.Unew[c][x] * A.tfrac;
}

HB2A_HD void amr_refine_thread(const AmrRefineArgs& A, long long t)
{
    const int ex = A.hi[0] - A.lo[0], ey = A.hi[1] - A.lo[1];
    int f[3];
    f[0] = A.lo[0] + (int)(t % ex);
    f[1] = A.lo[1] + (int)((t / ex) % ey);
    f[2] = A.lo[2] + (int)(t / ((long long)ex * ey));
    int ic[3] = {0, 0, 0};
    double delta[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < A.C.dim; d++) {
        const int q = amr_floor_div(f[d], A.ratio[d]);
        const int ir = f[d] - q * A.ratio[d];
        ic[d] = q + A.origin[d];
        delta[d] = ((double)ir + 0.5) * A.dxf[d] - A.dxc[d] * 0.5;
    }
    const long long xc = amr_cidx(A.C, ic[0], ic[1], ic[2]);
    const long long xf = amr_cidx(A.F, f[0], f[1], f[2]);
    for (int c = 0; c < A.ncomp; c++) {
        const double v0 = amr_coarse_value(A, c, xc);
        double val = v0;
        for (int d = 0; d < A.C.dim; d++) {
            const double vR = amr_coarse_value(A, c, xc + A.C.cs[d]);
            const double vL = amr_coarse_value(A, c, xc - A.C.cs[d]);
            const double dR = vR - v0, dL = v0 - vL;
            const double coef2 = 0.5 * (dR + dL);
            const double bound = 2.0 * fmin(fabs(dR), fabs(dL));
            double slope = 0.0;
            if (dL * dR > 0.0) slope = copysign(fmin(fabs(coef2), bound), coef2) / A.dxc[d];
            val = val + slope * delta[d];
        }
        A.Uf[c][xf] = val;
    }
}

struct AmrCoarsenArgs {
    AmrLayout C, F;
    int ratio[3];
    int origin[3];          /* as above */
    int lo[3], hi[3];       /* COARSE cells [lo, hi), in coarse patch indices, overwritten by the averages */
    int ncomp;
    double dxc[3], dxf[3];
    const double* Uf[HB2A_MAXC];
    double* Uc[HB2A_MAXC];
};

HB2A_HD void amr_coarsen_thread(const AmrCoarsenArgs& A, long long t)
{
    const int ex = A.hi[0] - A.lo[0], ey = A.hi[1] - A.lo[1];
    int c3[3];
    c3[0] = A.lo[0] + (int)(t % ex);
    c3[1] = A.lo[1] + (int)((t / ex) % ey);
    c3[2] = A.lo[2] + (int)(t / ((long long)ex * ey));
    const int dim = A.C.dim;
    double dVf = A.dxf[0] * A.dxf[1], dVc = A.dxc[0] * A.dxc[1];
    if (dim == 3) {
        dVf = dVf * A.dxf[2];
        dVc = dVc * A.dxc[2];
    }
    const int r0 = A.ratio[0], r1 = A.ratio[1], r2 = (dim == 3) ? A.ratio[2] : 1;
    const int f0 = (c3[0] - A.origin[0]) * r0, f1 = (c3[1] - A.origin[1]) * r1, f2 = (dim == 3) ? (c3[2] - A.origin[2]) * r2 : 0;
    const long long xc = amr_cidx(A.C, c3[0], c3[1], c3[2]);
    for (int c = 0; c < A.ncomp; c++) {
        double spv = 0.0;
        for (int i2 = 0; i2 < r2; i2++)
            for (int i1 = 0; i1 < r1; i1++)
                for (int i0 = 0; i0 < r0; i0++) spv = spv + A.Uf[c][amr_cidx(A.F, f0 + i0, f1 + i1, f2 + i2)] * dVf;
        A.Uc[c][xc] = spv / dVc;
    }
}

/* Outer-side flux integrals of ONE fine patch: fsum[(2 dir + side) * neq + e] is a dense array over the tangential cells
 * of the patch (x fastest among the remaining directions). */
struct AmrFluxsumArgs {
    int dim, neq;
    int n[3];               /* fine patch interior */
    const double* F[3 * 12]; /* side fluxes [dir * neq + e], ghost 0 */
    double* fsum[6 * 12];
};

HB2A_HD long long amr_side_index(const int n[3], int dir, int i, int j, int k)
{
    const long long e0 = n[0] + (dir == 0 ? 1 : 0), e1 = n[1] + (dir == 1 ? 1 : 0);
    return i + e0 * (j + e1 * (long long)k);
}

/* thread t of (dir, side): tangential cell t */
HB2A_HD void amr_fluxsum_thread(const AmrFluxsumArgs& A, int dir, int side, long long t)
{
    const int ta = (dir == 0) ? 1 : 0, tb = (dir == 2) ? 1 : 2;     /* tangential directions, ta faster */
    const int na = A.n[ta];
    int idx[3];
    idx[dir] = side ? A.n[dir] : 0;
    idx[ta] = (int)(t % na);
    idx[tb] = (int)(t / na);
    const long long f = amr_side_index(A.n, dir, idx[0], idx[1], idx[2]);
    for (int e = 0; e < A.neq; e++) {
        double* s = A.fsum[(2 * dir + side) * A.neq + e];
        s[t] = s[t] + A.F[dir * A.neq + e][f];
    }
}

struct AmrCoarsenFluxsumArgs {
    int dim, neq;
    int nf[3];              /* fine patch interior */
    int nc[3];              /* coarse patch interior */
    int ratio[3];
    int origin[3];          /* coarse cell that holds fine cell 0 */
    double dxc[3], dxf[3];
    const double* fsum[6 * 12];
    double* Fc[3 * 12];      /* coarse side fluxes [dir * neq + e], ghost 0: faces on the fine patch's boundary are overwritten */
};

/* thread t of (dir, side): coarse tangential face t of the fine patch's boundary */
HB2A_HD void amr_coarsen_fluxsum_thread(const AmrCoarsenFluxsumArgs& A, int dir, int side, long long t)
{
    const int ta = (dir == 0) ? 1 : 0, tb = (dir == 2) ? 1 : 2;
    const int nca = A.nf[ta] / A.ratio[ta];
    const int ca = (int)(t % nca), cb = (int)(t / nca);
    const int ra = A.ratio[ta], rb = (A.dim == 3) ? A.ratio[tb] : 1;
    double areaf = A.dxf[ta], areac = A.dxc[ta];
    if (A.dim == 3) {
        areaf = areaf * A.dxf[tb];
        areac = areac * A.dxc[tb];
    }
    int ci[3] = {0, 0, 0};
    ci[dir] = A.origin[dir] + (side ? A.nf[dir] / A.ratio[dir] : 0);
    ci[ta] = A.origin[ta] + ca;
    if (A.dim == 3 || tb < A.dim) ci[tb] = A.origin[tb] + cb;
    if (A.dim == 2) ci[2] = 0;
    const long long fc = amr_side_index(A.nc, dir, ci[0], ci[1], ci[2]);
    for (int e = 0; e < A.neq; e++) {
        const double* s = A.fsum[(2 * dir + side) * A.neq + e];
        double spv = 0.0;
        for (int ib = 0; ib < rb; ib++)
            for (int ia = 0; ia < ra; ia++) spv = spv + s[(ca * ra + ia) + (long long)A.nf[ta] * (cb * rb + ib)] * areaf;
        A.Fc[dir * A.neq + e][fc] = spv / areac;
    }
}

struct AmrExtrapArgs {
    AmrLayout L;
    int dir, side, ncomp;
    double* U[HB2A_MAXC];
};

/* thread t: ghost cell t of the slab beyond face (dir, side); the other directions run over the INTERIOR */
HB2A_HD void amr_extrapolate_thread(const AmrExtrapArgs& A, long long t)
{
    const int dir = A.dir;
    int ext[3] = {A.L.n[0], A.L.n[1], A.L.n[2]};
    ext[dir] = A.L.g[dir];
    int c[3];
    c[0] = (int)(t % ext[0]);
    c[1] = (int)((t / ext[0]) % ext[1]);
    c[2] = (int)(t / ((long long)ext[0] * ext[1]));
    int p[3] = {c[0], c[1], c[2]};
    if (A.side == 0) {
        c[dir] = c[dir] - A.L.g[dir];
        p[dir] = 0;
    } else {
        c[dir] = A.L.n[dir] + c[dir];
        p[dir] = A.L.n[dir] - 1;
    }
    const long long xd = amr_cidx(A.L, c[0], c[1], c[2]), xs = amr_cidx(A.L, p[0], p[1], p[2]);
    for (int q = 0; q < A.ncomp; q++) A.U[q][xd] = A.U[q][xs];
}

inline void amr_make_layout(int dim, const int* n, int g, AmrLayout* L)
{
    L->dim = dim;
    for (int a = 0; a < 3; a++) {
        L->n[a] = (a < dim) ? n[a] : 1;
        L->g[a] = (a < dim) ? g : 0;
    }
    L->cs[0] = 1;
    L->cs[1] = L->n[0] + 2 * L->g[0];
    L->cs[2] = L->cs[1] * (L->n[1] + 2 * L->g[1]);
    L->ncell_g = L->cs[2] * (L->n[2] + 2 * L->g[2]);
}

}  // namespace hb2
