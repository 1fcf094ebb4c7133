/*
 * hb2_amr.cu -- sm_100a kernels and C ABI of the two-level AMR operators around the convective hot path.
 * The arithmetic lives in hb2_amr.cuh (thread functions shared with the host emulation); every kernel is a grid-stride
 * loop over a small index range (ghost slabs, the cells under a fine patch, patch-boundary faces): HBM-bound copies.
 * Compiled with -fmad=false.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "hamers_b200.h"
#include "hb2_amr.cuh"

namespace hb2 {
int set_error(int code, const std::string& msg);      /* hb2_abi.cu: records the message for hb2_last_error() */
void* plan_stream(hb2_plan_t plan);
int plan_layout(hb2_plan_t plan, int* dim, int n[3], int* ghosts, int* ncomp);
void plan_count_launch(hb2_plan_t plan);
}

using namespace hb2;

static int fail(int code, const char* msg) { return hb2::set_error(code, msg); }

namespace {

__global__ void __launch_bounds__(256) k_amr_refine(const __grid_constant__ AmrRefineArgs A, long long total)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        amr_refine_thread(A, t);
}

__global__ void __launch_bounds__(256) k_amr_coarsen(const __grid_constant__ AmrCoarsenArgs A, long long total)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        amr_coarsen_thread(A, t);
}

/* blockIdx.y = 2 dir + side */
__global__ void __launch_bounds__(256) k_amr_fluxsum(const __grid_constant__ AmrFluxsumArgs A)
{
    const int dir = blockIdx.y >> 1, side = blockIdx.y & 1;
    if (dir >= A.dim) return;
    long long total = 1;
    for (int a = 0; a < A.dim; a++)
        if (a != dir) total *= A.n[a];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        amr_fluxsum_thread(A, dir, side, t);
}

__global__ void __launch_bounds__(256) k_amr_coarsen_fluxsum(const __grid_constant__ AmrCoarsenFluxsumArgs A)
{
    const int dir = blockIdx.y >> 1, side = blockIdx.y & 1;
    if (dir >= A.dim) return;
    long long total = 1;
    for (int a = 0; a < A.dim; a++)
        if (a != dir) total *= A.nf[a] / A.ratio[a];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        amr_coarsen_fluxsum_thread(A, dir, side, t);
}

__global__ void __launch_bounds__(256) k_amr_extrapolate(const __grid_constant__ AmrExtrapArgs A, long long total)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        amr_extrapolate_thread(A, t);
}

unsigned blocks_for(long long total)
{
    long long b = (total + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148LL * 16) b = 148LL * 16;
    return (unsigned)b;
}

int check_pair(const hb2_amr_pair* p)
{
    if (!p) return fail(-1, "null hb2_amr_pair");
    if (p->dim != 2 && p->dim != 3) return fail(-2, "hb2_amr_pair: dim must be 2 or 3");
    if (p->ncomp < 1 || p->ncomp > HB2A_MAXC || p->neq < 1 || p->neq > 12) return fail(-30, "hb2_amr_pair: ncomp / neq out of range");
    for (int a = 0; a < p->dim; a++) {
        if (p->ratio[a] < 1 || p->nf[a] < 1 || p->nc[a] < 1) return fail(-31, "hb2_amr_pair: ratio and patch dims must be positive");
        if (p->nf[a] % p->ratio[a] != 0) return fail(-32, "hb2_amr_pair: the fine patch must cover whole coarse cells");
        if (!(p->dxc[a] > 0.0) || !(p->dxf[a] > 0.0)) return fail(-4, "grid spacing must be positive");
    }
    if (p->ghosts_c < 0 || p->ghosts_f < 0) return fail(-33, "hb2_amr_pair: negative ghost width");
    return 0;
}

void pair_layouts(const hb2_amr_pair* p, AmrLayout* C, AmrLayout* F)
{
    amr_make_layout(p->dim, p->nc, p->ghosts_c, C);
    amr_make_layout(p->dim, p->nf, p->ghosts_f, F);
}

int launch_status(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof msg, "%s: %s", what, cudaGetErrorString(e));
        return fail(-20, msg);
    }
    return 0;
}

}  // namespace

extern "C" {

int hb2_amr_refine_dev(const hb2_amr_pair* p, const double* const* Uc_old, const double* const* Uc_new, double tfrac,
                       const int32_t lo[3], const int32_t hi[3], double* const* Uf, void* cuda_stream)
{
    if (int rc = check_pair(p)) return rc;
    if (!Uc_old || !Uf || !lo || !hi) return fail(-1, "hb2_amr_refine_dev: null argument");
    AmrRefineArgs A;
    memset(&A, 0, sizeof A);
    pair_layouts(p, &A.C, &A.F);
    long long total = 1;
    for (int a = 0; a < 3; a++) {
        const bool on = a < p->dim;
        A.ratio[a] = on ? p->ratio[a] : 1;
        A.origin[a] = on ? p->origin[a] : 0;
        A.lo[a] = on ? lo[a] : 0;
        A.hi[a] = on ? hi[a] : 1;
        A.dxc[a] = on ? p->dxc[a] : 1.0;
        A.dxf[a] = on ? p->dxf[a] : 1.0;
        if (A.hi[a] <= A.lo[a]) return 0;          /* empty box */
        if (on) {
            if (A.lo[a] < -p->ghosts_f || A.hi[a] > p->nf[a] + p->ghosts_f) return fail(-34, "hb2_amr_refine_dev: box outside the fine ghost box");
            /* the coarse cells read, one more on each side for the slopes, must lie in the coarse ghost box */
            const int c_lo = amr_floor_div(A.lo[a], A.ratio[a]) + A.origin[a] - 1;
            const int c_hi = amr_floor_div(A.hi[a] - 1, A.ratio[a]) + A.origin[a] + 1;
            if (c_lo < -p->ghosts_c || c_hi > p->nc[a] - 1 + p->ghosts_c) return fail(-35, "hb2_amr_refine_dev: the coarse stencil leaves the coarse ghost box");
        }
        total *= (A.hi[a] - A.lo[a]);
    }
    A.ncomp = p->ncomp;
    A.has_new = Uc_new ? 1 : 0;
    A.tfrac = tfrac;
    for (int c = 0; c < p->ncomp; c++) {
        A.Uold[c] = Uc_old[c];
        A.Unew[c] = Uc_new ? Uc_new[c] : nullptr;
        A.Uf[c] = Uf[c];
    }
    k_amr_refine<<<blocks_for(total), 256, 0, (cudaStream_t)cuda_stream>>>(A, total);
    return launch_status("hb2_amr_refine_dev");
}

int hb2_amr_coarsen_dev(const hb2_amr_pair* p, const double* const* Uf, const int32_t lo[3], const int32_t hi[3],
                        double* const* Uc, void* cuda_stream)
{
    if (int rc = check_pair(p)) return rc;
    if (!Uf || !Uc || !lo || !hi) return fail(-1, "hb2_amr_coarsen_dev: null argument");
    AmrCoarsenArgs A;
    memset(&A, 0, sizeof A);
    pair_layouts(p, &A.C, &A.F);
    long long total = 1;
    for (int a = 0; a < 3; a++) {
        const bool on = a < p->dim;
        A.ratio[a] = on ? p->ratio[a] : 1;
        A.origin[a] = on ? p->origin[a] : 0;
        A.lo[a] = on ? lo[a] : 0;
        A.hi[a] = on ? hi[a] : 1;
        A.dxc[a] = on ? p->dxc[a] : 1.0;
        A.dxf[a] = on ? p->dxf[a] : 1.0;
        if (A.hi[a] <= A.lo[a]) return 0;
        if (on && (A.lo[a] < A.origin[a] || A.hi[a] > A.origin[a] + p->nf[a] / p->ratio[a]))
            return fail(-36, "hb2_amr_coarsen_dev: coarse box not covered by the fine patch");
        total *= (A.hi[a] - A.lo[a]);
    }
    A.ncomp = p->ncomp;
    for (int c = 0; c < p->ncomp; c++) {
        A.Uf[c] = Uf[c];
        A.Uc[c] = Uc[c];
    }
    k_amr_coarsen<<<blocks_for(total), 256, 0, (cudaStream_t)cuda_stream>>>(A, total);
    return launch_status("hb2_amr_coarsen_dev");
}

int hb2_amr_fluxsum_update_dev(const hb2_amr_pair* p, const double* const* F_fine, double* const* fluxsum, void* cuda_stream)
{
    if (int rc = check_pair(p)) return rc;
    if (!F_fine || !fluxsum) return fail(-1, "hb2_amr_fluxsum_update_dev: null argument");
    AmrFluxsumArgs A;
    memset(&A, 0, sizeof A);
    A.dim = p->dim;
    A.neq = p->neq;
    long long most = 1;
    for (int a = 0; a < 3; a++) A.n[a] = (a < p->dim) ? p->nf[a] : 1;
    for (int d = 0; d < p->dim; d++) {
        long long tot = 1;
        for (int a = 0; a < p->dim; a++)
            if (a != d) tot *= A.n[a];
        if (tot > most) most = tot;
        for (int e = 0; e < p->neq; e++) {
            A.F[d * p->neq + e] = F_fine[d * p->neq + e];
            for (int s = 0; s < 2; s++) A.fsum[(2 * d + s) * p->neq + e] = fluxsum[(2 * d + s) * p->neq + e];
        }
    }
    dim3 grid(blocks_for(most), 2 * p->dim);
    k_amr_fluxsum<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(A);
    return launch_status("hb2_amr_fluxsum_update_dev");
}

int hb2_amr_coarsen_fluxsum_dev(const hb2_amr_pair* p, const double* const* fluxsum, double* const* F_coarse, void* cuda_stream)
{
    if (int rc = check_pair(p)) return rc;
    if (!fluxsum || !F_coarse) return fail(-1, "hb2_amr_coarsen_fluxsum_dev: null argument");
    AmrCoarsenFluxsumArgs A;
    memset(&A, 0, sizeof A);
    A.dim = p->dim;
    A.neq = p->neq;
    long long most = 1;
    for (int a = 0; a < 3; a++) {
        const bool on = a < p->dim;
        A.nf[a] = on ? p->nf[a] : 1;
        A.nc[a] = on ? p->nc[a] : 1;
        A.ratio[a] = on ? p->ratio[a] : 1;
        A.origin[a] = on ? p->origin[a] : 0;
        A.dxc[a] = on ? p->dxc[a] : 1.0;
        A.dxf[a] = on ? p->dxf[a] : 1.0;
        if (on && (A.origin[a] < 0 || A.origin[a] + A.nf[a] / A.ratio[a] > A.nc[a]))
            return fail(-37, "hb2_amr_coarsen_fluxsum_dev: the fine patch must lie inside the coarse patch");
    }
    for (int d = 0; d < p->dim; d++) {
        long long tot = 1;
        for (int a = 0; a < p->dim; a++)
            if (a != d) tot *= A.nf[a] / A.ratio[a];
        if (tot > most) most = tot;
        for (int e = 0; e < p->neq; e++) {
            A.Fc[d * p->neq + e] = F_coarse[d * p->neq + e];
            for (int s = 0; s < 2; s++) A.fsum[(2 * d + s) * p->neq + e] = fluxsum[(2 * d + s) * p->neq + e];
        }
    }
    dim3 grid(blocks_for(most), 2 * p->dim);
    k_amr_coarsen_fluxsum<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(A);
    return launch_status("hb2_amr_coarsen_fluxsum_dev");
}

int hb2_fill_ghosts_extrapolate_dev(hb2_plan_t plan, double* const* U, int32_t dir, int32_t side)
{
    if (!plan || !U) return fail(-1, "hb2_fill_ghosts_extrapolate_dev: null argument");
    int dim, n[3], g, ncomp;
    if (int rc = plan_layout(plan, &dim, n, &g, &ncomp)) return rc;
    if (dir < 0 || dir >= dim || (side != 0 && side != 1)) return fail(-38, "hb2_fill_ghosts_extrapolate_dev: bad direction / side");
    AmrExtrapArgs A;
    memset(&A, 0, sizeof A);
    amr_make_layout(dim, n, g, &A.L);
    A.dir = dir;
    A.side = side;
    A.ncomp = ncomp;
    for (int c = 0; c < ncomp; c++) A.U[c] = U[c];
    long long total = 1;
    for (int a = 0; a < 3; a++) total *= (a == dir) ? A.L.g[a] : A.L.n[a];
    k_amr_extrapolate<<<blocks_for(total), 256, 0, (cudaStream_t)plan_stream(plan)>>>(A, total);
    plan_count_launch(plan);
    return launch_status("hb2_fill_ghosts_extrapolate_dev");
}

}  // extern "C"
