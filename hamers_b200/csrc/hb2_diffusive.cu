/*
 * hb2_diffusive.cu -- SURVEY.md row f4: kernels and C ABI of the node-based sixth-order diffusive flux
 * (DiffusiveFluxReconstructorNodeSixthOrder of the reference, "SIXTH_ORDER" in its input decks) and of the
 * Navier-Stokes stage update that consumes it.  Entry points are declared in include/hamers_b200.h.
 *
 * HBM-bound streaming kernels with x-contiguous (coalesced) accesses and grid-stride loops over a grid sized as a multiple
 * of the SM count:
 *   k_diff_primitives   ghost box: 5 (4) conservative doubles in, velocity + temperature out
 *   k_diff_node_all     cells extended by 3 in every direction: the twelve sixth-order derivatives of the primitives, each
 *                       evaluated once (stencil reads served by L1/L2: a primitive value is used by 18 neighbouring
 *                       nodes), diffusivities on the fly, node fluxes of the momentum and energy equations of ALL flux
 *                       directions out
 *   k_diff_face<FDIR>   faces: six-node reconstruction, times dt, all equations out (the continuity flux is +0.0 like
 *                       the reference's fillAll(0))                                  -- the materialised route
 *   k_diff_divergence_accumulate   both faces of a cell in every direction, differenced on the spot: U += beta (-div F_d)
 *                                                                                    -- the flux-free route
 * No CPU fallback: every entry point needs a CUDA device.  Built with -fmad=false: reference operation order.
 */
#include "../../include/hamers_b200.h"
#include "hb2_diffusive.cuh"
#include "hb2_diffusive_march.cuh"

#include <cuda_runtime.h>
#include <cstdlib>
#include <string>

namespace hb2 {
int set_error(int code, const std::string& msg);   /* hb2_abi.cu: what hb2_last_error() returns */
}
using namespace hb2;

#define HB2D_CUDA(call)                                                                             \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return set_error(-100 - (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

struct hb2_diff_plan_s {
    hb2_diffusive_desc d;
    DiffGeom G;
    DiffConsts K;
    int device, neq, sm_count;
    cudaStream_t stream;
    double* P[4];                 /* primitive scratch on the ghost box */
    double* Fn[5];                /* node-flux scratch (equation 0 unused) */
    double* FnDir[3][5];          /* one node-flux set per direction, for the flux-free update (allocated on first use;
                                     FnDir[0] aliases Fn) */
    double* stQ[5];               /* staging of the host-buffer entry point */
    double* stF[15];
    long long nside[3];
    long long launches;
    int reconstructor;            /* HB2_DIFF_NODE_SIXTH_ORDER (default) / HB2_DIFF_MIDPOINT_SIXTH_ORDER */
    int math;                     /* arithmetic of the flux-free route: HB2_MATH_EXACT (default) / HB2_MATH_FAST */
    int marching;                 /* 3-D: marching kernels with an asynchronous load pipeline (hb2_diffusive_march.cuh; HB2_DIFF_MARCH,
                                     default 1); 0: the grid-stride forms */
    int bricks;                   /* 3-D: 32 x 4 x 2 brick index map of the grid-stride kernels (HB2_DIFF_BRICKS, default 1) */
};

namespace {

template <int DIM>
__global__ void __launch_bounds__(256) k_diff_primitives(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffConsts K,
                                                         const __grid_constant__ DiffPtrs A)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < G.ncell_g; x += stride)
        diff_primitives_thread<DIM>(K, A, x);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_diff_node_all(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffConsts K,
                                                       const __grid_constant__ DiffAllPtrs A)
{
    const long long total = diff_node_all_count<DIM>(G);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
        diff_node_all_thread<DIM>(G, K, A, t);
}

/* (defined before its users below) */
/* The same threads with a block-tiled index map (3-D): a block of 256 threads owns a 32 x 4 x 2 brick of the index space, so
 * that the y- and z-neighbours of the sixth-order stencils are loaded by the same block and hit in L1 (a linear map gives a
 * block one x-row: every y / z neighbour comes from L2).  ext: extents of the index space the thread function decodes. */
constexpr int BX = 32, BY = 4, BZ = 2;
__device__ __forceinline__ bool brick_index(long long brick, int e0, int e1, int e2, long long& t)
{
    const int b0 = (e0 + BX - 1) / BX, b1 = (e1 + BY - 1) / BY;
    const int bz = (int)(brick / ((long long)b0 * b1));
    const int r = (int)(brick % ((long long)b0 * b1));
    const int i = (r % b0) * BX + (threadIdx.x & 31);
    const int j = (r / b0) * BY + ((threadIdx.x >> 5) & 3);
    const int k = bz * BZ + (threadIdx.x >> 7);
    t = i + (long long)e0 * (j + (long long)e1 * k);
    return i < e0 && j < e1 && k < e2;
}
__host__ __device__ inline long long brick_count(int e0, int e1, int e2)
{
    return (long long)((e0 + BX - 1) / BX) * ((e1 + BY - 1) / BY) * ((e2 + BZ - 1) / BZ);
}

__global__ void __launch_bounds__(256) k_diff_node_all_bricks(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffConsts K,
                                                              const __grid_constant__ DiffAllPtrs A)
{
    const int e0 = G.n[0] + 6, e1 = G.n[1] + 6, e2 = G.n[2] + 6;
    const long long nb = brick_count(e0, e1, e2);
    for (long long b = blockIdx.x; b < nb; b += gridDim.x) {
        long long t;
        if (brick_index(b, e0, e1, e2, t)) diff_node_all_thread<3>(G, K, A, t);
    }
}

__global__ void __launch_bounds__(256) k_diff_divergence_accumulate_bricks(const __grid_constant__ NsDivArgs A)
{
    const int e0 = A.G6.n[0], e1 = A.G6.n[1], e2 = A.G6.n[2];
    const long long nb = brick_count(e0, e1, e2);
    for (long long b = blockIdx.x; b < nb; b += gridDim.x) {
        long long t;
        if (brick_index(b, e0, e1, e2, t)) diff_divergence_accumulate_thread<3>(A, t);
    }
}

template <int DIM, int FDIR>
__global__ void __launch_bounds__(256) k_diff_face(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffPtrs A, double dt)
{
    const long long total = diff_face_count<DIM, FDIR>(G);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
        diff_face_thread<DIM, FDIR>(G, A, dt, t);
}

template <int DIM, int FDIR>
__global__ void __launch_bounds__(256) k_diff_mid_flux(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffConsts K,
                                                       const __grid_constant__ DiffMidPtrs A)
{
    const long long total = diff_mid_count<DIM, FDIR>(G);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
        diff_mid_flux_thread<DIM, FDIR>(G, K, A, t);
}

template <int DIM, int FDIR>
__global__ void __launch_bounds__(256) k_diff_mid_face(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffMidPtrs A, double dt)
{
    const long long total = diff_face_count<DIM, FDIR>(G);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
        diff_mid_face_thread<DIM, FDIR>(G, A, dt, t);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_advance_ns(const __grid_constant__ NsArgs A)
{
    const long long total = (long long)A.G.n[0] * A.G.n[1] * A.G.n[2];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) advance_ns_thread<DIM>(A, t);
}

/* ghost cells only: the 2 g_z full planes, then the 2 g_y rows of every interior plane, then the 2 g_x cells of every interior
 * row (a loop over the whole ghost box spends most of its threads on interior cells: 0.14 ms at 256^3 for 3 % of the cells) */
__host__ __device__ inline long long fill_ghost_count(const DiffGeom& G)
{
    return 2LL * G.g[2] * G.gd[0] * G.gd[1] + (long long)G.n[2] * 2 * G.g[1] * G.gd[0] + (long long)G.n[2] * G.n[1] * 2 * G.g[0];
}
__global__ void __launch_bounds__(256) k_diff_fill_periodic(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffStatePtrs A,
                                                            int ncomp, int mask)
{
    const long long nz = 2LL * G.g[2] * G.gd[0] * G.gd[1], ny = (long long)G.n[2] * 2 * G.g[1] * G.gd[0];
    const long long total = fill_ghost_count(G);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        int i, j, k;                                     /* ghost-box indices */
        if (t < nz) {
            const int p = (int)(t / ((long long)G.gd[0] * G.gd[1]));
            const int r = (int)(t % ((long long)G.gd[0] * G.gd[1]));
            k = p < G.g[2] ? p : p + G.n[2];
            j = r / G.gd[0];
            i = r % G.gd[0];
        } else if (t < nz + ny) {
            const long long u = t - nz;
            const int r = (int)((u / G.gd[0]) % (2 * G.g[1]));
            k = G.g[2] + (int)(u / ((long long)2 * G.g[1] * G.gd[0]));
            j = r < G.g[1] ? r : r + G.n[1];
            i = (int)(u % G.gd[0]);
        } else {
            const long long u = t - nz - ny;
            const int c = (int)(u % (2 * G.g[0]));
            k = G.g[2] + (int)(u / ((long long)2 * G.g[0] * G.n[1]));
            j = G.g[1] + (int)((u / (2 * G.g[0])) % G.n[1]);
            i = c < G.g[0] ? c : c + G.n[0];
        }
        diff_fill_periodic_thread(G, A, ncomp, mask, i + G.cs[1] * j + G.cs[2] * k);
    }
}

__global__ void __launch_bounds__(256) k_diff_extract_view(const __grid_constant__ DiffGeom Gs, const __grid_constant__ DiffGeom Gd,
                                                           const __grid_constant__ DiffStatePtrs A, int ncomp)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < Gd.ncell_g; t += stride)
        diff_extract_view_thread(Gs, Gd, A, ncomp, t);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_diff_accumulate(const __grid_constant__ NsAccArgs A)
{
    const long long total = (long long)A.G.n[0] * A.G.n[1] * A.G.n[2];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) diff_accumulate_thread<DIM>(A, t);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_diff_divergence_accumulate(const __grid_constant__ NsDivArgs A)
{
    const long long total = (long long)A.G6.n[0] * A.G6.n[1] * A.G6.n[2];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
        diff_divergence_accumulate_thread<DIM>(A, t);
}

/* max over the ghost box of the diffusive spectral radius.  Non-negative doubles order like their bit patterns. */
template <int DIM>
__global__ void __launch_bounds__(256) k_diff_spectral_radius(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffConsts K,
                                                              double c_p_eos, const double* __restrict__ rho, unsigned long long* out)
{
    double m = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < G.ncell_g; x += stride)
        m = fmax(m, diff_spectral_radius_cell<DIM>(G, K, c_p_eos, rho[x]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

/* dynamic shared memory of the marching kernels (function attributes are per device) */
cudaError_t march_attr(int dev)
{
    static unsigned long long attr_set = 0;
    if ((attr_set >> (dev & 63)) & 1ull) return cudaSuccess;
    const int nb = (int)(march::NODE_SMEM_DOUBLES * sizeof(double)), db = (int)(march::DIV_SMEM_DOUBLES * sizeof(double));
    cudaError_t e = cudaFuncSetAttribute(march::k_diff_node_march<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, nb);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(march::k_diff_node_march<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, nb);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(march::k_diff_div_march<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, db);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(march::k_diff_div_march<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, db);
    if (e != cudaSuccess) return e;
    attr_set |= 1ull << (dev & 63);
    return cudaSuccess;
}

int grid_for(long long work, int sm_count)
{
    const long long blocks = (work + 255) / 256;
    const long long cap = (long long)sm_count * 8;       /* 8 resident 256-thread CTAs per SM, grid-stride beyond */
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template <int DIM, int FDIR>
int launch_dir(hb2_diff_plan_t p, const DiffPtrs& A, double dt)
{
    const DiffGeom& G = p->G;
    k_diff_face<DIM, FDIR><<<grid_for(p->nside[FDIR], p->sm_count), 256, 0, p->stream>>>(G, A, dt);
    p->launches += 1;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

/* Re-associated 3-D route: F^f of momentum a < f shares the array of F^a of momentum f (symmetric stress, bit for bit; see
 * k_diff_node_march<1>).  Fn[f][e]: e = 1 + a. */
template <class T>
void diff_alias_symmetric(T (&Fn)[3][5])
{
    Fn[1][1] = Fn[0][2];
    Fn[2][1] = Fn[0][3];
    Fn[2][2] = Fn[1][3];
}

/* primitives, then the node fluxes of all directions in one pass (every derivative evaluated once) into the plan's
 * per-direction scratch sets */
template <int DIM>
int node_stage(hb2_diff_plan_t p, const double* const* Q, bool fast = false)
{
    const size_t bytes = sizeof(double) * (size_t)p->G.ncell_g;
    for (int f = 0; f < DIM; f++)
        for (int e = 1; e < DIM + 2; e++) {
            if (f == 0) p->FnDir[0][e] = p->Fn[e];
            if (!p->FnDir[f][e]) {
                HB2D_CUDA(cudaMalloc(&p->FnDir[f][e], bytes));
                HB2D_CUDA(cudaMemsetAsync(p->FnDir[f][e], 0, bytes, p->stream));
            }
        }
    DiffPtrs A{};
    for (int c = 0; c < DIM + 2; c++) A.Q[c] = Q[c];
    for (int v = 0; v < DIM + 1; v++) A.P[v] = p->P[v];
    DiffAllPtrs N{};
    for (int v = 0; v < DIM + 1; v++) N.P[v] = p->P[v];
    for (int f = 0; f < DIM; f++)
        for (int e = 0; e < DIM + 2; e++) N.Fn[f][e] = p->FnDir[f][e];
    if constexpr (DIM == 3) {
        if (p->marching) {
            using namespace march;
            HB2D_CUDA(march_attr(p->device));
            dim3 grid((p->G.n[0] + 9 + TX - 1) / TX, (p->G.n[1] + 6 + TY - 1) / TY);
            const int seg = march_seg_len((long long)grid.x * grid.y, p->G.n[2] + 6, p->sm_count);
            grid.z = (p->G.n[2] + 6 + seg - 1) / seg;
            DiffFast FK;
            make_diff_fast(p->G, p->K, 0.0, 0.0, &FK);
            if (fast) diff_alias_symmetric(N.Fn);
            if (fast)
                k_diff_node_march<1><<<grid, NT, NODE_SMEM_DOUBLES * sizeof(double), p->stream>>>(p->G, p->K, FK, A, N, seg);
            else
                k_diff_node_march<0><<<grid, NT, NODE_SMEM_DOUBLES * sizeof(double), p->stream>>>(p->G, p->K, FK, A, N, seg);
            p->launches += 1;
            HB2D_CUDA(cudaGetLastError());
            return 0;
        }
    }
    k_diff_primitives<DIM><<<grid_for(p->G.ncell_g, p->sm_count), 256, 0, p->stream>>>(p->G, p->K, A);
    if (DIM == 3 && p->bricks) {
        const long long nb = brick_count(p->G.n[0] + 6, p->G.n[1] + 6, p->G.n[2] + 6);
        k_diff_node_all_bricks<<<(unsigned)(nb < p->sm_count * 8LL ? nb : p->sm_count * 8LL), 256, 0, p->stream>>>(p->G, p->K, N);
    } else {
        k_diff_node_all<DIM><<<grid_for(diff_node_all_count<DIM>(p->G), p->sm_count), 256, 0, p->stream>>>(p->G, p->K, N);
    }
    p->launches += 2;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

/* midpoint family: primitives, then per direction the midpoint flux of every equation (plan scratch) and the faces */
template <int DIM, int FDIR>
int launch_mid_dir(hb2_diff_plan_t p, double dt, double* const* flux)
{
    DiffMidPtrs A{};
    for (int v = 0; v < DIM + 1; v++) A.P[v] = p->P[v];
    for (int e = 0; e < DIM + 2; e++) {
        A.Fm[e] = p->Fn[e];
        A.F[e] = flux[FDIR * (DIM + 2) + e];
    }
    k_diff_mid_flux<DIM, FDIR><<<grid_for(diff_mid_count<DIM, FDIR>(p->G), p->sm_count), 256, 0, p->stream>>>(p->G, p->K, A);
    k_diff_mid_face<DIM, FDIR><<<grid_for(p->nside[FDIR], p->sm_count), 256, 0, p->stream>>>(p->G, A, dt);
    p->launches += 2;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

template <int DIM>
int run_flux_midpoint(hb2_diff_plan_t p, const double* const* Q, double dt, double* const* flux)
{
    DiffPtrs A{};
    for (int c = 0; c < DIM + 2; c++) A.Q[c] = Q[c];
    for (int v = 0; v < DIM + 1; v++) A.P[v] = p->P[v];
    k_diff_primitives<DIM><<<grid_for(p->G.ncell_g, p->sm_count), 256, 0, p->stream>>>(p->G, p->K, A);
    p->launches += 1;
    HB2D_CUDA(cudaGetLastError());
    int rc = launch_mid_dir<DIM, 0>(p, dt, flux);
    if (!rc) rc = launch_mid_dir<DIM, 1>(p, dt, flux);
    if (!rc && DIM == 3) rc = launch_mid_dir<DIM, (DIM == 3 ? 2 : 1)>(p, dt, flux);
    return rc;
}

template <int DIM>
int run_flux(hb2_diff_plan_t p, const double* const* Q, double dt, double* const* flux)
{
    if (p->reconstructor == HB2_DIFF_MIDPOINT_SIXTH_ORDER) return run_flux_midpoint<DIM>(p, Q, dt, flux);
    int rc = node_stage<DIM>(p, Q);
    if (rc) return rc;
    for (int f = 0; f < DIM; f++) {
        DiffPtrs A{};
        for (int e = 0; e < DIM + 2; e++) {
            A.Fn[e] = p->FnDir[f][e];
            A.F[e] = flux[f * (DIM + 2) + e];
        }
        rc = f == 0 ? launch_dir<DIM, 0>(p, A, dt) : f == 1 ? launch_dir<DIM, 1>(p, A, dt) : launch_dir<DIM, (DIM == 3 ? 2 : 1)>(p, A, dt);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace

namespace {
template <int DIM>
int run_divergence(hb2_diff_plan_t p, const double* const* Q, double dt, int num_ghosts, double beta, double* const* U)
{
    const bool fast = DIM == 3 && p->marching && p->math == HB2_MATH_FAST;
    int rc = node_stage<DIM>(p, Q, fast);
    if (rc) return rc;
    NsDivArgs D{};
    D.G6 = p->G;
    make_diff_geom(DIM, p->d.n, p->d.dx, num_ghosts, &D.GU);
    D.neq = DIM + 2;
    D.beta = beta;
    D.dt = dt;
    for (int f = 0; f < DIM; f++)
        for (int e = 0; e < DIM + 2; e++) D.Fn[f][e] = p->FnDir[f][e];
    for (int e = 0; e < DIM + 2; e++) D.U[e] = U[e];
    const long long total = (long long)p->G.n[0] * p->G.n[1] * p->G.n[2];
    if (DIM == 3 && p->marching) {
        using namespace march;
        HB2D_CUDA(march_attr(p->device));
        dim3 grid((p->G.n[0] + TX - 1) / TX, (p->G.n[1] + TY - 1) / TY);
        const int seg = march_seg_len((long long)grid.x * grid.y, p->G.n[2], p->sm_count);
        grid.z = (p->G.n[2] + seg - 1) / seg;
        DiffFast FK;
        make_diff_fast(p->G, p->K, dt, beta, &FK);
        if (fast) diff_alias_symmetric(D.Fn);
        if (fast)
            k_diff_div_march<1><<<grid, NT, DIV_SMEM_DOUBLES * sizeof(double), p->stream>>>(D, FK, seg);
        else
            k_diff_div_march<0><<<grid, NT, DIV_SMEM_DOUBLES * sizeof(double), p->stream>>>(D, FK, seg);
    } else if (DIM == 3 && p->bricks) {
        const long long nb = brick_count(p->G.n[0], p->G.n[1], p->G.n[2]);
        k_diff_divergence_accumulate_bricks<<<(unsigned)(nb < p->sm_count * 8LL ? nb : p->sm_count * 8LL), 256, 0, p->stream>>>(D);
    } else {
        k_diff_divergence_accumulate<DIM><<<grid_for(total, p->sm_count), 256, 0, p->stream>>>(D);
    }
    p->launches += 1;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace

extern "C" {

int hb2_diffusive_plan_create(const hb2_diffusive_desc* d, hb2_diff_plan_t* out)
{
    if (!d || !out) return set_error(-1, "null argument");
    if (d->dim != 2 && d->dim != 3) return set_error(-2, "dim must be 2 or 3");
    for (int a = 0; a < d->dim; a++) {
        if (d->n[a] < 1) return set_error(-3, "patch dims must be positive");
        if (!(d->dx[a] > 0.0)) return set_error(-4, "grid spacing must be positive");
    }
    if (!(d->species_gamma > 1.0)) return set_error(-8, "species_gamma must be > 1");
    if (!(d->species_c_v > 0.0) || !(d->species_c_p > 0.0) || !(d->species_Pr > 0.0))
        return set_error(-30, "species_c_v, species_c_p and species_Pr must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return set_error(-20, "no CUDA device: hamers_b200 has no CPU fallback (the CPU oracle lives under oracle/ and is test infrastructure)");
    hb2_diff_plan_t p = new hb2_diff_plan_s();
    p->d = *d;
    p->device = d->device;
    if (p->device < 0) HB2D_CUDA(cudaGetDevice(&p->device));      /* -1: the calling thread's current device */
    if (p->device >= ndev) {
        delete p;
        return set_error(-21, "device index out of range");
    }
    HB2D_CUDA(cudaSetDevice(p->device));
    cudaDeviceProp prop;
    HB2D_CUDA(cudaGetDeviceProperties(&prop, p->device));
    p->sm_count = prop.multiProcessorCount;
    make_diff_geom(d->dim, d->n, d->dx, HB2_DIFF_G, &p->G);
    p->K.gamma = d->species_gamma;
    p->K.c_v = d->species_c_v;
    p->K.mu = d->species_mu;
    p->K.mu_v = d->species_mu_v;
    p->K.kappa = d->species_c_p * d->species_mu / d->species_Pr;      /* EquationOfThermalConductivityPrandtl.cpp:309 */
    p->neq = d->dim + 2;
    p->stream = nullptr;
    p->math = HB2_MATH_EXACT;
    p->reconstructor = HB2_DIFF_NODE_SIXTH_ORDER;
    {
        const char* v = getenv("HB2_DIFF_MARCH");
        p->marching = (v && *v) ? atoi(v) : 1;
        v = getenv("HB2_DIFF_BRICKS");
        p->bricks = (v && *v) ? atoi(v) : 1;
    }
    for (int a = 0; a < 3; a++) {
        long long s = 1;
        for (int b = 0; b < 3; b++) s *= p->G.n[b] + ((a == b && a < d->dim) ? 1 : 0);
        p->nside[a] = s;
    }
    const size_t bytes = sizeof(double) * (size_t)p->G.ncell_g;
    for (int v = 0; v < d->dim + 1; v++) HB2D_CUDA(cudaMalloc(&p->P[v], bytes));
    for (int e = 1; e < p->neq; e++) {
        HB2D_CUDA(cudaMalloc(&p->Fn[e], bytes));
        HB2D_CUDA(cudaMemset(p->Fn[e], 0, bytes));
    }
    *out = p;
    return 0;
}

int hb2_diffusive_plan_destroy(hb2_diff_plan_t p)
{
    if (!p) return 0;
    cudaSetDevice(p->device);
    for (int v = 0; v < 4; v++) cudaFree(p->P[v]);
    for (int e = 0; e < 5; e++) {
        cudaFree(p->Fn[e]);
        cudaFree(p->stQ[e]);
        for (int f = 1; f < 3; f++) cudaFree(p->FnDir[f][e]);     /* FnDir[0] aliases Fn */
    }
    for (int e = 0; e < 15; e++) cudaFree(p->stF[e]);
    delete p;
    return 0;
}

int hb2_diffusive_plan_set_stream(hb2_diff_plan_t p, void* stream)
{
    if (!p) return set_error(-1, "null plan");
    p->stream = (cudaStream_t)stream;
    return 0;
}

int hb2_diffusive_plan_set_math(hb2_diff_plan_t p, int32_t math)
{
    if (!p) return set_error(-1, "null plan");
    if (math != HB2_MATH_EXACT && math != HB2_MATH_FAST) return set_error(-32, "math must be HB2_MATH_EXACT or HB2_MATH_FAST");
    p->math = math;
    return 0;
}

int hb2_diffusive_plan_set_reconstructor(hb2_diff_plan_t p, int32_t reconstructor)
{
    if (!p) return set_error(-1, "null plan");
    if (reconstructor != HB2_DIFF_NODE_SIXTH_ORDER && reconstructor != HB2_DIFF_MIDPOINT_SIXTH_ORDER)
        return set_error(-33, "reconstructor must be HB2_DIFF_NODE_SIXTH_ORDER or HB2_DIFF_MIDPOINT_SIXTH_ORDER");
    p->reconstructor = reconstructor;
    return 0;
}

int hb2_diffusive_plan_launches(hb2_diff_plan_t p, int64_t* launches)
{
    if (!p || !launches) return set_error(-1, "null argument");
    *launches = p->launches;
    return 0;
}

int hb2_compute_diffusive_flux_dev(hb2_diff_plan_t p, const double* const* Q, double dt, double* const* flux)
{
    if (!p || !Q || !flux) return set_error(-1, "null argument");
    HB2D_CUDA(cudaSetDevice(p->device));
    return p->d.dim == 2 ? run_flux<2>(p, Q, dt, flux) : run_flux<3>(p, Q, dt, flux);
}

int hb2_compute_diffusive_flux_host(hb2_diff_plan_t p, const double* const* Q_host, double dt, double* const* flux_host)
{
    if (!p || !Q_host || !flux_host) return set_error(-1, "null argument");
    HB2D_CUDA(cudaSetDevice(p->device));
    const size_t cbytes = sizeof(double) * (size_t)p->G.ncell_g;
    const int dim = p->d.dim, neq = p->neq;
    const double* Qd[5];
    double* Fd[15];
    for (int c = 0; c < neq; c++) {
        if (!p->stQ[c]) HB2D_CUDA(cudaMalloc(&p->stQ[c], cbytes));
        HB2D_CUDA(cudaMemcpyAsync(p->stQ[c], Q_host[c], cbytes, cudaMemcpyHostToDevice, p->stream));
        Qd[c] = p->stQ[c];
    }
    for (int f = 0; f < dim; f++)
        for (int e = 0; e < neq; e++) {
            double*& slot = p->stF[f * neq + e];
            if (!slot) HB2D_CUDA(cudaMalloc(&slot, sizeof(double) * (size_t)p->nside[f]));
            Fd[f * neq + e] = slot;
        }
    int rc = hb2_compute_diffusive_flux_dev(p, Qd, dt, Fd);
    if (rc) return rc;
    for (int f = 0; f < dim; f++)
        for (int e = 0; e < neq; e++)
            HB2D_CUDA(cudaMemcpyAsync(flux_host[f * neq + e], Fd[f * neq + e], sizeof(double) * (size_t)p->nside[f],
                                      cudaMemcpyDeviceToHost, p->stream));
    HB2D_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int hb2_diffusive_accumulate_dev(hb2_diff_plan_t p, int32_t num_ghosts, double beta, const double* const* Fd, double* const* U)
{
    if (!p || !Fd || !U) return set_error(-1, "null argument");
    if (num_ghosts < 0) return set_error(-31, "num_ghosts must be >= 0");
    HB2D_CUDA(cudaSetDevice(p->device));
    const int dim = p->d.dim, neq = p->neq;
    NsAccArgs A{};
    make_diff_geom(dim, p->d.n, p->d.dx, num_ghosts, &A.G);
    A.neq = neq;
    A.beta = beta;
    for (int f = 0; f < dim * neq; f++) {
        A.Fd[f] = Fd[f];
        if (!A.Fd[f]) return set_error(-12, "diffusive flux row is NULL");
    }
    for (int e = 0; e < neq; e++) A.U[e] = U[e];
    const long long total = (long long)A.G.n[0] * A.G.n[1] * A.G.n[2];
    if (dim == 2)
        k_diff_accumulate<2><<<grid_for(total, p->sm_count), 256, 0, p->stream>>>(A);
    else
        k_diff_accumulate<3><<<grid_for(total, p->sm_count), 256, 0, p->stream>>>(A);
    p->launches++;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

int hb2_diffusive_divergence_accumulate_dev(hb2_diff_plan_t p, const double* const* Q, double dt, int32_t num_ghosts, double beta,
                                            double* const* U)
{
    if (!p || !Q || !U) return set_error(-1, "null argument");
    if (num_ghosts < 0) return set_error(-31, "num_ghosts must be >= 0");
    if (p->reconstructor != HB2_DIFF_NODE_SIXTH_ORDER)
        return set_error(-34, "the flux-free update exists for the node reconstructor; use hb2_compute_diffusive_flux_dev + "
                              "hb2_diffusive_accumulate_dev with the midpoint one");
    HB2D_CUDA(cudaSetDevice(p->device));
    return p->d.dim == 2 ? run_divergence<2>(p, Q, dt, num_ghosts, beta, U) : run_divergence<3>(p, Q, dt, num_ghosts, beta, U);
}

int hb2_diffusive_max_spectral_radius_dev(hb2_diff_plan_t p, const double* const* Q, double species_c_p_eos, double* out_dev)
{
    if (!p || !Q || !out_dev) return set_error(-1, "null argument");
    if (!(species_c_p_eos > 0.0)) return set_error(-30, "species_c_p_eos must be positive");
    HB2D_CUDA(cudaSetDevice(p->device));
    HB2D_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), p->stream));
    unsigned long long* o = (unsigned long long*)out_dev;
    if (p->d.dim == 2)
        k_diff_spectral_radius<2><<<grid_for(p->G.ncell_g, p->sm_count), 256, 0, p->stream>>>(p->G, p->K, species_c_p_eos, Q[0], o);
    else
        k_diff_spectral_radius<3><<<grid_for(p->G.ncell_g, p->sm_count), 256, 0, p->stream>>>(p->G, p->K, species_c_p_eos, Q[0], o);
    p->launches++;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

int hb2_diffusive_fill_ghosts_periodic_dev(hb2_diff_plan_t p, double* const* U, int32_t periodic_mask)
{
    if (!p || !U) return set_error(-1, "null argument");
    HB2D_CUDA(cudaSetDevice(p->device));
    DiffStatePtrs A{};
    for (int c = 0; c < p->neq; c++) A.U[c] = U[c];
    k_diff_fill_periodic<<<grid_for(fill_ghost_count(p->G), p->sm_count), 256, 0, p->stream>>>(p->G, A, p->neq, periodic_mask);
    p->launches++;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

int hb2_diffusive_extract_view_dev(hb2_diff_plan_t p, const double* const* U, int32_t num_ghosts, double* const* U_view)
{
    if (!p || !U || !U_view) return set_error(-1, "null argument");
    if (num_ghosts < 0 || num_ghosts > HB2_DIFF_G) return set_error(-31, "num_ghosts must be in 0..6");
    HB2D_CUDA(cudaSetDevice(p->device));
    DiffGeom Gd;
    make_diff_geom(p->d.dim, p->d.n, p->d.dx, num_ghosts, &Gd);
    DiffStatePtrs A{};
    for (int c = 0; c < p->neq; c++) {
        A.src[c] = U[c];
        A.U[c] = U_view[c];
    }
    k_diff_extract_view<<<grid_for(Gd.ncell_g, p->sm_count), 256, 0, p->stream>>>(p->G, Gd, A, p->neq);
    p->launches++;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

int hb2_advance_stage_ns_dev(hb2_diff_plan_t p, int32_t num_ghosts, int32_t ncoef, const double* alpha, const double* beta,
                             const double* const* U_int, const double* const* Fc_int, const double* const* Fd_int,
                             const double* const* S_int, double* const* U_out)
{
    if (!p || !alpha || !beta || !U_int || !Fc_int || !Fd_int || !S_int || !U_out) return set_error(-1, "null argument");
    if (ncoef < 1 || ncoef > HB2_MAXS) return set_error(-10, "ncoef out of range");
    if (num_ghosts < 0) return set_error(-31, "num_ghosts must be >= 0");
    HB2D_CUDA(cudaSetDevice(p->device));
    const int dim = p->d.dim, neq = p->neq;
    NsArgs A{};
    make_diff_geom(dim, p->d.n, p->d.dx, num_ghosts, &A.G);
    A.neq = neq;
    A.ncoef = ncoef;
    for (int m = 0; m < ncoef; m++) {
        A.alpha[m] = alpha[m];
        A.beta[m] = beta[m];
        for (int e = 0; e < neq; e++) {
            A.U[m][e] = U_int[m * neq + e];
            A.S[m][e] = S_int[m * neq + e];
            if (alpha[m] != 0.0 && !A.U[m][e]) return set_error(-11, "U_int row with alpha != 0 is NULL");
            if (beta[m] != 0.0 && !A.S[m][e]) return set_error(-12, "S_int row with beta != 0 is NULL");
        }
        for (int f = 0; f < dim * neq; f++) {
            A.Fc[m][f] = Fc_int[m * dim * neq + f];
            A.Fd[m][f] = Fd_int[m * dim * neq + f];
            if (beta[m] != 0.0 && (!A.Fc[m][f] || !A.Fd[m][f])) return set_error(-12, "flux row with beta != 0 is NULL");
        }
    }
    for (int e = 0; e < neq; e++) A.Uout[e] = U_out[e];
    const long long total = (long long)A.G.n[0] * A.G.n[1] * A.G.n[2];
    if (dim == 2)
        k_advance_ns<2><<<grid_for(total, p->sm_count), 256, 0, p->stream>>>(A);
    else
        k_advance_ns<3><<<grid_for(total, p->sm_count), 256, 0, p->stream>>>(A);
    p->launches++;
    HB2D_CUDA(cudaGetLastError());
    return 0;
}

}  /* extern "C" */
