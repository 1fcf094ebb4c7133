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
    int tiled;                    /* 3-D: marching tiled kernels, HB2_DIFF_TILED=1 (measured slower than the grid-stride forms) */
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

template <int DIM>
__global__ void __launch_bounds__(256) k_advance_ns(const __grid_constant__ NsArgs A)
{
    const long long total = (long long)A.G.n[0] * A.G.n[1] * A.G.n[2];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) advance_ns_thread<DIM>(A, t);
}

__global__ void __launch_bounds__(256) k_diff_fill_periodic(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffStatePtrs A,
                                                            int ncomp, int mask)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < G.ncell_g; t += stride)
        diff_fill_periodic_thread(G, A, ncomp, mask, t);
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

/* ---- tiled 3-D kernels of the node stage and of the flux-free update -------------------------------------------------------
 * The grid-stride forms above leave every stencil load to L1/L2 (ncu, 256^3: k_diff_node_all 1.06 ms at 44 % of HBM,
 * k_diff_divergence_accumulate 1.50 ms at 28 %, both > 75 % long-scoreboard stalls).  Here a 32 x 8 thread block owns a
 * 32 x 8 column of nodes (cells) and MARCHES along z: values along z live in a per-thread register ring of seven planes,
 * values of the current plane in a shared-memory tile with a halo of three, the next plane's global loads are issued
 * before the current plane is processed.  Same operations in the same order as the grid-stride kernels: bit-identical.
 *
 * k_diff_node_tiled: conservative variables in (the primitives never reach HBM: k_diff_primitives is fused away), the
 * twelve node-flux arrays out.  Algorithmic traffic 40 B read (+ halo re-reads from L2) + 96 B written per node. */
constexpr int TX = 32, TY = 8, HALO = 3;
constexpr int SX = TX + 2 * HALO, SY = TY + 2 * HALO;          /* 38 x 14 halo tile */
constexpr int NHALO = SX * SY - TX * TY;                       /* 276 halo cells per plane */

__device__ __forceinline__ void halo_cell(int h, int& hx, int& hy)
{
    /* h-th halo cell of the 38 x 14 tile: the three full rows below, the three above, then the side columns */
    if (h < HALO * SX) {
        hx = h % SX;
        hy = h / SX;
    } else if (h < 2 * HALO * SX) {
        const int q = h - HALO * SX;
        hx = q % SX;
        hy = TY + HALO + q / SX;
    } else {
        const int q = h - 2 * HALO * SX;
        const int row = q / (2 * HALO), c = q % (2 * HALO);
        hy = HALO + row;
        hx = (c < HALO) ? c : TX + c;
    }
}

__global__ void __launch_bounds__(256, 2) k_diff_node_tiled(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffConsts K,
                                                            const __grid_constant__ DiffPtrs Q, const __grid_constant__ DiffAllPtrs A)
{
    __shared__ double sP[4][SY][SX];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    /* node coordinates of the tile origin in the domain extended by three (nodes -3 .. n + 2) */
    const int i0 = (int)blockIdx.x * TX - 3, j0 = (int)blockIdx.y * TY - 3;
    const int i = i0 + tx, j = j0 + ty;
    const bool own = i < G.n[0] + 3 && j < G.n[1] + 3;           /* the node exists */
    const bool cell = i < G.n[0] + 6 && j < G.n[1] + 6;          /* the cell exists (a neighbour of an existing node may not be a node) */
    const int kz_lo = -3, kz_hi = G.n[2] + 3;
    const long long col = (i + G.g[0]) + G.cs[1] * (j + G.g[1]);
    /* halo cells this thread converts: h = tid and tid + 256 */
    int hx[2], hy[2];
    bool hok[2];
    long long hcol[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int h = (int)threadIdx.x + 256 * r;
        hok[r] = h < NHALO;
        hx[r] = hy[r] = 0;
        if (hok[r]) halo_cell(h, hx[r], hy[r]);
        const int ci = i0 - HALO + hx[r], cj = j0 - HALO + hy[r];
        hok[r] = hok[r] && ci < G.n[0] + 6 && cj < G.n[1] + 6;      /* inside the ghost box (>= -6 by construction) */
        hcol[r] = (ci + G.g[0]) + G.cs[1] * (cj + G.g[1]);
    }
    auto load_prims = [&](long long x, double (&P)[4]) {
        double q[5];
#pragma unroll
        for (int c = 0; c < 5; c++) q[c] = Q.Q[c][x];
        diff_primitives<3>(q, K, P);
    };
    /* register ring: ring[v][m] = primitive v of the own column at plane kz - 3 + m */
    double ring[4][7];
#pragma unroll
    for (int v = 0; v < 4; v++)
#pragma unroll
        for (int m = 0; m < 7; m++) ring[v][m] = 0.0;
    if (cell) {
#pragma unroll
        for (int m = 1; m < 7; m++) {           /* planes kz_lo - 3 + (m - 1) .. : after the first shift they sit at m - 1 */
            double P[4];
            load_prims(col + G.cs[2] * (kz_lo - 3 + (m - 1) + G.g[2]), P);
#pragma unroll
            for (int v = 0; v < 4; v++) ring[v][m] = P[v];
        }
    }
    for (int kz = kz_lo; kz < kz_hi; kz++) {
        /* shift the ring and take in plane kz + 3 */
#pragma unroll
        for (int v = 0; v < 4; v++)
#pragma unroll
            for (int m = 0; m < 6; m++) ring[v][m] = ring[v][m + 1];
        if (cell) {
            double P[4];
            load_prims(col + G.cs[2] * (kz + 3 + G.g[2]), P);
#pragma unroll
            for (int v = 0; v < 4; v++) ring[v][6] = P[v];
        }
        /* the plane of the nodes: own cell from the ring, halo cells from global memory */
        if (cell) {
#pragma unroll
            for (int v = 0; v < 4; v++) sP[v][ty + HALO][tx + HALO] = ring[v][3];
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            if (hok[r]) {
                double P[4];
                load_prims(hcol[r] + G.cs[2] * (kz + G.g[2]), P);
#pragma unroll
                for (int v = 0; v < 4; v++) sP[v][hy[r]][hx[r]] = P[v];
            }
        }
        __syncthreads();
        if (own) {
            double der[4][3];
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const double* c = &sP[v][ty + HALO][tx + HALO];
                der[v][0] = diff_first_derivative6(c[-3], c[-2], c[-1], c[1], c[2], c[3], G.dx_inv[0]);
                der[v][1] = diff_first_derivative6(c[-3 * SX], c[-2 * SX], c[-SX], c[SX], c[2 * SX], c[3 * SX], G.dx_inv[1]);
                der[v][2] = diff_first_derivative6(ring[v][0], ring[v][1], ring[v][2], ring[v][4], ring[v][5], ring[v][6], G.dx_inv[2]);
            }
            const double vel[3] = {ring[0][3], ring[1][3], ring[2][3]};
            double Fn[3][5];
            diff_node_flux_from_derivatives<3>(K, vel, der, Fn);
            const long long x = col + G.cs[2] * (kz + G.g[2]);
#pragma unroll
            for (int f = 0; f < 3; f++)
#pragma unroll
                for (int e = 1; e < 5; e++) A.Fn[f][e][x] = Fn[f][e];
        }
        __syncthreads();
    }
}

/* k_diff_div_tiled: U += beta (-div F_d) from the node fluxes of the three directions; F^z along z in a register ring,
 * F^x / F^y of the current plane in shared-memory tiles with a halo of three along their own direction.  Algorithmic
 * traffic 96 B of node fluxes + 64 B read-modify-write of the state per cell. */
__global__ void __launch_bounds__(256, 2) k_diff_div_tiled(const __grid_constant__ NsDivArgs A)
{
    __shared__ double sX[4][TY][SX];
    __shared__ double sY[4][SY][TX];
    const DiffGeom &G = A.G6, &GU = A.GU;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = (int)blockIdx.x * TX, j0 = (int)blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const bool own = i < G.n[0] && j < G.n[1];
    const bool cx_ok = i < G.n[0] + 3 && j < G.n[1];             /* nodes beside the last cells are somebody's neighbours */
    const bool cy_ok = i < G.n[0] && j < G.n[1] + 3;
    const long long col = (i + G.g[0]) + G.cs[1] * (j + G.g[1]);
    const long long colU = (i + GU.g[0]) + GU.cs[1] * (j + GU.g[1]);
    /* halo loads of this thread: x halo 6 x 8 = 48 cells (threads 0..47), y halo 6 x 32 = 192 cells (threads 48..239) */
    const int t = (int)threadIdx.x;
    const bool hx_on = t < 2 * HALO * TY;
    const int hxr = t / (2 * HALO), hxc = t % (2 * HALO);                  /* row, column index 0..5 */
    const int hx_s = (hxc < HALO) ? hxc : TX + hxc;                          /* position in the 38-wide row */
    const int hx_i = i0 - HALO + hx_s, hx_j = j0 + hxr;
    const bool hx_ok = hx_on && hx_i < G.n[0] + 3 && hx_j < G.n[1];
    const long long hx_col = (hx_i + G.g[0]) + G.cs[1] * (hx_j + G.g[1]);
    const int u = t - 2 * HALO * TY;
    const bool hy_on = u >= 0 && u < 2 * HALO * TX;
    const int hyr = hy_on ? u / TX : 0, hyc = hy_on ? u % TX : 0;           /* halo row 0..5, column */
    const int hy_s = (hyr < HALO) ? hyr : TY + hyr;                          /* position in the 14-high column */
    const int hy_i = i0 + hyc, hy_j = j0 - HALO + hy_s;
    const bool hy_ok = hy_on && hy_i < G.n[0] && hy_j < G.n[1] + 3;
    const long long hy_col = (hy_i + G.g[0]) + G.cs[1] * (hy_j + G.g[1]);

    double ring[4][7];                       /* F^z of equation e + 1 at planes k - 3 + m */
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
        for (int m = 0; m < 7; m++) ring[e][m] = 0.0;
    if (own) {
#pragma unroll
        for (int m = 1; m < 7; m++)
#pragma unroll
            for (int e = 0; e < 4; e++) ring[e][m] = A.Fn[2][e + 1][col + G.cs[2] * (-3 + (m - 1) + G.g[2])];
    }
    for (int k = 0; k < G.n[2]; k++) {
        const long long zoff = G.cs[2] * (k + G.g[2]);
#pragma unroll
        for (int e = 0; e < 4; e++)
#pragma unroll
            for (int m = 0; m < 6; m++) ring[e][m] = ring[e][m + 1];
        if (own) {
#pragma unroll
            for (int e = 0; e < 4; e++) ring[e][6] = A.Fn[2][e + 1][col + G.cs[2] * (k + 3 + G.g[2])];
        }
        if (cx_ok) {
#pragma unroll
            for (int e = 0; e < 4; e++) sX[e][ty][tx + HALO] = A.Fn[0][e + 1][col + zoff];
        }
        if (cy_ok) {
#pragma unroll
            for (int e = 0; e < 4; e++) sY[e][ty + HALO][tx] = A.Fn[1][e + 1][col + zoff];
        }
        if (hx_ok) {
#pragma unroll
            for (int e = 0; e < 4; e++) sX[e][hxr][hx_s] = A.Fn[0][e + 1][hx_col + zoff];
        }
        if (hy_ok) {
#pragma unroll
            for (int e = 0; e < 4; e++) sY[e][hy_s][hyc] = A.Fn[1][e + 1][hy_col + zoff];
        }
        __syncthreads();
        if (own) {
            const long long xu = colU + GU.cs[2] * (k + GU.g[2]);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const double* cx = &sX[e][ty][tx + HALO];
                const double* cy = &sY[e][ty + HALO][tx];
                const double FxL = diff_reconstruct6(cx[-3], cx[-2], cx[-1], cx[0], cx[1], cx[2], A.dt);
                const double FxR = diff_reconstruct6(cx[-2], cx[-1], cx[0], cx[1], cx[2], cx[3], A.dt);
                const double FyB = diff_reconstruct6(cy[-3 * TX], cy[-2 * TX], cy[-TX], cy[0], cy[TX], cy[2 * TX], A.dt);
                const double FyT = diff_reconstruct6(cy[-2 * TX], cy[-TX], cy[0], cy[TX], cy[2 * TX], cy[3 * TX], A.dt);
                double div = -(FxR - FxL) / G.dx[0] - (FyT - FyB) / G.dx[1];
                const double FzB = diff_reconstruct6(ring[e][0], ring[e][1], ring[e][2], ring[e][3], ring[e][4], ring[e][5], A.dt);
                const double FzF = diff_reconstruct6(ring[e][1], ring[e][2], ring[e][3], ring[e][4], ring[e][5], ring[e][6], A.dt);
                div -= (FzF - FzB) / G.dx[2];
                A.U[e + 1][xu] += A.beta * div;
            }
        }
        __syncthreads();
    }
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

/* primitives, then the node fluxes of all directions in one pass (every derivative evaluated once) into the plan's
 * per-direction scratch sets */
template <int DIM>
int node_stage(hb2_diff_plan_t p, const double* const* Q)
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
        if (p->tiled) {
            dim3 grid((p->G.n[0] + 6 + TX - 1) / TX, (p->G.n[1] + 6 + TY - 1) / TY);
            k_diff_node_tiled<<<grid, 256, 0, p->stream>>>(p->G, p->K, A, N);
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

template <int DIM>
int run_flux(hb2_diff_plan_t p, const double* const* Q, double dt, double* const* flux)
{
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
    int rc = node_stage<DIM>(p, Q);
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
    if (DIM == 3 && p->tiled) {
        dim3 grid((p->G.n[0] + TX - 1) / TX, (p->G.n[1] + TY - 1) / TY);
        k_diff_div_tiled<<<grid, 256, 0, p->stream>>>(D);
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
    {
        const char* v = getenv("HB2_DIFF_TILED");
        p->tiled = (v && *v) ? atoi(v) : 0;
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
    k_diff_fill_periodic<<<grid_for(p->G.ncell_g, p->sm_count), 256, 0, p->stream>>>(p->G, A, p->neq, periodic_mask);
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
