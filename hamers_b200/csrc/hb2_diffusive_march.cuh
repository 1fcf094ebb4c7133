/*
 * hb2_diffusive_march.cuh -- SURVEY.md row f4, 3-D: the node stage and the flux-free update of the sixth-order diffusive
 * flux as MARCHING kernels with an asynchronous load pipeline (included by hb2_diffusive.cu only).
 *
 * The grid-stride forms leave every stencil load to L1 / L2 (ncu at 256^3: node fluxes 1.06 ms, divergence update 1.50 ms,
 * both > 75 % long-scoreboard stalls at 44 % / 28 % of HBM).  Here a 32 x 8 thread block owns a 32 x 8 column of nodes
 * (cells) and marches along z over a segment of planes:
 *   - values along z live in a per-thread register ring of seven planes (the thread's own column),
 *   - values of the current plane in a shared-memory tile with a halo of three in x and in y (no corners: neither the
 *     derivatives nor the reconstructions have cross stencils),
 *   - the global loads of the NEXT plane are cp.async copies issued before the current plane is processed, so a block
 *     always has one plane of loads in flight and ONE barrier per plane (double-buffered tiles).
 * Same operations in the same order as the thread functions of hb2_diffusive.cuh (the unit is built with -fmad=false):
 * bit-identical to the grid-stride kernels and to the oracle.
 *
 * Reference behaviour: DiffusiveFluxReconstructorNodeSixthOrder.cpp:65-939, DiffusiveFluxReconstructorNode.cpp:888-1055,
 * NavierStokes.cpp:2085-2092 (see hb2_diffusive.cuh).
 */
#pragma once
#include "hb2_diffusive.cuh"

namespace hb2 {
namespace march {

constexpr int TX = 32, TY = 8, NT = TX * TY, HALO = 3;
constexpr int SX = TX + 2 * HALO, SY = TY + 2 * HALO;          /* 38 x 14 tile (corners unused) */
constexpr int NHX = 2 * HALO * TY;                             /* 48 x-halo cells per plane: threads 0..47 */
constexpr int NHY = 2 * HALO * TX;                             /* 192 y-halo cells per plane: threads 48..239 */
static_assert(NHX + NHY <= NT, "one halo cell per thread");

__device__ __forceinline__ void cp_async8(unsigned dst_shared, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
/* all but the most recent group of the thread's copies have landed */
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

/* planes of loads a block keeps in flight (HB2_DIFF_DEPTH = 1: the next plane only, staged behind the barrier of the current
 * one; 2: two planes, one commit group each, three tile sets).  Measured on B200, Navier-Stokes 512^3: 91.06 ms per step with
 * one plane, 91.10 with two (diffusive calls alone 28.2 / 29.0 ms; profiles/r02_ap_ns_depth.txt) -- the kernels are not
 * waiting for a plane that was requested too late, so the default stays 1 (less shared memory). */
#ifndef HB2_DIFF_DEPTH
#define HB2_DIFF_DEPTH 1
#endif
constexpr int DEPTH = HB2_DIFF_DEPTH;
static_assert(DEPTH == 1 || DEPTH == 2, "one or two planes in flight");

/* the halo cell of a thread: position (hx, hy) in the 38 x 14 tile; returns false for threads without one */
__device__ __forceinline__ bool halo_of_thread(int t, int& hx, int& hy)
{
    if (t < NHX) {                       /* x halo: rows 3..10, columns 0,1,2 and 35,36,37 */
        const int c = t % (2 * HALO);
        hy = HALO + t / (2 * HALO);
        hx = (c < HALO) ? c : TX + c;
        return true;
    }
    const int u = t - NHX;
    if (u < NHY) {                       /* y halo: rows 0,1,2 and 11,12,13, columns 3..34 */
        const int r = u / TX;
        hx = HALO + u % TX;
        hy = (r < HALO) ? r : TY + r;
        return true;
    }
    hx = hy = 0;
    return false;
}

/* ---- node stage: conservative variables in (the primitives never reach HBM), the twelve node-flux arrays out.
 * Algorithmic traffic 40 B read (+ halo re-reads served by L2) + 96 B written per node.
 * Tile origin: ghost-box cell (32 bx, 8 by + 3), so that rows start on 256-byte boundaries of the arrays' rows; the nodes are the
 * cells of the interior extended by three (node coordinate = ghost-box index - 6). */
constexpr int NODE_SMEM_DOUBLES = 2 * 4 * SY * SX + DEPTH * (5 * NT + 5 * NT);

template <int MATH>
__global__ void __launch_bounds__(NT, 2) k_diff_node_march(const __grid_constant__ DiffGeom G, const __grid_constant__ DiffConsts K,
                                                           const __grid_constant__ DiffFast FK, const __grid_constant__ DiffPtrs Q,
                                                           const __grid_constant__ DiffAllPtrs A, int seg_len)
{
    auto prims = [&](const double (&q)[5], double (&P)[4]) {
        if constexpr (MATH == 0)
            diff_primitives<3>(q, K, P);
        else
            diff_primitives_fast(q, FK, P);
    };
    extern __shared__ double smem[];
    double* sP = smem;                                  /* [2][4][SY][SX] primitives of the node plane */
    double* sQo = smem + 2 * 4 * SY * SX;               /* [DEPTH][5][NT] staged own cell, plane kz + 3 */
    double* sQh = sQo + DEPTH * 5 * NT;                 /* [DEPTH][5][NT] staged halo cell, plane kz */
    const int t = (int)threadIdx.x, tx = t & 31, ty = t >> 5;
    const int i = (int)blockIdx.x * TX - 6 + tx, j = (int)blockIdx.y * TY - 3 + ty;      /* node coordinates */
    const bool cell = i < G.n[0] + 6 && j < G.n[1] + 6;                                  /* the cell exists (i >= -6, j >= -3) */
    const bool own = cell && i >= -3 && i < G.n[0] + 3 && j < G.n[1] + 3;                /* the node exists */
    const int kz_lo = -3 + (int)blockIdx.z * seg_len;
    const int kz_hi = (kz_lo + seg_len < G.n[2] + 3) ? kz_lo + seg_len : G.n[2] + 3;
    const long long col = (i + G.g[0]) + G.cs[1] * (j + G.g[1]);
    int hx, hy;
    bool hok = halo_of_thread(t, hx, hy);
    const int hi = i - tx - HALO + hx, hj = j - ty - HALO + hy;                          /* cell coordinates of the halo cell */
    hok = hok && hi >= -6 && hi < G.n[0] + 6 && hj < G.n[1] + 6;
    const long long hcol = (hi + G.g[0]) + G.cs[1] * (hj + G.g[1]);
    const unsigned so = (unsigned)__cvta_generic_to_shared(sQo + t), sh = (unsigned)__cvta_generic_to_shared(sQh + t);

    /* register ring: ring[v][m] = primitive v of the own column at plane kz - 3 + m */
    double ring[4][7];
#pragma unroll
    for (int v = 0; v < 4; v++)
#pragma unroll
        for (int m = 0; m < 7; m++) ring[v][m] = 0.0;
    if (cell) {
#pragma unroll
        for (int m = 1; m < 7; m++) {                   /* planes kz_lo - 3 .. kz_lo + 2: after the first shift they sit at m - 1 */
            double q[5], P[4];
            const long long x = col + G.cs[2] * (kz_lo - 3 + (m - 1) + G.g[2]);
#pragma unroll
            for (int c = 0; c < 5; c++) q[c] = Q.Q[c][x];
            prims(q, P);
#pragma unroll
            for (int v = 0; v < 4; v++) ring[v][m] = P[v];
        }
    }
    /* the copies of plane kz go to staging slot sb; one commit group per plane (empty behind the last plane) */
    auto stage = [&](int kz, int sb) {
        if (kz < kz_hi) {
            const unsigned off = (unsigned)(sb * 5 * NT * sizeof(double));
            if (cell) {
                const long long x = col + G.cs[2] * (kz + 3 + G.g[2]);
#pragma unroll
                for (int c = 0; c < 5; c++) cp_async8(so + off + (unsigned)(c * NT * sizeof(double)), Q.Q[c] + x);
            }
            if (hok) {
                const long long x = hcol + G.cs[2] * (kz + G.g[2]);
#pragma unroll
                for (int c = 0; c < 5; c++) cp_async8(sh + off + (unsigned)(c * NT * sizeof(double)), Q.Q[c] + x);
            }
        }
        cp_async_commit();
    };
    stage(kz_lo, 0);
    if (DEPTH == 2) stage(kz_lo + 1, 1);
    int b = 0;
    for (int kz = kz_lo; kz < kz_hi; kz++, b ^= 1) {
        double* P_b = sP + b * (4 * SY * SX);
        const int sb = (DEPTH == 2) ? b : 0;            /* staging slot of plane kz */
#pragma unroll
        for (int v = 0; v < 4; v++)
#pragma unroll
            for (int m = 0; m < 6; m++) ring[v][m] = ring[v][m + 1];
        /* the thread's own copies: no barrier needed to read them back */
        if (DEPTH == 2)
            cp_async_wait_but_one();
        else
            cp_async_wait_all();
        if (cell) {
            double q[5], P[4];
#pragma unroll
            for (int c = 0; c < 5; c++) q[c] = sQo[(sb * 5 + c) * NT + t];
            prims(q, P);
#pragma unroll
            for (int v = 0; v < 4; v++) {
                ring[v][6] = P[v];
                P_b[(v * SY + ty + HALO) * SX + tx + HALO] = ring[v][3];
            }
        }
        if (hok) {
            double q[5], P[4];
#pragma unroll
            for (int c = 0; c < 5; c++) q[c] = sQh[(sb * 5 + c) * NT + t];
            prims(q, P);
#pragma unroll
            for (int v = 0; v < 4; v++) P_b[(v * SY + hy) * SX + hx] = P[v];
        }
        stage(kz + DEPTH, sb);                          /* the slot just read back; in flight while DEPTH planes are processed */
        __syncthreads();
        if (own) {
            double der[4][3];
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const double* c = P_b + (v * SY + ty + HALO) * SX + tx + HALO;
                if constexpr (MATH == 0) {
                    der[v][0] = diff_first_derivative6(c[-3], c[-2], c[-1], c[1], c[2], c[3], G.dx_inv[0]);
                    der[v][1] = diff_first_derivative6(c[-3 * SX], c[-2 * SX], c[-SX], c[SX], c[2 * SX], c[3 * SX], G.dx_inv[1]);
                    der[v][2] = diff_first_derivative6(ring[v][0], ring[v][1], ring[v][2], ring[v][4], ring[v][5], ring[v][6], G.dx_inv[2]);
                } else {
                    der[v][0] = diff_first_derivative6_fast(c[-3], c[-2], c[-1], c[1], c[2], c[3], FK.cd[0]);
                    der[v][1] = diff_first_derivative6_fast(c[-3 * SX], c[-2 * SX], c[-SX], c[SX], c[2 * SX], c[3 * SX], FK.cd[1]);
                    der[v][2] = diff_first_derivative6_fast(ring[v][0], ring[v][1], ring[v][2], ring[v][4], ring[v][5], ring[v][6], FK.cd[2]);
                }
            }
            const double vel[3] = {ring[0][3], ring[1][3], ring[2][3]};
            double Fn[3][5];
            if constexpr (MATH == 0)
                diff_node_flux_from_derivatives<3>(K, vel, der, Fn);
            else
                diff_node_flux_fast(FK, vel, der, Fn);
            const long long x = col + G.cs[2] * (kz + G.g[2]);
#pragma unroll
            for (int f = 0; f < 3; f++)
#pragma unroll
                for (int e = 1; e < 5; e++) {
                    /* re-associated route: the viscous stress is symmetric BIT FOR BIT (D2 (d_a u_f + d_f u_a), an IEEE sum
                     * commutes), so F^f of momentum a < f is the array F^a of momentum f -- the host aliases the pointers
                     * (diff_alias_symmetric), nine arrays are written and read instead of twelve */
                    if (MATH == 1 && e <= 3 && e - 1 < f) continue;
                    A.Fn[f][e][x] = Fn[f][e];
                }
        }
        /* the tile written next is the other one; this one is rewritten two iterations on, behind the next barrier */
    }
}

/* ---- flux-free update: U += beta (-div F_d) from the node fluxes of the three directions; F^z of the own column along z in
 * a register ring, F^x / F^y of the current plane in shared-memory tiles with a halo of three along their own direction.
 * The high z face of a cell is the low z face of the next one: reconstructed once and carried.  Algorithmic traffic 96 B of
 * node fluxes + 64 B read-modify-write of the state per cell. */
constexpr int DIV_X = 4 * TY * SX, DIV_Y = 4 * SY * TX;       /* doubles per tile set */
constexpr int DIV_NBUF = DEPTH + 1;                           /* tile sets: the one being read + DEPTH in flight */
constexpr int DIV_SMEM_DOUBLES = DIV_NBUF * (DIV_X + DIV_Y);

template <int MATH>
__global__ void __launch_bounds__(NT, 2) k_diff_div_march(const __grid_constant__ NsDivArgs A, const __grid_constant__ DiffFast FK, int seg_len)
{
    extern __shared__ double smem[];
    const DiffGeom &G = A.G6, &GU = A.GU;
    const int t = (int)threadIdx.x, tx = t & 31, ty = t >> 5;
    const int i = (int)blockIdx.x * TX + tx, j = (int)blockIdx.y * TY + ty;
    const bool own = i < G.n[0] && j < G.n[1];
    const bool cx_ok = i < G.n[0] + 3 && j < G.n[1];             /* nodes beside the last cells are somebody's neighbours */
    const bool cy_ok = i < G.n[0] && j < G.n[1] + 3;
    const int k0 = (int)blockIdx.z * seg_len;
    const int k1 = (k0 + seg_len < G.n[2]) ? k0 + seg_len : G.n[2];
    const long long col = (i + G.g[0]) + G.cs[1] * (j + G.g[1]);
    const long long colU = (i + GU.g[0]) + GU.cs[1] * (j + GU.g[1]);
    /* the thread's halo node: x halo of F^x (threads 0..47) or y halo of F^y (threads 48..239) */
    int hx, hy;
    bool hok = halo_of_thread(t, hx, hy);
    const bool is_x = t < NHX;
    const int hi = i - tx - HALO + hx, hj = j - ty - HALO + hy;
    hok = hok && hi >= -3 && hj >= -3 && (is_x ? (hi < G.n[0] + 3 && hj < G.n[1]) : (hi < G.n[0] && hj < G.n[1] + 3));
    const long long hcol = (hi + G.g[0]) + G.cs[1] * (hj + G.g[1]);
    const int h_off = is_x ? (hy - HALO) * SX + hx : DIV_X + hy * TX + (hx - HALO);      /* inside one buffer, equation 0 */
    const int h_es = is_x ? TY * SX : SY * TX;
    const double* const* h_src = A.Fn[is_x ? 0 : 1];
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem);
    const int ox = ty * SX + tx + HALO, oy = DIV_X + (ty + HALO) * TX + tx;             /* own slots, equation 0 */

    /* one commit group per plane (empty behind the last plane) */
    auto stage = [&](int k, int b) {
        if (k < k1) {
            const long long zoff = G.cs[2] * (k + G.g[2]);
            const unsigned base = s0 + (unsigned)(b * (DIV_X + DIV_Y) * sizeof(double));
            if (cx_ok) {
#pragma unroll
                for (int e = 0; e < 4; e++) cp_async8(base + (unsigned)((ox + e * TY * SX) * sizeof(double)), A.Fn[0][e + 1] + col + zoff);
            }
            if (cy_ok) {
#pragma unroll
                for (int e = 0; e < 4; e++) cp_async8(base + (unsigned)((oy + e * SY * TX) * sizeof(double)), A.Fn[1][e + 1] + col + zoff);
            }
            if (hok) {
#pragma unroll
                for (int e = 0; e < 4; e++) cp_async8(base + (unsigned)((h_off + e * h_es) * sizeof(double)), h_src[e + 1] + hcol + zoff);
            }
        }
        cp_async_commit();
    };
    if (k0 >= k1) return;
    stage(k0, 0);
    if (DEPTH == 2) stage(k0 + 1, 1);
    double ring[4][7];                       /* F^z of equation e + 1 at planes k - 3 + m */
    double fz_next[4], fzf[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
#pragma unroll
        for (int m = 0; m < 7; m++) ring[e][m] = 0.0;
        fz_next[e] = 0.0;
        fzf[e] = 0.0;
    }
    if (own) {
#pragma unroll
        for (int m = 1; m < 7; m++)
#pragma unroll
            for (int e = 0; e < 4; e++) ring[e][m] = A.Fn[2][e + 1][col + G.cs[2] * (k0 - 3 + (m - 1) + G.g[2])];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            fz_next[e] = A.Fn[2][e + 1][col + G.cs[2] * (k0 + 3 + G.g[2])];
            /* low z face of the first cell: nodes k0 - 3 .. k0 + 2 */
            if constexpr (MATH == 0) fzf[e] = diff_reconstruct6(ring[e][1], ring[e][2], ring[e][3], ring[e][4], ring[e][5], ring[e][6], A.dt);
        }
    }
    /* the state of the cell is fetched one plane ahead as well (only its own thread ever touches it) */
    double un[4] = {0.0, 0.0, 0.0, 0.0};
    if (own) {
#pragma unroll
        for (int e = 0; e < 4; e++) un[e] = A.U[e + 1][colU + GU.cs[2] * (k0 + GU.g[2])];
    }
    int b = 0;
    for (int k = k0; k < k1; k++, b = (b + 1 == DIV_NBUF) ? 0 : b + 1) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
#pragma unroll
            for (int m = 0; m < 6; m++) ring[e][m] = ring[e][m + 1];
            ring[e][6] = fz_next[e];
        }
        double uo[4];
#pragma unroll
        for (int e = 0; e < 4; e++) uo[e] = un[e];
        if (DEPTH == 2)
            cp_async_wait_but_one();
        else
            cp_async_wait_all();
        __syncthreads();                     /* plane k has landed in buffer b; every thread is done with the buffer of plane k - 1 */
        stage(k + DEPTH, (b + DEPTH) % DIV_NBUF);   /* = the buffer of plane k - 1 */
        const long long xu = colU + GU.cs[2] * (k + GU.g[2]);
        if (own) {
            if (k + 1 < k1) {
#pragma unroll
                for (int e = 0; e < 4; e++) fz_next[e] = A.Fn[2][e + 1][col + G.cs[2] * (k + 4 + G.g[2])];
#pragma unroll
                for (int e = 0; e < 4; e++) un[e] = A.U[e + 1][xu + GU.cs[2]];
            }
            const double* buf = smem + b * (DIV_X + DIV_Y);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const double* cx = buf + ox + e * TY * SX;
                const double* cy = buf + oy + e * SY * TX;
                if constexpr (MATH == 1) {
                    double acc = uo[e];
                    acc = diff_divergence_fast(acc, cx[-3], cx[-2], cx[-1], cx[1], cx[2], cx[3], FK.kd[0]);
                    acc = diff_divergence_fast(acc, cy[-3 * TX], cy[-2 * TX], cy[-TX], cy[TX], cy[2 * TX], cy[3 * TX], FK.kd[1]);
                    acc = diff_divergence_fast(acc, ring[e][0], ring[e][1], ring[e][2], ring[e][4], ring[e][5], ring[e][6], FK.kd[2]);
                    A.U[e + 1][xu] = acc;
                    continue;
                }
                const double FxL = diff_reconstruct6(cx[-3], cx[-2], cx[-1], cx[0], cx[1], cx[2], A.dt);
                const double FxR = diff_reconstruct6(cx[-2], cx[-1], cx[0], cx[1], cx[2], cx[3], A.dt);
                const double FyB = diff_reconstruct6(cy[-3 * TX], cy[-2 * TX], cy[-TX], cy[0], cy[TX], cy[2 * TX], A.dt);
                const double FyT = diff_reconstruct6(cy[-2 * TX], cy[-TX], cy[0], cy[TX], cy[2 * TX], cy[3 * TX], A.dt);
                double div = -(FxR - FxL) / G.dx[0] - (FyT - FyB) / G.dx[1];
                const double FzB = fzf[e];
                const double FzF = diff_reconstruct6(ring[e][1], ring[e][2], ring[e][3], ring[e][4], ring[e][5], ring[e][6], A.dt);
                fzf[e] = FzF;
                div -= (FzF - FzB) / G.dx[2];
                A.U[e + 1][xu] = uo[e] + A.beta * div;
            }
        }
    }
}

/* planes per segment of a march over `planes` planes by `tiles` tiles: about four waves of two resident blocks per SM, at
 * least 32 planes each (every segment re-reads six planes of its own column) */
inline int march_seg_len(long long tiles, int planes, int sm_count)
{
    long long nseg = ((long long)sm_count * 2 * 4 + tiles - 1) / tiles;
    const long long maxseg = planes / 32 > 0 ? planes / 32 : 1;
    if (nseg > maxseg) nseg = maxseg;
    if (nseg < 1) nseg = 1;
    return (int)((planes + nseg - 1) / nseg);
}

}  // namespace march
}  // namespace hb2
