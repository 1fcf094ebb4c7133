/*
 * hb2_sweeps.cu -- sm_100a kernels of the WCNS5-JS / HLLC-HLL path.
 *
 * Compiled TWICE into the product library:
 *   -DHB2_MATH=0 -fmad=false : reference operation order, bit-identical to the oracle
 *   -DHB2_MATH=1 -fmad=true  : FMA contraction + one-division WENO weights (<= 1e-12 relative)
 * The arithmetic lives in hb2_core.cuh; this file only maps threads onto it.
 *
 * Kernel shapes
 *   k_sensor  : thread per cell of the (N+4)^d box, dilatation + vorticity magnitude.
 *   k_xsweep  : block = BX consecutive linear positions of one k-plane staged in shared memory
 *               (primitive variables + node fluxes, converted once per cell), thread per midpoint
 *               flux, midpoint/face fluxes exchanged through shared memory.
 *   k_march   : y / z sweeps; lanes along x (coalesced 8-byte loads, 256 B per warp row), each thread
 *               marches a pencil segment with a register-rotating 6-cell stencil.
 *   k_advance : RK update from materialised side fluxes (API-preserving mode).
 */
#include "hb2_ops.h"

#ifndef HB2_MATH
#error "compile with -DHB2_MATH=0 (exact) or -DHB2_MATH=1 (fast)"
#endif

namespace hb2 {
namespace {

constexpr int MATH = HB2_MATH;

template <class Tr>
__global__ void __launch_bounds__(256) k_sensor(const __grid_constant__ Geom G, const __grid_constant__ QTab Qtab,
                                                double* __restrict__ theta, double* __restrict__ Omega)
{
    const double* const* Q = Qtab.p;
    const int e0 = G.n[0] + 4, e1 = G.n[1] + 4, e2 = (Tr::DIM == 3) ? G.n[2] + 4 : 1;
    const long long total = (long long)e0 * e1 * e2;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(id % e0) - 2;
        const int j = (int)((id / e0) % e1) - 2;
        const int k = (Tr::DIM == 3) ? (int)(id / ((long long)e0 * e1)) - 2 : 0;
        const long long x = cidx(G, i, j, k);
        double th, Om;
        sensor_cell<Tr>(G, Q, x, th, Om);
        theta[x] = th;
        Omega[x] = Om;
    }
}

template <class Tr>
__global__ void __launch_bounds__(128) k_xsweep(const __grid_constant__ DirArgs A, int BX, long long tiles_per_plane)
{
    extern __shared__ double smem[];
    XSmem<Tr> sm(smem, BX);
    const long long tile = blockIdx.x % tiles_per_plane;
    const int k = (int)(blockIdx.x / tiles_per_plane);
    const long long pstart = cidx(A.G, -A.G.g[0], 0, k);
    const long long p0 = pstart - 1 + tile * (long long)(BX - 3);
    const int t = threadIdx.x;
    for (int idx = t; idx < BX + 5; idx += blockDim.x) xsweep_phase_load<Tr, MATH>(A, sm, p0, idx);
    __syncthreads();
    xsweep_phase_mid<Tr, MATH>(A, sm, p0, k, t);
    __syncthreads();
    xsweep_phase_face<Tr, MATH>(A, sm, p0, k, t);
    __syncthreads();
    xsweep_phase_cell<Tr, MATH>(A, sm, p0, k, t);
}

template <class Tr, int DIR>
__global__ void __launch_bounds__(128) k_march(const __grid_constant__ DirArgs A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.G.n[0]) return;
    march_pencil<Tr, DIR, MATH>(A, i, (int)blockIdx.y, (int)blockIdx.z);
}

__global__ void __launch_bounds__(256) k_advance(const __grid_constant__ AdvanceArgs P)
{
    const Geom& G = P.G;
    const long long ncell = (long long)G.n[0] * G.n[1] * G.n[2];
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < ncell;
         id += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(id % G.n[0]);
        const int j = (int)((id / G.n[0]) % G.n[1]);
        const int k = (int)(id / ((long long)G.n[0] * G.n[1]));
        const long long x = cidx(G, i, j, k);
        const long long fxL = i + (long long)(G.n[0] + 1) * (j + (long long)G.n[1] * k);
        const long long fyB = i + (long long)G.n[0] * (j + (long long)(G.n[1] + 1) * k);
        const long long fzB = id;
        double Unew[HB2_MAXE];
        for (int e = 0; e < P.neq; e++) {
            double u = 0.0;
            for (int n = 0; n < P.ncoef; n++) {
                if (P.alpha[n] != 0.0) u += P.alpha[n] * P.Uint[n][e][x];
                if (P.beta[n] != 0.0) {
                    const double* Fx = P.Fint[n][0 * P.neq + e];
                    const double* Fy = P.Fint[n][1 * P.neq + e];
                    const double* Sn = P.Sint[n][e];
                    const double s = Sn ? Sn[id] : 0.0;
                    if (G.dim == 2) {
                        u += P.beta[n] * (-(Fx[fxL + 1] - Fx[fxL]) / G.dx[0] - (Fy[fyB + G.n[0]] - Fy[fyB]) / G.dx[1] + s);
                    } else {
                        const double* Fz = P.Fint[n][2 * P.neq + e];
                        u += P.beta[n] * (-(Fx[fxL + 1] - Fx[fxL]) / G.dx[0] - (Fy[fyB + G.n[0]] - Fy[fyB]) / G.dx[1] -
                                          (Fz[fzB + (long long)G.n[0] * G.n[1]] - Fz[fzB]) / G.dx[2] + s);
                    }
                }
            }
            Unew[e] = u;
            P.Uout[e][x] = u;
        }
        if (P.model == FE) {
            double zl = 1.0;
            for (int si = 0; si < P.ns - 1; si++) zl -= Unew[P.ns + G.dim + 1 + si];
            P.Uout[P.neq][x] = zl;
        }
        /* gamma-weighted source accumulation (cell part; side part below by a second grid-stride loop) */
        for (int e = 0; e < P.neq; e++) {
            if (!P.Sacc[e]) continue;
            double s = P.Sacc[e][id];
            for (int n = 0; n < P.ncoef; n++)
                if (P.gamma[n] != 0.0 && P.Sint[n][e]) s += P.gamma[n] * P.Sint[n][e][id];
            P.Sacc[e][id] = s;
        }
    }
    /* gamma-weighted flux accumulation, Euler.cpp:1555-1640 */
    for (int d = 0; d < G.dim; d++) {
        long long ext[3] = {G.n[0], G.n[1], G.n[2]};
        ext[d] += 1;
        const long long nside = ext[0] * ext[1] * ext[2];
        for (int e = 0; e < P.neq; e++) {
            double* acc = P.Facc[d * P.neq + e];
            if (!acc) continue;
            for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < nside;
                 id += (long long)gridDim.x * blockDim.x) {
                double f = acc[id];
                for (int n = 0; n < P.ncoef; n++)
                    if (P.gamma[n] != 0.0 && P.Fint[n][d * P.neq + e]) f += P.gamma[n] * P.Fint[n][d * P.neq + e][id];
                acc[id] = f;
            }
        }
    }
}

/* ---- host-side launchers ------------------------------------------------------------- */

template <class Tr>
int launch_sensor_t(const Geom& G, const QTab& Qtab_dev, double* theta, double* Omega, cudaStream_t st)
{
    const long long total = (long long)(G.n[0] + 4) * (G.n[1] + 4) * (Tr::DIM == 3 ? G.n[2] + 4 : 1);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    k_sensor<Tr><<<(unsigned)blocks, 256, 0, st>>>(G, Qtab_dev, theta, Omega);
    return (int)cudaGetLastError();
}

template <class Tr>
int launch_sweep_t(const LaunchCfg& cfg, int dir, const DirArgs& A, cudaStream_t st)
{
    const Geom& G = A.G;
    if (dir == 0) {
        const int BX = cfg.bx;
        const long long run = (long long)G.n[1] * G.gd[0];
        const long long tiles = (run + (BX - 3) - 1) / (BX - 3);
        const long long planes = (Tr::DIM == 3) ? G.n[2] : 1;
        const size_t smem = (size_t)XSmem<Tr>::doubles(BX) * sizeof(double);
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(k_xsweep<Tr>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_set = true;
        }
        k_xsweep<Tr><<<(unsigned)(tiles * planes), BX, smem, st>>>(A, BX, tiles);
        return (int)cudaGetLastError();
    }
    const int nthr = cfg.march_block;
    const int N = G.n[dir];
    const int nseg = (N + A.seg_len - 1) / A.seg_len;
    dim3 grid((G.n[0] + nthr - 1) / nthr, 1, nseg);
    if (dir == 1) {
        grid.y = (Tr::DIM == 3) ? G.n[2] : 1;
        k_march<Tr, 1><<<grid, nthr, 0, st>>>(A);
    } else {
        grid.y = G.n[1];
        k_march<Tr, (Tr::DIM == 3 ? 2 : 1)><<<grid, nthr, 0, st>>>(A);
    }
    return (int)cudaGetLastError();
}

#define HB2_DISPATCH(cfg, CALL)                                                       \
    do {                                                                              \
        if ((cfg).model == SS && (cfg).dim == 2) { using Tr = Traits<SS, 2, 1>; CALL; } \
        if ((cfg).model == SS && (cfg).dim == 3) { using Tr = Traits<SS, 3, 1>; CALL; } \
        if ((cfg).model == FE && (cfg).dim == 2 && (cfg).ns == 2) { using Tr = Traits<FE, 2, 2>; CALL; } \
        if ((cfg).model == FE && (cfg).dim == 3 && (cfg).ns == 2) { using Tr = Traits<FE, 3, 2>; CALL; } \
    } while (0)

int op_sensor(const LaunchCfg& cfg, const Geom& G, const QTab& Qtab_dev, double* theta, double* Omega, cudaStream_t st)
{
    HB2_DISPATCH(cfg, return launch_sensor_t<Tr>(G, Qtab_dev, theta, Omega, st));
    return -1;
}

int op_sweep(const LaunchCfg& cfg, int dir, const DirArgs& A, cudaStream_t st)
{
    HB2_DISPATCH(cfg, return launch_sweep_t<Tr>(cfg, dir, A, st));
    return -1;
}

int op_advance(const AdvanceArgs& P, cudaStream_t st)
{
    const long long ncell = (long long)P.G.n[0] * P.G.n[1] * P.G.n[2];
    long long blocks = (ncell + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    k_advance<<<(unsigned)blocks, 256, 0, st>>>(P);
    return (int)cudaGetLastError();
}

const Ops g_ops = {op_sensor, op_sweep, op_advance};

}  // namespace

#if HB2_MATH == 0
const Ops* ops_exact() { return &g_ops; }
#else
const Ops* ops_fast() { return &g_ops; }
#endif

}  // namespace hb2
