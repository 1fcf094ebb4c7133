/*
 * hb2_sweeps.cu -- sm_100a kernels of the WCNS5-JS / HLLC-HLL path.
 *
 * Compiled TWICE into the product library:
 *   -DHB2_MATH=0 -fmad=false : reference operation order, bit-identical to the oracle
 *   -DHB2_MATH=1 -fmad=true  : FP64-instruction-minimal re-association (hb2_fast.cuh), <= 1e-12 relative
 * The arithmetic lives in hb2_core.cuh / hb2_fast.cuh, the thread mapping in hb2_sweep.cuh; this file holds the
 * __global__ entry points and their launchers.
 *
 * Kernels
 *   k_sensor  : thread per cell of the (N+4)^d box: dilatation + vorticity magnitude.
 *   k_flags   : thread per cell of the (N+3)^d box: the s > 0.65 decision of the cell's three low faces (one byte).
 *   k_sweep   : one direction: a 256-thread block marches along the sweep axis, one barrier per iteration
 *               (hb2_sweep.cuh).
 *   k_advance : RK update from materialised side fluxes (API-preserving mode).
 */
#include "hb2_ops.h"
#include "hb2_sweep.cuh"

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
    const double hidx[3] = {0.5 / G.dx[0], 0.5 / G.dx[1], 0.5 / G.dx[2]};
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(id % e0) - 2;
        const long long r = id / e0;
        const int j = (int)(r % e1) - 2;
        const int k = (Tr::DIM == 3) ? (int)(r / e1) - 2 : 0;
        const long long x = cidx(G, i, j, k);
        double th, Om;
        if (MATH == 0)
            sensor_cell<Tr>(G, Q, x, th, Om);
        else
            sensor_cell_fast<Tr>(G, Q, x, hidx, th, Om);
        theta[x] = th;
        Omega[x] = Om;
    }
}

/* bit d of hyb[cell] = shock-sensor decision of the face between cells (cell - e_d) and cell, for cells -1..N+1 */
template <int DIM>
__global__ void __launch_bounds__(256) k_flags(const __grid_constant__ Geom G, const double* __restrict__ theta,
                                               const double* __restrict__ Omega, unsigned char* __restrict__ hyb)
{
    const int e0 = G.n[0] + 3, e1 = G.n[1] + 3, e2 = (DIM == 3) ? G.n[2] + 3 : 1;
    const long long total = (long long)e0 * e1 * e2;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(id % e0) - 1;
        const long long r = id / e0;
        const int j = (int)(r % e1) - 1;
        const int k = (DIM == 3) ? (int)(r / e1) - 1 : 0;
        const long long x = cidx(G, i, j, k);
        const double th = theta[x], Om = Omega[x];
        unsigned char f = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const long long xl = x - G.cs[d];
            if (face_sensor(theta[xl], th, Omega[xl], Om)) f |= (unsigned char)(1u << d);
        }
        hyb[x] = f;
    }
}

#ifndef HB2_MINB
#define HB2_MINB 2
#endif

template <class Tr, int DIR, int NTERM>
__global__ void __launch_bounds__(256, (MATH == 1 && Tr::MODEL == SS) ? HB2_MINB : 1) k_sweep(const __grid_constant__ DirArgs A)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    extern __shared__ double smem[];
    const BlockId b = {(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z};
    const PencilCtx c = pencil_ctx<Tr, DIR, MATH>(A, b, (int)threadIdx.x);
    const int nsteps = Sh::nsteps(c.c1 - c.c0);
    PipeRegs<Tr> pr;
    pipeline_prologue<Tr, DIR, MATH>(A, smem, c, pr);
    __syncthreads();
    for (int t = 0; t <= nsteps; t++) {
        pipeline_iteration<Tr, DIR, MATH, NTERM>(A, smem, c, t, nsteps, pr);
        __syncthreads();
    }
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
int launch_sensor_t(const Geom& G, const QTab& Qtab_dev, double* theta, double* Omega, unsigned char* hyb, cudaStream_t st)
{
    const long long total = (long long)(G.n[0] + 4) * (G.n[1] + 4) * (Tr::DIM == 3 ? G.n[2] + 4 : 1);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    k_sensor<Tr><<<(unsigned)blocks, 256, 0, st>>>(G, Qtab_dev, theta, Omega);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    k_flags<Tr::DIM><<<(unsigned)blocks, 256, 0, st>>>(G, theta, Omega, hyb);
    return (int)cudaGetLastError();
}

template <class Tr, int DIR, int NTERM>
int launch_dir_n(const DirArgs& A, cudaStream_t st)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const Geom& G = A.G;
    const size_t smem = (size_t)Sh::SMEM_DOUBLES * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_sweep<Tr, DIR, NTERM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_sweep<Tr, DIR, NTERM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int nseg = (G.n[DIR] + A.seg_len - 1) / A.seg_len;
    dim3 grid(1, 1, nseg);
    if (DIR == 0) {
        grid.x = (G.n[1] + Sh::P - 1) / Sh::P;
        grid.y = (Tr::DIM == 3) ? G.n[2] : 1;
    } else if (DIR == 1) {
        grid.x = (G.n[0] + 31) / 32;
        grid.y = (Tr::DIM == 3) ? G.n[2] : 1;
    } else {
        grid.x = (G.n[0] + 31) / 32;
        grid.y = G.n[1];
    }
    k_sweep<Tr, DIR, NTERM><<<grid, Sh::NT, smem, st>>>(A);
    return (int)cudaGetLastError();
}

template <class Tr, int DIR>
int launch_dir(const DirArgs& A, cudaStream_t st)
{
    /* the RK linear combination exists only in the last direction of a fused stage */
    if (DIR == Tr::DIM - 1 && A.mode == MODE_FUSED) {
        if (A.nterm == 1) return launch_dir_n<Tr, DIR, 1>(A, st);
        if (A.nterm == 2) return launch_dir_n<Tr, DIR, 2>(A, st);
        if (A.nterm == 3) return launch_dir_n<Tr, DIR, 3>(A, st);
        return (int)cudaErrorInvalidValue;
    }
    return launch_dir_n<Tr, DIR, 0>(A, st);
}

template <class Tr>
int launch_sweep_t(const LaunchCfg&, int dir, const DirArgs& A, cudaStream_t st)
{
    if (dir == 0) return launch_dir<Tr, 0>(A, st);
    if (dir == 1) return launch_dir<Tr, 1>(A, st);
    return launch_dir<Tr, (Tr::DIM == 3 ? 2 : 1)>(A, st);
}

#define HB2_DISPATCH(cfg, CALL)                                                       \
    do {                                                                              \
        if ((cfg).model == SS && (cfg).dim == 2) { using Tr = Traits<SS, 2, 1>; CALL; } \
        if ((cfg).model == SS && (cfg).dim == 3) { using Tr = Traits<SS, 3, 1>; CALL; } \
        if ((cfg).model == FE && (cfg).dim == 2 && (cfg).ns == 2) { using Tr = Traits<FE, 2, 2>; CALL; } \
        if ((cfg).model == FE && (cfg).dim == 3 && (cfg).ns == 2) { using Tr = Traits<FE, 3, 2>; CALL; } \
    } while (0)

int op_sensor(const LaunchCfg& cfg, const Geom& G, const QTab& Qtab_dev, double* theta, double* Omega, unsigned char* hyb,
              cudaStream_t st)
{
    HB2_DISPATCH(cfg, return launch_sensor_t<Tr>(G, Qtab_dev, theta, Omega, hyb, st));
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
