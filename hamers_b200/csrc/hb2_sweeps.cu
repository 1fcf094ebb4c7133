/*
 * hb2_sweeps.cu -- sm_100a kernels of the WCNS5-JS / HLLC-HLL path.
 *
 * Compiled SIX times into the product library:
 *   -DHB2_MATH=0 -fmad=false : reference operation order, bit-identical to the oracle (WCNS5-JS), and again with
 *                              -DHB2_SCHEME=1 (WCNS5-Z) and -DHB2_SCHEME=2 (WCNS6-LD), SURVEY row f2
 *   -DHB2_MATH=1 -fmad=true  : FP64-instruction-minimal re-association (hb2_fast.cuh), <= 1e-12 relative, WCNS5-JS and
 *                              (-DHB2_SCHEME=1, 2) WCNS5-Z, WCNS6-LD
 * The arithmetic lives in hb2_core.cuh / hb2_fast.cuh, the thread mapping in hb2_sweep.cuh; this file holds the
 * __global__ entry points and their launchers.
 *
 * Kernels
 *   k_sensor  : the s > 0.65 decision of every cell's low faces (one byte per cell) in one pass: a block marches
 *               along z with velocity and theta/Omega planes in shared-memory rings (hb2_sensor.cuh).
 *   k_sweep   : one direction: a 256-thread block marches along the sweep axis, one barrier per iteration
 *               (hb2_sweep.cuh).
 *   k_advance : RK update from materialised side fluxes (API-preserving mode).
 */
#include "hb2_ops.h"
#include "hb2_sweep.cuh"
#include "hb2_sensor.cuh"

#ifndef HB2_MATH
#error "compile with -DHB2_MATH=0 (exact) or -DHB2_MATH=1 (fast)"
#endif

namespace hb2 {
namespace {

constexpr int MATH = HB2_MATH;

/* the shock-sensor decisions of the whole patch in one pass (hb2_sensor.cuh) */
template <class Tr>
__global__ void __launch_bounds__(384, 2) k_sensor(const __grid_constant__ SensorArgs A)
{
    extern __shared__ double smem[];
    const int tid = (int)threadIdx.x;
    const SensorTile T = sensor_tile<Tr>(A, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
    SensorRegs<Tr> R;
    if (Tr::DIM == 3) {
        for (int t = T.kb - 2; t < T.kb; t++) {
            sensor_phase_fetch<Tr>(A, T, tid, t, R);
            sensor_phase_velocity<Tr, MATH>(smem, tid, t, R);
        }
        sensor_phase_fetch<Tr>(A, T, tid, T.kb, R);
        for (int t = T.kb; t <= T.ke + 1; t++) {
            if (t <= T.ke) sensor_phase_velocity<Tr, MATH>(smem, tid, t, R);
            if (t < T.ke) sensor_phase_fetch<Tr>(A, T, tid, t + 1, R);
            __syncthreads();
            if (t <= T.ke) sensor_phase_gradient<Tr, MATH>(A, smem, T, tid, t - 1);
            if (t >= T.kb + 2) sensor_phase_decision<Tr, MATH>(A, smem, T, tid, t - 2);
        }
    } else {
        sensor_phase_fetch<Tr>(A, T, tid, 0, R);
        sensor_phase_velocity<Tr, MATH>(smem, tid, 0, R);
        __syncthreads();
        sensor_phase_gradient<Tr, MATH>(A, smem, T, tid, 0);
        __syncthreads();
        sensor_phase_decision<Tr, MATH>(A, smem, T, tid, 0);
    }
}

#ifndef HB2_MINB
#define HB2_MINB 2
#endif
/* resident warps per SM the other sweeps (reference-order build, multi-species fast) are compiled for, where their shared
 * memory allows as many; 0: one block.  Reference-order single-species sweeps at 384^3, ms x / y / z
 * (profiles/r02_ar_exact_blocks_ab.txt): 256 threads, one block 14.0 / 14.5 / 15.0; 128 threads, as many blocks as fit at
 * <= 255 registers 16.2 / 14.7 / 15.6; 128 threads compiled for 12 warps (168 registers, three blocks) 12.5 / 12.0 / 13.1 */
#ifndef HB2_WARPS_EXACT
#define HB2_WARPS_EXACT 12
#endif
/* resident warps per SM the WCNS5 fast single-species sweeps are compiled for (16: 128 registers; 12: 168) */
#ifndef HB2_WARPS_X
#define HB2_WARPS_X (HB2_MINB * 8)
#endif
#ifndef HB2_WARPS_YZ
#define HB2_WARPS_YZ (HB2_MINB * 8)
#endif
/* resident warps per SM the WCNS6-LD fast sweeps are compiled for (8: up to 255 registers; 12: 168; 16: 128).  Measured at
 * 512^3, ms per sweep x / y / z (profiles/r02_ao_ld_warps_ab.txt): 8 warps 16.5 / 15.7 / 16.5, 12 warps 22.5 (spills) / 14.9 /
 * 15.5, 16 warps 20.8 / 15.0 / 16.2 */
#ifndef HB2_WARPS_LD
#define HB2_WARPS_LD 12
#endif
#ifndef HB2_WARPS_LD_X
#define HB2_WARPS_LD_X 8
#endif

/* minimum resident blocks per SM a sweep kernel is compiled for (its register budget): the warps asked for by the switches
 * above, but never more blocks than the shared memory of the SM holds */
template <class Tr, int DIR>
constexpr int sweep_min_blocks()
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const int warps = (MATH == 1 && Tr::MODEL == SS)
                          ? ((HB2_SCHEME != HB2_WCNS6_LD) ? ((DIR == 0) ? HB2_WARPS_X : HB2_WARPS_YZ) : ((DIR == 0) ? HB2_WARPS_LD_X : HB2_WARPS_LD))
                          : (MATH == 0 && HB2_SCHEME == HB2_WCNS6_LD) ? 8 /* the WCNS6-LD reference-order kernels spill at 168 registers */
                          : HB2_WARPS_EXACT;
    const int by_warps = warps * 32 / Sh::NT;
    const int by_smem = (227 * 1024) / (Sh::SMEM_DOUBLES * 8 + 1024);
    const int b = (by_warps < by_smem) ? by_warps : by_smem;
    return b > 0 ? b : 1;
}

template <class Tr, int DIR, int NTERM>
__global__ void __launch_bounds__((SweepShape<Tr, DIR, MATH>::NT), (sweep_min_blocks<Tr, DIR>())) k_sweep(const __grid_constant__ DirArgs A)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    extern __shared__ double smem[];
    const BlockId b = {(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z};
    PencilCtx c = pencil_ctx<Tr, DIR, MATH>(A, b, (int)threadIdx.x);
    c.sq = (unsigned)__cvta_generic_to_shared(smem + Sh::OFF_Q + threadIdx.x);
    const int nsteps = Sh::nsteps(c.c1 - c.c0);
    PipeRegs<Tr> pr;
    pr.mbar = (unsigned)__cvta_generic_to_shared(smem + Sh::OFF_B);
    if (A.bulk && threadIdx.x == 0) {
        mbar_init(pr.mbar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pipeline_prologue<Tr, DIR, MATH>(A, smem, c, pr);
    int t_lo, t_hi;
    steady_range<Tr, DIR, MATH>(c, t_lo, t_hi);
    __syncthreads();
    for (int t = 0; t <= nsteps; t++) {
        if (A.bulk) pipeline_issue<Tr, DIR, MATH>(A, smem, c, t + 1, pr.mbar);
        pipeline_step<Tr, DIR, MATH, NTERM>(A, smem, c, t, nsteps, t_lo, t_hi, pr);
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
int launch_sensor_t(const SensorArgs& A, cudaStream_t st)
{
    using Sh = SensorShape<Tr>;
    const size_t smem = (size_t)Sh::SMEM_DOUBLES * sizeof(double);
    /* function attributes are per device: one process may hold plans on several GPUs */
    static unsigned long long attr_set = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_sensor<Tr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set |= 1ull << (dev & 63);
    }
    dim3 grid(Sh::tiles_x(A.G), Sh::tiles_y(A.G), Sh::segments(A.G, A.seg_len));
    k_sensor<Tr><<<grid, Sh::NT, smem, st>>>(A);
    return (int)cudaGetLastError();
}

template <class Tr, int DIR, int NTERM>
int launch_dir_n(const DirArgs& A, cudaStream_t st)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const Geom& G = A.G;
    const size_t smem = (size_t)Sh::SMEM_DOUBLES * sizeof(double);
    static unsigned long long attr_set = 0; /* per device, see launch_sensor_t */
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_sweep<Tr, DIR, NTERM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_sweep<Tr, DIR, NTERM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        attr_set |= 1ull << (dev & 63);
    }
    const int nseg = (G.n[DIR] + A.seg_len - 1) / A.seg_len;
    dim3 grid(1, 1, nseg);
    if (DIR == 0) {
        grid.x = (G.n[1] + Sh::P - 1) / Sh::P;
        grid.y = (Tr::DIM == 3) ? G.n[2] : 1;
    } else if (DIR == 1) {
        grid.x = (G.n[0] + Sh::P - 1) / Sh::P;
        grid.y = (Tr::DIM == 3) ? G.n[2] : 1;
    } else {
        grid.x = (G.n[0] + Sh::P - 1) / Sh::P;
        grid.y = G.n[1];
    }
    k_sweep<Tr, DIR, NTERM><<<grid, Sh::NT, smem, st>>>(A);
    return (int)cudaGetLastError();
}

template <class Tr, int DIR>
int launch_dir(const DirArgs& A, cudaStream_t st)
{
    if (A.mode == MODE_EMIT) return launch_dir_n<Tr, DIR, HB2_NTERM_EMIT>(A, st);
    /* the RK linear combination exists only in the last direction of a fused stage: only the variants the ABI asks
     * for are instantiated (fast build: the flux state rebuilt + 0..2 loaded states; exact build: 1..3 loaded states) */
    if constexpr (DIR == Tr::DIM - 1) {
        if constexpr (MATH == 1) {
            if (A.nterm == HB2_NTERM_QREC) return launch_dir_n<Tr, DIR, HB2_NTERM_QREC>(A, st);
            if (A.nterm == HB2_NTERM_QREC + 1) return launch_dir_n<Tr, DIR, HB2_NTERM_QREC + 1>(A, st);
            if (A.nterm == HB2_NTERM_QREC + 2) return launch_dir_n<Tr, DIR, HB2_NTERM_QREC + 2>(A, st);
        } else {
            if (A.nterm == 1) return launch_dir_n<Tr, DIR, 1>(A, st);
            if (A.nterm == 2) return launch_dir_n<Tr, DIR, 2>(A, st);
            if (A.nterm == 3) return launch_dir_n<Tr, DIR, 3>(A, st);
        }
        return (int)cudaErrorInvalidValue;
    } else {
        return launch_dir_n<Tr, DIR, 0>(A, st);
    }
}

template <class Tr>
int launch_sweep_t(const LaunchCfg&, int dir, const DirArgs& A, cudaStream_t st)
{
    if (dir == 0) return launch_dir<Tr, 0>(A, st);
    if (dir == 1) return launch_dir<Tr, 1>(A, st);
    return launch_dir<Tr, (Tr::DIM == 3 ? 2 : 1)>(A, st);
}

/* the four-eqn conservative model and the models with three species exist in the reference-order translation
 * units only (the reference is generic in d_num_species, FlowModelFiveEqnAllaire.cpp:29; its shipped decks use two) */
#if HB2_MATH == 0
#define HB2_DISPATCH_FC(cfg, CALL)                                                    \
    if ((cfg).model == FC && (cfg).dim == 2 && (cfg).ns == 2) { using Tr = Traits<FC, 2, 2>; CALL; } \
    if ((cfg).model == FC && (cfg).dim == 3 && (cfg).ns == 2) { using Tr = Traits<FC, 3, 2>; CALL; } \
    if ((cfg).model == FC && (cfg).dim == 2 && (cfg).ns == 3) { using Tr = Traits<FC, 2, 3>; CALL; } \
    if ((cfg).model == FC && (cfg).dim == 3 && (cfg).ns == 3) { using Tr = Traits<FC, 3, 3>; CALL; } \
    if ((cfg).model == FE && (cfg).dim == 2 && (cfg).ns == 3) { using Tr = Traits<FE, 2, 3>; CALL; } \
    if ((cfg).model == FE && (cfg).dim == 3 && (cfg).ns == 3) { using Tr = Traits<FE, 3, 3>; CALL; }
#else
#define HB2_DISPATCH_FC(cfg, CALL)
#endif

#define HB2_DISPATCH(cfg, CALL)                                                       \
    do {                                                                              \
        if ((cfg).model == SS && (cfg).dim == 2) { using Tr = Traits<SS, 2, 1>; CALL; } \
        if ((cfg).model == SS && (cfg).dim == 3) { using Tr = Traits<SS, 3, 1>; CALL; } \
        if ((cfg).model == FE && (cfg).dim == 2 && (cfg).ns == 2) { using Tr = Traits<FE, 2, 2>; CALL; } \
        if ((cfg).model == FE && (cfg).dim == 3 && (cfg).ns == 2) { using Tr = Traits<FE, 3, 2>; CALL; } \
        HB2_DISPATCH_FC(cfg, CALL)                                                    \
    } while (0)

int op_sensor(const LaunchCfg& cfg, const SensorArgs& A, cudaStream_t st)
{
    HB2_DISPATCH(cfg, return launch_sensor_t<Tr>(A, st));
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
#if HB2_SCHEME == HB2_WCNS5_Z
const Ops* ops_exact_z() { return &g_ops; }
#elif HB2_SCHEME == HB2_WCNS6_LD
const Ops* ops_exact_ld() { return &g_ops; }
#else
const Ops* ops_exact() { return &g_ops; }
#endif
#else
#if HB2_SCHEME == HB2_WCNS5_Z
const Ops* ops_fast_z() { return &g_ops; }
#elif HB2_SCHEME == HB2_WCNS6_LD
const Ops* ops_fast_ld() { return &g_ops; }
#else
const Ops* ops_fast() { return &g_ops; }
#endif
#endif

}  // namespace hb2
