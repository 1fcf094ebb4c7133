/*
 * hb2_sensor.cuh -- the shock-sensor decisions of a patch in ONE pass over the conservative variables.
 *
 * Reference: the velocity gradients (DerivativeFirstOrder.cpp:382, 601), dilatation theta and vorticity magnitude Omega
 * (ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:1523-1659, 2D :663-731) are written to patch-sized temporaries and
 * re-read per direction by the face loops that form s = -theta_avg/(|theta_avg| + Omega_avg + eps) and select HLLC-HLL
 * where s > 0.65 (:2072-2134).  Here a 256-thread block owns a TX x TY tile of cells and MARCHES along z:
 *
 *     A(t)  velocities of plane t      on the tile + halo (2 low, 1 high)  -> 4-plane shared-memory ring
 *           (density and momentum of plane t+1 are fetched into registers right behind, so the HBM latency hides
 *           behind the work on plane t)
 *     ---- the ONE barrier of the iteration ----
 *     B(t)  theta, Omega of plane t-1  on the tile + 1 low halo            -> 3-plane shared-memory ring
 *     C(t)  decisions of plane t-2     bit d = face between cells (c - e_d) and c, one byte per cell  -> HBM
 *           (from the theta/Omega planes t-2 and t-3 of earlier iterations: B and C share a phase)
 *
 * so the only HBM traffic is one read of density and momentum (L2 absorbs the halo overlap of neighbouring tiles) and
 * one byte written per cell: 33 B/cell instead of the 17 patch-sized double arrays the reference streams.
 *
 * The phase bodies are `__host__ __device__` functions of (block coordinates, thread id): the kernel puts barriers
 * between them, tests/host_emu calls them from loops.
 */
#pragma once
#include "hb2_fast.cuh"

namespace hb2 {

struct SensorArgs {
    Geom G;
    const double* Q[HB2_MAXC];
    unsigned char* hyb; /* ghost-box layout, valid on cells -1..N+1 after the pass */
    int seg_len;        /* planes per marching segment (3D) */
};

/* thread layout of the sensor pass: NX threads along x, 384 / NX rows of threads, VY rows of the velocity tile.
 *   64 x 6 threads, 11 rows (default): a 61 x 8 tile of decisions from 64 x 11 velocities (69 %), second row pass 5/6 busy
 *   48 x 8 threads, 16 rows: a 45 x 13 tile from 48 x 16 velocities (76 %), both row passes full, fewer idle columns on a
 *     256-cell box (4 % instead of 15 %) -- measured SLOWER on B200: 2.05 vs 1.73 ms at 512^3, 0.290 vs 0.264 ms at 256^3
 *     (rows of 48 threads straddle the warps: 1.5 warps per row, and the tile rows are no longer bank-aligned)
 *   32 x 12 threads, 24 rows: a 29 x 21 tile from 32 x 24 velocities (79 %), one warp per row -- SLOWER as well: 2.03 vs 1.82 ms
 *     at 512^3, 0.284 vs 0.264 ms at 256^3; 32 x 12 threads, 12 rows (29 x 9, one row per thread): 1.93 ms
 *     (profiles/r02_au_sensor_tiles_ab.txt).  256-byte rows that start 29 cells apart are never sector-aligned. */
#ifndef HB2_SENSOR_NX
#define HB2_SENSOR_NX 64
#endif
#ifndef HB2_SENSOR_VY
#define HB2_SENSOR_VY 11
#endif
#define HB2_SENSOR_TX (HB2_SENSOR_NX - 3)
#define HB2_SENSOR_TY (HB2_SENSOR_VY - 3)

template <class Tr>
struct SensorShape {
    static constexpr int NT = 384;
    static constexpr int NX = HB2_SENSOR_NX, NY = NT / NX; /* thread layout: tx = tid % NX along x, ty = tid / NX */
    static constexpr int VX = NX, VY = HB2_SENSOR_VY;      /* velocity tile: cells i0-2 .. i0+TX, one row of threads wide */
    static constexpr int TX = VX - 3, TY = VY - 3; /* cells whose decisions the block produces per plane */
    static_assert(NT % NX == 0, "whole rows of threads");
    static constexpr int SX = TX + 1, SY = TY + 1; /* theta/Omega tile: cells i0-1 .. i0+TX-1 */
    static constexpr int VP = VX * VY, SP = SX * SY;
    static constexpr int KY = (VY + NY - 1) / NY;  /* rows per thread */
    /* ring depths for ONE barrier per plane: a velocity plane is overwritten two barriers after its last reader, a
     * theta/Omega plane is read (decisions of planes t-2) while plane t-1 is being written */
    static constexpr int NVP = (Tr::DIM == 3) ? 4 : 1;
    static constexpr int NSP = (Tr::DIM == 3) ? 3 : 1;
    static constexpr int OFF_S = Tr::DIM * NVP * VP;
    static constexpr int SMEM_DOUBLES = OFF_S + 2 * NSP * SP;
    HB2_HD static int vslot(int t) { return (Tr::DIM == 3) ? ((t + 8) & 3) : 0; } /* t >= -3 */
    HB2_HD static int sslot(int t) { return (Tr::DIM == 3) ? (t + 6) % 3 : 0; }  /* t >= -3 */
    /* decisions are produced on cells -1..N+1 */
    HB2_HD static int tiles_x(const Geom& G) { return (G.n[0] + 3 + TX - 1) / TX; }
    HB2_HD static int tiles_y(const Geom& G) { return (G.n[1] + 3 + TY - 1) / TY; }
    HB2_HD static int segments(const Geom& G, int seg_len) { return (Tr::DIM == 3) ? (G.n[2] + 3 + seg_len - 1) / seg_len : 1; }
};

struct SensorTile {
    int i0, j0; /* first cell of the tile */
    int kb, ke; /* planes [kb, ke) of decisions */
};

/* density and momentum of the cells a thread fetched for the next velocity plane (in flight while the block works on
 * the current plane) */
template <class Tr>
struct SensorRegs {
    double rho[SensorShape<Tr>::KY];
    double m[SensorShape<Tr>::KY][Tr::DIM];
};

template <class Tr>
HB2_HD SensorTile sensor_tile(const SensorArgs& A, int bx, int by, int bz)
{
    using Sh = SensorShape<Tr>;
    SensorTile T;
    T.i0 = -1 + bx * Sh::TX;
    T.j0 = -1 + by * Sh::TY;
    if (Tr::DIM == 3) {
        T.kb = -1 + bz * A.seg_len;
        T.ke = (T.kb + A.seg_len < A.G.n[2] + 2) ? T.kb + A.seg_len : A.G.n[2] + 2;
    } else {
        T.kb = 0;
        T.ke = 1;
    }
    return T;
}

/* A1: fetch density and momentum of plane t on the tile + halo into registers */
template <class Tr>
HB2_HD void sensor_phase_fetch(const SensorArgs& A, const SensorTile& T, int tid, int t, SensorRegs<Tr>& R)
{
    using Sh = SensorShape<Tr>;
    constexpr int DIM = Tr::DIM, NM = Tr::NM;
    const Geom& G = A.G;
    const int tx = tid % Sh::NX, ty = tid / Sh::NX;
    const int i = T.i0 - 2 + tx;
#pragma unroll
    for (int k = 0; k < Sh::KY; k++) {
        const int lj = ty + k * Sh::NY;
        const int j = T.j0 - 2 + lj;
        R.rho[k] = 1.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) R.m[k][a] = 0.0;
        if (lj >= Sh::VY || i > G.n[0] + 3 || j > G.n[1] + 3) continue;
        const long long x = cidx(G, i, j, t);
        double rho = A.Q[0][x];
#pragma unroll
        for (int si = 1; si < NM; si++) rho += A.Q[si][x];
        R.rho[k] = rho;
#pragma unroll
        for (int a = 0; a < DIM; a++) R.m[k][a] = A.Q[NM + a][x];
    }
}

/* A2: velocities of plane t into the ring.  Exact build: the reference's quotients m/rho. */
template <class Tr, int MATH>
HB2_HD void sensor_phase_velocity(double* smem, int tid, int t, const SensorRegs<Tr>& R)
{
    using Sh = SensorShape<Tr>;
    constexpr int DIM = Tr::DIM;
    double* sV = smem + Sh::vslot(t) * DIM * Sh::VP;
    const int tx = tid % Sh::NX, ty = tid / Sh::NX;
#pragma unroll
    for (int k = 0; k < Sh::KY; k++) {
        const int lj = ty + k * Sh::NY;
        if (lj >= Sh::VY) continue;
        const int idx = lj * Sh::VX + tx;
        if (MATH == 0) {
#pragma unroll
            for (int a = 0; a < DIM; a++) sV[a * Sh::VP + idx] = R.m[k][a] / R.rho[k];
        } else {
            const double r = rcp_fast(R.rho[k]);
#pragma unroll
            for (int a = 0; a < DIM; a++) sV[a * Sh::VP + idx] = R.m[k][a] * r;
        }
    }
}

/* B: theta, Omega of plane tc from the velocity planes tc-1, tc, tc+1 */
template <class Tr, int MATH>
HB2_HD void sensor_phase_gradient(const SensorArgs& A, double* smem, const SensorTile& T, int tid, int tc)
{
    using Sh = SensorShape<Tr>;
    constexpr int DIM = Tr::DIM;
    const Geom& G = A.G;
    const double* Vc = smem + Sh::vslot(tc) * DIM * Sh::VP;
    const double* Vm = smem + Sh::vslot(tc - 1) * DIM * Sh::VP;
    const double* Vp = smem + Sh::vslot(tc + 1) * DIM * Sh::VP;
    double* sS = smem + Sh::OFF_S + Sh::sslot(tc) * 2 * Sh::SP;
    const double hidx[3] = {0.5 / G.dx[0], 0.5 / G.dx[1], 0.5 / G.dx[2]};
    const int li = tid % Sh::NX, ty = tid / Sh::NX;
    const int i = T.i0 - 1 + li;
    if (li >= Sh::SX || i > G.n[0] + 1) return;
#pragma unroll
    for (int k = 0; k < Sh::KY; k++) {
        const int lj = ty + k * Sh::NY;
        const int j = T.j0 - 1 + lj;
        if (lj >= Sh::SY || j > G.n[1] + 1) continue;
        const int vc = (lj + 1) * Sh::VX + (li + 1);
        double grad[DIM][DIM]; /* grad[a][b] = d u_a / d x_b */
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const double* U = Vc + a * Sh::VP + vc;
            double up[DIM], um[DIM];
            up[0] = U[1];
            um[0] = U[-1];
            up[1] = U[Sh::VX];
            um[1] = U[-Sh::VX];
            if (DIM == 3) {
                up[DIM - 1] = Vp[a * Sh::VP + vc];
                um[DIM - 1] = Vm[a * Sh::VP + vc];
            }
#pragma unroll
            for (int b = 0; b < DIM; b++)
                grad[a][b] = (MATH == 0) ? (0.5 * (up[b] - um[b])) / G.dx[b] : (up[b] - um[b]) * hidx[b];
        }
        double theta, Omega;
        if (DIM == 2) {
            theta = grad[0][0] + grad[1][1];
            Omega = fabs(grad[1][0] - grad[0][1]);
        } else {
            theta = grad[0][0] + grad[1][1] + grad[2 % DIM][2 % DIM];
            const double omega_x = grad[2 % DIM][1] - grad[1][2 % DIM];
            const double omega_y = grad[0][2 % DIM] - grad[2 % DIM][0];
            const double omega_z = grad[1][0] - grad[0][1];
            if (MATH == 0)
                Omega = sqrt(omega_x * omega_x + omega_y * omega_y + omega_z * omega_z);
            else
                Omega = sqrt_fast<true>(fma(omega_x, omega_x, fma(omega_y, omega_y, omega_z * omega_z)));
        }
        const int idx = lj * Sh::SX + li;
        sS[idx] = theta;
        sS[Sh::SP + idx] = Omega;
    }
}

/* fast build: s > 0.65 without the division (the denominator is positive):
 * -theta_avg/(|theta_avg| + Omega_avg + eps) > 0.65  <=>  -(thL+thR) > 0.65 (|thL+thR| + (OmL+OmR) + 2 eps) */
HB2_HD bool face_sensor_fast(double th_L, double th_R, double Om_L, double Om_R)
{
    const double ts = th_L + th_R;
    return -ts > HB2_SENSOR_THRESHOLD * (fabs(ts) + (Om_L + Om_R) + 2.0 * HB2_EPS);
}

/* C: decisions of plane tc from theta/Omega of planes tc-1, tc */
template <class Tr, int MATH>
HB2_HD void sensor_phase_decision(const SensorArgs& A, const double* smem, const SensorTile& T, int tid, int tc)
{
    using Sh = SensorShape<Tr>;
    constexpr int DIM = Tr::DIM;
    const Geom& G = A.G;
    const double* cur = smem + Sh::OFF_S + Sh::sslot(tc) * 2 * Sh::SP;
    const double* prev = smem + Sh::OFF_S + Sh::sslot(tc - 1) * 2 * Sh::SP;
    const int li = tid % Sh::NX, ty = tid / Sh::NX;
    const int i = T.i0 + li;
    if (li >= Sh::TX || i > G.n[0] + 1) return;
#pragma unroll
    for (int k = 0; k < Sh::KY; k++) {
        const int lj = ty + k * Sh::NY;
        const int j = T.j0 + lj;
        if (lj >= Sh::TY || j > G.n[1] + 1) continue;
        const int sc = (lj + 1) * Sh::SX + (li + 1);
        const double th = cur[sc], Om = cur[Sh::SP + sc];
        double thl[DIM], Oml[DIM];
        thl[0] = cur[sc - 1];
        Oml[0] = cur[Sh::SP + sc - 1];
        thl[1] = cur[sc - Sh::SX];
        Oml[1] = cur[Sh::SP + sc - Sh::SX];
        if (DIM == 3) {
            thl[DIM - 1] = prev[sc];
            Oml[DIM - 1] = prev[Sh::SP + sc];
        }
        unsigned int f = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const bool on = (MATH == 0) ? face_sensor(thl[d], th, Oml[d], Om) : face_sensor_fast(thl[d], th, Oml[d], Om);
            if (on) f |= (1u << d);
        }
        A.hyb[cidx(G, i, j, tc)] = (unsigned char)f;
    }
}

}  // namespace hb2
