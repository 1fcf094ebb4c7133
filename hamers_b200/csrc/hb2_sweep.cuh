/*
 * hb2_sweep.cuh -- thread mapping of one direction sweep: a thread block MARCHES along the sweep axis.
 *
 * One block owns P pencils (lines of cells along the sweep axis) and advances along them in chunks of C cells;
 * P*C = 256 threads, one thread per (pencil, position in the chunk):
 *     y / z sweep : P = 32 pencils = 32 consecutive x (lanes -> 256-byte coalesced rows), C = 8 (one row per warp)
 *     x sweep     : P = 8 pencils = 8 consecutive y rows (one per warp), C = 32 consecutive x (lanes)
 * Per step t the block runs three phases separated by barriers; all hand-over goes through shared-memory RINGS
 * indexed by the position along the sweep axis (slot = position mod RING):
 *     L(t)  load    cells  c0-4+tC+o  : conservative -> primitive variables, sound speed, node flux   -> sV, sN
 *     F(t)  face    faces  c0-6+tC+o  : characteristic projection, WCNS5-JS, bounds check, HLLC/HLLC-HLL -> sM
 *     U(t)  update  cells  c0-8+tC+o  : 6th-order midpoint-and-node flux difference, advective source, then either
 *                                       the side flux (EMIT) or the running right-hand side / fused RK update (FUSED)
 * so every cell is converted once and every midpoint flux is computed once per sweep (the reference writes ~70
 * patch-sized temporaries instead), and nothing but the final result leaves the SM.
 * Global loads of step t+1 are issued before phase F(t) and committed to the ring after it (latency hidden
 * behind the FP64-bound phase).
 *
 * The phase bodies are `__host__ __device__` functions of (block coordinates, thread id): the CUDA kernel calls them
 * with barriers in between, tests/host_emu calls them from loops.
 *
 * Reference behaviour: ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:1378-2655 (3D), :546-1376 (2D);
 * Euler.cpp:1424-1655 (RK update).
 */
#pragma once
#include "hb2_fast.cuh"

namespace hb2 {

template <class Tr, int DIR>
struct SweepShape {
    static constexpr int NT = 256;
    static constexpr int NW = NT / 32;
    static constexpr int P = (DIR == 0) ? NW : 32;    /* pencils per block */
    static constexpr int C = (DIR == 0) ? 32 : NW;    /* cells per chunk along the sweep axis */
    static constexpr int RING = (DIR == 0) ? 64 : 16; /* ring slots along the sweep axis (>= C + 5, power of two) */
    static constexpr int CS = RING * P;               /* doubles per ring component */
    static constexpr int NV = Tr::NEQ + 1;            /* primitive variables + sound speed */
    static constexpr int NN = Tr::NEQ;                /* node flux */
    static constexpr int NMID = Tr::NEQ + (Tr::ADV ? 1 : 0); /* midpoint flux (+ HLLC midpoint velocity) */
    static constexpr int SMEM_DOUBLES = (NV + NN + NMID) * CS;
    HB2_HD static int slot(int pp, int s) { return (DIR == 0) ? pp * RING + (s & (RING - 1)) : (s & (RING - 1)) * 32 + pp; }
    HB2_HD static int nsteps(int ncells) { return (ncells + 8 + C - 1) / C; }
};

struct BlockId {
    int x, y, z;
};

/* what one thread needs to know about its pencil */
struct PencilCtx {
    int pp;          /* pencil index inside the block */
    int o;           /* position inside the chunk */
    bool valid;      /* pencil exists */
    int c0, c1;      /* segment [c0, c1) of cells along the sweep axis */
    int i, j, k;     /* coordinates of sweep cell 0 of the pencil */
    long long base;  /* ghost-box index of sweep cell 0 */
    long long st;    /* ghost-box stride along the sweep axis */
};

template <class Tr, int DIR>
HB2_HD PencilCtx pencil_ctx(const DirArgs& A, const BlockId& b, int tid)
{
    using Sh = SweepShape<Tr, DIR>;
    const Geom& G = A.G;
    PencilCtx c;
    const int lane = tid & 31, w = tid >> 5;
    c.pp = (DIR == 0) ? w : lane;
    c.o = (DIR == 0) ? lane : w;
    c.i = c.j = c.k = 0;
    if (DIR == 0) {
        c.j = b.x * Sh::P + c.pp;
        c.k = b.y;
        c.valid = c.j < G.n[1];
    } else if (DIR == 1) {
        c.i = b.x * 32 + c.pp;
        c.k = b.y;
        c.valid = c.i < G.n[0];
    } else {
        c.i = b.x * 32 + c.pp;
        c.j = b.y;
        c.valid = c.i < G.n[0];
    }
    const int N = G.n[DIR];
    c.c0 = b.z * A.seg_len;
    c.c1 = (c.c0 + A.seg_len < N) ? c.c0 + A.seg_len : N;
    c.base = cidx(G, c.i, c.j, c.k);
    c.st = G.cs[DIR];
    return c;
}

/* ---- phase L: load ---------------------------------------------------------------------- */
template <class Tr, int DIR>
HB2_HD bool load_wanted(const PencilCtx& c, int t, int& s)
{
    using Sh = SweepShape<Tr, DIR>;
    s = c.c0 - 4 + t * Sh::C + c.o;
    return c.valid && s <= c.c1 + 3;
}

template <class Tr, int DIR, int MATH>
HB2_HD void phase_commit(const DirArgs& A, double* smem, const PencilCtx& c, int s, const double (&q)[Tr::NCOMP])
{
    using Sh = SweepShape<Tr, DIR>;
    constexpr int NEQ = Tr::NEQ;
    double V[NEQ], cs, Fn[NEQ];
    if (MATH == 0)
        cons_to_prim<Tr>(q, A.K, V, cs);
    else
        cons_to_prim_fast<Tr>(q, A.K, V, cs);
    node_flux<Tr, DIR>(q, V, Fn);
    double* sV = smem;
    double* sN = smem + Sh::NV * Sh::CS;
    const int sl = Sh::slot(c.pp, s);
#pragma unroll
    for (int e = 0; e < NEQ; e++) {
        sV[e * Sh::CS + sl] = V[e];
        sN[e * Sh::CS + sl] = Fn[e];
    }
    sV[NEQ * Sh::CS + sl] = cs;
}

/* ---- phase F: midpoint flux --------------------------------------------------------------- */
template <class Tr, int DIR, int MATH>
HB2_HD void phase_face(const DirArgs& A, double* smem, const PencilCtx& c, int t)
{
    using Sh = SweepShape<Tr, DIR>;
    constexpr int NEQ = Tr::NEQ;
    const int f = c.c0 - 6 + t * Sh::C + c.o;
    if (!c.valid || f < c.c0 - 1 || f > c.c1 + 1) return;
    const bool hybrid = (A.hyb[c.base + (long long)f * c.st] >> DIR) & 1;
    const double* sV = smem;
    double* sM = smem + (Sh::NV + Sh::NN) * Sh::CS;
    int so[6];
#pragma unroll
    for (int m = 0; m < 6; m++) so[m] = Sh::slot(c.pp, f - 3 + m);
    double Fm[NEQ], um;
    if (MATH == 0) {
        double V[6][NEQ];
#pragma unroll
        for (int m = 0; m < 6; m++)
#pragma unroll
            for (int e = 0; e < NEQ; e++) V[m][e] = sV[e * Sh::CS + so[m]];
        face_midpoint<Tr, DIR, 0>(V, sV[NEQ * Sh::CS + so[2]], sV[NEQ * Sh::CS + so[3]], hybrid, A.K, Fm, um);
    } else {
        face_midpoint_fast<Tr, DIR>(sV, Sh::CS, so, hybrid, A.K, Fm, um);
    }
    const int sl = Sh::slot(c.pp, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) sM[e * Sh::CS + sl] = Fm[e];
    if (Tr::ADV) sM[NEQ * Sh::CS + sl] = um;
}

/* ---- phase U: flux difference, source, output ------------------------------------------------ */
template <class Tr, int DIR, int MATH>
HB2_HD void phase_update(const DirArgs& A, const double* smem, const PencilCtx& c, int t)
{
    using Sh = SweepShape<Tr, DIR>;
    constexpr int DIM = Tr::DIM, NEQ = Tr::NEQ, NS = Tr::NS, IV = Tr::IV, IP = Tr::IP;
    constexpr bool LAST = (DIR == DIM - 1);
    const int cc = c.c0 - 8 + t * Sh::C + c.o;
    if (!c.valid || cc < c.c0 || cc >= c.c1) return;
    const Geom& G = A.G;
    const double* sV = smem;
    const double* sN = smem + Sh::NV * Sh::CS;
    const double* sM = smem + (Sh::NV + Sh::NN) * Sh::CS;
    const int m_m1 = Sh::slot(c.pp, cc - 1), m_0 = Sh::slot(c.pp, cc), m_p1 = Sh::slot(c.pp, cc + 1),
              m_p2 = Sh::slot(c.pp, cc + 2);
    const bool fused = (A.mode == MODE_FUSED);
    const double dxd = G.dx[DIR];
    const int ci = (DIR == 0) ? cc : c.i, cj = (DIR == 1) ? cc : c.j, ck = (DIR == 2) ? cc : c.k;
    const long long ix = iidx(G, ci, cj, ck);

    /* velocity-divergence contribution of this direction (advective equations of the five-eqn model) */
    double Tsum = 0.0;
    if (Tr::ADV) {
        const double* um = sM + NEQ * Sh::CS;
        const double* un = sV + (IV + DIR) * Sh::CS;
        const double Td = (3.0 / 2.0 * (um[m_p1] - um[m_0]) - 3.0 / 10.0 * (un[m_p1] - un[m_m1]) +
                           1.0 / 30.0 * (um[m_p2] - um[m_m1])) / dxd;
        Tsum = (DIR == 0) ? Td : A.T[ix] + Td;
        if (!LAST) A.T[ix] = Tsum;
    }

    double rhs[NEQ];
    if (MATH == 1 && fused) {
        /* difference form: F[c+1] - F[c] = dt (3/2 (M[c+1]-M[c]) + 1/30 (M[c+2]-M[c-1]) - 3/10 (N[c+1]-N[c-1])) */
        const double k0 = A.dt / dxd;
        const double k1 = 1.5 * k0, k2 = (1.0 / 30.0) * k0, k3 = (3.0 / 10.0) * k0;
#pragma unroll
        for (int e = 0; e < NEQ; e++) {
            const double* M = sM + e * Sh::CS;
            const double* Nf = sN + e * Sh::CS;
            const double r0 = (DIR == 0) ? 0.0 : A.R[e][ix];
            rhs[e] = fma(k3, Nf[m_p1] - Nf[m_m1], fma(-k2, M[m_p2] - M[m_m1], fma(-k1, M[m_p1] - M[m_0], r0)));
        }
    } else {
#pragma unroll
        for (int e = 0; e < NEQ; e++) {
            const double* M = sM + e * Sh::CS;
            const double* Nf = sN + e * Sh::CS;
            const double F_lo = A.dt * (1.0 / 30.0 * (M[m_p1] + M[m_m1]) - 3.0 / 10.0 * (Nf[m_0] + Nf[m_m1]) + 23.0 / 15.0 * M[m_0]);
            const double F_hi = A.dt * (1.0 / 30.0 * (M[m_p2] + M[m_0]) - 3.0 / 10.0 * (Nf[m_p1] + Nf[m_0]) + 23.0 / 15.0 * M[m_p1]);
            if (!fused) {
                A.F[e][sidx<DIR>(G, ci, cj, ck)] = F_lo;
                if (cc + 1 == G.n[DIR])
                    A.F[e][sidx<DIR>(G, ci + (DIR == 0), cj + (DIR == 1), ck + (DIR == 2))] = F_hi;
            } else {
                const double dF = (F_hi - F_lo) / dxd;
                rhs[e] = (DIR == 0) ? -dF : A.R[e][ix] - dF;
            }
        }
    }

    if (fused) {
        if (LAST) {
            if (Tr::ADV) {
#pragma unroll
                for (int si = 0; si < NS - 1; si++) {
                    const int e = IP + 1 + si;
                    rhs[e] = rhs[e] + A.dt * sV[e * Sh::CS + m_0] * Tsum;
                }
            }
            rk_update_cell<Tr>(A, c.base + (long long)cc * c.st, rhs);
        } else {
#pragma unroll
            for (int e = 0; e < NEQ; e++) A.R[e][ix] = rhs[e];
        }
    } else if (LAST && Tr::ADV) {
#pragma unroll
        for (int si = 0; si < NS - 1; si++) {
            const int e = IP + 1 + si;
            A.S[e][ix] += A.dt * sV[e * Sh::CS + m_0] * Tsum;
        }
    }
}

}  // namespace hb2
