/*
 * hb2_sweep.cuh -- thread mapping of one direction sweep: a thread block MARCHES along the sweep axis.
 *
 * One block owns P pencils (lines of cells along the sweep axis) and advances along them in chunks of C cells;
 * P*C = 256 threads, one thread per (pencil, position in the chunk):
 *     y / z sweep : P = 32 pencils = 32 consecutive x (lanes -> 256-byte coalesced rows), C = 8 (one row per warp)
 *     x sweep     : P = 16 pencils = 16 consecutive y rows (two per warp), C = 16 consecutive x (half a warp), HB2_XC
 * The work of a pencil is a three-stage software pipeline whose hand-over goes through shared-memory RINGS indexed
 * by the position along the sweep axis (slot = position mod RING):
 *     L(t)  load    cells  c0-4+tC+o  : conservative -> primitive variables, sound speed              -> sV (, sN)
 *     F(t)  face    faces  c0-6+tC+o  : characteristic projection, WCNS5-JS, bounds check, HLLC/HLLC-HLL -> sM
 *     U(t)  update  cells  c0-8+tC+o  : 6th-order midpoint-and-node flux difference, advective source, then either
 *                                       the side flux (EMIT) or the running right-hand side / fused RK update (FUSED)
 * so every cell is converted once and every midpoint flux is computed once per sweep (the reference writes ~70
 * patch-sized temporaries instead), and nothing but the final result leaves the SM.
 * Iteration t of a block executes  L(t+1) ; F(t) ; U(t-1)  in every thread and ends with ONE barrier: the rings are
 * long enough (3C+5 cells, 2C+3 faces) that nothing written in an iteration is read by another thread before the
 * barrier.  The global loads of L(t+2) and of U(t-1) are issued before F(t), so their latency hides behind the
 * FP64-bound face phase, and every phase keeps the FP64 pipe busy (no load-only or store-only phases).
 *
 * The phase bodies are `__host__ __device__` functions of (block coordinates, thread id): the CUDA kernel calls them
 * with the barrier in between, tests/host_emu calls them from loops (in forward and in reverse thread order, to
 * expose any intra-iteration hazard).
 *
 * Reference behaviour: ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:1378-2655 (3D), :546-1376 (2D);
 * Euler.cpp:1424-1655 (RK update).
 */
#pragma once
#include "hb2_fast.cuh"

/* Prefetch switches, chosen by measurement on B200 at 256^3 (ms per sweep x / y / z; the kernels sit at the
 * 128-register limit of two resident blocks, so every extra value held across the face phase can tip into spills):
 *   HB2_PREFETCH_R    1: fetch the running right-hand side of the update phase before the face phase, 0: after it
 *   HB2_PREFETCH_FLAG 1: fetch the sensor byte of the face one iteration ahead, 0: at the start of its face phase
 *     FLAG=0 R=1 : 1.10 / 1.06 / 1.25      FLAG=0 R=0 : 1.10 / 1.11 / 1.26      FLAG=1 R=1 : 1.07 / 1.36 / 1.21
 * Since the conservative variables of the load phase are staged with cp.async (no register held across the face
 * phase) FLAG=1 no longer spills and is the default: 0.94 / 0.99 / 1.13.  Staging the sensor byte with cp.async as
 * well was measured too (0.95 / 1.00 / 1.14) and dropped.  With the spills gone R=0 became the better choice at 512^3
 * (ms per sweep x / y / z: R=1 6.85 / 7.62 / 8.75, R=0 6.84 / 7.27 / 8.52; the evict-first hints of
 * HB2_STREAM_HINTS make no difference). */
#ifndef HB2_PREFETCH_FLAG
#define HB2_PREFETCH_FLAG 1
#endif
#ifndef HB2_PREFETCH_R
#define HB2_PREFETCH_R 0
#endif
/* ... and per direction once the blocks became smaller (64 / 128 / 128 threads, below): at 512^3 R=1 measures
 * x 5.92 / y 6.15 / z 6.64 ms against 5.88 / 6.30 / 6.84 with R=0 (profiles/r02_am_flags_ab.txt): x keeps 0, y and z take 1 */
#ifndef HB2_PREFETCH_R_YZ
#define HB2_PREFETCH_R_YZ 1
#endif
/* steady-state variant of the iteration body (no "wanted" tests) for the interior iterations of full blocks; 0 = off */
#ifndef HB2_STEADY
#define HB2_STEADY 1
#endif
/* L2 prefetch distance (iterations) of the update phase's HBM inputs; 0 = off */
#ifndef HB2_PREFETCH_L2
#define HB2_PREFETCH_L2 1
#endif

/* x sweep: cells of a row one iteration advances (lanes along x).  A pencil of n cells takes ceil((n + 8)/XC)
 * iterations: with 32 a 512-cell row needs 17 for 16.25 (4 % of the lanes idle) and a 256-cell row 9 for 8.25 (8 %);
 * with 16 (two rows per warp) it is 33 for 32.5 and 17 for 16.5. */
#ifndef HB2_XC
#define HB2_XC 16
#endif

/* Threads per block of the fast sweeps (single-species / multi-species models), per sweep direction.  The resident warps
 * per SM stay the same (shared memory and the 128-register limit allow 16, five-eqn 8), but smaller blocks are more
 * independent barrier domains whose FP64-heavy face phases and load / update phases interleave.  Measured on B200, ms per
 * sweep x / y / z (profiles/r02_ak_ss_compact_ab.txt, r02_aj_fe_compact_ab.txt):
 *   single-species 512^3:  256 threads 6.50 / 6.69 / 7.19   128: 6.20 / 6.41 / 6.95   64: 5.92 / 6.62 / 9.90 (x at 32: 5.94)
 *   five-eqn 384^3:        256 threads 5.96 / 6.56 / 6.58   128: 5.45 / 5.78 / 6.03   64: 5.42 / 5.47 / 6.81 (x at 32: 6.55)
 * (y / z blocks of 64 threads own rows of 8 cells = 64 B: the z sweep's plane-strided accesses no longer fill DRAM bursts).
 * ..._Z is the LAST direction of the model's dimension (the sweep with the RK update): the 2-D five-eqn y sweep takes 5.22 ms
 * with 128 threads against 6.25 with 64 at 8192^2.
 * Reference-order (MATH == 0) kernels: 128 threads (HB2_NT_EXACT), three resident blocks where the rings allow
 * (hb2_sweeps.cu: HB2_WARPS_EXACT). */
#ifndef HB2_NT_EXACT
#define HB2_NT_EXACT 128
#endif
#ifndef HB2_NT_EXACT_X
#define HB2_NT_EXACT_X HB2_NT_EXACT
#endif
#ifndef HB2_NT_SS_X
#define HB2_NT_SS_X 64
#endif
#ifndef HB2_NT_SS_Y
#define HB2_NT_SS_Y 128
#endif
#ifndef HB2_NT_SS_Z
#define HB2_NT_SS_Z 128
#endif
#ifndef HB2_NT_MS_X
#define HB2_NT_MS_X 64
#endif
#ifndef HB2_NT_MS_Y
#define HB2_NT_MS_Y 64
#endif
#ifndef HB2_NT_MS_Z
#define HB2_NT_MS_Z 128
#endif

namespace hb2 {

/* shared memory of a block of nt threads (doubles): rings of NV primitive, NN node-flux and NMID midpoint-flux components,
 * the staging slots and the push table (the layout of SweepShape below) */
template <class Tr, int DIR, int MATH>
constexpr int sweep_smem_doubles(int nt)
{
    const int nv = Tr::NEQ + 1 + ((MATH == 1 && Tr::MODEL == FE) ? 1 : 0), nn = (MATH == 0) ? Tr::NEQ : 0;
    const int nmid = Tr::NEQ + (Tr::ADV ? 1 : 0);
    const int p = (DIR == 0) ? nt / HB2_XC : nt / 8;
    const int ring = (DIR == 0) ? 4 * HB2_XC : 32, dup = (DIR == 0) ? 8 : 5;
    return nv * (ring + dup) * p + (nn + nmid) * ring * p + Tr::NCOMP * nt + 27 * Tr::NCOMP + 3;
}

template <class Tr, int DIR, int MATH>
struct SweepShape {
    /* 256 threads; models with so many equations that the rings of a 256-thread block exceed the 227 KB of shared memory
     * (five-eqn with three species, reference-order build: 273 KB) run COMPACT blocks of 128 threads with half the pencils --
     * y / z sweeps: 16 consecutive x per row (two rows per warp), x sweep: 8 rows */
    static constexpr int NT_FIT = (sweep_smem_doubles<Tr, DIR, MATH>(256) * 8 <= 227 * 1024) ? 256 : 128;
    static constexpr int NT_CAP = (MATH != 1) ? ((DIR == 0) ? HB2_NT_EXACT_X : HB2_NT_EXACT)
        : (Tr::MODEL == SS) ? ((DIR == 0) ? HB2_NT_SS_X : (DIR != Tr::DIM - 1) ? HB2_NT_SS_Y : HB2_NT_SS_Z)
                            : ((DIR == 0) ? HB2_NT_MS_X : (DIR != Tr::DIM - 1) ? HB2_NT_MS_Y : HB2_NT_MS_Z);
    static constexpr int NT = (NT_FIT < NT_CAP) ? NT_FIT : NT_CAP;
    static constexpr int NW = NT / 32;
    static constexpr int P = (DIR == 0) ? NT / HB2_XC : NT / 8;   /* pencils per block */
    static constexpr int C = (DIR == 0) ? HB2_XC : 8;             /* cells per chunk along the sweep axis */
    /* ring slots along the sweep axis: one iteration touches 3C+5 consecutive cells and 2C+3 consecutive faces */
    static constexpr int RING = (DIR == 0) ? 4 * HB2_XC : 32;
    static constexpr int CS = RING * P;                /* doubles per ring component (midpoint flux, node flux) */
    /* The primitive-variable ring is MIRRORED: its first DUP slots are stored a second time behind the last slot, so
     * that the six stencil cells of a face are always at base + m*MS (no wrap inside a stencil window). */
    static constexpr int DUP = (DIR == 0) ? 8 : 5;
    static constexpr int RINGV = RING + DUP;
    static constexpr int CSV = RINGV * P;              /* doubles per component of the primitive-variable ring */
    static constexpr int MS = (DIR == 0) ? 1 : P;      /* stride between consecutive cells of a pencil */
    /* primitive variables + sound speed (+ total energy for the five-eqn model, whose node flux needs the stored one) */
    static constexpr int IC = Tr::NEQ;                 /* sound speed */
    static constexpr int IE = Tr::NEQ + 1;             /* total energy (five-eqn only) */
    static constexpr int NV = Tr::NEQ + 1 + ((MATH == 1 && Tr::MODEL == FE) ? 1 : 0);
    /* exact arithmetic keeps the node flux of the conservative variables (bit-identical to the reference); the fast
     * variant re-evaluates it from the primitive ring in the update phase and saves the ring */
    static constexpr int NN = (MATH == 0) ? Tr::NEQ : 0;
    static constexpr int NMID = Tr::NEQ + (Tr::ADV ? 1 : 0); /* midpoint flux (+ HLLC midpoint velocity) */
    static constexpr int OFF_N = NV * CSV;
    static constexpr int OFF_M = OFF_N + NN * CS;
    /* staging area of the load phase: the conservative variables of the cell a thread fetched with cp.async one
     * iteration ago, [component][thread]; every thread reads back only its own slots, so no barrier is involved and no
     * register is held across the face phase */
    static constexpr int OFF_Q = OFF_M + NMID * CS;
    /* the push addresses of the fused ghost fill (DirArgs::push), 27 neighbours x components */
    static constexpr int OFF_P = OFF_Q + Tr::NCOMP * NT;
    /* mbarrier of the bulk-copy staging (one 64-bit word) */
    static constexpr int OFF_B = OFF_P + 27 * Tr::NCOMP + ((OFF_P + 27 * Tr::NCOMP) & 1);
    static constexpr int SMEM_DOUBLES = OFF_B + 2;
    HB2_HD static int slot(int pp, int s) { return (DIR == 0) ? pp * RING + (s & (RING - 1)) : (s & (RING - 1)) * P + pp; }
    /* primitive-variable ring: r = ring position in [0, RINGV) */
    HB2_HD static int slotv(int pp, int r) { return (DIR == 0) ? pp * RINGV + r : r * P + pp; }
    HB2_HD static int nsteps(int ncells) { return (ncells + 8 + C - 1) / C; }
};

struct BlockId {
    int x, y, z;
};

/* what one thread needs to know about its pencil */
struct PencilCtx {
    int tid;         /* thread index inside the block */
    unsigned sq;     /* device: shared-window address of the thread's first staging slot (set once by the kernel; the
                        generic -> shared conversion costs an S2R when left inside the loop) */
    int pp;          /* pencil index inside the block */
    int o;           /* position inside the chunk */
    bool valid;      /* pencil exists */
    bool full;       /* every pencil of the block exists (block-uniform) */
    int c0, c1;      /* segment [c0, c1) of cells along the sweep axis */
    int i, j, k;     /* coordinates of sweep cell 0 of the pencil */
    long long base;  /* ghost-box index of sweep cell 0 */
    long long st;    /* ghost-box stride along the sweep axis */
    long long ibase; /* ghost-0 (interior layout) index of sweep cell 0 */
    long long ist;   /* interior-layout stride along the sweep axis */
};

template <class Tr, int DIR, int MATH>
HB2_HD PencilCtx pencil_ctx(const DirArgs& A, const BlockId& b, int tid)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const Geom& G = A.G;
    PencilCtx c;
    c.tid = tid;
    c.sq = 0;
    c.pp = (DIR == 0) ? tid / Sh::C : tid % Sh::P;      /* y / z: lanes along x (a full warp row, or two half rows when compact) */
    c.o = (DIR == 0) ? tid % Sh::C : tid / Sh::P;
    c.i = c.j = c.k = 0;
    if (DIR == 0) {
        c.j = b.x * Sh::P + c.pp;
        c.k = b.y;
        c.valid = c.j < G.n[1];
        c.full = (b.x + 1) * Sh::P <= G.n[1];
    } else if (DIR == 1) {
        c.i = b.x * Sh::P + c.pp;
        c.k = b.y;
        c.valid = c.i < G.n[0];
        c.full = (b.x + 1) * Sh::P <= G.n[0];
    } else {
        c.i = b.x * Sh::P + c.pp;
        c.j = b.y;
        c.valid = c.i < G.n[0];
        c.full = (b.x + 1) * Sh::P <= G.n[0];
    }
    const int N = G.n[DIR];
    c.c0 = b.z * A.seg_len;
    c.c1 = (c.c0 + A.seg_len < N) ? c.c0 + A.seg_len : N;
    c.base = cidx(G, c.i, c.j, c.k);
    c.st = G.cs[DIR];
    c.ibase = iidx(G, c.i, c.j, c.k);
    c.ist = (DIR == 0) ? 1 : ((DIR == 1) ? (long long)G.n[0] : (long long)G.n[0] * G.n[1]);
    return c;
}

/* ---- asynchronous global -> shared staging (LDGSTS): 8 bytes per call, completion awaited by the issuing thread ---- */
HB2_HD void stage_async8(double* dst_smem, unsigned dst_shared, const double* src)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_shared), "l"(src) : "memory");
#else
    *dst_smem = *src;
#endif
}
HB2_HD void stage_wait_all()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

template <class Tr, int DIR, int MATH>
HB2_HD void stage_cons(const DirArgs& A, double* smem, const PencilCtx& c, long long x)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    double* sQ = smem + Sh::OFF_Q + c.tid;
#pragma unroll
    for (int cix = 0; cix < Tr::NCOMP; cix++)
        stage_async8(sQ + cix * Sh::NT, c.sq + (unsigned)(cix * Sh::NT * sizeof(double)), A.Q[cix] + x);
}

template <class Tr, int DIR, int MATH>
HB2_HD void staged_cons(const double* smem, const PencilCtx& c, double (&q)[Tr::NCOMP])
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const double* sQ = smem + Sh::OFF_Q + c.tid;
#pragma unroll
    for (int cix = 0; cix < Tr::NCOMP; cix++) q[cix] = sQ[cix * Sh::NT];
}

/* ---- load / commit: chunk t = cells c0-4+tC+o ------------------------------------------------- */
template <class Tr, int DIR, int MATH>
HB2_HD bool load_wanted(const PencilCtx& c, int t, int& s)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    s = c.c0 - 4 + t * Sh::C + c.o;
    return c.valid && s <= c.c1 + 3;
}

template <class Tr, int DIR, int MATH>
HB2_HD void phase_commit(const DirArgs& A, double* smem, const PencilCtx& c, int s, const double (&q)[Tr::NCOMP])
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    constexpr int NEQ = Tr::NEQ;
    double V[NEQ], cs;
    if constexpr (MATH == 0)
        cons_to_prim<Tr>(q, A.K, V, cs);
    else
        cons_to_prim_fast<Tr>(q, A.K, V, cs);
    double* sV = smem;
    const int r = s & (Sh::RING - 1);
    const int sv = Sh::slotv(c.pp, r);
#pragma unroll
    for (int e = 0; e < NEQ; e++) sV[e * Sh::CSV + sv] = V[e];
    sV[Sh::IC * Sh::CSV + sv] = cs;
    if (Sh::NV > NEQ + 1) sV[Sh::IE * Sh::CSV + sv] = q[Tr::IP];
    if (r < Sh::DUP) {
        const int sv2 = Sh::slotv(c.pp, r + Sh::RING);
#pragma unroll
        for (int e = 0; e < NEQ; e++) sV[e * Sh::CSV + sv2] = V[e];
        sV[Sh::IC * Sh::CSV + sv2] = cs;
    }
    if (MATH == 0) {
        double Fn[NEQ];
        node_flux<Tr, DIR>(q, V, Fn);
        double* sN = smem + Sh::OFF_N;
        const int sl = Sh::slot(c.pp, s);
#pragma unroll
        for (int e = 0; e < NEQ; e++) sN[e * Sh::CS + sl] = Fn[e];
    }
}

/* ---- bulk-copy staging (DirArgs::bulk): the conservative variables of a whole chunk travel as ROWS with cp.async.bulk
 * (SASS UBLKCP, the TMA unit's non-tensor form), issued by the lanes of warp 0 and completed on one mbarrier -- instead of
 * five 8-byte LDGSTS plus their address arithmetic in EVERY thread (in kernels where each non-FP64 instruction competes
 * with the FP64 issue slot).  A row is the chunk's cells of one (component, sweep position) -- y / z sweeps: 32 pencils
 * contiguous in x (256 B) -- or of one (component, pencil) -- x sweep: 16 cells (128 B); the staging slots keep the layout
 * [component][thread].  The iteration order changes with it: the copy of chunk t+1 is issued at the TOP of iteration t
 * (after the barrier that ends iteration t-1: every thread has read its slot by then) and consumed at the END of
 * iteration t, behind the face and update phases; face(t) never needs chunk t+1 (it reads cells up to c0 + (t+1)C - 5).
 * Needs 16-byte aligned rows: even n[0], even ghost width, 16-byte aligned component pointers, even segment length in x
 * (the host checks; otherwise the LDGSTS path runs). */
template <class Tr, int DIR, int MATH>
HB2_HD void bulk_rows(const PencilCtx& c, const Geom& G, int t, int& nrow, int& rowlen)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const int s0 = c.c0 - 4 + t * Sh::C;
    int along = c.c1 + 4 - s0;                           /* cells of the chunk that exist */
    along = along < 0 ? 0 : (along > Sh::C ? Sh::C : along);
    if (DIR == 0) {
        const int across = G.n[1] - (c.j - c.pp);        /* pencils (rows) of the block that exist */
        nrow = across > Sh::P ? Sh::P : across;
        rowlen = along;
    } else {
        const int across = G.n[0] - (c.i - c.pp);
        nrow = along;
        rowlen = across > Sh::P ? Sh::P : across;
    }
    if (nrow < 0) nrow = 0;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(mbar),
        "r"(parity)
        : "memory");
}
#endif

/* top of iteration t - 1 (and of the prologue's successor): stage chunk t.  Every thread calls it; warp 0 works. */
template <class Tr, int DIR, int MATH>
HB2_HD void pipeline_issue(const DirArgs& A, double* smem, const PencilCtx& c, int t, unsigned mbar)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    if (c.tid >= 32) return;
    int nrow, rowlen;
    bulk_rows<Tr, DIR, MATH>(c, A.G, t, nrow, rowlen);
    if (nrow <= 0 || rowlen <= 0) return;
    const int s0 = c.c0 - 4 + t * Sh::C;
    double* sQ = smem + Sh::OFF_Q;
#if defined(__CUDA_ARCH__)
    if (c.tid == 0) mbar_expect_tx(mbar, (unsigned)(rowlen * 8 * nrow * Tr::NCOMP));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const unsigned sq0 = (unsigned)__cvta_generic_to_shared(sQ);
#else
    (void)mbar;
#endif
    for (int r = c.tid; r < nrow * Tr::NCOMP; r += 32) {
        const int cix = r / nrow, rr = r - cix * nrow;
        long long x;
        int slot;
        if (DIR == 0) {
            x = c.base + (long long)(rr - c.pp) * A.G.cs[1] + s0;
            slot = cix * Sh::NT + rr * Sh::C;
        } else {
            x = (c.base - c.pp) + (long long)(s0 + rr) * c.st;
            slot = cix * Sh::NT + rr * Sh::P;
        }
#if defined(__CUDA_ARCH__)
        bulk_g2s(sq0 + (unsigned)(slot * sizeof(double)), A.Q[cix] + x, (unsigned)(rowlen * 8), mbar);
#else
        for (int q = 0; q < rowlen; q++) sQ[slot + q] = A.Q[cix][x + q];
#endif
    }
}

/* end of iteration t: chunk t + 1 has landed (mbarrier phase), every thread takes its own cell and commits it */
template <class Tr, int DIR, int MATH>
HB2_HD void pipeline_consume(const DirArgs& A, double* smem, const PencilCtx& c, int t, unsigned mbar, unsigned& parity)
{
    int nrow, rowlen;
    bulk_rows<Tr, DIR, MATH>(c, A.G, t, nrow, rowlen);
    if (nrow <= 0 || rowlen <= 0) return;                /* block-uniform: nothing was issued */
#if defined(__CUDA_ARCH__)
    mbar_wait(mbar, parity);
#else
    (void)mbar;
#endif
    parity ^= 1u;
    int s;
    if (load_wanted<Tr, DIR, MATH>(c, t, s)) {
        double q[Tr::NCOMP];
        staged_cons<Tr, DIR, MATH>(smem, c, q);
        phase_commit<Tr, DIR, MATH>(A, smem, c, s, q);
    }
}

/* ---- face phase: midpoint fluxes of faces c0-6+tC+o --------------------------------------------- */
template <class Tr, int DIR, int MATH>
HB2_HD bool face_wanted(const PencilCtx& c, int t, int& f)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    f = c.c0 - 6 + t * Sh::C + c.o;
    return c.valid && f >= c.c0 - 1 && f <= c.c1 + 1;
}

/* the shock-sensor decision byte of the face's right cell (fetched one iteration ahead) */
template <class Tr, int DIR, int MATH, bool STEADY = false>
HB2_HD unsigned int face_flag_fetch(const DirArgs& A, const PencilCtx& c, int t)
{
    int f;
    const bool wanted = face_wanted<Tr, DIR, MATH>(c, t, f);
    if (!STEADY && !wanted) return 0;
    return A.hyb[c.base + (long long)f * c.st];
}

template <class Tr, int DIR, int MATH, bool STEADY = false>
HB2_HD void phase_face(const DirArgs& A, double* smem, const PencilCtx& c, int t, unsigned int flag)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    constexpr int NEQ = Tr::NEQ;
    int f;
    const bool wanted = face_wanted<Tr, DIR, MATH>(c, t, f);
    if (!STEADY && !wanted) return;
    const bool hybrid = (flag >> DIR) & 1;
    double* sM = smem + Sh::OFF_M;
    /* stencil window: cells f-3..f+2 at win[comp*CSV + m*MS] */
    const double* win = smem + Sh::slotv(c.pp, (f - 3) & (Sh::RING - 1));
    double Fm[NEQ], um;
    if constexpr (MATH == 0) {
        double V[6][NEQ];
#pragma unroll
        for (int m = 0; m < 6; m++)
#pragma unroll
            for (int e = 0; e < NEQ; e++) V[m][e] = win[e * Sh::CSV + m * Sh::MS];
        face_midpoint<Tr, DIR, 0>(V, win[Sh::IC * Sh::CSV + 2 * Sh::MS], win[Sh::IC * Sh::CSV + 3 * Sh::MS], hybrid, A.K, Fm, um);
    } else {
        face_midpoint_fast<Tr, DIR, Sh::CSV, Sh::MS>(win, hybrid, A.K, Fm, um);
    }
    const int sl = Sh::slot(c.pp, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) sM[e * Sh::CS + sl] = Fm[e];
    if (Tr::ADV) sM[NEQ * Sh::CS + sl] = um;
}

/* ---- update phase: cells c0-8+tC+o ---------------------------------------------------------------- */
template <class Tr>
struct UpdateIn {          /* global inputs of the update phase, fetched before the face phase of the same iteration */
    double R[Tr::NEQ];     /* running right-hand side of the previous directions (FUSED, DIR > 0) */
    double T;              /* running velocity-divergence sum (five-eqn, DIR > 0) */
};

template <class Tr, int DIR, int MATH>
HB2_HD bool update_wanted(const PencilCtx& c, int t, int& cc)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    cc = c.c0 - 8 + t * Sh::C + c.o;
    return t >= 0 && c.valid && cc >= c.c0 && cc < c.c1;
}

template <class Tr, int DIR>
HB2_HD long long update_index(const DirArgs& A, const PencilCtx& c, int cc)
{
    return c.ibase + (long long)cc * c.ist;
}

template <class Tr, int DIR, bool FUSED>
HB2_HD void update_fetch(const DirArgs& A, const PencilCtx& c, int cc, UpdateIn<Tr>& in)
{
    const long long ix = update_index<Tr, DIR>(A, c, cc);
    if (DIR > 0 && FUSED) {
#pragma unroll
        for (int e = 0; e < Tr::NEQ; e++) in.R[e] = load_stream(A.R[e] + ix);
    }
    in.T = (Tr::ADV && DIR > 0) ? A.T[ix] : 0.0;
}

HB2_HD void prefetch_l2(const void* p)
{
#if defined(__CUDA_ARCH__)
#if defined(HB2_PREFETCH_TO_L1)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
#else
    (void)p;
#endif
}

template <class Tr, int DIR, int NTERM>
HB2_HD void update_prefetch(const DirArgs& A, const PencilCtx& c, int cc)
{
    constexpr bool QREC = (NTERM >= HB2_NTERM_QREC);
    constexpr int NLOAD = QREC ? NTERM - HB2_NTERM_QREC : (NTERM > 0 ? NTERM : 0);
    if (DIR > 0) {
        const long long ix = c.ibase + (long long)cc * c.ist;
#pragma unroll
        for (int e = 0; e < Tr::NEQ; e++) prefetch_l2(A.R[e] + ix);
    }
    if (NLOAD > 0) {
        const long long x = c.base + (long long)cc * c.st;
#pragma unroll
        for (int k = 0; k < NLOAD; k++)
#pragma unroll
            for (int e = 0; e < Tr::NEQ; e++) prefetch_l2(A.Ut[k][e] + x);
    }
}

/* NTERM: number of states in the RK linear combination (compile-time, so that their loads are unconditional and are
 * all issued up front); 0 for the directions / modes without the RK update */
template <class Tr, int DIR, int MATH, int NTERM>
HB2_HD void phase_update(const DirArgs& A, const double* smem, const PencilCtx& c, int cc, const UpdateIn<Tr>& in)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    constexpr int DIM = Tr::DIM, NEQ = Tr::NEQ, NS = Tr::NS, IV = Tr::IV, IP = Tr::IP;
    constexpr bool LAST = (DIR == DIM - 1);
    const Geom& G = A.G;
    const double* sV = smem;
    const double* sM = smem + Sh::OFF_M;
    const int m_m1 = Sh::slot(c.pp, cc - 1), m_0 = Sh::slot(c.pp, cc), m_p1 = Sh::slot(c.pp, cc + 1),
              m_p2 = Sh::slot(c.pp, cc + 2);
    /* the same cells in the primitive-variable ring (primary copies) */
    const int v_m1 = Sh::slotv(c.pp, (cc - 1) & (Sh::RING - 1)), v_0 = Sh::slotv(c.pp, cc & (Sh::RING - 1)),
              v_p1 = Sh::slotv(c.pp, (cc + 1) & (Sh::RING - 1));
    constexpr bool fused = (NTERM != HB2_NTERM_EMIT);
    const double dxd = G.dx[DIR];
    const int ci = (DIR == 0) ? cc : c.i, cj = (DIR == 1) ? cc : c.j, ck = (DIR == 2) ? cc : c.k;
    const long long ix = c.ibase + (long long)cc * c.ist;

    /* last direction of a fused stage: the states of the RK linear combination (the states with alpha != 0, compacted
     * by the host).  All loads are issued here, before the shared-memory work below, and consumed at the very end. */
    constexpr bool QREC = (NTERM >= HB2_NTERM_QREC);
    constexpr int NLOAD = QREC ? NTERM - HB2_NTERM_QREC : NTERM;
    double ut[NLOAD > 0 ? NLOAD : 1][NEQ];
    if (NLOAD > 0) {
        const long long x = c.base + (long long)cc * c.st;
#pragma unroll
        for (int k = 0; k < NLOAD; k++)
#pragma unroll
            for (int e = 0; e < NEQ; e++) ut[k][e] = load_stream(A.Ut[k][e] + x);
    }

    /* velocity-divergence contribution of this direction (advective equations of the five-eqn model) */
    double Tsum = 0.0;
    if (Tr::ADV) {
        const double* um = sM + NEQ * Sh::CS;
        const double* un = sV + (IV + DIR) * Sh::CSV;
        const double Td = (3.0 / 2.0 * (um[m_p1] - um[m_0]) - 3.0 / 10.0 * (un[v_p1] - un[v_m1]) +
                           1.0 / 30.0 * (um[m_p2] - um[m_m1])) / dxd;
        Tsum = (DIR == 0) ? Td : in.T + Td;
        if (!LAST) A.T[ix] = Tsum;
    }

    double rhs[NEQ];
    if constexpr (MATH == 1) {
        /* node fluxes from the primitive ring */
        double Np[NEQ], Nm[NEQ];
        node_flux_prim<Tr, DIR, Sh::CSV>(sV + v_p1, sV[(Sh::NV - 1) * Sh::CSV + v_p1], A.K, Np);
        node_flux_prim<Tr, DIR, Sh::CSV>(sV + v_m1, sV[(Sh::NV - 1) * Sh::CSV + v_m1], A.K, Nm);
        if (fused) {
            /* difference form: F[c+1] - F[c] = dt (3/2 (M[c+1]-M[c]) + 1/30 (M[c+2]-M[c-1]) - 3/10 (N[c+1]-N[c-1])) */
            const double k1 = A.kf[DIR][0], k2 = A.kf[DIR][1], k3 = A.kf[DIR][2];
#pragma unroll
            for (int e = 0; e < NEQ; e++) {
                const double* M = sM + e * Sh::CS;
                const double r0 = (DIR == 0) ? 0.0 : in.R[e];
                rhs[e] = fma(k3, Np[e] - Nm[e], fma(-k2, M[m_p2] - M[m_m1], fma(-k1, M[m_p1] - M[m_0], r0)));
            }
        } else {
            double N0[NEQ];
            node_flux_prim<Tr, DIR, Sh::CSV>(sV + v_0, sV[(Sh::NV - 1) * Sh::CSV + v_0], A.K, N0);
#pragma unroll
            for (int e = 0; e < NEQ; e++) {
                const double* M = sM + e * Sh::CS;
                A.F[e][sidx<DIR>(G, ci, cj, ck)] =
                    A.dt * (1.0 / 30.0 * (M[m_p1] + M[m_m1]) - 3.0 / 10.0 * (N0[e] + Nm[e]) + 23.0 / 15.0 * M[m_0]);
                if (cc + 1 == G.n[DIR])
                    A.F[e][sidx<DIR>(G, ci + (DIR == 0), cj + (DIR == 1), ck + (DIR == 2))] =
                        A.dt * (1.0 / 30.0 * (M[m_p2] + M[m_0]) - 3.0 / 10.0 * (Np[e] + N0[e]) + 23.0 / 15.0 * M[m_p1]);
            }
        }
    } else {
        const double* sN = smem + Sh::OFF_N;
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
                rhs[e] = (DIR == 0) ? -dF : in.R[e] - dF;
            }
        }
    }

    if (fused) {
        if (LAST) {
            if (Tr::ADV) {
#pragma unroll
                for (int si = 0; si < Tr::NZ; si++) {
                    const int e = IP + 1 + si;
                    rhs[e] = rhs[e] + A.dt * sV[e * Sh::CSV + v_0] * Tsum;
                }
            }
            /* Euler.cpp:1479, 1544-1548: zero; += alpha_n*U_n for alpha_n != 0; += beta*(...) */
            double ua[NEQ];
            double qc[NEQ];
            if constexpr (QREC) prim_to_cons<Tr, Sh::CSV>(sV + v_0, sV[(Sh::NV - 1) * Sh::CSV + v_0], A.K, qc);
#pragma unroll
            for (int e = 0; e < NEQ; e++) {
                double u = 0.0;
#pragma unroll
                for (int k = 0; k < NLOAD; k++) u += A.alpha_t[k] * ut[k][e];
                /* the flux state is the newest one: last term of the reference's sum */
                if (QREC) u = fma(A.alpha_q, qc[e], u);
                ua[e] = u;
            }
            rk_update_cell<Tr>(A, reinterpret_cast<double* const*>(smem + Sh::OFF_P), c.base + (long long)cc * c.st, ci, cj,
                               ck, ua, rhs);
        } else {
#pragma unroll
            for (int e = 0; e < NEQ; e++) store_stream(A.R[e] + ix, rhs[e]);
        }
    } else if (LAST && Tr::ADV) {
#pragma unroll
        for (int si = 0; si < Tr::NZ; si++) {
            const int e = IP + 1 + si;
            A.S[e][ix] += A.dt * sV[e * Sh::CSV + v_0] * Tsum;
        }
    }
}

/* ---- one iteration of the marching pipeline (no barrier inside) -------------------------------------
 * iteration t, t = 0..nsteps:  commit chunk t+1 (loaded during iteration t-1), fetch chunk t+2 and the update inputs,
 * face phase t, update phase t-1.  Everything an iteration reads from the rings was written in EARLIER iterations,
 * and nothing it writes is read by another thread in the same iteration, so ONE barrier per iteration suffices. */
template <class Tr>
struct PipeRegs {
    int s;
    int have;
    unsigned int flag;    /* sensor byte of the NEXT face phase (32-bit: a byte would be packed with `have`, which
                             makes the pack instruction wait for the load) */
    unsigned int parity;  /* bulk staging: phase parity of the mbarrier */
    unsigned int mbar;    /* bulk staging, device: shared-window address of the mbarrier */
};

template <class Tr, int DIR, int MATH>
HB2_HD void pipeline_prologue(const DirArgs& A, double* smem, const PencilCtx& c, PipeRegs<Tr>& pr)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    if (DIR == Tr::DIM - 1 && A.push) {
        double** sP = reinterpret_cast<double**>(smem + Sh::OFF_P);
        for (int n = c.tid; n < 27 * Tr::NCOMP; n += Sh::NT) sP[n] = A.push[n];
    }
    int s;
    if (load_wanted<Tr, DIR, MATH>(c, 0, s)) {
        double q[Tr::NCOMP];
        load_cons<Tr>(A, c.base + (long long)s * c.st, q);
        phase_commit<Tr, DIR, MATH>(A, smem, c, s, q);
    }
    pr.parity = 0u;
    if (A.bulk) {
        pr.have = 0;                     /* chunk 1 is issued at the top of iteration 0 (pipeline_issue) */
    } else {
        pr.have = load_wanted<Tr, DIR, MATH>(c, 1, pr.s);
        if (pr.have) stage_cons<Tr, DIR, MATH>(A, smem, c, c.base + (long long)pr.s * c.st);
    }
    pr.flag = face_flag_fetch<Tr, DIR, MATH>(A, c, 0);
}

/* STEADY iterations: the block has all its pencils and iteration t lies in steady_range() -- every thread loads, commits,
 * computes a face and updates a cell, so the per-phase "is it wanted" tests (and the basic-block boundaries they put between
 * the phases) are compiled out.  Same arithmetic on the same data: the variant only removes tests that are known to pass. */
template <class Tr, int DIR, int MATH>
HB2_HD void steady_range(const PencilCtx& c, int& t_lo, int& t_hi)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const int len = c.c1 - c.c0;
    t_lo = 1 + (8 + Sh::C - 1) / Sh::C;          /* update(t - 1) starts at cell c0: (t - 1) C >= 8 */
    t_hi = (len + 8) / Sh::C - 3;                /* load(t + 2) still inside the pencil for every position of the chunk:
                                                    (t + 3) C <= len + 8; implies face(t), face(t + 1), update(t - 1), update(t) */
}

template <class Tr, int DIR, int MATH, int NTERM, bool STEADY = false>
HB2_HD void pipeline_iteration(const DirArgs& A, double* smem, const PencilCtx& c, int t, int nsteps, PipeRegs<Tr>& pr)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    if (STEADY || !A.bulk) {
        if (STEADY || pr.have) {
            double q[Tr::NCOMP];
            stage_wait_all();
            staged_cons<Tr, DIR, MATH>(smem, c, q);
            phase_commit<Tr, DIR, MATH>(A, smem, c, pr.s, q);
        }
        const bool have = load_wanted<Tr, DIR, MATH>(c, t + 2, pr.s);
        pr.have = STEADY ? 1 : have;
        if (pr.have) stage_cons<Tr, DIR, MATH>(A, smem, c, c.base + (long long)pr.s * c.st);
    }
    int cc;
    const bool upd = update_wanted<Tr, DIR, MATH>(c, t - 1, cc);
    const bool do_update = STEADY ? true : upd;
    UpdateIn<Tr> uin;
    constexpr bool FUSED = (NTERM != HB2_NTERM_EMIT);
    /* The HBM inputs of the update phase (running right-hand side, RK states) are loaded where they are consumed: held
     * in registers across the face phase they spill, staged with cp.async they load the shared-memory pipe.  Their DRAM
     * latency is taken out one iteration ahead with L2 prefetches instead (no register, no shared memory): HB2_PREFETCH_L2. */
    if (HB2_PREFETCH_L2 && FUSED) {
        int cn;
        const bool pre = update_wanted<Tr, DIR, MATH>(c, t + HB2_PREFETCH_L2 - 1, cn);
        if (STEADY || pre) update_prefetch<Tr, DIR, NTERM>(A, c, cn);
    }
    constexpr bool PREFETCH_R = (DIR == 0) ? (HB2_PREFETCH_R != 0) : (HB2_PREFETCH_R_YZ != 0);
    if (PREFETCH_R && do_update) update_fetch<Tr, DIR, FUSED>(A, c, cc, uin);
    unsigned int flag;
    if (HB2_PREFETCH_FLAG) {
        flag = pr.flag;
        pr.flag = (STEADY || t + 1 < nsteps) ? face_flag_fetch<Tr, DIR, MATH, STEADY>(A, c, t + 1) : 0u;
    } else {
        flag = (STEADY || t < nsteps) ? face_flag_fetch<Tr, DIR, MATH, STEADY>(A, c, t) : 0u;
    }
    if (STEADY || t < nsteps) phase_face<Tr, DIR, MATH, STEADY>(A, smem, c, t, flag);
    if (!PREFETCH_R && do_update) update_fetch<Tr, DIR, FUSED>(A, c, cc, uin);
    if (do_update) phase_update<Tr, DIR, MATH, NTERM>(A, smem, c, cc, uin);
    if (!STEADY && A.bulk) pipeline_consume<Tr, DIR, MATH>(A, smem, c, t + 1, pr.mbar, pr.parity);
}

/* SKEWED steady iteration: the phases of one iteration only read what EARLIER iterations wrote to the rings, so a warp may run
 * them in any order.  Odd warps run [commit, stage, update] before [face], even warps the other way round: at any time half
 * the warps of a block are in the FP64-dense face phase and half in the memory / integer-heavy ones, instead of all sixteen
 * warps of an SM hitting the FP64 pipe together and leaving it idle together.  The two parts exist once in the code; a
 * two-trip loop with a warp-uniform test picks the order. */
#ifndef HB2_SKEW
#define HB2_SKEW 0
#endif
template <class Tr, int DIR, int MATH, int NTERM>
HB2_HD void pipeline_iteration_skewed(const DirArgs& A, double* smem, const PencilCtx& c, int t, int nsteps, PipeRegs<Tr>& pr)
{
    constexpr bool FUSED = (NTERM != HB2_NTERM_EMIT);
    const bool face_first = ((c.tid >> 5) & 1) == 0;
#pragma unroll 1
    for (int part = 0; part < 2; part++) {
        if ((part == 0) == face_first) {
            const unsigned int flag = pr.flag;
            pr.flag = face_flag_fetch<Tr, DIR, MATH, true>(A, c, t + 1);
            phase_face<Tr, DIR, MATH, true>(A, smem, c, t, flag);
        } else {
            double q[Tr::NCOMP];
            stage_wait_all();
            staged_cons<Tr, DIR, MATH>(smem, c, q);
            phase_commit<Tr, DIR, MATH>(A, smem, c, pr.s, q);
            load_wanted<Tr, DIR, MATH>(c, t + 2, pr.s);
            pr.have = 1;
            stage_cons<Tr, DIR, MATH>(A, smem, c, c.base + (long long)pr.s * c.st);
            int cc, cn;
            update_wanted<Tr, DIR, MATH>(c, t - 1, cc);
            if (HB2_PREFETCH_L2 && FUSED) {
                update_wanted<Tr, DIR, MATH>(c, t + HB2_PREFETCH_L2 - 1, cn);
                update_prefetch<Tr, DIR, NTERM>(A, c, cn);
            }
            UpdateIn<Tr> uin;
            update_fetch<Tr, DIR, FUSED>(A, c, cc, uin);
            phase_update<Tr, DIR, MATH, NTERM>(A, smem, c, cc, uin);
        }
    }
    (void)nsteps;
}

/* one iteration, steady or general (block-uniform choice): what the kernel and the host emulation both call */
template <class Tr, int DIR, int MATH, int NTERM>
HB2_HD void pipeline_step(const DirArgs& A, double* smem, const PencilCtx& c, int t, int nsteps, int t_lo, int t_hi, PipeRegs<Tr>& pr)
{
#if HB2_STEADY
    if (c.full && !A.bulk && t >= t_lo && t <= t_hi) {
        if (HB2_SKEW)
            pipeline_iteration_skewed<Tr, DIR, MATH, NTERM>(A, smem, c, t, nsteps, pr);
        else
            pipeline_iteration<Tr, DIR, MATH, NTERM, true>(A, smem, c, t, nsteps, pr);
    } else
#endif
        pipeline_iteration<Tr, DIR, MATH, NTERM, false>(A, smem, c, t, nsteps, pr);
}

}  // namespace hb2
