/*
 * emu.cpp -- TEST-ONLY host emulation of the CUDA sweeps (never part of the product library).
 *
 * Compiles hamers_b200/csrc/hb2_core.cuh with g++ and drives the very same per-thread functions
 * the kernels call (pencil_ctx, pipeline_prologue, pipeline_iteration, sensor_phase_*) from plain loops that stand in
 * for the CUDA grid.  It exists because this container has no GPU: it lets the CPU test suite
 * check the kernels' indexing and arithmetic against the oracle before any GPU time is spent.
 * The GPU parity tests (pytest -m gpu) remain the parity tests proper.
 */
#include "../../hamers_b200/csrc/hb2_sweep.cuh"
#include "../../hamers_b200/csrc/hb2_sensor.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace hb2;

struct EmuDesc {
    int dim, n[3], model, ns;
    double gamma[4], dx[3];
    int weno_p, math, bx, seg_len;
    int weno_q;            /* WCNS6-LD constants (the interpolator itself is compiled in: -DHB2_SCHEME) */
    double weno_C, weno_alpha_tau;
    int ghosts;            /* ghost width of the cell-data layout (0 = 4); the kernels read 4 layers whatever it is */
    double R[4];           /* four-eqn conservative model: species gas constants */
};

static int emu_neq(const EmuDesc* d) { return d->model == SS ? d->dim + 2 : (d->model == FC ? d->dim + 1 + d->ns : d->dim + 2 * d->ns); }
static int emu_ncomp(const EmuDesc* d) { return d->model == FE ? emu_neq(d) + 1 : emu_neq(d); }

static void make_geom(const EmuDesc* d, Geom* G)
{
    G->dim = d->dim;
    for (int a = 0; a < 3; a++) {
        G->n[a] = (a < d->dim) ? d->n[a] : 1;
        G->g[a] = (a < d->dim) ? (d->ghosts > 0 ? d->ghosts : HB2_G) : 0;
        G->gd[a] = G->n[a] + 2 * G->g[a];
        G->dx[a] = (a < d->dim) ? d->dx[a] : 1.0;
    }
    G->cs[0] = 1;
    G->cs[1] = G->gd[0];
    G->cs[2] = (long long)G->gd[0] * G->gd[1];
    G->ncell_g = (long long)G->gd[0] * G->gd[1] * G->gd[2];
}

/* one direction: the loops below stand in for the CUDA grid, the barrier becomes the boundary between iterations.
 * Within an iteration the threads run one after the other -- in forward order for even blocks and in REVERSE order for
 * odd blocks -- so that a write that another thread still reads in the same iteration (a ring hazard) changes results. */
template <class Tr, int DIR, int MATH, int NTERM>
static void run_dir_n(const DirArgs& A)
{
    using Sh = SweepShape<Tr, DIR, MATH>;
    const Geom& G = A.G;
    const int nseg = (G.n[DIR] + A.seg_len - 1) / A.seg_len;
    int gx, gy;
    if (DIR == 0) {
        gx = (G.n[1] + Sh::P - 1) / Sh::P;
        gy = (Tr::DIM == 3) ? G.n[2] : 1;
    } else if (DIR == 1) {
        gx = (G.n[0] + Sh::P - 1) / Sh::P;
        gy = (Tr::DIM == 3) ? G.n[2] : 1;
    } else {
        gx = (G.n[0] + Sh::P - 1) / Sh::P;
        gy = G.n[1];
    }
    std::vector<double> smem(Sh::SMEM_DOUBLES);
    std::vector<PipeRegs<Tr>> regs(Sh::NT);
    int nblock = 0;
    for (int bz = 0; bz < nseg; bz++)
        for (int by = 0; by < gy; by++)
            for (int bx = 0; bx < gx; bx++, nblock++) {
                const BlockId b = {bx, by, bz};
                /* poison the rings: reading a slot that was never written must show up */
                for (auto& v : smem) v = std::nan("");
                const PencilCtx c0 = pencil_ctx<Tr, DIR, MATH>(A, b, 0);
                const int nsteps = Sh::nsteps(c0.c1 - c0.c0);
                int t_lo, t_hi;
                steady_range<Tr, DIR, MATH>(c0, t_lo, t_hi);
                const bool rev = (nblock & 1);
                for (int n = 0; n < Sh::NT; n++) {
                    const int tid = rev ? Sh::NT - 1 - n : n;
                    pipeline_prologue<Tr, DIR, MATH>(A, smem.data(), pencil_ctx<Tr, DIR, MATH>(A, b, tid), regs[tid]);
                }
                for (int t = 0; t <= nsteps; t++) {
                    /* bulk staging: warp 0 issues the copy of chunk t + 1 at the top of the iteration; the mbarrier wait of
                     * the consumers orders it before their reads, whatever the thread order */
                    if (A.bulk)
                        for (int n = 0; n < Sh::NT; n++) {
                            const int tid = rev ? Sh::NT - 1 - n : n;
                            pipeline_issue<Tr, DIR, MATH>(A, smem.data(), pencil_ctx<Tr, DIR, MATH>(A, b, tid), t + 1, 0u);
                        }
                    for (int n = 0; n < Sh::NT; n++) {
                        const int tid = rev ? Sh::NT - 1 - n : n;
                        pipeline_step<Tr, DIR, MATH, NTERM>(A, smem.data(), pencil_ctx<Tr, DIR, MATH>(A, b, tid), t, nsteps, t_lo, t_hi, regs[tid]);
                    }
                }
            }
}

template <class Tr, int DIR, int MATH>
static void run_dir(const DirArgs& A)
{
    if (A.mode == MODE_EMIT) return run_dir_n<Tr, DIR, MATH, HB2_NTERM_EMIT>(A);
    if (DIR == Tr::DIM - 1) {
        if (A.nterm == 1) return run_dir_n<Tr, DIR, MATH, 1>(A);
        if (A.nterm == 2) return run_dir_n<Tr, DIR, MATH, 2>(A);
        if (MATH == 1 && A.nterm >= HB2_NTERM_QREC) {
            constexpr int Q0 = (MATH == 1) ? HB2_NTERM_QREC : 1;
            if (A.nterm == HB2_NTERM_QREC) return run_dir_n<Tr, DIR, MATH, Q0>(A);
            if (A.nterm == HB2_NTERM_QREC + 1) return run_dir_n<Tr, DIR, MATH, Q0 + 1>(A);
            return run_dir_n<Tr, DIR, MATH, Q0 + 2>(A);
        }
        return run_dir_n<Tr, DIR, MATH, 3>(A);
    }
    run_dir_n<Tr, DIR, MATH, 0>(A);
}

template <class Tr, int MATH>
static void run_sweeps(DirArgs A0, int sensor_seg_len, int seg_len, double* const* F_all, int mode)
{
    const Geom& G = A0.G;
    /* sensor decisions on cells -1..N+1: the phase functions of k_sensor, the loops stand in for grid and barriers
     * (threads in reverse order in odd blocks; rings poisoned) */
    std::vector<unsigned char> hyb(G.ncell_g, 0);
    {
        using Sh = SensorShape<Tr>;
        SensorArgs S;
        memset(&S, 0, sizeof(S));
        S.G = G;
        for (int c = 0; c < Tr::NCOMP; c++) S.Q[c] = A0.Q[c];
        S.hyb = hyb.data();
        S.seg_len = sensor_seg_len > 0 ? sensor_seg_len : G.n[2] + 3;
        std::vector<double> sm(Sh::SMEM_DOUBLES);
        int nblock = 0;
        for (int bz = 0; bz < Sh::segments(G, S.seg_len); bz++)
            for (int by = 0; by < Sh::tiles_y(G); by++)
                for (int bx = 0; bx < Sh::tiles_x(G); bx++, nblock++) {
                    for (auto& v : sm) v = std::nan("");
                    const SensorTile T = sensor_tile<Tr>(S, bx, by, bz);
                    const bool rev = (nblock & 1);
                    auto each = [&](auto fn) {
                        for (int n = 0; n < Sh::NT; n++) fn(rev ? Sh::NT - 1 - n : n);
                    };
                    std::vector<SensorRegs<Tr>> R(Sh::NT);
                    if (Tr::DIM == 3) {
                        for (int t = T.kb - 2; t < T.kb; t++)
                            each([&](int tid) {
                                sensor_phase_fetch<Tr>(S, T, tid, t, R[tid]);
                                sensor_phase_velocity<Tr, MATH>(sm.data(), tid, t, R[tid]);
                            });
                        each([&](int tid) { sensor_phase_fetch<Tr>(S, T, tid, T.kb, R[tid]); });
                        for (int t = T.kb; t <= T.ke + 1; t++) {
                            each([&](int tid) {
                                if (t <= T.ke) sensor_phase_velocity<Tr, MATH>(sm.data(), tid, t, R[tid]);
                                if (t < T.ke) sensor_phase_fetch<Tr>(S, T, tid, t + 1, R[tid]);
                            });
                            /* the one barrier */
                            each([&](int tid) {
                                if (t <= T.ke) sensor_phase_gradient<Tr, MATH>(S, sm.data(), T, tid, t - 1);
                                if (t >= T.kb + 2) sensor_phase_decision<Tr, MATH>(S, sm.data(), T, tid, t - 2);
                            });
                        }
                    } else {
                        each([&](int tid) {
                            sensor_phase_fetch<Tr>(S, T, tid, 0, R[tid]);
                            sensor_phase_velocity<Tr, MATH>(sm.data(), tid, 0, R[tid]);
                        });
                        each([&](int tid) { sensor_phase_gradient<Tr, MATH>(S, sm.data(), T, tid, 0); });
                        each([&](int tid) { sensor_phase_decision<Tr, MATH>(S, sm.data(), T, tid, 0); });
                    }
                }
    }
    A0.hyb = hyb.data();
    A0.mode = mode;

    for (int dir = 0; dir < Tr::DIM; dir++) {
        DirArgs A = A0;
        if (mode == MODE_EMIT)
            for (int e = 0; e < Tr::NEQ; e++) A.F[e] = F_all[dir * Tr::NEQ + e];
        if (dir != Tr::DIM - 1) A.ncoef = 0;
        A.seg_len = seg_len > 0 ? seg_len : G.n[dir];
        {
            /* the library's condition for the bulk-copy staging (hb2_abi.cu: bulk_ok) */
            const char* v = getenv("HB2_BULK_STAGE");
            const int x_seg = seg_len > 0 ? seg_len : G.n[0];
            A.bulk = (v && atoi(v) != 0) && (G.n[0] % 2 == 0) && (G.g[0] % 2 == 0) && (x_seg % 2 == 0 || x_seg >= G.n[0]);
        }
        if (dir == 0)
            run_dir<Tr, 0, MATH>(A);
        else if (dir == 1)
            run_dir<Tr, 1, MATH>(A);
        else
            run_dir<Tr, (Tr::DIM == 3 ? 2 : 1), MATH>(A);
    }
}

#define EMU_DISPATCH(d, CALL)                                                                     \
    do {                                                                                          \
        if ((d)->model == SS && (d)->dim == 2) { using Tr = Traits<SS, 2, 1>; if ((d)->math == 0) { constexpr int MATH = 0; CALL; } else { constexpr int MATH = 1; CALL; } return 0; } \
        if ((d)->model == SS && (d)->dim == 3) { using Tr = Traits<SS, 3, 1>; if ((d)->math == 0) { constexpr int MATH = 0; CALL; } else { constexpr int MATH = 1; CALL; } return 0; } \
        if ((d)->model == FE && (d)->dim == 2 && (d)->ns == 3) { using Tr = Traits<FE, 2, 3>; constexpr int MATH = 0; CALL; return 0; } \
        if ((d)->model == FE && (d)->dim == 3 && (d)->ns == 3) { using Tr = Traits<FE, 3, 3>; constexpr int MATH = 0; CALL; return 0; } \
        if ((d)->model == FC && (d)->dim == 2 && (d)->ns == 3) { using Tr = Traits<FC, 2, 3>; constexpr int MATH = 0; CALL; return 0; } \
        if ((d)->model == FC && (d)->dim == 3 && (d)->ns == 3) { using Tr = Traits<FC, 3, 3>; constexpr int MATH = 0; CALL; return 0; } \
        if ((d)->model == FE && (d)->dim == 2) { using Tr = Traits<FE, 2, 2>; if ((d)->math == 0) { constexpr int MATH = 0; CALL; } else { constexpr int MATH = 1; CALL; } return 0; } \
        if ((d)->model == FE && (d)->dim == 3) { using Tr = Traits<FE, 3, 2>; if ((d)->math == 0) { constexpr int MATH = 0; CALL; } else { constexpr int MATH = 1; CALL; } return 0; } \
        /* four-eqn conservative: reference-order kernels only, like the product library */ \
        if ((d)->model == FC && (d)->dim == 2) { using Tr = Traits<FC, 2, 2>; constexpr int MATH = 0; CALL; return 0; } \
        if ((d)->model == FC && (d)->dim == 3) { using Tr = Traits<FC, 3, 2>; constexpr int MATH = 0; CALL; return 0; } \
    } while (0)

static void fill_common(const EmuDesc* d, DirArgs* A, const double* const* Q, double dt, std::vector<double>& T)
{
    memset(A, 0, sizeof(*A));
    make_geom(d, &A->G);
    for (int s = 0; s < 4; s++) {
        A->K.gamma[s] = (s < d->ns) ? d->gamma[s] : 1.4;
        A->K.inv_gm1[s] = 1.0 / (A->K.gamma[s] - 1.0);
        const double R = (d->model == FC && s < d->ns) ? d->R[s] : 1.0;
        A->K.cp[s] = A->K.gamma[s] / (A->K.gamma[s] - 1.0) * R;
        A->K.cv[s] = 1.0 / (A->K.gamma[s] - 1.0) * R;
    }
    A->K.weno_p = d->weno_p > 0 ? d->weno_p : 2;
    A->K.weno_q = d->weno_q > 0 ? d->weno_q : 4;
    A->K.weno_C = d->weno_C > 0.0 ? d->weno_C : 1.0e9;
    A->K.weno_alpha_tau = d->weno_alpha_tau > 0.0 ? d->weno_alpha_tau : 35.0;
    const int ncomp = emu_ncomp(d);
    for (int c = 0; c < ncomp; c++) A->Q[c] = Q[c];
    dir_args_set_dt(A, dt);
    T.assign((size_t)A->G.n[0] * A->G.n[1] * A->G.n[2], 0.0);
    A->T = T.data();
}

extern "C" int emu_flux_and_source(const EmuDesc* d, const double* const* Q, double dt, double* const* F, double* const* S)
{
    DirArgs A;
    std::vector<double> T;
    fill_common(d, &A, Q, dt, T);
    const int neq = emu_neq(d);
    for (int e = 0; e < neq; e++) A.S[e] = S ? S[e] : nullptr;
    EMU_DISPATCH(d, (run_sweeps<Tr, MATH>(A, d->bx, d->seg_len, F, MODE_EMIT)));
    return -1;
}

extern "C" int emu_fused_stage_push(const EmuDesc* d, int ncoef, const double* alpha, const double* beta,
                                    const double* const* U_int, double dt, double* const* U_out, int push);

extern "C" int emu_fused_stage(const EmuDesc* d, int ncoef, const double* alpha, const double* beta,
                               const double* const* U_int, double dt, double* const* U_out)
{
    return emu_fused_stage_push(d, ncoef, alpha, beta, U_int, dt, U_out, 0);
}

/* push != 0: the patch is its own periodic neighbour in every direction (self-push table, like a one-box level) */
extern "C" int emu_fused_stage_push(const EmuDesc* d, int ncoef, const double* alpha, const double* beta,
                                    const double* const* U_int, double dt, double* const* U_out, int push)
{
    const int ncomp = emu_ncomp(d);
    const int neq = emu_neq(d);
    DirArgs A;
    std::vector<double> T;
    fill_common(d, &A, U_int + (size_t)(ncoef - 1) * ncomp, dt, T);
    std::vector<std::vector<double>> R(neq, std::vector<double>(T.size(), 0.0));
    for (int e = 0; e < neq; e++) A.R[e] = R[e].data();
    A.ncoef = ncoef;
    for (int m = 0; m < ncoef; m++) {
        A.alpha[m] = alpha[m];
        for (int c = 0; c < ncomp; c++) A.Uint[m][c] = U_int[m * ncomp + c];
    }
    A.beta = beta[ncoef - 1];
    A.nterm = 0;
    /* fast build: like hb2_fused_stage_dev, the flux state is rebuilt from the primitive ring */
    const bool qrec = (d->math == 1);
    A.alpha_q = alpha[ncoef - 1];
    for (int m = 0; m < (qrec ? ncoef - 1 : ncoef); m++)
        if (alpha[m] != 0.0) {
            A.alpha_t[A.nterm] = alpha[m];
            for (int c = 0; c < ncomp; c++) A.Ut[A.nterm][c] = U_int[m * ncomp + c];
            A.nterm++;
        }
    if (qrec) A.nterm += HB2_NTERM_QREC;
    for (int c = 0; c < ncomp; c++) A.Uout[c] = U_out[c];
    std::vector<double*> table(27 * ncomp, nullptr);
    if (push) {
        for (int oz = (d->dim == 3 ? -1 : 0); oz <= (d->dim == 3 ? 1 : 0); oz++)
            for (int oy = -1; oy <= 1; oy++)
                for (int ox = -1; ox <= 1; ox++) {
                    if (!(ox | oy | oz)) continue;
                    const int code = (ox + 1) + 3 * (oy + 1) + 9 * (oz + 1);
                    for (int c = 0; c < ncomp; c++) table[code * ncomp + c] = U_out[c];
                }
        A.push = table.data();
    }
    EMU_DISPATCH(d, (run_sweeps<Tr, MATH>(A, d->bx, d->seg_len, nullptr, MODE_FUSED)));
    return -1;
}
