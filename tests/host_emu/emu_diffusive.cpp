/*
 * emu_diffusive.cpp -- TEST-ONLY host emulation of the diffusive-flux kernels (never part of the product library).
 *
 * Compiles hamers_b200/csrc/hb2_diffusive.cuh with g++ and calls the very per-thread functions the kernels of
 * hb2_diffusive.cu call (diff_primitives_thread, diff_node_all_thread, diff_face_thread, advance_ns_thread, ...) from plain loops
 * that stand in for the grid-stride loops.  Lets the CPU test suite check indexing and arithmetic against the oracle in a
 * container without a GPU; the GPU parity tests (pytest -m gpu) remain the parity tests proper.
 */
#include "../../hamers_b200/csrc/hb2_diffusive.cuh"
#include <cmath>
#include <limits>
#include <vector>

using namespace hb2;

struct EmuDiffDesc {
    int dim, n[3];
    double dx[3];
    double gamma, c_v, mu, mu_v, c_p, Pr;
};

template <int DIM>
struct NodeStage {
    DiffGeom G;
    DiffConsts K;
    std::vector<std::vector<double>> P, Fn;
    /* scratch starts as NaN: a node or primitive the kernels read without having written it shows up in the output */
    NodeStage(const EmuDiffDesc* d, const double* const* Q)
    {
        make_diff_geom(d->dim, d->n, d->dx, HB2_DIFF_G, &G);
        K = DiffConsts{d->gamma, d->c_v, d->mu, d->mu_v, d->c_p * d->mu / d->Pr};
        const double nan = std::numeric_limits<double>::quiet_NaN();
        P.assign(DIM + 1, std::vector<double>((size_t)G.ncell_g, nan));
        Fn.assign(3 * (DIM + 2), std::vector<double>((size_t)G.ncell_g, nan));
        DiffPtrs A{};
        for (int c = 0; c < DIM + 2; c++) A.Q[c] = Q[c];
        for (int v = 0; v < DIM + 1; v++) A.P[v] = P[v].data();
        for (long long x = 0; x < G.ncell_g; x++) diff_primitives_thread<DIM>(K, A, x);
        DiffAllPtrs N{};
        for (int v = 0; v < DIM + 1; v++) N.P[v] = P[v].data();
        for (int f = 0; f < DIM; f++)
            for (int e = 0; e < DIM + 2; e++) N.Fn[f][e] = Fn[f * (DIM + 2) + e].data();
        for (long long t = 0; t < diff_node_all_count<DIM>(G); t++) diff_node_all_thread<DIM>(G, K, N, t);
    }
};

template <int DIM, int FDIR>
static void run_faces(NodeStage<DIM>& S, double dt, double* const* F)
{
    DiffPtrs A{};
    for (int e = 0; e < DIM + 2; e++) {
        A.Fn[e] = S.Fn[FDIR * (DIM + 2) + e].data();
        A.F[e] = F[FDIR * (DIM + 2) + e];
    }
    for (long long t = 0; t < diff_face_count<DIM, FDIR>(S.G); t++) diff_face_thread<DIM, FDIR>(S.G, A, dt, t);
}

template <int DIM>
static int run(const EmuDiffDesc* d, const double* const* Q, double dt, double* const* F)
{
    NodeStage<DIM> S(d, Q);
    run_faces<DIM, 0>(S, dt, F);
    run_faces<DIM, 1>(S, dt, F);
    if (DIM == 3) run_faces<DIM, (DIM == 3 ? 2 : 1)>(S, dt, F);
    return 0;
}

extern "C" int emu_diffusive_flux(const EmuDiffDesc* d, const double* const* Q, double dt, double* const* F)
{
    return d->dim == 2 ? run<2>(d, Q, dt, F) : run<3>(d, Q, dt, F);
}

extern "C" int emu_advance_stage_ns(const EmuDiffDesc* d, int g, int ncoef, const double* alpha, const double* beta,
                                    const double* const* U_int, const double* const* Fc_int, const double* const* Fd_int,
                                    const double* const* S_int, double* const* U_out)
{
    const int dim = d->dim, neq = dim + 2;
    NsArgs A{};
    make_diff_geom(dim, d->n, d->dx, g, &A.G);
    A.neq = neq;
    A.ncoef = ncoef;
    for (int m = 0; m < ncoef; m++) {
        A.alpha[m] = alpha[m];
        A.beta[m] = beta[m];
        for (int e = 0; e < neq; e++) {
            A.U[m][e] = U_int[m * neq + e];
            A.S[m][e] = S_int[m * neq + e];
        }
        for (int f = 0; f < dim * neq; f++) {
            A.Fc[m][f] = Fc_int[m * dim * neq + f];
            A.Fd[m][f] = Fd_int[m * dim * neq + f];
        }
    }
    for (int e = 0; e < neq; e++) A.Uout[e] = U_out[e];
    const long long total = (long long)A.G.n[0] * A.G.n[1] * A.G.n[2];
    for (long long t = 0; t < total; t++) {
        if (dim == 2)
            advance_ns_thread<2>(A, t);
        else
            advance_ns_thread<3>(A, t);
    }
    return 0;
}

extern "C" int emu_diff_fill_periodic(const EmuDiffDesc* d, double* const* U, int mask)
{
    DiffGeom G;
    make_diff_geom(d->dim, d->n, d->dx, HB2_DIFF_G, &G);
    DiffStatePtrs A{};
    for (int c = 0; c < d->dim + 2; c++) A.U[c] = U[c];
    for (long long t = 0; t < G.ncell_g; t++) diff_fill_periodic_thread(G, A, d->dim + 2, mask, t);
    return 0;
}

extern "C" int emu_diff_extract_view(const EmuDiffDesc* d, const double* const* U, int g, double* const* V)
{
    DiffGeom Gs, Gd;
    make_diff_geom(d->dim, d->n, d->dx, HB2_DIFF_G, &Gs);
    make_diff_geom(d->dim, d->n, d->dx, g, &Gd);
    DiffStatePtrs A{};
    for (int c = 0; c < d->dim + 2; c++) {
        A.src[c] = U[c];
        A.U[c] = V[c];
    }
    for (long long t = 0; t < Gd.ncell_g; t++) diff_extract_view_thread(Gs, Gd, A, d->dim + 2, t);
    return 0;
}

extern "C" int emu_diff_accumulate(const EmuDiffDesc* d, int g, double beta, const double* const* Fd, double* const* U)
{
    const int dim = d->dim, neq = dim + 2;
    NsAccArgs A{};
    make_diff_geom(dim, d->n, d->dx, g, &A.G);
    A.neq = neq;
    A.beta = beta;
    for (int f = 0; f < dim * neq; f++) A.Fd[f] = Fd[f];
    for (int e = 0; e < neq; e++) A.U[e] = U[e];
    const long long total = (long long)A.G.n[0] * A.G.n[1] * A.G.n[2];
    for (long long t = 0; t < total; t++) {
        if (dim == 2)
            diff_accumulate_thread<2>(A, t);
        else
            diff_accumulate_thread<3>(A, t);
    }
    return 0;
}

/* the compile-time term tables the kernels are unrolled from */
extern "C" int emu_diff_terms(int dim, int f, int d, int e, int var[4], int diff[4])
{
    const DiffTermList tl = dim == 2 ? DiffTerms<2>::get(f, d, e) : DiffTerms<3>::get(f, d, e);
    for (int i = 0; i < tl.n; i++) {
        var[i] = tl.t[i].var;
        diff[i] = tl.t[i].diff;
    }
    return tl.n;
}

template <int DIM>
static int run_div(const EmuDiffDesc* d, const double* const* Q, double dt, int g, double beta, double* const* U)
{
    NodeStage<DIM> S(d, Q);
    NsDivArgs D{};
    D.G6 = S.G;
    make_diff_geom(d->dim, d->n, d->dx, g, &D.GU);
    D.neq = DIM + 2;
    D.beta = beta;
    D.dt = dt;
    for (int f = 0; f < DIM; f++)
        for (int e = 0; e < DIM + 2; e++) D.Fn[f][e] = S.Fn[f * (DIM + 2) + e].data();
    for (int e = 0; e < DIM + 2; e++) D.U[e] = U[e];
    const long long total = (long long)S.G.n[0] * S.G.n[1] * S.G.n[2];
    for (long long t = 0; t < total; t++) diff_divergence_accumulate_thread<DIM>(D, t);
    return 0;
}

extern "C" int emu_diff_divergence_accumulate(const EmuDiffDesc* d, const double* const* Q, double dt, int g, double beta,
                                              double* const* U)
{
    return d->dim == 2 ? run_div<2>(d, Q, dt, g, beta, U) : run_div<3>(d, Q, dt, g, beta, U);
}

/* midpoint family: the thread functions of k_diff_primitives / k_diff_mid_flux / k_diff_mid_face from plain loops */
template <int DIM, int FDIR>
static void run_mid_dir(const DiffGeom& G, const DiffConsts& K, std::vector<std::vector<double>>& P, std::vector<std::vector<double>>& Fm,
                        double dt, double* const* F)
{
    DiffMidPtrs A{};
    for (int v = 0; v < DIM + 1; v++) A.P[v] = P[v].data();
    for (int e = 0; e < DIM + 2; e++) {
        A.Fm[e] = Fm[e].data();
        A.F[e] = F[FDIR * (DIM + 2) + e];
    }
    for (long long t = 0; t < diff_mid_count<DIM, FDIR>(G); t++) diff_mid_flux_thread<DIM, FDIR>(G, K, A, t);
    for (long long t = 0; t < diff_face_count<DIM, FDIR>(G); t++) diff_mid_face_thread<DIM, FDIR>(G, A, dt, t);
}

template <int DIM>
static int run_mid(const EmuDiffDesc* d, const double* const* Q, double dt, double* const* F)
{
    DiffGeom G;
    make_diff_geom(d->dim, d->n, d->dx, HB2_DIFF_G, &G);
    DiffConsts K{d->gamma, d->c_v, d->mu, d->mu_v, d->c_p * d->mu / d->Pr};
    const double nan = std::numeric_limits<double>::quiet_NaN();
    std::vector<std::vector<double>> P(DIM + 1, std::vector<double>((size_t)G.ncell_g, nan)), Fm(DIM + 2, std::vector<double>((size_t)G.ncell_g, nan));
    DiffPtrs A{};
    for (int c = 0; c < DIM + 2; c++) A.Q[c] = Q[c];
    for (int v = 0; v < DIM + 1; v++) A.P[v] = P[v].data();
    for (long long x = 0; x < G.ncell_g; x++) diff_primitives_thread<DIM>(K, A, x);
    run_mid_dir<DIM, 0>(G, K, P, Fm, dt, F);
    run_mid_dir<DIM, 1>(G, K, P, Fm, dt, F);
    if (DIM == 3) run_mid_dir<DIM, (DIM == 3 ? 2 : 1)>(G, K, P, Fm, dt, F);
    return 0;
}

extern "C" int emu_diffusive_flux_midpoint(const EmuDiffDesc* d, const double* const* Q, double dt, double* const* F)
{
    return d->dim == 2 ? run_mid<2>(d, Q, dt, F) : run_mid<3>(d, Q, dt, F);
}

/* The re-associated arithmetic of the flux-free route (HB2_MATH_FAST; the point functions the marching kernels of
 * hb2_diffusive_march.cuh call with MATH = 1), 3-D, from plain loops: primitives on the ghost box, node fluxes on the interior
 * extended by three, then the update of every interior cell. */
extern "C" int emu_diff_divergence_accumulate_fast(const EmuDiffDesc* d, const double* const* Q, double dt, int g, double beta,
                                                   double* const* U)
{
    if (d->dim != 3) return -1;
    DiffGeom G, GU;
    make_diff_geom(3, d->n, d->dx, HB2_DIFF_G, &G);
    make_diff_geom(3, d->n, d->dx, g, &GU);
    DiffConsts K{d->gamma, d->c_v, d->mu, d->mu_v, d->c_p * d->mu / d->Pr};
    DiffFast F;
    make_diff_fast(G, K, dt, beta, &F);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    std::vector<std::vector<double>> P(4, std::vector<double>((size_t)G.ncell_g, nan)), Fn(15, std::vector<double>((size_t)G.ncell_g, nan));
    for (long long x = 0; x < G.ncell_g; x++) {
        const double q[5] = {Q[0][x], Q[1][x], Q[2][x], Q[3][x], Q[4][x]};
        double p[4];
        diff_primitives_fast(q, F, p);
        for (int v = 0; v < 4; v++) P[v][(size_t)x] = p[v];
    }
    for (int k = -3; k < G.n[2] + 3; k++)
        for (int j = -3; j < G.n[1] + 3; j++)
            for (int i = -3; i < G.n[0] + 3; i++) {
                const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
                double der[4][3];
                for (int v = 0; v < 4; v++)
                    for (int a = 0; a < 3; a++) {
                        const double* c = P[v].data() + x;
                        const long long s = G.cs[a];
                        der[v][a] = diff_first_derivative6_fast(c[-3 * s], c[-2 * s], c[-s], c[s], c[2 * s], c[3 * s], F.cd[a]);
                    }
                const double vel[3] = {P[0][(size_t)x], P[1][(size_t)x], P[2][(size_t)x]};
                double fn[3][5];
                diff_node_flux_fast(F, vel, der, fn);
                for (int f = 0; f < 3; f++)
                    for (int e = 1; e < 5; e++) Fn[f * 5 + e][(size_t)x] = fn[f][e];
            }
    for (int k = 0; k < G.n[2]; k++)
        for (int j = 0; j < G.n[1]; j++)
            for (int i = 0; i < G.n[0]; i++) {
                const long long x = (i + G.g[0]) + G.cs[1] * (j + G.g[1]) + G.cs[2] * (k + G.g[2]);
                const long long xu = (i + GU.g[0]) + GU.cs[1] * (j + GU.g[1]) + GU.cs[2] * (k + GU.g[2]);
                for (int e = 1; e < 5; e++) {
                    double acc = U[e][xu];
                    for (int a = 0; a < 3; a++) {
                        const double* c = Fn[a * 5 + e].data() + x;
                        const long long s = G.cs[a];
                        acc = diff_divergence_fast(acc, c[-3 * s], c[-2 * s], c[-s], c[s], c[2 * s], c[3 * s], F.kd[a]);
                    }
                    U[e][xu] = acc;
                }
            }
    return 0;
}

extern "C" double emu_diff_max_spectral_radius(const EmuDiffDesc* d, double c_p_eos, const double* rho)
{
    DiffGeom G;
    make_diff_geom(d->dim, d->n, d->dx, HB2_DIFF_G, &G);
    DiffConsts K{d->gamma, d->c_v, d->mu, d->mu_v, d->c_p * d->mu / d->Pr};
    double m = 0.0;
    for (long long x = 0; x < G.ncell_g; x++)
        m = std::fmax(m, d->dim == 2 ? diff_spectral_radius_cell<2>(G, K, c_p_eos, rho[x]) : diff_spectral_radius_cell<3>(G, K, c_p_eos, rho[x]));
    return m;
}
