/*
 * hb2_fast.cuh -- re-associated ("fast") arithmetic of the WCNS5-JS / HLLC-HLL path for the FP64 pipe of B200.
 *
 * Measured on B200 (tools/ubench_fp64.cu, profiles/r01_b_ubench_fp64.txt): DFMA latency 8.2 cycles, one warp
 * instruction per 2 cycles per SM sub-partition (64 lanes/clk/SM, 36.7 TFLOP/s); `1.0/x` costs 7.4 issue slots,
 * `a/x` 11.7, `sqrt` 12.5, while MUFU.RCP64H / RSQ64H + two Newton steps cost 4.5 / 7.  The path is bound by the
 * FP64 pipe, so the fast variant is written to MINIMISE FP64 INSTRUCTIONS while staying an
 * algebraic re-statement of the reference formulas (<= 1e-12 relative of the oracle on well-conditioned faces):
 *   - WCNS5-JS: the smoothness indicators in their sum-of-squares form 13/12 (a-2b+c)^2 + 1/4 (..)^2 (the
 *     polynomial the reference expands, ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:31-44, :58-71), scaled by 4
 *     (the normalised weights are invariant when beta and epsilon are scaled together); the second differences are
 *     shared between the minus and the plus side; the six divisions of the normalised weights become ONE reciprocal:
 *     omega_k = d_k prod_{j!=k} b_j / sum_k(d_k prod_{j!=k} b_j), b_k = (beta_k + eps)^p;
 *   - characteristic projection constants from one reciprocal of rho c^2;
 *   - HLLC written on the upwind side only; passive components (mass, tangential momentum, volume fractions) are
 *     q_K * u_mid with u_mid = u_K + s(chi - 1) (identical to F_K + s(Q*_K - Q_K)); rho_K eps_K = p_K/(gamma-1);
 *   - reciprocals / square roots by MUFU seed + two Newton steps (~1 ulp, no special-case branches);
 *   - flux differencing in difference form 3/2 (Fm[c+1]-Fm[c]) + 1/30 (Fm[c+2]-Fm[c-1]) - 3/10 (Fn[c+1]-Fn[c-1]).
 * Everything is `__host__ __device__` (tests/host_emu runs it on the CPU; there the seeds are plain 1/x, 1/sqrt).
 *
 * Reference formulas restated: see the citations in hb2_core.cuh (same functions, reference operation order).
 */
#pragma once
#include "hb2_core.cuh"

/* Unroll factor of the characteristic-field loop of face_midpoint_fast.  Measured on B200 (256^3, x sweep): the rolled
 * loop (1) costs ~115 M extra warp instructions for the field switch and register shuffles (783 M vs 667 M) and the
 * warp schedulers sustain only ~0.57 instructions/clk here, so the fully unrolled loop is 14 % faster; its code
 * (29 KB) still fits the 32 KB instruction cache. */
#ifndef HB2_FIELD_UNROLL
#define HB2_FIELD_UNROLL 8
#endif

namespace hb2 {

constexpr int kFieldUnroll = HB2_FIELD_UNROLL;

/* min / max without the NaN bookkeeping of fmin / fmax (3 instead of ~6 instructions; the operands are finite) */
HB2_HD double min_fast(double a, double b) { return (a < b) ? a : b; }
HB2_HD double max_fast(double a, double b) { return (a > b) ? a : b; }

/* true if the predicate holds for ANY thread of the (converged part of the) warp: lets a warp skip a rarely needed,
 * if-converted block with one vote.  On the host (tests/host_emu) it is the predicate itself. */
HB2_HD bool warp_any(bool pred)
{
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), pred);
#else
    return pred;
#endif
}

/* 1/x: MUFU.RCP64H seed r0 (relative error e0 ~ 2^-22) + ONE third-order step r0 (1 + e + e^2), e = 1 - x r0
 * (remaining error e0^3 ~ 2^-66, i.e. the result is rounded from a value good to ~1 ulp): 3 FP64 instructions instead
 * of the 4 of two Newton steps, and ~10 instead of the library division */
/* keeps a warp-uniform, rarely taken block a real branch (the compiler otherwise if-converts it into predicated
 * instructions that are issued every time) */
HB2_HD void uniform_branch_fence()
{
#if defined(__CUDA_ARCH__)
    asm volatile("" ::: "memory");
#endif
}

HB2_HD double rcp_fast(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#else
    return 1.0 / x;
#endif
}

/* sqrt(x) for x > 0: MUFU.RSQ64H seed y0 = (1 + e0)/sqrt(x), e0 ~ 2^-22; g = x y0, r = 1/2 - g y0/2 = -(e0 + e0^2/2);
 * sqrt(x) = g/(1 + e0) = g (1 + r + 3/2 r^2) + O(e0^3): 6 FP64 instructions instead of the 7 of two coupled steps.
 * GUARD: also return 0 for x == 0 (the seed is +inf there). */
template <bool GUARD>
HB2_HD double sqrt_fast(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y;
    const double h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    const double t = fma(1.5 * r, r, r);
    const double q = fma(g, t, g);
    return (GUARD && x == 0.0) ? 0.0 : q;
#else
    return sqrt(x);
#endif
}

/* ------------------------------------------------------------------------------------------
 * cell stage
 * ---------------------------------------------------------------------------------------- */
template <class Tr>
HB2_HD void cons_to_prim_fast(const double (&q)[Tr::NCOMP], const Consts& K, double (&V)[Tr::NEQ], double& c)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS;
    if (Tr::MODEL == SS) {
        const double rho = q[0];
        const double r = rcp_fast(rho);
        V[0] = rho;
        double ke = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            V[1 + a] = q[1 + a] * r;
            ke = fma(V[1 + a], V[1 + a], ke);
        }
        const double p = (K.gamma[0] - 1.0) * fma(-0.5 * rho, ke, q[DIM + 1]);
        V[DIM + 1] = p;
        c = sqrt_fast<true>(K.gamma[0] * p * r);
    } else {
        double rho = q[0];
#pragma unroll
        for (int si = 1; si < NS; si++) rho += q[si];
        const double r = rcp_fast(rho);
#pragma unroll
        for (int si = 0; si < NS; si++) V[si] = q[si];
        double ke = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            V[NS + a] = q[NS + a] * r;
            ke = fma(V[NS + a], V[NS + a], ke);
        }
        double xi = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) xi = fma(q[NS + DIM + 1 + si], K.inv_gm1[si], xi);
        const double Gamma = rcp_fast(xi); /* gamma_m - 1 */
        const double p = Gamma * fma(-0.5 * rho, ke, q[NS + DIM]);
        V[NS + DIM] = p;
        /* c^2 = Gamma p/rho + sum_i Y_i p/rho with sum_i Y_i = 1 */
        c = sqrt_fast<true>((Gamma + 1.0) * p * r);
#pragma unroll
        for (int si = 0; si < Tr::NZ; si++) V[NS + DIM + 1 + si] = q[NS + DIM + 1 + si];
    }
}

/* node flux of a cell from its PRIMITIVE variables (cell[comp*CS]); E_stored: the stored total energy (five-eqn) */
template <class Tr, int DIR, int CS>
HB2_HD void node_flux_prim(const double* cell, double E_stored, const Consts& K, double (&Fn)[Tr::NEQ])
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, IV = Tr::IV, IP = Tr::IP;
    double u[DIM];
    double ke = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
        u[a] = cell[(IV + a) * CS];
        ke = fma(u[a], u[a], ke);
    }
    const double p = cell[IP * CS];
    const double un = u[DIR];
    double rho, E;
    if (Tr::MODEL == SS) {
        rho = cell[0];
        E = fma(0.5 * rho, ke, p * K.inv_gm1[0]);
        Fn[0] = rho * un;
    } else {
        rho = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) {
            const double zr = cell[si * CS];
            rho += zr;
            Fn[si] = un * zr;
        }
        E = E_stored;
#pragma unroll
        for (int si = 0; si < Tr::NZ; si++) Fn[IP + 1 + si] = un * cell[(IP + 1 + si) * CS];
    }
    const double ru = rho * un;
#pragma unroll
    for (int a = 0; a < DIM; a++) Fn[IV + a] = (a == DIR) ? fma(ru, u[a], p) : ru * u[a];
    Fn[IP] = un * (E + p);
}

/* conservative variables of a cell from its PRIMITIVE variables (cell[comp*CS]); E_stored: the stored total energy
 * (five-eqn: kept in the ring; single-species: rebuilt, p/(gamma-1) + rho |u|^2/2, good to an ulp or two) */
template <class Tr, int CS>
HB2_HD void prim_to_cons(const double* cell, double E_stored, const Consts& K, double (&q)[Tr::NEQ])
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, IV = Tr::IV, IP = Tr::IP;
    double rho = 0.0;
#pragma unroll
    for (int si = 0; si < NS; si++) {
        q[si] = cell[si * CS];
        rho += q[si];
    }
    double ke = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
        const double u = cell[(IV + a) * CS];
        q[IV + a] = rho * u;
        ke = fma(u, u, ke);
    }
    q[IP] = (Tr::MODEL == SS) ? fma(0.5 * rho, ke, cell[IP * CS] * K.inv_gm1[0]) : E_stored;
#pragma unroll
    for (int si = 0; si < Tr::NZ; si++) q[IP + 1 + si] = cell[(IP + 1 + si) * CS];
}

/* ------------------------------------------------------------------------------------------
 * WCNS5-JS: minus-side and plus-side midpoint values from the six stencil values w0..w5
 * ---------------------------------------------------------------------------------------- */
HB2_HD void weno5js_pair_fast(double w0, double w1, double w2, double w3, double w4, double w5, double eps4,
                              double& wm, double& wp)
{
    /* eps4 = 4*eps*(scale of w)^2: the caller may hand in characteristic variables scaled by a constant (the
     * normalised weights are invariant when beta and epsilon are scaled together) */
    /* first differences, then the second differences of the four 3-cell sub-stencils (shared by both sides) */
    const double e01 = w1 - w0, e12 = w2 - w1, e23 = w3 - w2, e34 = w4 - w3, e45 = w5 - w4;
    const double s012 = e12 - e01, s123 = e23 - e12, s234 = e34 - e23, s345 = e45 - e34;
    /* 4*beta_k + 4*eps = 13/3 s^2 + f^2 + eps4: the s-part of the two middle sub-stencils serves both sides */
    const double q012 = fma((13.0 / 3.0) * s012, s012, eps4);
    const double q123 = fma((13.0 / 3.0) * s123, s123, eps4);
    const double q234 = fma((13.0 / 3.0) * s234, s234, eps4);
    const double q345 = fma((13.0 / 3.0) * s345, s345, eps4);
    /* minus side (cell 2 is the upwind cell): f0 = w0 - 4 w1 + 3 w2 = s012 + 2 e12, f1 = w1 - w3,
     * f2 = 3 w2 - 4 w3 + w4 = s234 - 2 e23 */
    const double f0 = fma(2.0, e12, s012);
    const double f1 = e12 + e23; /* sign irrelevant: squared */
    const double f2 = fma(-2.0, e23, s234);
    double b0 = fma(f0, f0, q012);
    double b1 = fma(f1, f1, q123);
    double b2 = fma(f2, f2, q234);
    /* plus side (cell 3 is the upwind cell), mirrored: g0 = 3 w3 - 4 w4 + w5 = s345 - 2 e34, g1 = w4 - w2,
     * g2 = w1 - 4 w2 + 3 w3 = s123 + 2 e23 */
    const double g0 = fma(-2.0, e34, s345);
    const double g1 = e23 + e34;
    const double g2 = fma(2.0, e23, s123);
    double c0 = fma(g0, g0, q345);
    double c1 = fma(g1, g1, q234);
    double c2 = fma(g2, g2, q123);
    /* constant_p = 2 (the reference's default, ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:188-191); plans with
     * another exponent run the exact-arithmetic kernels */
    b0 *= b0; b1 *= b1; b2 *= b2;
    c0 *= c0; c1 *= c1; c2 *= c2;
    /* un-normalised weights (x16): a0 = b1 b2, 10 a1 = 10 b0 b2, 5 a2 = 5 b0 b1;
     * value = P1 + (a0 (P0 - P1) + 5 a2 (P2 - P1))/sum, the sub-stencil differences are third differences:
     * P0 - P1 = 3/8 (s012 - s123), P2 - P1 = 1/8 (s123 - s234) (mirrored on the plus side), and the central
     * sub-stencil value is P1 = (w2 + w3)/2 - s123/8 (plus side: - s234/8) */
    const double d1 = s012 - s123, d2 = s123 - s234, d3 = s234 - s345;
    const double h = fma(0.5, e23, w2);
    {
        const double a0 = b1 * b2, a1 = b0 * b2, a2 = b0 * b1;
        const double sum = fma(5.0, a2, fma(10.0, a1, a0));
        const double P1 = fma(-0.125, s123, h);
        const double u0 = a0 * d1, u2 = a2 * d2;
        wm = fma(fma(0.625, u2, 0.375 * u0), rcp_fast(sum), P1);
    }
    {
        const double a0 = c1 * c2, a1 = c0 * c2, a2 = c0 * c1;
        const double sum = fma(5.0, a2, fma(10.0, a1, a0));
        const double P1 = fma(-0.125, s234, h);
        const double u0 = a0 * d3, u2 = a2 * d2;
        wp = fma(-fma(0.625, u2, 0.375 * u0), rcp_fast(sum), P1);
    }
}

/* WCNS5-Z (ConvectiveFluxReconstructorWCNS5-Z-HLLC-HLL.cpp:78-165) in the same shared form: with b_k = 4(beta_k + eps),
 * B_k = b_k^2 and T = (b_0 - b_2)^2 = (4 tau_5)^2 the un-normalised weights d_k (1 + (tau_5/(beta_k + eps))^2) become
 * d_k (B_k + T) prod_{j != k} B_j over the common denominator prod_j B_j: one reciprocal per side, 82 FP64 instructions
 * per minus/plus pair.  constant_p = 2 only (other exponents run the reference-order kernels). */
HB2_HD void weno5z_pair_fast(double w0, double w1, double w2, double w3, double w4, double w5, double eps4,
                             double& wm, double& wp)
{
    const double e01 = w1 - w0, e12 = w2 - w1, e23 = w3 - w2, e34 = w4 - w3, e45 = w5 - w4;
    const double s012 = e12 - e01, s123 = e23 - e12, s234 = e34 - e23, s345 = e45 - e34;
    const double q012 = fma((13.0 / 3.0) * s012, s012, eps4);
    const double q123 = fma((13.0 / 3.0) * s123, s123, eps4);
    const double q234 = fma((13.0 / 3.0) * s234, s234, eps4);
    const double q345 = fma((13.0 / 3.0) * s345, s345, eps4);
    const double f0 = fma(2.0, e12, s012);
    const double f1 = e12 + e23;
    const double f2 = fma(-2.0, e23, s234);
    const double b0 = fma(f0, f0, q012);
    const double b1 = fma(f1, f1, q123);
    const double b2 = fma(f2, f2, q234);
    const double g0 = fma(-2.0, e34, s345);
    const double g1 = e23 + e34;
    const double g2 = fma(2.0, e23, s123);
    const double c0 = fma(g0, g0, q345);
    const double c1 = fma(g1, g1, q234);
    const double c2 = fma(g2, g2, q123);
    const double d1 = s012 - s123, d2 = s123 - s234, d3 = s234 - s345;
    const double h = fma(0.5, e23, w2);
    {
        const double tb = b0 - b2;
        const double T = tb * tb;
        const double B0 = b0 * b0, B1 = b1 * b1, B2 = b2 * b2;
        const double a0 = (B0 + T) * (B1 * B2), a1 = (B1 + T) * (B0 * B2), a2 = (B2 + T) * (B0 * B1);
        const double sum = fma(5.0, a2, fma(10.0, a1, a0));
        const double P1 = fma(-0.125, s123, h);
        const double u0 = a0 * d1, u2 = a2 * d2;
        wm = fma(fma(0.625, u2, 0.375 * u0), rcp_fast(sum), P1);
    }
    {
        const double tb = c0 - c2;
        const double T = tb * tb;
        const double B0 = c0 * c0, B1 = c1 * c1, B2 = c2 * c2;
        const double a0 = (B0 + T) * (B1 * B2), a1 = (B1 + T) * (B0 * B2), a2 = (B2 + T) * (B0 * B1);
        const double sum = fma(5.0, a2, fma(10.0, a1, a0));
        const double P1 = fma(-0.125, s234, h);
        const double u0 = a0 * d3, u2 = a2 * d2;
        wp = fma(-fma(0.625, u2, 0.375 * u0), rcp_fast(sum), P1);
    }
}

/* WCNS6-LD (ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp:23-339), one side, from the six values (a..f, upwind
 * cell c), the three shared smoothness indicators b_k = beta_k + eps and sigma.  constant_p = 2, constant_q = 4.
 * The reference's 3 + 4 + 1 divisions by (beta_k + eps) and (beta_avg + eps) become five reciprocals shared by the
 * upwind and the central weights, the two normalisations one reciprocal each; the R_tau > alpha_tau switch is a select.
 * beta_3 is the reference's polynomial as written (FMA-contracted). */
HB2_HD double weno6ld_side_fast(double a, double b, double c, double d, double e, double f, double b0, double b1,
                                double b2, double sigma, double eps_b, const Consts& K)
{
    const double beta_3 = (1.0 / 232243200.0) *
        fma(a, fma(525910327.0, a, fma(-4562164630.0, b, fma(7799501420.0, c, fma(-6610694540.0, d, fma(2794296070.0, e, -472758974.0 * f))))),
        fma(5.0 * b, fma(2146987907.0, b, fma(-7722406988.0, c, fma(6763559276.0, d, fma(-2926461814.0, e, 503766638.0 * f)))),
        fma(20.0 * c, fma(1833221603.0, c, fma(-3358664662.0, d, fma(1495974539.0, e, -263126407.0 * f))),
        fma(20.0 * d, fma(1607794163.0, d, fma(-1486026707.0, e, 268747951.0 * f)),
        fma(5.0 * e, fma(1432381427.0, e, -536951582.0 * f), 263126407.0 * f * f)))));
    const double b3 = beta_3 + eps_b;
    const double r0 = rcp_fast(b0), r1 = rcp_fast(b1), r2 = rcp_fast(b2), r3 = rcp_fast(b3);
    /* upwind (WCNS5-Z) weights */
    const double tau_5 = fabs(b0 - b2);
    const double y0 = tau_5 * r0, y1 = tau_5 * r1, y2 = tau_5 * r2;
    const double u0 = (1.0 / 16.0) * fma(y0, y0, 1.0), u1 = (5.0 / 8.0) * fma(y1, y1, 1.0), u2 = (5.0 / 16.0) * fma(y2, y2, 1.0);
    const double ur = rcp_fast(u0 + u1 + u2);
    /* central weights: (beta_0 + beta_2 + 6 beta_1)/8 + eps = (b0 + b2 + 6 b1)/8 */
    const double bavg = 0.125 * fma(6.0, b1, b0 + b2);
    const double tau_6 = fabs(b3 - bavg);
    const double x0 = tau_6 * r0, x1 = tau_6 * r1, x2 = tau_6 * r2, x3 = tau_6 * r3;
    const double x0s = x0 * x0, x1s = x1 * x1, x2s = x2 * x2, x3s = x3 * x3;
    const double c0 = (1.0 / 32.0) * fma(x0s, x0s, K.weno_C), c1 = (15.0 / 32.0) * fma(x1s, x1s, K.weno_C);
    const double c2 = (15.0 / 32.0) * fma(x2s, x2s, K.weno_C), c3 = (1.0 / 32.0) * fma(x3s, x3s, K.weno_C);
    const double cr = rcp_fast((c0 + c1) + (c2 + c3));
    double o0 = c0 * cr, o1 = c1 * cr, o2 = c2 * cr, o3 = c3 * cr;
    const double R_tau = tau_6 * rcp_fast(bavg);
    const bool blend = R_tau > K.weno_alpha_tau;
    const double sg = blend ? sigma : 0.0;      /* omega = sigma omega_upwind + (1 - sigma) omega_central where blended */
    const double su = sg * ur, om = 1.0 - sg;
    o0 = fma(su, u0, om * o0);
    o1 = fma(su, u1, om * o1);
    o2 = fma(su, u2, om * o2);
    o3 = om * o3;
    return fma(0.375 * o0, a, fma(fma(-1.25, o0, -0.125 * o1), b, fma(fma(1.875, o0, fma(0.75, o1, 0.375 * o2)), c,
           fma(fma(0.375, o1, fma(0.75, o2, 1.875 * o3)), d, fma(fma(-0.125, o2, -1.25 * o3), e, 0.375 * o3 * f)))));
}

HB2_HD void weno6ld_pair_fast(double w0, double w1, double w2, double w3, double w4, double w5, double eps4,
                              const Consts& K, double& wm, double& wp)
{
    /* eps4 = 4 eps scale^2 (the caller hands in characteristic variables scaled by 1 or 2) */
    const double eps_b = 0.25 * eps4;
    const double eps_s = (eps4 > 6.0 * HB2_EPS) ? 2.0 * HB2_EPS : HB2_EPS;
    const double e01 = w1 - w0, e12 = w2 - w1, e23 = w3 - w2, e34 = w4 - w3, e45 = w5 - w4;
    const double s012 = e12 - e01, s123 = e23 - e12, s234 = e34 - e23, s345 = e45 - e34;
    /* sigma of the unmirrored stencil, both sides (:23-39): alpha_1..3 = e12, e23, e34 */
    const double a23 = fabs(e23);
    const double theta_1 = fabs(s123) * rcp_fast(fabs(e12) + a23 + eps_s);
    const double theta_2 = fabs(s234) * rcp_fast(a23 + fabs(e34) + eps_s);
    const double sigma = max_fast(theta_1, theta_2);
    /* beta_k + eps = 13/12 s^2 + (f/2)^2 + eps */
    const double q012 = fma((13.0 / 12.0) * s012, s012, eps_b);
    const double q123 = fma((13.0 / 12.0) * s123, s123, eps_b);
    const double q234 = fma((13.0 / 12.0) * s234, s234, eps_b);
    const double q345 = fma((13.0 / 12.0) * s345, s345, eps_b);
    const double hm = 0.5 * (e12 + e23), hp = 0.5 * (e23 + e34);
    const double f0 = fma(0.5, s012, e12), f2 = fma(0.5, s234, -e23);
    const double g0 = fma(0.5, s345, -e34), g2 = fma(0.5, s123, e23);
    const double b0 = fma(f0, f0, q012), b1 = fma(hm, hm, q123), b2 = fma(f2, f2, q234);
    const double c0 = fma(g0, g0, q345), c1 = fma(hp, hp, q234), c2 = fma(g2, g2, q123);
    wm = weno6ld_side_fast(w0, w1, w2, w3, w4, w5, b0, b1, b2, sigma, eps_b, K);
    wp = weno6ld_side_fast(w5, w4, w3, w2, w1, w0, c0, c1, c2, sigma, eps_b, K);
}

/* the fast pair of this translation unit's interpolator (HB2_SCHEME) */
HB2_HD void weno_pair_fast(double w0, double w1, double w2, double w3, double w4, double w5, double eps4, const Consts& K,
                           double& wm, double& wp)
{
#if HB2_SCHEME == HB2_WCNS5_Z
    weno5z_pair_fast(w0, w1, w2, w3, w4, w5, eps4, wm, wp);
#elif HB2_SCHEME == HB2_WCNS6_LD
    weno6ld_pair_fast(w0, w1, w2, w3, w4, w5, eps4, K, wm, wp);
#else
    weno5js_pair_fast(w0, w1, w2, w3, w4, w5, eps4, wm, wp);
#endif
}

/* ------------------------------------------------------------------------------------------
 * HLLC (always) and HLLC-HLL (when `hybrid`) midpoint flux from the two interpolated states
 * ---------------------------------------------------------------------------------------- */
template <class Tr, int DIR>
HB2_HD void riemann_fast(const double (&V_L)[Tr::NEQ], const double (&V_R)[Tr::NEQ], const Consts& K, bool hybrid,
                         double (&Fm)[Tr::NEQ], double& vel_mid)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, NEQ = Tr::NEQ, IV = Tr::IV, IP = Tr::IP;

    /* side thermodynamics: rho, c, and pe = rho*eps = p/(gamma_m - 1) */
    double rho_L, rho_R, g1_L, g1_R; /* g1 = 1/(gamma_m - 1) */
    if (Tr::MODEL == SS) {
        rho_L = V_L[0];
        rho_R = V_R[0];
        g1_L = g1_R = K.inv_gm1[0];
    } else {
        rho_L = V_L[0];
        rho_R = V_R[0];
#pragma unroll
        for (int si = 1; si < NS; si++) {
            rho_L += V_L[si];
            rho_R += V_R[si];
        }
        double zl_L = 1.0, zl_R = 1.0;
        g1_L = 0.0;
        g1_R = 0.0;
#pragma unroll
        for (int si = 0; si < NS - 1; si++) {
            g1_L = fma(V_L[IP + 1 + si], K.inv_gm1[si], g1_L);
            g1_R = fma(V_R[IP + 1 + si], K.inv_gm1[si], g1_R);
            zl_L -= V_L[IP + 1 + si];
            zl_R -= V_R[IP + 1 + si];
        }
        g1_L = fma(zl_L, K.inv_gm1[NS - 1], g1_L);
        g1_R = fma(zl_R, K.inv_gm1[NS - 1], g1_R);
    }
    const double p_L = V_L[IP], p_R = V_R[IP];
    const double un_L = V_L[IV + DIR], un_R = V_R[IV + DIR];
    /* c^2 = gamma_m p/rho, gamma_m = 1 + 1/g1 */
    const double rr = rcp_fast(rho_L * rho_R);
    double gam_L, gam_R;
    if (Tr::MODEL == SS) {
        gam_L = gam_R = K.gamma[0];
    } else {
        const double rg = rcp_fast(g1_L * g1_R);
        gam_L = fma(rg, g1_R, 1.0);
        gam_R = fma(rg, g1_L, 1.0);
    }
    const double c_L = sqrt_fast<false>(gam_L * p_L * (rr * rho_R));
    const double c_R = sqrt_fast<false>(gam_R * p_R * (rr * rho_L));

    const double u_average = 0.5 * (un_L + un_R);
    const double c_average = 0.5 * (c_L + c_R);
    const double s_L = min_fast(u_average - c_average, un_L - c_L);
    const double s_R = max_fast(u_average + c_average, un_R + c_R);
    const double m_L = rho_L * (s_L - un_L);
    const double m_R = rho_R * (s_R - un_R);
    const double s_star = fma(m_L, un_L, fma(-m_R, un_R, p_R - p_L)) * rcp_fast(m_L - m_R);

    /* upwind side of the contact */
    const bool left = s_star > 0.0;
    const double rho_K = left ? rho_L : rho_R;
    const double un_K = left ? un_L : un_R;
    const double p_K = left ? p_L : p_R;
    const double s_K = left ? s_L : s_R;
    const double g1_K = left ? g1_L : g1_R;
    const double s_mp = left ? min_fast(0.0, s_L) : max_fast(0.0, s_R);
    const double d_K = s_K - un_K;
    const double d_S = s_K - s_star;
    const double rdd = rcp_fast(d_K * d_S);
    const double Chi = d_K * d_K * rdd;  /* (s_K - u_K)/(s_K - s*) */
    const double pd = p_K * d_S * rdd;   /* p_K/(s_K - u_K) */
    const double u_mid = fma(s_mp, Chi - 1.0, un_K);
    vel_mid = u_mid;

    double ke = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
        const double v = left ? V_L[IV + a] : V_R[IV + a];
        ke = fma(v, v, ke);
    }
    const double E_K = fma(0.5 * rho_K, ke, p_K * g1_K);
    const double ru = rho_K * u_mid;
    /* passive components: q_K * u_mid */
#pragma unroll
    for (int si = 0; si < Tr::NM; si++) Fm[si] = (Tr::MODEL == SS ? rho_K : (left ? V_L[si] : V_R[si])) * u_mid;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
        if (a == DIR) continue;
        Fm[IV + a] = ru * (left ? V_L[IV + a] : V_R[IV + a]);
    }
#pragma unroll
    for (int e = IP + 1; e < NEQ; e++) Fm[e] = (left ? V_L[e] : V_R[e]) * u_mid;
    /* normal momentum: rho u^2 + p + s(chi rho s* - rho u) */
    const double run = rho_K * un_K;
    const double cr = Chi * rho_K;
    Fm[IV + DIR] = fma(s_mp, fma(cr, s_star, -run), fma(run, un_K, p_K));
    /* energy: u(E + p) + s(chi(E + (s* - u)(rho s* + p/(s_K - u))) - E) */
    const double Es = Chi * fma(s_star - un_K, fma(rho_K, s_star, pd), E_K);
    Fm[IP] = fma(s_mp, Es - E_K, un_K * (E_K + p_K));

    if (!hybrid) return;

    /* HLLC-HLL: blend the passive components with the HLL flux */
    double diff[DIM];
    double mag2 = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
        diff[a] = V_R[IV + a] - V_L[IV + a];
        mag2 = fma(diff[a], diff[a], mag2);
    }
    const double vel_mag = sqrt(mag2);
    double alpha_1, alpha_2;
    if (vel_mag < HB2_EPS) {
        alpha_1 = 1.0;
        alpha_2 = 0.0;
    } else {
        alpha_1 = fabs(diff[DIR]) / vel_mag;
        alpha_2 = sqrt(1.0 - alpha_1 * alpha_1);
    }
    const double beta_1 = 0.5 * (1.0 + alpha_1 / (alpha_1 + alpha_2));
    const double beta_2 = 1.0 - beta_1;
    const double rs = 1.0 / (s_R - s_L);
#pragma unroll
    for (int e = 0; e < NEQ; e++) {
        if (e == IV + DIR || e == IP) continue;
        /* conservative value of the component on each side */
        double Q_L, Q_R;
        if (e < Tr::NM) {
            Q_L = (Tr::MODEL == SS) ? rho_L : V_L[e];
            Q_R = (Tr::MODEL == SS) ? rho_R : V_R[e];
        } else if (e < IP) {
            Q_L = rho_L * V_L[e];
            Q_R = rho_R * V_R[e];
        } else {
            Q_L = V_L[e];
            Q_R = V_R[e];
        }
        const double F_L = un_L * Q_L, F_R = un_R * Q_R;
        double F_HLL = (s_R * F_L - s_L * F_R + s_R * s_L * (Q_R - Q_L)) * rs;
        if (s_L > 0.0) F_HLL = F_L;
        if (s_R < 0.0) F_HLL = F_R;
        Fm[e] = beta_1 * Fm[e] + beta_2 * F_HLL;
    }
}

/* ------------------------------------------------------------------------------------------
 * One midpoint flux, stencil read from shared memory.
 *   win : points at component 0 of stencil cell 0 (cell f-3); component `comp` of stencil cell m is
 *         win[comp*CS + m*MS] (CS, MS compile-time: the loads become LDS [base + uniform + immediate]);
 *         component NEQ is the sound speed.
 * The characteristic fields are processed by one loop (unrolled by kFieldUnroll) whose body is the same for every
 * field: w = X + b*Y on the six stencil cells, WCNS5-JS pair, accumulate into the primitive variables.
 * ---------------------------------------------------------------------------------------- */
template <class Tr, int DIR, int CS, int MS>
HB2_HD void face_midpoint_fast(const double* win, bool hybrid, const Consts& K, double (&Fm)[Tr::NEQ], double& vel_mid)
{
    constexpr int DIM = Tr::DIM, NS = Tr::NS, NEQ = Tr::NEQ, IV = Tr::IV, IP = Tr::IP;
    static_assert(NS <= 2, "fast path is written for at most two species");
    const double eps4 = 4.0 * HB2_EPS;
    /* tangential velocity components in index order */
    constexpr int T0 = (DIR == 0) ? 1 : 0;
    constexpr int T1 = (DIR == 2) ? 1 : 2;

    double V_minus[NEQ], V_plus[NEQ];
    const double c_avg = 0.5 * (win[NEQ * CS + 2 * MS] + win[NEQ * CS + 3 * MS]);

    if (Tr::MODEL == SS) {
        /* characteristic variables scaled by 2 where that saves a multiplication:
         *   W0' = p - rho c u_n, W1 = rho - p/c^2, tangential velocities, W4' = p + rho c u_n  (W0 = W0'/2, W4 = W4'/2) */
        const double rho_avg = 0.5 * (win[2 * MS] + win[3 * MS]);
        const double rc = rho_avg * c_avg;
        const double r_rcc = rcp_fast(rc * c_avg);
        const double r_cc = r_rcc * rho_avg;      /* 1/c^2 */
        const double h_cc = 0.5 * r_cc;
        const double h_rc = 0.5 * r_rcc * c_avg;  /* 1/(2 rho c) */
        double rm = 0.0, rp = 0.0, um = 0.0, up = 0.0, pm = 0.0, pp = 0.0, t0m = 0.0, t0p = 0.0, t1m = 0.0, t1p = 0.0;
#pragma unroll kFieldUnroll
        for (int f = 0; f < NEQ; f++) {
            int xc, yc = 1 + DIR;
            double b = 0.0, e4 = eps4;
            bool has_y = true;
            if (f == 0) {
                xc = IP; b = -rc; e4 = 4.0 * eps4;
            } else if (f == 1) {
                xc = 0; yc = IP; b = -r_cc;
            } else if (f == NEQ - 1) {
                xc = IP; b = rc; e4 = 4.0 * eps4;
            } else {
                xc = 1 + ((f == 2) ? T0 : T1); has_y = false;
            }
            const double* X = win + xc * CS;
            double w[6];
#pragma unroll
            for (int m = 0; m < 6; m++) w[m] = X[m * MS];
            if (has_y) {
                const double* Y = win + yc * CS;
#pragma unroll
                for (int m = 0; m < 6; m++) w[m] = fma(b, Y[m * MS], w[m]);
            }
            double wm, wp;
            weno_pair_fast(w[0], w[1], w[2], w[3], w[4], w[5], e4, K, wm, wp);
            if (f == 0) {
                rm = h_cc * wm; rp = h_cc * wp;
                um = -h_rc * wm; up = -h_rc * wp;
                pm = 0.5 * wm; pp = 0.5 * wp;
            } else if (f == 1) {
                rm += wm; rp += wp;
            } else if (f == NEQ - 1) {
                rm = fma(h_cc, wm, rm); rp = fma(h_cc, wp, rp);
                um = fma(h_rc, wm, um); up = fma(h_rc, wp, up);
                pm = fma(0.5, wm, pm); pp = fma(0.5, wp, pp);
            } else if (f == 2) {
                t0m = wm; t0p = wp;
            } else {
                t1m = wm; t1p = wp;
            }
        }
        V_minus[0] = rm; V_plus[0] = rp;
        V_minus[1 + DIR] = um; V_plus[1 + DIR] = up;
        V_minus[IP] = pm; V_plus[IP] = pp;
        V_minus[1 + T0] = t0m; V_plus[1 + T0] = t0p;
        if (DIM == 3) {
            V_minus[1 + (T1 % DIM)] = t1m; V_plus[1 + (T1 % DIM)] = t1p;
        }
    } else {
        double Zr_avg[2];
        double rho_avg = 0.0;
#pragma unroll
        for (int si = 0; si < NS; si++) {
            Zr_avg[si] = 0.5 * (win[si * CS + 2 * MS] + win[si * CS + 3 * MS]);
            rho_avg += Zr_avg[si];
        }
        const double rc = rho_avg * c_avg;
        const double r_rcc = rcp_fast(rc * c_avg);
        const double r_rc = r_rcc * c_avg;  /* 1/(rho c) */
        const double r_c = r_rcc * rc;      /* 1/c */
        const double zc0 = Zr_avg[0] * r_rcc, zc1 = Zr_avg[NS - 1] * r_rcc; /* (Z rho)_avg/(rho c^2) */
        const double yh0 = 0.5 * Zr_avg[0] * r_c, yh1 = 0.5 * Zr_avg[NS - 1] * r_c;
        double z0m = 0.0, z0p = 0.0, z1m = 0.0, z1p = 0.0, um = 0.0, up = 0.0, pm = 0.0, pp = 0.0;
        double t0m = 0.0, t0p = 0.0, t1m = 0.0, t1p = 0.0, zzm = 0.0, zzp = 0.0;
        /* field order of the reference: u_n - p/(rho c) | Z_i rho_i - .. p (NS) | tangential (DIM-1) | Z_i (NS-1) | u_n + p/(rho c) */
        constexpr int F_T = 1 + NS, F_Z = F_T + (DIM - 1);
#pragma unroll kFieldUnroll
        for (int f = 0; f < NEQ; f++) {
            int xc;
            double b = 0.0;
            bool has_y = true;
            if (f == 0) {
                xc = IV + DIR; b = -r_rc;
            } else if (f <= NS) {
                xc = f - 1; b = (f == 1) ? -zc0 : -zc1;
            } else if (f < F_Z) {
                xc = IV + ((f == F_T) ? T0 : T1); has_y = false;
            } else if (f < NEQ - 1) {
                xc = IP + 1 + (f - F_Z); has_y = false;
            } else {
                xc = IV + DIR; b = r_rc;
            }
            const double* X = win + xc * CS;
            double w[6];
#pragma unroll
            for (int m = 0; m < 6; m++) w[m] = X[m * MS];
            if (has_y) {
                const double* Y = win + IP * CS;
#pragma unroll
                for (int m = 0; m < 6; m++) w[m] = fma(b, Y[m * MS], w[m]);
            }
            double wm, wp;
            weno_pair_fast(w[0], w[1], w[2], w[3], w[4], w[5], eps4, K, wm, wp);
            if (f == 0) {
                z0m = -yh0 * wm; z0p = -yh0 * wp;
                z1m = -yh1 * wm; z1p = -yh1 * wp;
                um = 0.5 * wm; up = 0.5 * wp;
                pm = -0.5 * rc * wm; pp = -0.5 * rc * wp;
            } else if (f == 1) {
                z0m += wm; z0p += wp;
            } else if (f <= NS) {
                z1m += wm; z1p += wp;
            } else if (f == F_T) {
                t0m = wm; t0p = wp;
            } else if (f < F_Z) {
                t1m = wm; t1p = wp;
            } else if (f < NEQ - 1) {
                zzm = wm; zzp = wp;
            } else {
                z0m = fma(yh0, wm, z0m); z0p = fma(yh0, wp, z0p);
                z1m = fma(yh1, wm, z1m); z1p = fma(yh1, wp, z1p);
                um = fma(0.5, wm, um); up = fma(0.5, wp, up);
                pm = fma(0.5 * rc, wm, pm); pp = fma(0.5 * rc, wp, pp);
            }
        }
        V_minus[0] = z0m; V_plus[0] = z0p;
        if (NS == 2) {
            V_minus[NS - 1] = z1m; V_plus[NS - 1] = z1p;
            V_minus[IP + 1] = zzm; V_plus[IP + 1] = zzp;
        }
        V_minus[IV + DIR] = um; V_plus[IV + DIR] = up;
        V_minus[IP] = pm; V_plus[IP] = pp;
        V_minus[IV + T0] = t0m; V_plus[IV + T0] = t0p;
        if (DIM == 3) {
            V_minus[IV + (T1 % DIM)] = t1m; V_plus[IV + (T1 % DIM)] = t1p;
        }
    }

    /* bounds check and first-order fallback (rare) */
    const int ok = side_bounded<Tr, DIR>(V_minus, K) & side_bounded<Tr, DIR>(V_plus, K);
    if (warp_any(!ok)) {
        uniform_branch_fence();
        if (!ok) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) {
                V_minus[e] = win[e * CS + 2 * MS];
                V_plus[e] = win[e * CS + 3 * MS];
            }
        }
    }
    riemann_fast<Tr, DIR>(V_minus, V_plus, K, hybrid, Fm, vel_mid);
}

}  // namespace hb2
