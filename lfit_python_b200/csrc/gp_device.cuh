// Gaussian-process likelihood of the residuals of one eclipse.
//
// The reference (CVModel.py:603-696) builds, per eclipse,
//     K_ij = a_in M(|x_i - x_j|) + sum_gaps a_out M(|x_i - x_j|) [x_i and x_j inside the same gap]
//            + delta_ij (ye_i^2 + 1.25e-12)
// with M the Matern-3/2 kernel of george (metric tau: M(d) = (1 + sqrt(3 d^2 / tau)) exp(-sqrt(3 d^2 / tau)))
// and asks george.GP(kernel, solver=HODLRSolver).log_likelihood(residuals, quiet=True).  The
// gaps are the out-of-eclipse stretches between the change points of calcChangepoints
// (CVModel.py:527-601).  george's HODLR solver approximates K^-1 (default tolerance 0.1); here
// the likelihood is evaluated exactly and in O(n): a Matern-3/2 process is a two-state linear
// SDE, so the sum of the global process and the per-gap process is a four-state Kalman filter
// run over the points in time order (the per-gap process is re-drawn from its prior whenever a
// new gap begins, which makes different gaps independent, as the block kernels are).
//
// Plain C++ with LFB_HD so that the same code runs in the kernel and in a host test harness.
#pragma once
#include "roche_device.cuh"

namespace lfb {

constexpr int kMaxGaps = 8;
constexpr double kGpWhiteNoise = 1.25e-12;  // george.GP's default white-noise term (TINY)
constexpr double kLn2Pi = 1.8378770664093454836;

struct GpPars {
    double a_in, a_out, tau;  // amplitudes (variances) and george metric (squared time scale)
    int n_gaps;
    double gap[kMaxGaps][2];  // [start, end], inclusive, disjoint, ascending
};

// index of the gap holding x, -1 if none
LFB_HD int gp_gap_of(const GpPars& G, double x)
{
    int found = -1;
#pragma unroll
    for (int k = kMaxGaps - 1; k >= 0; --k)  // fixed trip count: the table stays in registers
        if (k < G.n_gaps && x >= G.gap[k][0] && x <= G.gap[k][1]) found = k;
    return found;
}

// The change points of an eclipse's light curve (CVModel.py:580-599): one gap per cycle number e
// with x_min < e < 1 + x_max, from the end of eclipse e - 1 to the start of eclipse e.
LFB_HD void gp_changepoints(double x_min, double x_max, double dist_cp, double phi0, GpPars& G)
{
    G.n_gaps = 0;
    const int e_lo = (int)floor(x_min), e_hi = (int)ceil(x_max);
    for (int e = e_lo; e <= e_hi && G.n_gaps < kMaxGaps; ++e) {
        if (!((double)e > x_min && (double)e < 1.0 + x_max)) continue;
        G.gap[G.n_gaps][0] = (double)(e - 1) + dist_cp + phi0;
        G.gap[G.n_gaps][1] = (double)e - dist_cp + phi0;
        ++G.n_gaps;
    }
}

// The filter, one data point at a time (so that a kernel can feed it from staged tiles).
struct GpFilter {
    double c, c2;
    double m0, m1, m2, m3;                                    // state mean (f1, f1', f2, f2')
    double p00, p01, p11, p02, p03, p12, p13, p22, p23, p33;  // state covariance, upper triangle
    double xp, ll, sprod;
    double t_dt, t_a, t_b, t_c, t_d, t_q00, t_q01, t_q11;  // the transition over the last spacing (see advance)
    double g_lo, g_hi;  // the first gap that does not end before the current point (times ascend)
    int gi, cur_gap, last_gap, k, nprod, dir;  // last_gap: gap of the last point stepped over (-1: none)
    bool bad;

    // dir = +1: points arrive in ascending time; -1: in descending time (the process is
    // time-reversible: the reversed filter carries (f, -f') and runs the same recursion).  Inside,
    // the filter works with times multiplied by dir, which always ascend -- both directions run the
    // same instructions (they share warps).
    LFB_HD void init(const GpPars& G, int direction = 1)
    {
        dir = direction;
        bad = !(G.a_in > 0.0) || !(G.a_out > 0.0) || !(G.tau > 0.0) || !(G.a_in < 1e300) || !(G.a_out < 1e300) ||
              !(G.tau < 1e300);
        const double tau = bad ? 1.0 : G.tau;
        c = sqrt(3.0 / tau);
        c2 = c * c;
        // both processes at their priors
        m0 = m1 = m2 = m3 = 0.0;
        p00 = G.a_in; p01 = 0.0; p11 = G.a_in * c2;
        p02 = p03 = p12 = p13 = 0.0;
        p22 = G.a_out; p23 = 0.0; p33 = G.a_out * c2;
        cur_gap = -1;
        last_gap = -1;
        k = 0;
        xp = 0.0;
        ll = 0.0;
        sprod = 1.0;
        nprod = 0;
        gi = -1;
        t_dt = -1.0;
        t_a = t_d = 1.0;
        t_b = t_c = t_q00 = t_q01 = t_q11 = 0.0;
        g_lo = g_hi = -INFINITY;
        next_gap(G);
    }

    // bounds, in the filter's own time direction, of the next gap along the way
    LFB_HD void next_gap(const GpPars& G)
    {
        ++gi;
        if (gi < G.n_gaps) {
            const int idx = dir > 0 ? gi : G.n_gaps - 1 - gi;
            g_lo = dir > 0 ? G.gap[idx][0] : -G.gap[idx][1];
            g_hi = dir > 0 ? G.gap[idx][1] : -G.gap[idx][0];
        } else {
            g_lo = g_hi = INFINITY;
        }
    }

    // gap holding time xk (index into G.gap; -1: none); times only move in the filter's direction
    LFB_HD int gap_at(const GpPars& G, double xk)
    {
        const double xs = dir > 0 ? xk : -xk;
        while (xs > g_hi) next_gap(G);  // rarely: the point has left gap gi behind
        return xs >= g_lo ? (dir > 0 ? gi : G.n_gaps - 1 - gi) : -1;
    }

    // move the state to time xk without an observation
    LFB_HD void advance(const GpPars& G, double xk)
    {
        if (bad) return;
        if (k > 0) {
            const double dt = dir > 0 ? xk - xp : xp - xk;  // a select, not a branch
            if (!(dt >= 0.0)) {  // times must be monotonic
                bad = true;
                return;
            }
            // transition over dt: kept from the step before when the spacing repeats (regular sampling, up to the
            // rounding of the phases: 1e-11 relative, far below anything the likelihood can see)
            if (!(fabs(dt - t_dt) <= 1e-11 * dt)) {
                const double ed = exp(-c * dt);
                t_dt = dt;
                t_a = ed * (1.0 + c * dt);
                t_b = ed * dt;
                t_c = -ed * c2 * dt;
                t_d = ed * (1.0 - c * dt);
                t_q00 = 1.0 - (t_a * t_a + t_b * t_b * c2);
                t_q01 = -(t_a * t_c + t_b * t_d * c2);
                t_q11 = c2 - (t_c * t_c + t_d * t_d * c2);
            }
            const double a = t_a, b = t_b, cc = t_c, d = t_d;
            // mean
            const double n0 = a * m0 + b * m1, n1 = cc * m0 + d * m1, n2 = a * m2 + b * m3, n3 = cc * m2 + d * m3;
            m0 = n0; m1 = n1; m2 = n2; m3 = n3;
            // process noise of a unit-variance Matern-3/2 state: Pinf - A Pinf A^T, Pinf = diag(1, c^2)
            const double q00 = t_q00, q01 = t_q01, q11 = t_q11;
            // covariance blocks: X <- A X A^T (+ Q)
            {
                const double t00 = a * p00 + b * p01, t01 = a * p01 + b * p11, t10 = cc * p00 + d * p01, t11 = cc * p01 + d * p11;
                p00 = t00 * a + t01 * b + G.a_in * q00;
                p01 = t00 * cc + t01 * d + G.a_in * q01;
                p11 = t10 * cc + t11 * d + G.a_in * q11;
            }
            {
                const double t00 = a * p22 + b * p23, t01 = a * p23 + b * p33, t10 = cc * p22 + d * p23, t11 = cc * p23 + d * p33;
                p22 = t00 * a + t01 * b + G.a_out * q00;
                p23 = t00 * cc + t01 * d + G.a_out * q01;
                p33 = t10 * cc + t11 * d + G.a_out * q11;
            }
            {
                const double t00 = a * p02 + b * p12, t01 = a * p03 + b * p13, t10 = cc * p02 + d * p12, t11 = cc * p03 + d * p13;
                p02 = t00 * a + t01 * b;
                p03 = t00 * cc + t01 * d;
                p12 = t10 * a + t11 * b;
                p13 = t10 * cc + t11 * d;
            }
        }
        xp = xk;
        ++k;
    }

    // one data point: time xk (not before the previous one, in the filter's direction), noise
    // variance vk, residual rk
    LFB_HD void step(const GpPars& G, double xk, double vk, double rk)
    {
        advance(G, xk);
        if (bad) return;
        const int gk = gap_at(G, xk);
        last_gap = gk;
        if (gk >= 0 && gk != cur_gap) {
            // a new gap: its process is independent of everything before
            cur_gap = gk;
            m2 = m3 = 0.0;
            p02 = p03 = p12 = p13 = 0.0;
            p22 = G.a_out;
            p23 = 0.0;
            p33 = G.a_out * c2;
        }
        const double g = gk >= 0 ? 1.0 : 0.0;
        vk += kGpWhiteNoise;
        if (!(fabs(rk) < 1e300) || !(vk > 0.0)) {
            bad = true;
            return;
        }
        // observation H = (1, 0, g, 0)
        const double h0 = p00 + g * p02, h1 = p01 + g * p12, h2 = p02 + g * p22, h3 = p03 + g * p23;  // P H^T
        const double s = h0 + g * h2 + vk;
        if (!(s > 0.0) || !(s < 1e300)) {
            bad = true;
            return;
        }
        double is = fast_rcp(s);
        is = is * fma(-s, is, 2.0);  // fast_rcp is good to ~1e-11; one more Newton step
        const double v = rk - (m0 + g * m2);
        const double k0 = h0 * is, k1 = h1 * is, k2 = h2 * is, k3 = h3 * is;
        m0 += k0 * v; m1 += k1 * v; m2 += k2 * v; m3 += k3 * v;
        p00 -= k0 * h0; p01 -= k0 * h1; p02 -= k0 * h2; p03 -= k0 * h3;
        p11 -= k1 * h1; p12 -= k1 * h2; p13 -= k1 * h3;
        p22 -= k2 * h2; p23 -= k2 * h3;
        p33 -= k3 * h3;
        // ln s summed as the ln of a running product, eight factors at a time (s is a variance of order
        // 1e-10..1; the guard keeps the product far from under- and overflow)
        ll -= 0.5 * (v * v * is + kLn2Pi);
        sprod *= s;
        if (++nprod == 8 || !(sprod > 1e-250) || !(sprod < 1e250)) flush();
    }

    LFB_HD void flush()
    {
        ll -= 0.5 * log(sprod);
        sprod = 1.0;
        nprod = 0;
    }

    // -inf for a non-finite residual or an invalid kernel (what quiet=True returns)
    LFB_HD double result()
    {
        flush();
        return (!bad && ll == ll) ? ll : -INFINITY;
    }
};

// ---- two filters that meet in the middle ----
// The filter is a serial recursion, so its latency is what a thin batch of walkers pays.  Given the
// state z at a time t*, the data before and after t* are independent: with p(z | y1) = N(m1, P1) from
// a forward filter over the first half, p(z | y2) = N(m2, P2) from a backward filter over the second
// half (both started from the prior N(0, P0)),
//     ln p(y) = ln p(y1) + ln p(y2) + ln Int N(z; m1, P1) N(z; m2, P2) / N(z; 0, P0) dz,
// and the two halves run on neighbouring lanes.  The integral is
//     1/2 [ln|P0| - ln|P1| - ln|P2| - ln|L|] - 1/2 m1' P1^-1 m1 - 1/2 m2' P2^-1 m2 + 1/2 e' L^-1 e,
//     L = P1^-1 + P2^-1 - P0^-1,  e = P1^-1 m1 + P2^-1 m2.
// If t* lies inside a gap, z is all four states; otherwise the gap processes on the two sides are
// different (independent) ones and only the global process (f1, f1') couples the halves.

// Cholesky of a symmetric positive definite N x N matrix in place (lower triangle); false if not PD
template <int N>
LFB_HD bool gp_chol(double (&A)[N][N])
{
    for (int j = 0; j < N; ++j) {
        double d = A[j][j];
        for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        A[j][j] = d;
        for (int i = j + 1; i < N; ++i) {
            double v = A[i][j];
            for (int k = 0; k < j; ++k) v -= A[i][k] * A[j][k];
            A[i][j] = v / d;
        }
    }
    return true;
}

// with the Cholesky factor L of P: P^-1 (full, symmetric) into Inv, returns ln|P|
template <int N>
LFB_HD double gp_inverse_from_chol(const double (&L)[N][N], double (&Inv)[N][N])
{
    double Li[N][N];  // L^-1, lower triangular
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Li[i][j] = 0.0;
    double ld = 0.0;
    for (int j = 0; j < N; ++j) {
        Li[j][j] = 1.0 / L[j][j];
        ld += log(L[j][j]);
        for (int i = j + 1; i < N; ++i) {
            double v = 0.0;
            for (int k = j; k < i; ++k) v -= L[i][k] * Li[k][j];
            Li[i][j] = v / L[i][i];
        }
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double v = 0.0;
            for (int k = (i > j ? i : j); k < N; ++k) v += Li[k][i] * Li[k][j];
            Inv[i][j] = v;
        }
    return 2.0 * ld;
}

template <int N>
LFB_HD double gp_merge_n(const double (&P1)[N][N], const double (&m1)[N], const double (&P2)[N][N], const double (&m2)[N],
                         const double (&p0)[N] /* diagonal prior */)
{
    double L1[N][N], L2[N][N], I1[N][N], I2[N][N], Lam[N][N], e[N], z[N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            L1[i][j] = P1[i][j];
            L2[i][j] = P2[i][j];
        }
    if (!gp_chol<N>(L1) || !gp_chol<N>(L2)) return -INFINITY;
    const double ld1 = gp_inverse_from_chol<N>(L1, I1), ld2 = gp_inverse_from_chol<N>(L2, I2);
    double ld0 = 0.0, q1 = 0.0, q2 = 0.0;
    for (int i = 0; i < N; ++i) {
        ld0 += log(p0[i]);
        double a = 0.0, b = 0.0;
        for (int j = 0; j < N; ++j) {
            Lam[i][j] = I1[i][j] + I2[i][j] - (i == j ? 1.0 / p0[i] : 0.0);
            a += I1[i][j] * m1[j];
            b += I2[i][j] * m2[j];
        }
        e[i] = a + b;
        q1 += m1[i] * a;
        q2 += m2[i] * b;
    }
    if (!gp_chol<N>(Lam)) return -INFINITY;
    double ldl = 0.0, qe = 0.0;
    for (int i = 0; i < N; ++i) {  // z = Lchol^-1 e; e' Lam^-1 e = z' z
        double v = e[i];
        for (int k = 0; k < i; ++k) v -= Lam[i][k] * z[k];
        z[i] = v / Lam[i][i];
        qe += z[i] * z[i];
        ldl += log(Lam[i][i]);
    }
    return 0.5 * (ld0 - ld1 - ld2 - 2.0 * ldl) - 0.5 * (q1 + q2) + 0.5 * qe;
}

// F: forward filter advanced to t*; B: backward filter at t* (it carries (f, -f')); couple_gap: t* lies
// inside a gap whose process both halves observed
LFB_HD double gp_merge(const GpFilter& F, const GpFilter& B, const GpPars& G, bool couple_gap)
{
    if (F.bad || B.bad) return -INFINITY;
    if (!couple_gap) {
        const double P1[2][2] = {{F.p00, F.p01}, {F.p01, F.p11}}, m1[2] = {F.m0, F.m1};
        const double P2[2][2] = {{B.p00, -B.p01}, {-B.p01, B.p11}}, m2[2] = {B.m0, -B.m1};
        const double p0[2] = {G.a_in, G.a_in * F.c2};
        return gp_merge_n<2>(P1, m1, P2, m2, p0);
    }
    const double P1[4][4] = {{F.p00, F.p01, F.p02, F.p03}, {F.p01, F.p11, F.p12, F.p13},
                             {F.p02, F.p12, F.p22, F.p23}, {F.p03, F.p13, F.p23, F.p33}};
    const double m1[4] = {F.m0, F.m1, F.m2, F.m3};
    const double P2[4][4] = {{B.p00, -B.p01, B.p02, -B.p03}, {-B.p01, B.p11, -B.p12, B.p13},
                             {B.p02, -B.p12, B.p22, -B.p23}, {-B.p03, B.p13, -B.p23, B.p33}};
    const double m2[4] = {B.m0, -B.m1, B.m2, -B.m3};
    const double p0[4] = {G.a_in, G.a_in * F.c2, G.a_out, G.a_out * F.c2};
    return gp_merge_n<4>(P1, m1, P2, m2, p0);
}

// ln L of residuals r(k), k = 0..n-1, taken at ascending times x(k) with noise variances var(k).
template <class FX, class FV, class FR>
LFB_HD double gp_loglike(int n, FX x, FV var, FR r, const GpPars& G)
{
    GpFilter F;
    F.init(G);
    for (int k = 0; k < n; ++k) F.step(G, x(k), var(k), r(k));
    return F.result();
}

// The same number from two filters meeting at point m = n / 2 (the kernel runs them on two lanes)
template <class FX, class FV, class FR>
LFB_HD double gp_loglike_two_sided(int n, FX x, FV var, FR r, const GpPars& G)
{
    if (n < 4) return gp_loglike(n, x, var, r, G);
    const int m = n / 2;
    GpFilter F, B;
    F.init(G, 1);
    B.init(G, -1);
    for (int k = 0; k < m; ++k) F.step(G, x(k), var(k), r(k));
    for (int k = n - 1; k >= m; --k) B.step(G, x(k), var(k), r(k));
    const double lf = F.result(), lb = B.result();
    F.advance(G, x(m));
    return lf + lb + gp_merge(F, B, G, F.last_gap >= 0 && F.last_gap == B.last_gap);
}

}  // namespace lfb
