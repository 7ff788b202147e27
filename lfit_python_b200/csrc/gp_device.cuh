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
    double g_lo, g_hi;  // the first gap that does not end before the current point (times ascend)
    int gi, cur_gap, k, nprod;
    bool bad;

    LFB_HD void init(const GpPars& G)
    {
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
        k = 0;
        xp = 0.0;
        ll = 0.0;
        sprod = 1.0;
        nprod = 0;
        gi = 0;
        g_lo = G.n_gaps > 0 ? G.gap[0][0] : INFINITY;
        g_hi = G.n_gaps > 0 ? G.gap[0][1] : INFINITY;
    }

    // one data point: time xk (not before the previous one), noise variance vk, residual rk
    LFB_HD void step(const GpPars& G, double xk, double vk, double rk)
    {
        if (bad) return;
        if (k > 0) {
            const double dt = xk - xp;
            if (!(dt >= 0.0)) {  // times must ascend
                bad = true;
                return;
            }
            const double ed = exp(-c * dt);
            const double a = ed * (1.0 + c * dt), b = ed * dt, cc = -ed * c2 * dt, d = ed * (1.0 - c * dt);
            // mean
            const double n0 = a * m0 + b * m1, n1 = cc * m0 + d * m1, n2 = a * m2 + b * m3, n3 = cc * m2 + d * m3;
            m0 = n0; m1 = n1; m2 = n2; m3 = n3;
            // process noise of a unit-variance Matern-3/2 state: Pinf - A Pinf A^T, Pinf = diag(1, c^2)
            const double q00 = 1.0 - (a * a + b * b * c2), q01 = -(a * cc + b * d * c2), q11 = c2 - (cc * cc + d * d * c2);
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
        while (xk > g_hi) {  // rarely: the point has left gap gi behind
            ++gi;
            g_lo = gi < G.n_gaps ? G.gap[gi][0] : INFINITY;
            g_hi = gi < G.n_gaps ? G.gap[gi][1] : INFINITY;
        }
        const int gk = xk >= g_lo ? gi : -1;
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

// ln L of residuals r(k), k = 0..n-1, taken at ascending times x(k) with noise variances var(k).
template <class FX, class FV, class FR>
LFB_HD double gp_loglike(int n, FX x, FV var, FR r, const GpPars& G)
{
    GpFilter F;
    F.init(G);
    for (int k = 0; k < n; ++k) F.step(G, x(k), var(k), r(k));
    return F.result();
}

}  // namespace lfb
