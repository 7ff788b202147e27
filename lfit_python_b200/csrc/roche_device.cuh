// roche_device.cuh -- FP64 Roche geometry for the sm_100a kernels.
//
// Replaces, on the device, what the reference calls in trm.roche and inside
// lfit (neither is under /root/reference; call sites CVModel.py:222,288,460,561
// and CVModel.py:138): L1, the inclination for a given eclipse width, the
// ballistic stream, and the ingress/egress phases of a surface element behind
// the donor's critical lobe.
//
// Frame: a = 1, white dwarf at the origin, donor at (1,0,0), mu = q/(1+q),
//   Phi = -(1-mu)/r1 - mu/r2 - ((x-mu)^2 + y^2)/2,
//   earth(th) = (si cos th, -si sin th, ci), th = 2 pi phase.
// An element sees the donor when the potential along its line of sight (LOS),
// inside the sphere of radius 1-xl1 about the donor, dips below Phi(L1).
//
// Everything here is written so that it also compiles as plain C++ (LFB_HD
// empty): tests/ builds a host harness from this header to check the device
// arithmetic on a machine without a GPU.  That harness is not part of the
// product library.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LFB_HD __host__ __device__ __forceinline__
#define LFB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define LFB_HD inline
#define LFB_HD_NOINLINE inline
#endif

#ifndef __CUDACC__
inline double rsqrt(double v) { return 1.0 / sqrt(v); }
#endif

// path counters of the element solve (host experiments only; nothing in a normal build)
#ifndef LFB_CNT
#define LFB_CNT(i)
#endif

#ifdef __CUDACC__
// how often the last-resort solver ran on the device (diagnostics: lfb_robust_calls)
__device__ unsigned long long g_robust_calls = 0ull;
#endif

namespace lfb {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kTwoPi = 6.28318530717958647692528676655900577;
constexpr double kDeg = kPi / 180.0;
constexpr double kBig = 1e30;

struct Roche {
    double mu, omu, xl1, rs, phic, rin;  // rin: 0.9 x polar radius of the lobe (inscribed sphere)
};

struct Point {  // surface element: p0 fixed in the rotating frame + (xi, eta) fixed on the sky
    double x, y, z, xi, eta;
};

struct Derivs {
    double S, St, Sl, Stt, Stl, Sll;
};

LFB_HD void sincos_(double a, double* s, double* c)
{
#ifdef __CUDA_ARCH__
    sincos(a, s, c);
#else
    *s = sin(a);
    *c = cos(a);
#endif
}

LFB_HD double clampd(double v, double lim) { return v > lim ? lim : (v < -lim ? -lim : v); }

LFB_HD double pot(const Roche& R, double x, double y, double z)
{
    double yz = y * y + z * z, x2 = x - 1.0, xc = x - R.mu;
    return -R.omu / sqrt(x * x + yz) - R.mu / sqrt(x2 * x2 + yz) - 0.5 * (xc * xc + y * y);
}

// L1: Newton on dPhi/dx along the line of centres from the Hill-radius estimate.
LFB_HD bool roche_init(double q, Roche& R)
{
    if (!(q > 0.0) || !(q < 1e6)) return false;
    double mu = q / (1.0 + q), omu = 1.0 - mu;
    double m = mu < 0.5 ? mu : omu;
    double h = cbrt(m / 3.0);
    double rh = h * (1.0 - h / 3.0 - h * h / 9.0);
    double x = mu < 0.5 ? 1.0 - rh : rh;
    for (int it = 0; it < 10; ++it) {
        double omx = 1.0 - x;
        double f = omu / (x * x) - mu / (omx * omx) - (x - mu);
        double fp = -2.0 * omu / (x * x * x) - 2.0 * mu / (omx * omx * omx) - 1.0;
        const double dx = f / fp;
        x -= dx;
        x = x < 1e-4 ? 1e-4 : (x > 1.0 - 1e-4 ? 1.0 - 1e-4 : x);
        if (fabs(dx) < 1e-16) break;  // converged (quadratically: the step before was ~1e-8)
    }
    R.mu = mu;
    R.omu = omu;
    R.xl1 = x;
    R.rs = 1.0 - x;
    R.phic = pot(R, x, 0.0, 0.0);
    // polar radius of the critical lobe, Phi(1, 0, z) = Phi_c, approached from inside
    double q13 = cbrt(q), q23 = q13 * q13;
    double z = 0.8 * 0.49 * q23 / (0.6 * q23 + log(1.0 + q13));
    double cst = 0.5 * omu * omu + R.phic;
    for (int it = 0; it < 12; ++it) {
        double r1sq = 1.0 + z * z, ir1 = 1.0 / sqrt(r1sq);
        double f = -omu * ir1 - mu / z - cst;
        double fp = omu * z * ir1 * ir1 * ir1 + mu / (z * z);
        const double dz = f / fp;
        z -= dz;
        if (fabs(dz) < 1e-16) break;
    }
    R.rin = 0.9 * z;
    return true;
}

// Potential along LOS from the white-dwarf centre with sin(i) = u at cos(th) = c.
struct OriginPot {
    double P, Pu, Pc, Pl, Pll, Pul, Pcl;
};
LFB_HD void origin_pot(const Roche& R, double u, double c, double lam, OriginPot& o)
{
    double mu = R.mu;
    double Dd = 1.0 + lam * lam - 2.0 * lam * u * c;
    double isq = 1.0 / sqrt(Dd), i3 = isq * isq * isq, i5 = i3 * isq * isq;
    double Dl = 2.0 * lam - 2.0 * u * c, Du = -2.0 * lam * c, Dc = -2.0 * lam * u;
    double il = 1.0 / lam;
    o.P = -R.omu * il - mu * isq - 0.5 * lam * lam * u * u + mu * lam * u * c - 0.5 * mu * mu;
    o.Pl = R.omu * il * il + 0.5 * mu * i3 * Dl - lam * u * u + mu * u * c;
    o.Pu = 0.5 * mu * i3 * Du - lam * lam * u + mu * lam * c;
    o.Pc = 0.5 * mu * i3 * Dc + mu * lam * u;
    o.Pll = -2.0 * R.omu * il * il * il + 0.5 * mu * (-1.5 * i5 * Dl * Dl + 2.0 * i3) - u * u;
    o.Pul = 0.5 * mu * (-1.5 * i5 * Dl * Du - 2.0 * c * i3) - 2.0 * lam * u + mu * c;
    o.Pcl = 0.5 * mu * (-1.5 * i5 * Dl * Dc - 2.0 * u * i3) + mu * u;
}

// roche.findphi(q, 90): full phase width of the eclipse of the white-dwarf centre seen edge-on.
LFB_HD double findphi90(const Roche& R)
{
    double rl = R.rin / 0.9;
    double c = sqrt(1.0 - rl * rl), lam = c;
    OriginPot o;
    for (int it = 0; it < 12; ++it) {
        origin_pot(R, 1.0, c, lam, o);
        double F1 = o.P - R.phic, F2 = o.Pl;
        double det = o.Pc * o.Pll - o.Pl * o.Pcl;
        const double dc = (-F1 * o.Pll + F2 * o.Pl) / det, dl = (-o.Pc * F2 + o.Pcl * F1) / det;
        c += dc;
        lam += dl;
        if (fabs(dc) < 1e-16 && fabs(dl) < 1e-15) break;
    }
    return acos(c) / kPi;
}

// roche.findi(q, dphi): sin(i) for which the white-dwarf centre is eclipsed for dphi of the orbit.
LFB_HD bool findi(const Roche& R, double dphi, double maxphi, double& sini)
{
    if (!(dphi > 0.0) || !(dphi < maxphi)) return false;
    double c = cos(kPi * dphi);
    double u = cos(kPi * maxphi) / c, lam = u * c;
    OriginPot o;
    for (int it = 0; it < 12; ++it) {
        origin_pot(R, u, c, lam, o);
        double F1 = o.P - R.phic, F2 = o.Pl;
        double det = o.Pu * o.Pll - o.Pl * o.Pul;
        const double du = (-F1 * o.Pll + F2 * o.Pl) / det, dl = (-o.Pu * F2 + o.Pul * F1) / det;
        u += du;
        lam += dl;
        if (fabs(du) < 1e-16 && fabs(dl) < 1e-15) break;
    }
    if (!(u > 0.0) || !(u <= 1.0)) return false;
    sini = u;
    return true;
}

// 1/sqrt(x) for the well-scaled squared distances met here (no zero / denormal / inf handling):
// hardware seed + two Newton steps, a few ulp.
LFB_HD double fast_rsqrt(double x)
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}

// reciprocal good to ~1e-11 relative: enough for a Newton step (the residual, not the step,
// decides where the iteration converges to)
LFB_HD double fast_rcp(double x)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * fma(-x, r, 2.0);
    r = r * fma(-x, r, 2.0);
    return r;
#else
    return 1.0 / x;
#endif
}

// Potential and its derivatives over the (th, lam) family of LOS of one element, at the
// orbital angle whose cosine and sine are (c, s).
// PLANAR: the element lies in the orbital plane with no offset on the sky (disc, bright-spot strip: z = xi = eta
// = 0) -- the terms that vanish are left out (the compiler may not drop a multiplication by a 0.0 it cannot
// prove finite-safe).
template <bool PLANAR = false>
LFB_HD void ray_eval(const Roche& R, double si, double ci, const Point& T, double c, double s, double lam, Derivs& D)
{
    LFB_CNT(9);
    double ex = si * c, ey = -si * s;
    double dx = PLANAR ? lam * ex : fma(lam, ex, -T.xi * s - T.eta * ci * c);
    double dy = PLANAR ? lam * ey : fma(lam, ey, -T.xi * c + T.eta * ci * s);
    double dz = PLANAR ? lam * ci : fma(lam, ci, T.eta * si);
    double x = T.x + dx, y = T.y + dy, z = PLANAR ? dz : T.z + dz;
    double tx = dy, ty = -dx;  // d/dth of the part that turns with the observer
    double x2 = x - 1.0;
    double yz = y * y + z * z;
    double ir1 = fast_rsqrt(x * x + yz), ir2 = fast_rsqrt(x2 * x2 + yz);
    double a1 = R.omu * ir1 * ir1 * ir1, a2 = R.mu * ir2 * ir2 * ir2;
    double b1 = 3.0 * a1 * ir1 * ir1, b2 = 3.0 * a2 * ir2 * ir2;
    double a12 = a1 + a2, xc = x - R.mu;
    double gx = a1 * x + a2 * x2 - xc, gy = a12 * y - y, gz = a12 * z;
    double x_e = x * ex + y * ey + z * ci, d_e = x_e - ex;
    double x_t = x * tx + y * ty, d_t = x_t - tx;
    double t_t = tx * tx + ty * ty, e_t = ex * tx + ey * ty, e_xy = ex * ex + ey * ey;
    D.S = -R.omu * ir1 - R.mu * ir2 - 0.5 * (xc * xc + y * y);
    D.St = gx * tx + gy * ty;
    D.Sl = gx * ex + gy * ey + gz * ci;
    D.Sll = a12 - b1 * x_e * x_e - b2 * d_e * d_e - e_xy;
    D.Stt = (a12 - 1.0) * t_t - b1 * x_t * x_t - b2 * d_t * d_t - (gx * dx + gy * dy);
    D.Stl = (a12 - 1.0) * e_t - b1 * x_e * x_t - b2 * d_e * d_t + (gx * ey - gy * ex);
}

// Potential and its derivative along the LOS only (the residuals of the grazing-LOS system): what a Newton
// step with a frozen Jacobian needs -- about half of ray_eval.
template <bool PLANAR = false>
LFB_HD void ray_resid(const Roche& R, double si, double ci, const Point& T, double c, double s, double lam, double& S,
                      double& Sl)
{
    LFB_CNT(11);
    double ex = si * c, ey = -si * s;
    double x = PLANAR ? fma(lam, ex, T.x) : T.x + fma(lam, ex, -T.xi * s - T.eta * ci * c);
    double y = PLANAR ? fma(lam, ey, T.y) : T.y + fma(lam, ey, -T.xi * c + T.eta * ci * s);
    double z = PLANAR ? lam * ci : T.z + fma(lam, ci, T.eta * si);
    double x2 = x - 1.0;
    double yz = y * y + z * z;
    double ir1 = fast_rsqrt(x * x + yz), ir2 = fast_rsqrt(x2 * x2 + yz);
    double a1 = R.omu * ir1 * ir1 * ir1, a2 = R.mu * ir2 * ir2 * ir2;
    double a12 = a1 + a2, xc = x - R.mu;
    double gx = a1 * x + a2 * x2 - xc, gy = a12 * y - y, gz = a12 * z;
    S = -R.omu * ir1 - R.mu * ir2 - 0.5 * (xc * xc + y * y);
    Sl = gx * ex + gy * ey + gz * ci;
}

// (c, s) <- rotation by d radians, |d| <= 0.25: Taylor to d^13 (below 1e-17), no range reduction
LFB_HD void rotate_cs(double& c, double& s, double d)
{
    double d2 = d * d;
    if (d2 < 1e-6) {
        // late Newton steps: d^7 terms are below 1e-21
        double sd = d * fma(d2, fma(d2, 1.0 / 120.0, -1.0 / 6.0), 1.0);
        double cd = fma(d2, fma(d2, fma(d2, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0);
        double cn = c * cd - s * sd;
        s = s * cd + c * sd;
        c = cn;
        return;
    }
    double sd = d * fma(d2, fma(d2, fma(d2, fma(d2, fma(d2, fma(d2, 1.0 / 6227020800.0, -1.0 / 39916800.0), 1.0 / 362880.0),
                                                    -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
    double cd = fma(d2, fma(d2, fma(d2, fma(d2, fma(d2, fma(d2, 1.0 / 479001600.0, -1.0 / 3628800.0), 1.0 / 40320.0),
                                            -1.0 / 720.0), 1.0 / 24.0), -0.5), 1.0);
    double cn = c * cd - s * sd;
    s = s * cd + c * sd;
    c = cn;
}

// Angle of the unit vector (c, s), in (-pi, pi] give or take a rounding: what atan2(s, c) returns, for a third
// of its instructions.  An FP32 arctangent picks the nearest multiple of pi/64, whose cosine and sine come
// from a table; the rest is the arcsine series of a sine below 0.025 (terms to x^11: 5e-20).
#ifdef __CUDACC__
struct AngleRow {
    double c, s, a;
};
__device__ const AngleRow g_angle_rows[129] = {
#include "angle_table.inc"
};
#endif
LFB_HD double angle_of(double c, double s)
{
#ifdef __CUDA_ARCH__
    const float af = atan2f((float)s, (float)c);
    const int k = __float2int_rn(af * 20.371832715762602f);  // 64 / pi
    const AngleRow row = g_angle_rows[k + 64];
    const double x = s * row.c - c * row.s;  // sin(angle - row.a), |angle - row.a| < pi/128 + 1e-6
    const double x2 = x * x;
    const double p = fma(x2, fma(x2, fma(x2, fma(x2, 945.0 / 42240.0, 105.0 / 3456.0), 15.0 / 336.0), 3.0 / 40.0), 1.0 / 6.0);
    return row.a + fma(x * x2, p, x);
#else
    return atan2(s, c);
#endif
}

// min over the chord of the LOS inside the bounding sphere of Phi - Phi_c (robust path only)
LFB_HD_NOINLINE double chord_min(const Roche& R, double si, double ci, const Point& T, double th)
{
    double s, c;
    sincos_(th, &s, &c);
    double ex = si * c, ey = -si * s, ez = ci;
    double ox = T.x - T.xi * s - T.eta * ci * c, oy = T.y - T.xi * c + T.eta * ci * s, oz = T.z + T.eta * si;
    double wx = 1.0 - ox, wy = -oy, wz = -oz;
    double b = wx * ex + wy * ey + wz * ez;
    double d2 = wx * wx + wy * wy + wz * wz - b * b;
    if (d2 >= R.rs * R.rs) return 1.0;
    double half = sqrt(R.rs * R.rs - d2);
    double l1 = b - half, l2 = b + half;
    if (l1 < 0.0) l1 = 0.0;
    if (l2 <= l1) return 1.0;
    const int NS = 48;
    double best = 1e300;
    int kb = 0;
    for (int k = 0; k < NS; ++k) {
        double lam = l1 + (l2 - l1) * k / (NS - 1);
        double v = pot(R, ox + lam * ex, oy + lam * ey, oz + lam * ez);
        if (v < best) { best = v; kb = k; }
    }
    int ka = kb > 0 ? kb - 1 : 0, kc = kb < NS - 1 ? kb + 1 : NS - 1;
    double a = l1 + (l2 - l1) * ka / (NS - 1), cc = l1 + (l2 - l1) * kc / (NS - 1);
    const double gr = 0.6180339887498949;
    double x1 = cc - gr * (cc - a), x2 = a + gr * (cc - a);
    double f1 = pot(R, ox + x1 * ex, oy + x1 * ey, oz + x1 * ez);
    double f2 = pot(R, ox + x2 * ex, oy + x2 * ey, oz + x2 * ez);
    for (int it = 0; it < 70; ++it) {
        if (f1 < f2) {
            cc = x2; x2 = x1; f2 = f1;
            x1 = cc - gr * (cc - a);
            f1 = pot(R, ox + x1 * ex, oy + x1 * ey, oz + x1 * ez);
        } else {
            a = x1; x1 = x2; f1 = f2;
            x2 = a + gr * (cc - a);
            f2 = pot(R, ox + x2 * ex, oy + x2 * ey, oz + x2 * ez);
        }
    }
    double v = f1 < f2 ? f1 : f2;
    if (best < v) v = best;
    return v - R.phic;
}

// Last-resort ingress/egress: phase scan + golden section + bisection.  Taken by
// about one element in 10^6; kept out of line so it costs no registers on the fast path.
LFB_HD_NOINLINE int ingress_egress_robust(const Roche& R, double si, double ci, const Point& T, double* ph_in,
                                          double* ph_out)
{
    LFB_CNT(8);
#ifdef __CUDA_ARCH__
    atomicAdd(&g_robust_calls, 1ull);
#endif
    const int NSCAN = 384;
    double psi = atan2(T.y, 1.0 - T.x);
    double half = 0.5 * kPi, step = 2.0 * half / (NSCAN - 1), th_lo = psi - half;
    int kb = 0;
    double best = 1e300;
    for (int k = 0; k < NSCAN; ++k) {
        double g = chord_min(R, si, ci, T, th_lo + step * k);
        if (g < best) { best = g; kb = k; }
    }
    if (kb == 0 || kb == NSCAN - 1) return 0;
    double a = th_lo + step * (kb - 1), cc = th_lo + step * (kb + 1);
    const double gr = 0.6180339887498949;
    double x1 = cc - gr * (cc - a), x2 = a + gr * (cc - a);
    double f1 = chord_min(R, si, ci, T, x1), f2 = chord_min(R, si, ci, T, x2);
    for (int it = 0; it < 60; ++it) {
        if (f1 < f2) {
            cc = x2; x2 = x1; f2 = f1;
            x1 = cc - gr * (cc - a);
            f1 = chord_min(R, si, ci, T, x1);
        } else {
            a = x1; x1 = x2; f1 = f2;
            x2 = a + gr * (cc - a);
            f2 = chord_min(R, si, ci, T, x2);
        }
    }
    double thm = f1 < f2 ? x1 : x2, gm = f1 < f2 ? f1 : f2;
    if (best < gm) { gm = best; thm = th_lo + step * kb; }
    if (!(gm < 0.0)) return 0;
    int kl = (int)floor((thm - th_lo) / step);
    while (kl > 0 && chord_min(R, si, ci, T, th_lo + step * kl) < 0.0) --kl;
    int kr = (int)ceil((thm - th_lo) / step);
    while (kr < NSCAN - 1 && chord_min(R, si, ci, T, th_lo + step * kr) < 0.0) ++kr;
    if (chord_min(R, si, ci, T, th_lo + step * kl) < 0.0 || chord_min(R, si, ci, T, th_lo + step * kr) < 0.0)
        return 0;
    double lo = th_lo + step * kl, hi = thm;
    for (int it = 0; it < 90; ++it) {
        double m = 0.5 * (lo + hi);
        if (chord_min(R, si, ci, T, m) < 0.0) hi = m; else lo = m;
    }
    *ph_in = 0.5 * (lo + hi) / kTwoPi;
    lo = thm;
    hi = th_lo + step * kr;
    for (int it = 0; it < 90; ++it) {
        double m = 0.5 * (lo + hi);
        if (chord_min(R, si, ci, T, m) < 0.0) lo = m; else hi = m;
    }
    *ph_out = 0.5 * (lo + hi) / kTwoPi;
    return 1;
}

// ---- single-precision warm-up of the grazing-LOS Newton ----
// The FP64 pipe is what bounds the element solves.  The first Newton steps only have to get
// near the root, so they run in FP32 (its own pipe, twice the rate); the FP64 iteration then
// starts ~1e-6 from the root and converges in two or three steps instead of eight.  The root
// the FP64 iteration converges to is unchanged -- FP32 only picks the starting point, and a
// warm-up that misbehaves is dropped.
struct DerivsF {
    float S, St, Sl, Stt, Stl, Sll;
};
struct PointF {
    float x, y, z, xi, eta;
};

LFB_HD float rsqrt_f(float x)
{
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

// approximate FP32 reciprocal (one MUFU instruction; the warm-up only picks starting points)
LFB_HD float rcp_f(float x)
{
#ifdef __CUDA_ARCH__
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}

template <bool PLANAR = false>
LFB_HD void ray_eval_f(float mu, float omu, float si, float ci, const PointF& T, float c, float s, float lam, DerivsF& D)
{
    LFB_CNT(10);
    float ex = si * c, ey = -si * s;
    float dx = PLANAR ? lam * ex : fmaf(lam, ex, -T.xi * s - T.eta * ci * c);
    float dy = PLANAR ? lam * ey : fmaf(lam, ey, -T.xi * c + T.eta * ci * s);
    float dz = PLANAR ? lam * ci : fmaf(lam, ci, T.eta * si);
    float x = T.x + dx, y = T.y + dy, z = PLANAR ? dz : T.z + dz;
    float tx = dy, ty = -dx;
    float x2 = x - 1.0f;
    float yz = y * y + z * z;
    float ir1 = rsqrt_f(x * x + yz), ir2 = rsqrt_f(x2 * x2 + yz);
    float a1 = omu * ir1 * ir1 * ir1, a2 = mu * ir2 * ir2 * ir2;
    float b1 = 3.0f * a1 * ir1 * ir1, b2 = 3.0f * a2 * ir2 * ir2;
    float a12 = a1 + a2, xc = x - mu;
    float gx = a1 * x + a2 * x2 - xc, gy = a12 * y - y, gz = a12 * z;
    float x_e = x * ex + y * ey + z * ci, d_e = x_e - ex;
    float x_t = x * tx + y * ty, d_t = x_t - tx;
    float t_t = tx * tx + ty * ty, e_t = ex * tx + ey * ty, e_xy = ex * ex + ey * ey;
    D.S = -omu * ir1 - mu * ir2 - 0.5f * (xc * xc + y * y);
    D.St = gx * tx + gy * ty;
    D.Sl = gx * ex + gy * ey + gz * ci;
    D.Sll = a12 - b1 * x_e * x_e - b2 * d_e * d_e - e_xy;
    D.Stt = (a12 - 1.0f) * t_t - b1 * x_t * x_t - b2 * d_t * d_t - (gx * dx + gy * dy);
    D.Stl = (a12 - 1.0f) * e_t - b1 * x_e * x_t - b2 * d_e * d_t + (gx * ey - gy * ex);
}

// (c, s) <- rotation by d radians, |d| <= 0.2, to single precision
LFB_HD void rotate_cs_f(float& c, float& s, float d)
{
    float d2 = d * d;
    float sd = d * fmaf(d2, fmaf(d2, fmaf(d2, -1.0f / 5040.0f, 1.0f / 120.0f), -1.0f / 6.0f), 1.0f);
    float cd = fmaf(d2, fmaf(d2, fmaf(d2, -1.0f / 720.0f, 1.0f / 24.0f), -0.5f), 1.0f);
    float cn = c * cd - s * sd;
    s = s * cd + c * sd;
    c = cn;
}

constexpr int kWarmIters = 10;

// FP32 Newton towards the grazing LOS on side sg of the deepest LOS (cm, sm); true if it settled.
template <bool PLANAR = false>
LFB_HD bool warm_root(float mu, float omu, float phic, float si, float ci, const PointF& T, float cm, float sm, float sg,
                      float& c, float& s, float& lam)
{
    DerivsF D;
    for (int it = 0; it < kWarmIters; ++it) {
        LFB_CNT(13);
        ray_eval_f<PLANAR>(mu, omu, si, ci, T, c, s, lam, D);
        float F1 = D.S - phic, F2 = D.Sl;
        float idet = rcp_f(D.St * D.Sll - D.Sl * D.Stl);
        float dth = (-F1 * D.Sll + F2 * D.Sl) * idet, dl = (-D.St * F2 + D.Stl * F1) * idet;
        dth = dth > 0.2f ? 0.2f : (dth < -0.2f ? -0.2f : dth);
        dl = dl > 0.2f ? 0.2f : (dl < -0.2f ? -0.2f : dl);
        float cross = s * cm - c * sm;
        if (!(sg * (cross + dth * (c * cm + s * sm)) > 0.0f)) {
            dth = -0.5f * asinf(cross > 1.0f ? 1.0f : (cross < -1.0f ? -1.0f : cross));
            dl *= 0.5f;
        }
        rotate_cs_f(c, s, dth);
        lam += dl;
        if (!(fabsf(dth) < 1.0f)) return false;  // NaN
#ifndef LFB_WARM_TOL
#define LFB_WARM_TOL 3e-4f
#endif
        if (fabsf(dth) < LFB_WARM_TOL && fabsf(dl) < 10.0f * LFB_WARM_TOL) return true;
    }
    return false;
}

// FP32 search for the deepest LOS of an element (the minimum over (th, lam) of the potential along its lines of
// sight), from the conjunction LOS: Newton on lam alone first, then on both.  1: settled, g0 = depth below the
// critical potential there (single precision); 0: it did not settle -- the caller searches in FP64 from scratch.
template <bool PLANAR = false>
LFB_HD int warm_min(float mu, float omu, float phic, float si, float ci, const PointF& T, float& c, float& s, float& lam,
                    float& g0, DerivsF& D)
{
    for (int it = 0; it < 5; ++it) {
        LFB_CNT(14);
        ray_eval_f<PLANAR>(mu, omu, si, ci, T, c, s, lam, D);
        if (!(D.Sll > 0.0f)) return 0;
        float dl = -D.Sl * rcp_f(D.Sll);
        dl = dl > 0.1f ? 0.1f : (dl < -0.1f ? -0.1f : dl);
        lam += dl;
        if (fabsf(dl) < 1e-3f) break;
    }
    for (int it = 0; it < 12; ++it) {
        ray_eval_f<PLANAR>(mu, omu, si, ci, T, c, s, lam, D);
        float det = D.Stt * D.Sll - D.Stl * D.Stl;
        if (!(D.Sll > 0.0f) || !(det > 0.0f)) return 0;
        float idet = rcp_f(det);
        float dth = -(D.St * D.Sll - D.Sl * D.Stl) * idet, dl = -(D.Sl * D.Stt - D.St * D.Stl) * idet;
        dth = dth > 0.1f ? 0.1f : (dth < -0.1f ? -0.1f : dth);
        dl = dl > 0.1f ? 0.1f : (dl < -0.1f ? -0.1f : dl);
        rotate_cs_f(c, s, dth);
        lam += dl;
        if (!(fabsf(dth) < 1.0f)) return 0;  // NaN
        if (fabsf(dth) < 1e-4f && fabsf(dl) < 1e-4f) {
            g0 = D.S - phic;  // one small step old: good to ~1e-6
            return 1;
        }
    }
    return 0;
}

constexpr float kWarmRejectMargin = 3e-4f;  // depth above which the FP32 search alone says "never eclipsed"
constexpr int kMinIters = 12;   // deepest-LOS Newton (shallow elements only), early exit
constexpr int kRootIters = 16;  // grazing-LOS Newton, early exit
#ifndef LFB_FROZEN_D0
#define LFB_FROZEN_D0 2e-4
#endif
constexpr double kFrozenBelow = LFB_FROZEN_D0;  // Newton steps below this are followed by frozen-Jacobian steps

// The two grazing lines of sight of an element: orbital angle as (cos, sin) and distance along the LOS
struct Roots {
    double c[2], s[2], lam[2];
};

// 2-D Newton on (th, lam) for the grazing LOS (Phi = Phi_c, dPhi/dlam = 0) on either side of the
// deepest LOS (cm, sm), from the given starts (FP32 warm-up, then FP64).  True if both converged to
// grazing LOS of the right kind; res = their orbital angles (radians), ingress then egress.
template <bool PLANAR = false>
LFB_HD bool graze_roots(const Roche& R, double si, double ci, const Point& T, double cpsi, double spsi, double cm,
                        double sm, const Roots& start, double res[2], Roots* roots)
{
    Derivs D;
    Roots found;
    const float muf = (float)R.mu, omuf = (float)R.omu, phicf = (float)R.phic, sif = (float)si, cif = (float)ci;
    const PointF Tf = {(float)T.x, (float)T.y, (float)T.z, (float)T.xi, (float)T.eta};
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const double sg = side ? 1.0 : -1.0;
        double c = start.c[side], s = start.s[side], lam = start.lam[side];
#ifndef LFB_NO_WARMUP
        {
            float cf = (float)c, sf = (float)s, lf = (float)lam;
            if (warm_root<PLANAR>(muf, omuf, phicf, sif, cif, Tf, (float)cm, (float)sm, (float)sg, cf, sf, lf)) {
                const double cw = (double)cf, sw = (double)sf;
                const double nrm = fast_rsqrt(cw * cw + sw * sw);
                c = cw * nrm;
                s = sw * nrm;
                lam = (double)lf;
            }
        }
#endif
        bool conv = false;
        for (int it = 0; it < kRootIters; ++it) {
            ray_eval<PLANAR>(R, si, ci, T, c, s, lam, D);
            double F1 = D.S - R.phic, F2 = D.Sl;
            double idet = fast_rcp(D.St * D.Sll - D.Sl * D.Stl);
            double dth = clampd((-F1 * D.Sll + F2 * D.Sl) * idet, 0.2);
            double dl = clampd((-D.St * F2 + D.Stl * F1) * idet, 0.2);
            // stay on this side of the deepest LOS: sin(th + dth - thm) must keep the sign of sg
            double cross = s * cm - c * sm;  // sin(th - thm)
            bool newton_step = true;
            if (!(sg * (cross + dth * (c * cm + s * sm)) > 0.0)) {
                dth = -0.5 * asin(cross > 1.0 ? 1.0 : (cross < -1.0 ? -1.0 : cross));
                dl *= 0.5;
                newton_step = false;  // (pushed back from the dividing LOS: a small step here is no convergence)
            }
            rotate_cs(c, s, dth);
            lam += dl;
            // quadratic convergence: a Newton step below 1e-9 leaves an error far below 1e-15
            if (newton_step && fabs(dth) < 1e-9 && fabs(dl) < 1e-7) { conv = true; break; }
#ifndef LFB_NO_FROZEN_STEP
            // Close to the root (where the FP32 warm-up leaves the iteration) the next steps do not need a new
            // Jacobian: with the one just used the error shrinks by ~ |Newton step| per step, and each such step
            // costs the residuals only.  The contraction is measured (size of a step over the one before), and
            // the iteration ends when contraction x step -- the error left behind -- is below 1e-15.
            if (newton_step && fabs(dth) < kFrozenBelow && fabs(dl) < 10.0 * kFrozenBelow) {
                double before = fabs(dth) + 0.1 * fabs(dl);
                for (int k = 0; k < 2; ++k) {
                    double S2, Sl2;
                    ray_resid<PLANAR>(R, si, ci, T, c, s, lam, S2, Sl2);
                    const double G1 = S2 - R.phic;
                    const double dth2 = (-G1 * D.Sll + Sl2 * D.Sl) * idet, dl2 = (-D.St * Sl2 + D.Stl * G1) * idet;
                    const double now = fabs(dth2) + 0.1 * fabs(dl2);
                    if (!(now < 0.01 * before)) break;  // not what a frozen step near the root does: full steps again
                    rotate_cs(c, s, dth2);
                    lam += dl2;
                    if (now * now < 1e-15 * before) { conv = true; break; }
                    before = now;
                }
                if (conv) break;
            }
#endif
        }
        // accept only a converged grazing LOS of the right kind (D is one or two tiny steps old)
        double xx = T.x - T.xi * s - T.eta * ci * c + lam * si * c - 1.0;
        double yy = T.y - T.xi * c + T.eta * ci * s - lam * si * s;
        double zz = T.z + T.eta * si + lam * ci;
        bool ok = conv && D.Sll > 0.0 && lam > 0.0 && xx * xx + yy * yy + zz * zz <= R.rs * R.rs &&
                  (side ? D.St > 0.0 : D.St < 0.0) && c * cpsi + s * spsi > 0.0;
        if (!ok) return false;
        res[side] = angle_of(c, s);
        found.c[side] = c;
        found.s[side] = s;
        found.lam[side] = lam;
    }
    if (!(res[0] < res[1])) return false;
    if (roots) *roots = found;
    return true;
}

// Ingress/egress phases (cycles) of one element; 0 if it is never eclipsed.
// Fast path: 2-D Newton on (th, lam) for the two LOS that graze the critical surface
// (Phi = Phi_c, dPhi/dlam = 0), started from the tangents to a sphere inscribed in the lobe
// (deep elements) or from the osculating parabola at the deepest LOS (shallow elements).  The
// orbital angle is carried as (cos, sin) and advanced by small rotations, so the loop has no
// trigonometric calls; one atan2 per root at the end.  Every root is verified; anything
// unverified goes to ingress_egress_robust.
// hint: grazing LOS of the element at the origin (the white-dwarf centre, for white-dwarf tiles): tried
// first as the Newton starts when its eclipse is much wider than this element's offset; roots: where
// the element's own end up (fast path only).
template <bool PLANAR = false>
LFB_HD int ingress_egress(const Roche& R, double si, double ci, const Point& T, double* ph_in, double* ph_out,
                          const Roots* hint = nullptr, Roots* roots = nullptr)
{
    // conjunction: the LOS passes closest to the donor's centre
    const double X = 1.0 - T.x + T.eta * ci, Y = T.y - T.xi;
    const double ih = rsqrt(X * X + Y * Y);
    const double cpsi = X * ih, spsi = Y * ih;
    double c = cpsi, s = spsi;
    double ex = si * c, ey = -si * s;
    double wx = 1.0 - (T.x - T.xi * s - T.eta * ci * c);
    double wy = -(T.y - T.xi * c + T.eta * ci * s);
    double wz = -(T.z + T.eta * si);
    double lam = wx * ex + wy * ey + wz * ci;
    double w2 = wx * wx + wy * wy + wz * wz;
    LFB_CNT(0);
    if (w2 - lam * lam >= R.rs * R.rs || lam <= 0.0) { LFB_CNT(1); return 0; }
    double res[2];
    if (hint) {
        // The hint's LOS graze the lobe where this element's do, give or take the element's offset; and
        // the deepest LOS, which divides ingress from egress, is within that offset of conjunction.  When
        // the hinted eclipse is several offsets wide none of this can be confused: solve from the hint.
        const double off = sqrt(T.xi * T.xi + T.eta * T.eta + T.x * T.x + T.y * T.y + T.z * T.z);
        if (-hint->s[0] > 4.0 * off && hint->s[1] > 4.0 * off && hint->c[0] > 0.0 && hint->c[1] > 0.0 &&
            graze_roots<PLANAR>(R, si, ci, T, cpsi, spsi, cpsi, spsi, *hint, res, roots)) {
            LFB_CNT(2);
            *ph_in = res[0] * (1.0 / kTwoPi);
            *ph_out = res[1] * (1.0 / kTwoPi);
            return 1;
        }
    }
    double rxy = si * sqrt(wx * wx + wy * wy);
#ifndef LFB_START_F
#define LFB_START_F 1.15
#endif
    const double rstart = R.rin * LFB_START_F;
    double tang = sqrt(w2 - rstart * rstart);
    double cosd = (tang - wz * ci) / rxy;
    Derivs D;
    double c0, s0, c1, s1, lam0, lam1, cm = cpsi, sm = spsi;
    // How the search for the starts ends: -1 starts found, 0 never eclipsed, 2 hand over to the robust solver.
    // (No return inside the two branches: all lanes of a warp meet again before the grazing-LOS Newton, which is
    // the expensive part and the same code for deep and shallow elements.)
    int verdict = -1;
    c0 = c1 = s0 = s1 = lam0 = lam1 = 0.0;
    if (cosd < 0.995) {
        // the conjunction LOS passes well inside the inscribed sphere: start at its tangents
        LFB_CNT(3);
        cosd = cosd > -1.0 ? cosd : -1.0;
        double sind = sqrt(1.0 - cosd * cosd);
        c0 = cpsi * cosd + spsi * sind;  // psi - del
        s0 = spsi * cosd - cpsi * sind;
        c1 = cpsi * cosd - spsi * sind;  // psi + del
        s1 = spsi * cosd + cpsi * sind;
        lam0 = lam1 = tang;
    } else {
        LFB_CNT(4);
        bool warmed = false, starts = false;
#ifndef LFB_NO_WARMUP
        {
            // the search for the deepest LOS runs in FP32 first: an element that clears the lobe by a wide margin
            // is done, the others hand the FP64 iteration a start ~1e-4 from the minimum
            const PointF Tf = {(float)T.x, (float)T.y, (float)T.z, (float)T.xi, (float)T.eta};
            float cf = (float)c, sf = (float)s, lf = (float)lam, g0f = 0.0f;
            DerivsF Df;
            if (warm_min<PLANAR>((float)R.mu, (float)R.omu, (float)R.phic, (float)si, (float)ci, Tf, cf, sf, lf, g0f, Df)) {
                if (g0f > kWarmRejectMargin) {
                    LFB_CNT(6);
                    verdict = 0;
                } else {
                    const double cw = (double)cf, sw = (double)sf;
                    const double nrm = fast_rsqrt(cw * cw + sw * sw);
                    c = cw * nrm;
                    s = sw * nrm;
                    lam = (double)lf;
                    warmed = true;
#ifndef LFB_NO_WARM_STARTS
                    // An element this far inside the shadow is eclipsed beyond doubt, and the starts of the
                    // grazing-LOS iteration (osculating parabola at the deepest LOS) need no more than FP32.
                    const float isll = rcp_f(Df.Sll), kap = Df.Stt - Df.Stl * Df.Stl * isll;
                    const float del = sqrtf(-2.0f * g0f * rcp_f(kap)), slope = -Df.Stl * isll;
                    if (g0f < -kWarmRejectMargin && Df.Sll > 0.0f && kap > 0.0f && del < 0.25f) {
                        LFB_CNT(12);
                        float sdl, cdl;
#ifdef __CUDA_ARCH__
                        __sincosf(del, &sdl, &cdl);
#else
                        sdl = sinf(del);
                        cdl = cosf(del);
#endif
                        cm = c;
                        sm = s;
                        c0 = c * (double)cdl + s * (double)sdl;  // thm - del
                        s0 = s * (double)cdl - c * (double)sdl;
                        c1 = c * (double)cdl - s * (double)sdl;  // thm + del
                        s1 = s * (double)cdl + c * (double)sdl;
                        lam0 = lam - (double)(slope * del);
                        lam1 = lam + (double)(slope * del);
                        starts = true;
                    }
#endif
                }
            }
        }
#endif
        for (int it = 0; it < 5 && !warmed && verdict < 0; ++it) {
            ray_eval<PLANAR>(R, si, ci, T, c, s, lam, D);
            if (!(D.Sll > 0.0)) {  // no potential minimum along the closest LOS: out of reach
                LFB_CNT(5);
                verdict = 0;
                break;
            }
            double dl = clampd(-D.Sl * fast_rcp(D.Sll), 0.1);
            lam += dl;
            if (fabs(dl) < 1e-6) break;  // the 2-D Newton below finishes the job
        }
        bool conv = starts;
        for (int it = 0; it < kMinIters && verdict < 0 && !starts; ++it) {
            ray_eval<PLANAR>(R, si, ci, T, c, s, lam, D);
            double det = D.Stt * D.Sll - D.Stl * D.Stl;
            if (!(D.Sll > 0.0) || !(det > 0.0)) {
                verdict = D.S >= R.phic ? 0 : 2;
                break;
            }
            double dth = clampd(-(D.St * D.Sll - D.Sl * D.Stl) / det, 0.1);
            double dl = clampd(-(D.Sl * D.Stt - D.St * D.Stl) / det, 0.1);
            rotate_cs(c, s, dth);
            lam += dl;
            if (fabs(dth) < 1e-7 && fabs(dl) < 1e-7) { conv = true; break; }
        }
        if (verdict < 0 && !conv) verdict = 2;
        if (verdict < 0 && !starts) {
            ray_eval<PLANAR>(R, si, ci, T, c, s, lam, D);
            double g0 = D.S - R.phic;
            double kappa = D.Stt - D.Stl * D.Stl / D.Sll;
            if (!(g0 < 0.0)) {  // the deepest LOS clears the lobe
                LFB_CNT(6);
                verdict = 0;
            } else if (!(D.Sll > 0.0) || !(kappa > 0.0)) {
                verdict = 2;
            } else {
                LFB_CNT(7);
                double del = sqrt(-2.0 * g0 / kappa), slope = -D.Stl / D.Sll;
                if (!(del < 0.25)) {
                    verdict = 2;
                } else {
                    cm = c;
                    sm = s;
                    c0 = c1 = c;
                    s0 = s1 = s;
                    rotate_cs(c0, s0, -del);
                    rotate_cs(c1, s1, del);
                    lam0 = lam - slope * del;
                    lam1 = lam + slope * del;
                }
            }
        }
    }
    if (verdict < 0) {
        const Roots start = {{c0, c1}, {s0, s1}, {lam0, lam1}};
        if (graze_roots<PLANAR>(R, si, ci, T, cpsi, spsi, cm, sm, start, res, roots)) verdict = 1;
        else verdict = 2;
    }
    if (verdict == 1) {
        *ph_in = res[0] * (1.0 / kTwoPi);
        *ph_out = res[1] * (1.0 / kTwoPi);
        return 1;
    }
    if (verdict == 0) return 0;
    return ingress_egress_robust(R, si, ci, T, ph_in, ph_out);
}

// ---- ballistic stream from L1 (roche.bspot): fixed-sequence Gragg-Bulirsch-Stoer ----
constexpr double kStreamEps = 1e-5;
constexpr double kStreamH0 = 0.6;
constexpr int kStreamMaxSteps = 400;

LFB_HD void stream_rhs(const Roche& R, const double y[4], double f[4])
{
    double x = y[0], yy = y[1], x2 = x - 1.0;
    double ysq = yy * yy;
    double ir1 = rsqrt(x * x + ysq), ir2 = rsqrt(x2 * x2 + ysq);
    double a1 = R.omu * ir1 * ir1 * ir1, a2 = R.mu * ir2 * ir2 * ir2;
    f[0] = y[2];
    f[1] = y[3];
    f[2] = -(a1 * x + a2 * x2 - (x - R.mu)) + 2.0 * y[3];
    f[3] = -((a1 + a2) * yy - yy) - 2.0 * y[2];
}

LFB_HD_NOINLINE void gbs_step(const Roche& R, const double y0[4], double H, double yout[4])
{
    double T[6][4];
    double f0[4];
    stream_rhs(R, y0, f0);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int n = 2 * (k + 1);
        double h = H / n;
        double z0[4], z1[4], f[4];
        for (int j = 0; j < 4; ++j) { z0[j] = y0[j]; z1[j] = y0[j] + h * f0[j]; }
        for (int m = 1; m < n; ++m) {
            stream_rhs(R, z1, f);
            for (int j = 0; j < 4; ++j) {
                double t = z0[j] + 2.0 * h * f[j];
                z0[j] = z1[j];
                z1[j] = t;
            }
        }
        stream_rhs(R, z1, f);
        for (int j = 0; j < 4; ++j) T[k][j] = 0.5 * (z0[j] + z1[j] + h * f[j]);
#pragma unroll
        for (int m = k - 1; m >= 0; --m) {
            double ratio = (double)(k + 1) / (double)(m + 1);
            double fac = 1.0 / (ratio * ratio - 1.0);
            for (int j = 0; j < 4; ++j) T[m][j] = T[m + 1][j] + (T[m + 1][j] - T[m][j]) * fac;
        }
    }
    for (int j = 0; j < 4; ++j) yout[j] = T[0][j];
}

// (x, y, vx, vy) where the stream from L1 first reaches radius rad from the white dwarf;
// false if it passes closest approach without getting there (roche.bspot raises).
LFB_HD bool bspot(const Roche& R, double rad, double out[4])
{
    if (!(rad > 0.0) || !(rad < R.xl1 - 2.0 * kStreamEps)) return false;
    double A = R.omu / (R.xl1 * R.xl1 * R.xl1) + R.mu / (R.rs * R.rs * R.rs);
    double l2 = 0.5 * ((A - 2.0) + sqrt(A * (9.0 * A - 8.0)));
    double l1 = sqrt(l2);
    double m1 = (l2 - 2.0 * A - 1.0) / (2.0 * l1);
    double y[4] = {R.xl1 - kStreamEps, -m1 * kStreamEps, -l1 * kStreamEps, -l1 * m1 * kStreamEps};
    for (int step = 0; step < kStreamMaxSteps; ++step) {
        double x2 = y[0] - 1.0, ysq = y[1] * y[1];
        double r1sq = y[0] * y[0] + ysq, r2sq = x2 * x2 + ysq;
        double w2 = R.omu / (r1sq * sqrt(r1sq)) + R.mu / (r2sq * sqrt(r2sq)) + 1.0;
        double H = kStreamH0 / sqrt(w2);
        double yn[4];
        gbs_step(R, y, H, yn);
        double r0 = sqrt(r1sq), rn = sqrt(yn[0] * yn[0] + yn[1] * yn[1]);
        if (rn <= rad) {
            double h = H * (r0 - rad) / (r0 - rn);
            for (int it = 0; it < 8; ++it) {
                gbs_step(R, y, h, yn);
                rn = sqrt(yn[0] * yn[0] + yn[1] * yn[1]);
                const double dh = (rn - rad) * rn / (yn[0] * yn[2] + yn[1] * yn[3]);
                h -= dh;
                if (fabs(dh) <= 1e-8 * fabs(h)) break;  // quadratic convergence: the step after one this small is below rounding
            }
            gbs_step(R, y, h, yn);
            for (int j = 0; j < 4; ++j) out[j] = yn[j];
            return true;
        }
        if (yn[0] * yn[2] + yn[1] * yn[3] >= 0.0) return false;
        for (int j = 0; j < 4; ++j) y[j] = yn[j];
    }
    return false;
}

#ifdef __CUDACC__
// ---- the same integrator with one job spread over eight neighbouring lanes ----
// The six modified-midpoint sequences of a GBS step are independent: lane k of the group runs sequence k
// (n = 2 k + 2 substeps; lanes 6 and 7 repeat sequence 5), the six results are exchanged by shuffles and
// every lane finishes the extrapolation tableau in the order gbs_step uses -- same numbers, bit for bit,
// a third of the latency.  All lanes of the group must call it together (their inputs are identical).
__device__ __forceinline__ void gbs_step_lanes(const Roche& R, const double y0[4], double H, double yout[4])
{
    const int sub = threadIdx.x & 7, k_mine = sub < 6 ? sub : 5;
    const unsigned mask = 0xffu << (threadIdx.x & 24);
    double f0[4];
    stream_rhs(R, y0, f0);
    const int n = 2 * (k_mine + 1);
    const double h = H / n;
    double z0[4], z1[4], f[4], Tm[4];
    for (int j = 0; j < 4; ++j) { z0[j] = y0[j]; z1[j] = y0[j] + h * f0[j]; }
    for (int m = 1; m < n; ++m) {
        stream_rhs(R, z1, f);
        for (int j = 0; j < 4; ++j) {
            double t = z0[j] + 2.0 * h * f[j];
            z0[j] = z1[j];
            z1[j] = t;
        }
    }
    stream_rhs(R, z1, f);
    for (int j = 0; j < 4; ++j) Tm[j] = 0.5 * (z0[j] + z1[j] + h * f[j]);
    double T[6][4];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) T[k][j] = __shfl_sync(mask, Tm[j], k, 8);
    // the tableau, in gbs_step's order: after sequence k, columns k - 1 .. 0
#pragma unroll
    for (int k = 1; k < 6; ++k) {
#pragma unroll
        for (int m = k - 1; m >= 0; --m) {
            double ratio = (double)(k + 1) / (double)(m + 1);
            double fac = 1.0 / (ratio * ratio - 1.0);
            for (int j = 0; j < 4; ++j) T[m][j] = T[m + 1][j] + (T[m + 1][j] - T[m][j]) * fac;
        }
    }
    for (int j = 0; j < 4; ++j) yout[j] = T[0][j];
}

// bspot with gbs_step_lanes: every lane of the group of eight returns the same answer
__device__ __forceinline__ bool bspot_lanes(const Roche& R, double rad, double out[4])
{
    if (!(rad > 0.0) || !(rad < R.xl1 - 2.0 * kStreamEps)) return false;
    double A = R.omu / (R.xl1 * R.xl1 * R.xl1) + R.mu / (R.rs * R.rs * R.rs);
    double l2 = 0.5 * ((A - 2.0) + sqrt(A * (9.0 * A - 8.0)));
    double l1 = sqrt(l2);
    double m1 = (l2 - 2.0 * A - 1.0) / (2.0 * l1);
    double y[4] = {R.xl1 - kStreamEps, -m1 * kStreamEps, -l1 * kStreamEps, -l1 * m1 * kStreamEps};
    for (int step = 0; step < kStreamMaxSteps; ++step) {
        double x2 = y[0] - 1.0, ysq = y[1] * y[1];
        double r1sq = y[0] * y[0] + ysq, r2sq = x2 * x2 + ysq;
        double w2 = R.omu / (r1sq * sqrt(r1sq)) + R.mu / (r2sq * sqrt(r2sq)) + 1.0;
        double H = kStreamH0 / sqrt(w2);
        double yn[4];
        gbs_step_lanes(R, y, H, yn);
        double r0 = sqrt(r1sq), rn = sqrt(yn[0] * yn[0] + yn[1] * yn[1]);
        if (rn <= rad) {
            double h = H * (r0 - rad) / (r0 - rn);
            for (int it = 0; it < 8; ++it) {
                gbs_step_lanes(R, y, h, yn);
                rn = sqrt(yn[0] * yn[0] + yn[1] * yn[1]);
                const double dh = (rn - rad) * rn / (yn[0] * yn[2] + yn[1] * yn[3]);
                h -= dh;
                if (fabs(dh) <= 1e-8 * fabs(h)) break;  // (the same test on the same numbers in every lane of the group)
            }
            gbs_step_lanes(R, y, h, yn);
            for (int j = 0; j < 4; ++j) out[j] = yn[j];
            return true;
        }
        if (yn[0] * yn[2] + yn[1] * yn[3] >= 0.0) return false;
        for (int j = 0; j < 4; ++j) y[j] = yn[j];
    }
    return false;
}
#endif

// Radius of the critical surface from the donor's centre along unit vector d, plus the
// potential gradient there.  Newton from inside the lobe (monotone), then one safeguarded polish.
LFB_HD double donor_radius(const Roche& R, double dx, double dy, double dz, double g[3])
{
    double r = R.rin;
    double gx = 0, gy = 0, gz = 0;
    for (int it = 0; it < 40; ++it) {
        double x = 1.0 + r * dx, y = r * dy, z = r * dz;
        double yz = y * y + z * z;
        double ir1 = rsqrt(x * x + yz), ir2 = 1.0 / r;
        double a1 = R.omu * ir1 * ir1 * ir1, a2 = R.mu * ir2 * ir2 * ir2;
        double xc = x - R.mu;
        double f = -R.omu * ir1 - R.mu * ir2 - 0.5 * (xc * xc + y * y) - R.phic;
        gx = a1 * x + a2 * (x - 1.0) - xc;
        gy = (a1 + a2) * y - y;
        gz = (a1 + a2) * z;
        double fp = gx * dx + gy * dy + gz * dz;
        double dr = -f / fp;
        if (r + dr > R.rs) dr = 0.5 * (R.rs - r);
        r += dr;
        if (fabs(dr) < 1e-15) break;
    }
    double x = 1.0 + r * dx, y = r * dy, z = r * dz;
    double yz = y * y + z * z;
    double ir1 = rsqrt(x * x + yz), ir2 = 1.0 / r;
    double a1 = R.omu * ir1 * ir1 * ir1, a2 = R.mu * ir2 * ir2 * ir2;
    g[0] = a1 * x + a2 * (x - 1.0) - (x - R.mu);
    g[1] = (a1 + a2) * y - y;
    g[2] = (a1 + a2) * z;
    (void)gx; (void)gy; (void)gz;
    return r;
}

// Prior.ln_prob (model.py:83-113)
LFB_HD double prior_ln_prob(int type, double p1, double p2, double norm, double val)
{
    const double kLnSqrt2Pi = 0.91893853320467274178;
    const double kLnMinDenormal = -744.44007192138126;
    const double ninf = -INFINITY;
    switch (type) {
    case 1:
        if (val <= 0.0) return ninf;
        // fall through
    case 0: {
        double z = (val - p1) / p2;
        double t = -0.5 * z * z - kLnSqrt2Pi;
        if (!(t >= kLnMinDenormal)) return ninf;  // scipy's pdf underflows to 0 (model.py:85-89)
        return t - log(p2);
    }
    case 2:
        return (val > p1 && val < p2) ? log(1.0 / fabs(p1 - p2)) : ninf;
    case 3:
        return (val > p1 && val < p2) ? log(1.0 / norm / val) : ninf;
    case 4:
        return (val > 0.0 && val < p2) ? log(1.0 / norm / (val + p1)) : ninf;
    }
    return ninf;
}

}  // namespace lfb
