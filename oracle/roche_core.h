/*
 * oracle/roche_core.h -- Roche geometry for the CPU oracle (FP64, plain C).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lfit_python_b200/ may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in two
 * un-vendored, un-pinned third-party packages (`lfit`, `trm.roche`; imported at
 * /root/reference/CVModel.py:13,15) that are absent from /root/reference, so
 * there is no reference output to pin against.  This file restates the public
 * Roche-geometry definitions those call sites rely on:
 *   roche.xl1(q)            CVModel.py:222
 *   roche.findphi(q, 90)    CVModel.py:460
 *   roche.findi(q, dphi)    CVModel.py:561 (and inside lfit.CV.calcFlux)
 *   roche.bspot(q, rad)     CVModel.py:288
 *   ingress/egress of a point behind the donor's Roche lobe (inside lfit)
 * and is pinned only to first-principles known answers.
 *
 * Frame: separation a = 1, white dwarf at the origin, donor at (1,0,0), orbit in
 * the xy plane, angular velocity 1 about +z.  mu = q/(1+q).
 *   Phi(x,y,z) = -(1-mu)/r1 - mu/r2 - ((x-mu)^2 + y^2)/2
 *   earth(i,phase) = (sin i cos th, -sin i sin th, cos i),  th = 2 pi phase.
 * A point is eclipsed when the line of sight towards earth passes through the
 * donor's critical lobe, i.e. min over the chord of the LOS inside the sphere
 * |x-(1,0,0)| < 1-xl1 of Phi is below Phi(L1).
 */
#ifndef LFO_ROCHE_CORE_H
#define LFO_ROCHE_CORE_H

#include <math.h>

#define LFO_PI 3.14159265358979323846264338327950288
#define LFO_TWOPI 6.28318530717958647692528676655900577

typedef struct {
    double q, mu, omu; /* omu = 1 - mu */
    double xl1;        /* L1 distance from the white dwarf */
    double rs;         /* 1 - xl1: radius of the sphere about the donor bounding its lobe */
    double phic;       /* Phi(L1) */
} lfo_roche;

/* A surface element as the eclipse solver sees it: a part p0 fixed in the
 * rotating frame plus a part (xi, eta) fixed on the sky (xi along
 * uhat = (-sin th, -cos th, 0), eta along vhat = (-ci cos th, ci sin th, si)). */
typedef struct {
    double p0[3];
    double xi, eta;
} lfo_point;

typedef struct {
    double S, St, Sl, Stt, Stl, Sll;
} lfo_derivs;

static inline double lfo_pot(const lfo_roche *R, double x, double y, double z)
{
    double r1 = sqrt(x * x + y * y + z * z);
    double dx = x - 1.0;
    double r2 = sqrt(dx * dx + y * y + z * z);
    double xc = x - R->mu;
    return -R->omu / r1 - R->mu / r2 - 0.5 * (xc * xc + y * y);
}

static inline void lfo_grad(const lfo_roche *R, double x, double y, double z, double g[3])
{
    double r1sq = x * x + y * y + z * z;
    double dx = x - 1.0;
    double r2sq = dx * dx + y * y + z * z;
    double a1 = R->omu / (r1sq * sqrt(r1sq));
    double a2 = R->mu / (r2sq * sqrt(r2sq));
    g[0] = a1 * x + a2 * dx - (x - R->mu);
    g[1] = a1 * y + a2 * y - y;
    g[2] = a1 * z + a2 * z;
}

/* L1 point: root on (0,1) of dPhi/dx on the x axis (monotonic): bisection + polish */
static inline int lfo_xl1(double q, double *out)
{
    if (!(q > 0.0) || !isfinite(q)) return 1;
    double mu = q / (1.0 + q), omu = 1.0 - mu;
    double lo = 1e-9, hi = 1.0 - 1e-9;
    for (int it = 0; it < 200; ++it) {
        double x = 0.5 * (lo + hi);
        double f = omu / (x * x) - mu / ((1.0 - x) * (1.0 - x)) - (x - mu);
        if (f > 0.0) lo = x; else hi = x;
    }
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 3; ++it) {
        double omx = 1.0 - x;
        double f = omu / (x * x) - mu / (omx * omx) - (x - mu);
        double fp = -2.0 * omu / (x * x * x) - 2.0 * mu / (omx * omx * omx) - 1.0;
        x -= f / fp;
    }
    *out = x;
    return 0;
}

static inline int lfo_roche_init(lfo_roche *R, double q)
{
    double x;
    if (lfo_xl1(q, &x)) return 1;
    R->q = q;
    R->mu = q / (1.0 + q);
    R->omu = 1.0 - R->mu;
    R->xl1 = x;
    R->rs = 1.0 - x;
    R->phic = lfo_pot(R, x, 0.0, 0.0);
    return 0;
}

/* Potential and its first/second derivatives along the family of lines of
 * sight x(th, lam) = p0 + xi*uhat(th) + eta*vhat(th) + lam*earth(th). */
static inline void lfo_ray_eval(const lfo_roche *R, double si, double ci, const lfo_point *T,
                                double th, double lam, lfo_derivs *D)
{
    double c = cos(th), s = sin(th);
    double ex = si * c, ey = -si * s, ez = ci;
    /* rotating part d = xi*uhat + eta*vhat + lam*e */
    double dx = -T->xi * s - T->eta * ci * c + lam * ex;
    double dy = -T->xi * c + T->eta * ci * s + lam * ey;
    double dz = T->eta * si + lam * ez;
    double x = T->p0[0] + dx, y = T->p0[1] + dy, z = T->p0[2] + dz;
    /* d/dth of the rotating part is J d = (dy, -dx, 0); second is (-dx, -dy, 0) */
    double tx = dy, ty = -dx;
    double epx = ey, epy = -ex; /* d earth / d th */

    double x2 = x - 1.0;
    double r1sq = x * x + y * y + z * z;
    double r2sq = x2 * x2 + y * y + z * z;
    double ir1 = 1.0 / sqrt(r1sq), ir2 = 1.0 / sqrt(r2sq);
    double a1 = R->omu * ir1 * ir1 * ir1, a2 = R->mu * ir2 * ir2 * ir2;
    double b1 = 3.0 * a1 * ir1 * ir1, b2 = 3.0 * a2 * ir2 * ir2;
    double xc = x - R->mu;

    double gx = a1 * x + a2 * x2 - xc;
    double gy = (a1 + a2) * y - y;
    double gz = (a1 + a2) * z;

    double x_e = x * ex + y * ey + z * ez;       /* (pos rel WD) . e */
    double x_t = x * tx + y * ty;                /* . x_theta */
    double d_e = x2 * ex + y * ey + z * ez;      /* (pos rel donor) . e */
    double d_t = x2 * tx + y * ty;
    double t_t = tx * tx + ty * ty;
    double e_t = ex * tx + ey * ty;
    double e_exy = ex * ex + ey * ey;

    D->S = -R->omu * ir1 - R->mu * ir2 - 0.5 * (xc * xc + y * y);
    D->St = gx * tx + gy * ty;
    D->Sl = gx * ex + gy * ey + gz * ez;
    D->Sll = (a1 + a2) - b1 * x_e * x_e - b2 * d_e * d_e - e_exy;
    D->Stt = (a1 + a2) * t_t - b1 * x_t * x_t - b2 * d_t * d_t - t_t + (gx * (-dx) + gy * (-dy));
    D->Stl = (a1 + a2) * e_t - b1 * x_e * x_t - b2 * d_e * d_t - e_t + (gx * epx + gy * epy);
}

/* ---- robust (slow) ingress/egress: scan + golden section + bisection ---- */

static inline double lfo_minpot_along_los(const lfo_roche *R, double si, double ci, const lfo_point *T,
                                          double th, int *has_chord)
{
    double c = cos(th), s = sin(th);
    double ex = si * c, ey = -si * s, ez = ci;
    double ox = T->p0[0] - T->xi * s - T->eta * ci * c;
    double oy = T->p0[1] - T->xi * c + T->eta * ci * s;
    double oz = T->p0[2] + T->eta * si;
    double wx = 1.0 - ox, wy = -oy, wz = -oz;
    double b = wx * ex + wy * ey + wz * ez;
    double d2 = wx * wx + wy * wy + wz * wz - b * b;
    *has_chord = 0;
    if (d2 >= R->rs * R->rs) return 1.0;
    double half = sqrt(R->rs * R->rs - d2);
    double l1 = b - half, l2 = b + half;
    if (l1 < 0.0) l1 = 0.0;
    if (l2 <= l1) return 1.0;
    *has_chord = 1;
    enum { NS = 48 };
    double best = 1e300;
    int kb = 0;
    for (int k = 0; k < NS; ++k) {
        double lam = l1 + (l2 - l1) * k / (NS - 1);
        double v = lfo_pot(R, ox + lam * ex, oy + lam * ey, oz + lam * ez);
        if (v < best) { best = v; kb = k; }
    }
    int ka = kb > 0 ? kb - 1 : 0, kc = kb < NS - 1 ? kb + 1 : NS - 1;
    double a = l1 + (l2 - l1) * ka / (NS - 1), cc = l1 + (l2 - l1) * kc / (NS - 1);
    const double gr = 0.6180339887498949;
    double x1 = cc - gr * (cc - a), x2 = a + gr * (cc - a);
    double f1 = lfo_pot(R, ox + x1 * ex, oy + x1 * ey, oz + x1 * ez);
    double f2 = lfo_pot(R, ox + x2 * ex, oy + x2 * ey, oz + x2 * ez);
    for (int it = 0; it < 70; ++it) {
        if (f1 < f2) {
            cc = x2; x2 = x1; f2 = f1;
            x1 = cc - gr * (cc - a);
            f1 = lfo_pot(R, ox + x1 * ex, oy + x1 * ey, oz + x1 * ez);
        } else {
            a = x1; x1 = x2; f1 = f2;
            x2 = a + gr * (cc - a);
            f2 = lfo_pot(R, ox + x2 * ex, oy + x2 * ey, oz + x2 * ez);
        }
    }
    double v = f1 < f2 ? f1 : f2;
    if (best < v) v = best;
    return v - R->phic;
}

/* returns 1 if the point is ever eclipsed; phases (cycles) of ingress < egress */
static inline int lfo_ingress_egress_robust(const lfo_roche *R, double si, double ci, const lfo_point *T,
                                            double *ph_in, double *ph_out)
{
    enum { NSCAN = 384 };
    double psi = atan2(T->p0[1], 1.0 - T->p0[0]);
    double half = 0.5 * LFO_PI;
    double g[NSCAN];
    int hc;
    int kb = 0;
    double best = 1e300;
    for (int k = 0; k < NSCAN; ++k) {
        double th = psi - half + 2.0 * half * k / (NSCAN - 1);
        g[k] = lfo_minpot_along_los(R, si, ci, T, th, &hc);
        if (g[k] < best) { best = g[k]; kb = k; }
    }
    if (kb == 0 || kb == NSCAN - 1) return 0;
    double step = 2.0 * half / (NSCAN - 1);
    double a = psi - half + step * (kb - 1), cc = psi - half + step * (kb + 1);
    const double gr = 0.6180339887498949;
    double x1 = cc - gr * (cc - a), x2 = a + gr * (cc - a);
    double f1 = lfo_minpot_along_los(R, si, ci, T, x1, &hc);
    double f2 = lfo_minpot_along_los(R, si, ci, T, x2, &hc);
    for (int it = 0; it < 60; ++it) {
        if (f1 < f2) {
            cc = x2; x2 = x1; f2 = f1;
            x1 = cc - gr * (cc - a);
            f1 = lfo_minpot_along_los(R, si, ci, T, x1, &hc);
        } else {
            a = x1; x1 = x2; f1 = f2;
            x2 = a + gr * (cc - a);
            f2 = lfo_minpot_along_los(R, si, ci, T, x2, &hc);
        }
    }
    double thm = f1 < f2 ? x1 : x2, gm = f1 < f2 ? f1 : f2;
    if (best < gm) { gm = best; thm = psi - half + step * kb; }
    if (!(gm < 0.0)) return 0;
    /* brackets: walk outwards on the scan grid until the LOS clears the lobe */
    int kl = (int)floor((thm - (psi - half)) / step);
    while (kl > 0 && g[kl] < 0.0) --kl;
    int kr = (int)ceil((thm - (psi - half)) / step);
    while (kr < NSCAN - 1 && g[kr] < 0.0) ++kr;
    if (g[kl] < 0.0 || g[kr] < 0.0) return 0; /* eclipse wider than half an orbit: not physical here */
    double lo = psi - half + step * kl, hi = thm;
    for (int it = 0; it < 90; ++it) {
        double m = 0.5 * (lo + hi);
        if (lfo_minpot_along_los(R, si, ci, T, m, &hc) < 0.0) hi = m; else lo = m;
    }
    *ph_in = 0.5 * (lo + hi) / LFO_TWOPI;
    lo = thm; hi = psi - half + step * kr;
    for (int it = 0; it < 90; ++it) {
        double m = 0.5 * (lo + hi);
        if (lfo_minpot_along_los(R, si, ci, T, m, &hc) < 0.0) lo = m; else hi = m;
    }
    *ph_out = 0.5 * (lo + hi) / LFO_TWOPI;
    return 1;
}

/* ---- fast ingress/egress: 2-D Newton for the deepest LOS, then for the two
 * grazing lines of sight (Phi = Phi_c and dPhi/dlam = 0).  Same definition as
 * the robust solver; converges to the same roots. ---- */
#ifndef LFO_NEWTON_MIN_ITERS
#define LFO_NEWTON_MIN_ITERS 12
#endif
#ifndef LFO_NEWTON_ROOT_ITERS
#define LFO_NEWTON_ROOT_ITERS 16
#endif

static inline double lfo_eggleton(double q)
{
    /* volume-equivalent lobe radius of the donor (mass ratio q = M2/M1), used as a start value only */
    double q13 = cbrt(q), q23 = q13 * q13;
    return 0.49 * q23 / (0.6 * q23 + log(1.0 + q13));
}

static inline void lfo_newton_clamp(double *d, double lim)
{
    if (*d > lim) *d = lim;
    if (*d < -lim) *d = -lim;
}

/* polar radius of the donor's critical lobe: Phi(1,0,z) = Phi_c.  A sphere of
 * this radius about the donor lies inside the lobe (start values only). */
static inline double lfo_lobe_polar_radius(const lfo_roche *R)
{
    double z = 0.8 * lfo_eggleton(R->q);
    for (int it = 0; it < 20; ++it) {
        double r1sq = 1.0 + z * z;
        double f = -R->omu / sqrt(r1sq) - R->mu / z - 0.5 * R->omu * R->omu - R->phic;
        double fp = R->omu * z / (r1sq * sqrt(r1sq)) + R->mu / (z * z);
        z -= f / fp;
    }
    return z;
}

/* counters for tests: [0] calls, [1] deep starts, [2] shallow starts, [3] fell back to the robust solver */
#ifdef LFO_STATS
static long long lfo_ie_stats[8]; /* [4] ray evaluations in the Newton solver */
static long long lfo_ie_reason[16];
#define LFO_STAT(i) (lfo_ie_stats[i]++)
#define LFO_FB(k) do { lfo_ie_reason[k]++; goto fallback; } while (0)
#else
#define LFO_STAT(i) ((void)0)
#define LFO_FB(k) goto fallback
#endif

static inline int lfo_ingress_egress_newton(const lfo_roche *R, double si, double ci, const lfo_point *T,
                                            double *ph_in, double *ph_out)
{
    double psi = atan2(T->p0[1] - T->xi, 1.0 - T->p0[0] + T->eta * ci);
    double th = psi, c = cos(th), s = sin(th);
    double ex = si * c, ey = -si * s, ez = ci;
    double ox = T->p0[0] - T->xi * s - T->eta * ci * c;
    double oy = T->p0[1] - T->xi * c + T->eta * ci * s;
    double oz = T->p0[2] + T->eta * si;
    double wx = 1.0 - ox, wy = -oy, wz = -oz;
    double lam = wx * ex + wy * ey + wz * ez;
    double w2 = wx * wx + wy * wy + wz * wz;
    double d2 = w2 - lam * lam;
    LFO_STAT(0);
    if (d2 >= R->rs * R->rs || lam <= 0.0) return 0;
    double rin = 0.9 * lfo_lobe_polar_radius(R);
    double rxy = si * sqrt(wx * wx + wy * wy), c0 = wz * ci;
    double cosd = (sqrt(w2 - rin * rin) - c0) / rxy;
    lfo_derivs D;
    double th0[2], lam0[2], thm = psi;
    if (cosd < 0.995) { /* tangent phases at least 0.1 rad either side of conjunction */
        /* the LOS can pass inside the inscribed sphere: certainly eclipsed; start the two
         * grazing solves where the LOS touches that sphere */
        double del = acos(cosd > -1.0 ? cosd : -1.0);
        th0[0] = psi - del; th0[1] = psi + del;
        lam0[0] = lam0[1] = sqrt(w2 - rin * rin);
        LFO_STAT(1);
    } else {
        /* shallow: minimum of the potential along the conjunction LOS ... */
        LFO_STAT(2);
        for (int it = 0; it < 5; ++it) {
            { LFO_STAT(4); lfo_ray_eval(R, si, ci, T, th, lam, &D); }
            if (!(D.Sll > 0.0)) return 0; /* no potential minimum along the closest LOS: the lobe is out of reach */
            double dl = -D.Sl / D.Sll;
            lfo_newton_clamp(&dl, 0.1);
            lam += dl;
        }
        /* ... then the deepest LOS nearby (2-D Newton on the gradient of Phi(th, lam)) ... */
        int conv = 0;
        for (int it = 0; it < LFO_NEWTON_MIN_ITERS; ++it) {
            { LFO_STAT(4); lfo_ray_eval(R, si, ci, T, th, lam, &D); }
            double det = D.Stt * D.Sll - D.Stl * D.Stl;
            if (!(D.Sll > 0.0) || !(det > 0.0)) {
                if (D.S >= R->phic) return 0; /* no minimum nearby and not eclipsed here: out of reach */
                LFO_FB(1);
            }
            double dth = -(D.St * D.Sll - D.Sl * D.Stl) / det;
            double dl = -(D.Sl * D.Stt - D.St * D.Stl) / det;
            lfo_newton_clamp(&dth, 0.1);
            lfo_newton_clamp(&dl, 0.1);
            th += dth;
            lam += dl;
            if (fabs(dth) < 1e-7 && fabs(dl) < 1e-7) { conv = 1; break; }
        }
        if (!conv) LFO_FB(2);
        { LFO_STAT(4); lfo_ray_eval(R, si, ci, T, th, lam, &D); }
        double g0 = D.S - R->phic;
        if (!(g0 < 0.0)) return 0; /* the deepest LOS clears the lobe: never eclipsed */
        /* ... and the osculating parabola of g(th) = min_lam Phi - Phi_c there */
        double kappa = D.Stt - D.Stl * D.Stl / D.Sll;
        if (!(D.Sll > 0.0) || !(kappa > 0.0)) LFO_FB(3);
        double del = sqrt(-2.0 * g0 / kappa), slope = -D.Stl / D.Sll;
        thm = th;
        th0[0] = th - del; th0[1] = th + del;
        lam0[0] = lam - slope * del; lam0[1] = lam + slope * del;
    }
    double res[2];
    for (int side = 0; side < 2; ++side) {
        double sg = side ? 1.0 : -1.0;
        th = th0[side];
        lam = lam0[side];
        int conv = 0;
        for (int it = 0; it < LFO_NEWTON_ROOT_ITERS; ++it) {
            { LFO_STAT(4); lfo_ray_eval(R, si, ci, T, th, lam, &D); }
            double F1 = D.S - R->phic, F2 = D.Sl;
            double det = D.St * D.Sll - D.Sl * D.Stl;
            double dth = (-F1 * D.Sll + F2 * D.Sl) / det;
            double dl = (-D.St * F2 + D.Stl * F1) / det;
            lfo_newton_clamp(&dth, 0.2);
            lfo_newton_clamp(&dl, 0.2);
            /* stay on this side of the deepest LOS (a step pushed back from it is not a Newton step:
             * however small, it does not signal convergence) */
            int newton_step = 1;
            if (sg * (th + dth - thm) <= 0.0) { dth = 0.5 * (thm - th); dl *= 0.5; newton_step = 0; }
            th += dth;
            lam += dl;
            if (newton_step && fabs(dth) < 1e-13 && fabs(dl) < 1e-10) { conv = 1; break; }
        }
        /* accept only a converged grazing LOS of the right kind: a minimum along the LOS,
         * inside the bounding sphere, entering (side 0) or leaving (side 1) the lobe */
        { LFO_STAT(4); lfo_ray_eval(R, si, ci, T, th, lam, &D); }
        double xx = T->p0[0] - T->xi * sin(th) - T->eta * ci * cos(th) + lam * si * cos(th) - 1.0;
        double yy = T->p0[1] - T->xi * cos(th) + T->eta * ci * sin(th) - lam * si * sin(th);
        double zz = T->p0[2] + T->eta * si + lam * ci;
        int ok = conv && D.Sll > 0.0 && lam > 0.0 && xx * xx + yy * yy + zz * zz <= R->rs * R->rs &&
                 (side ? D.St > 0.0 : D.St < 0.0) && fabs(th - psi) < 0.5 * LFO_PI;
        if (!ok) LFO_FB(7 + side);
        res[side] = th;
    }
    if (!(res[0] < res[1])) LFO_FB(9);
    *ph_in = res[0] / LFO_TWOPI;
    *ph_out = res[1] / LFO_TWOPI;
    return 1;
fallback:
    LFO_STAT(3);
    return lfo_ingress_egress_robust(R, si, ci, T, ph_in, ph_out);
}

/* full phase width of the eclipse of the white-dwarf centre at i = 90 deg
 * (= roche.findphi(q, 90), CVModel.py:460).  Unknowns (c = cos th, lam) on the
 * LOS lam*(c, -s, 0):  Phi = -(1-mu)/lam - mu/sqrt(D) - lam^2/2 + mu lam c - mu^2/2,
 * D = 1 + lam^2 - 2 lam c. */
static inline void lfo_origin_pot(const lfo_roche *R, double u, double c, double lam, double *P, double *Pu,
                                  double *Pl, double *Pul, double *Pll, double *Pc, double *Pcl)
{
    /* LOS from the origin with sin(i) = u at orbital angle cos(th) = c */
    double mu = R->mu;
    double Dd = 1.0 + lam * lam - 2.0 * lam * u * c;
    double isq = 1.0 / sqrt(Dd);
    double i3 = isq * isq * isq, i5 = i3 * isq * isq;
    double Dl = 2.0 * lam - 2.0 * u * c, Du = -2.0 * lam * c, Dc = -2.0 * lam * u;
    *P = -R->omu / lam - mu * isq - 0.5 * lam * lam * u * u + mu * lam * u * c - 0.5 * mu * mu;
    *Pl = R->omu / (lam * lam) + 0.5 * mu * i3 * Dl - lam * u * u + mu * u * c;
    *Pu = 0.5 * mu * i3 * Du - lam * lam * u + mu * lam * c;
    *Pc = 0.5 * mu * i3 * Dc + mu * lam * u;
    *Pll = -2.0 * R->omu / (lam * lam * lam) + 0.5 * mu * (-1.5 * i5 * Dl * Dl + i3 * 2.0) - u * u;
    *Pul = 0.5 * mu * (-1.5 * i5 * Dl * Du + i3 * (-2.0 * c)) - 2.0 * lam * u + mu * c;
    *Pcl = 0.5 * mu * (-1.5 * i5 * Dl * Dc + i3 * (-2.0 * u)) + mu * u;
}

static inline double lfo_findphi90(const lfo_roche *R)
{
    double rl = lfo_eggleton(R->q);
    double c = sqrt(1.0 - rl * rl), lam = c;
    for (int it = 0; it < 12; ++it) {
        double P, Pu, Pl, Pul, Pll, Pc, Pcl;
        lfo_origin_pot(R, 1.0, c, lam, &P, &Pu, &Pl, &Pul, &Pll, &Pc, &Pcl);
        double F1 = P - R->phic, F2 = Pl;
        double det = Pc * Pll - Pl * Pcl;
        double dc = (-F1 * Pll + F2 * Pl) / det;
        double dl = (-Pc * F2 + Pcl * F1) / det;
        c += dc;
        lam += dl;
    }
    return acos(c) / LFO_PI;
}

/* sin(i) such that the white-dwarf centre is eclipsed for a full phase width
 * dphi (= roche.findi(q, dphi)); returns 1 if impossible */
static inline int lfo_findi(const lfo_roche *R, double dphi, double maxphi, double *sini)
{
    if (!(dphi > 0.0) || !(dphi < maxphi)) return 1;
    double c = cos(LFO_PI * dphi);
    double u = cos(LFO_PI * maxphi) / c;
    double lam = u * c;
    for (int it = 0; it < 12; ++it) {
        double P, Pu, Pl, Pul, Pll, Pc, Pcl;
        lfo_origin_pot(R, u, c, lam, &P, &Pu, &Pl, &Pul, &Pll, &Pc, &Pcl);
        double F1 = P - R->phic, F2 = Pl;
        double det = Pu * Pll - Pl * Pul;
        double du = (-F1 * Pll + F2 * Pl) / det;
        double dl = (-Pu * F2 + Pul * F1) / det;
        u += du;
        lam += dl;
    }
    if (!(u > 0.0) || !(u <= 1.0)) return 1;
    *sini = u;
    return 0;
}

/* ---- ballistic stream from L1 (roche.bspot): fixed-sequence Gragg-Bulirsch-Stoer ---- */
#ifndef LFO_STREAM_EPS
#define LFO_STREAM_EPS 1e-5   /* start offset from L1 along the unstable manifold */
#endif
#ifndef LFO_STREAM_H0
#define LFO_STREAM_H0 0.6     /* macro step in units of the local dynamical time */
#endif
#ifndef LFO_STREAM_MAXSTEPS
#define LFO_STREAM_MAXSTEPS 400
#endif
#define LFO_GBS_K 6

static inline void lfo_stream_rhs(const lfo_roche *R, const double y[4], double f[4])
{
    double x = y[0], yy = y[1];
    double x2 = x - 1.0;
    double r1sq = x * x + yy * yy, r2sq = x2 * x2 + yy * yy;
    double a1 = R->omu / (r1sq * sqrt(r1sq)), a2 = R->mu / (r2sq * sqrt(r2sq));
    f[0] = y[2];
    f[1] = y[3];
    f[2] = -(a1 * x + a2 * x2 - (x - R->mu)) + 2.0 * y[3];
    f[3] = -((a1 + a2) * yy - yy) - 2.0 * y[2];
}

static inline void lfo_gbs_step(const lfo_roche *R, const double y0[4], double H, double yout[4])
{
    static const int nseq[LFO_GBS_K] = {2, 4, 6, 8, 10, 12};
    double T[LFO_GBS_K][4];
    for (int k = 0; k < LFO_GBS_K; ++k) {
        int n = nseq[k];
        double h = H / n;
        double z0[4], z1[4], f[4];
        lfo_stream_rhs(R, y0, f);
        for (int j = 0; j < 4; ++j) { z0[j] = y0[j]; z1[j] = y0[j] + h * f[j]; }
        for (int m = 1; m < n; ++m) {
            lfo_stream_rhs(R, z1, f);
            for (int j = 0; j < 4; ++j) {
                double t = z0[j] + 2.0 * h * f[j];
                z0[j] = z1[j];
                z1[j] = t;
            }
        }
        lfo_stream_rhs(R, z1, f);
        for (int j = 0; j < 4; ++j) T[k][j] = 0.5 * (z0[j] + z1[j] + h * f[j]);
        for (int m = k - 1; m >= 0; --m) {
            /* in-place Neville in h^2: after this loop T[m] holds the (k-m)-th extrapolation */
            double ratio = (double)nseq[k] / (double)nseq[m];
            double fac = 1.0 / (ratio * ratio - 1.0);
            for (int j = 0; j < 4; ++j) T[m][j] = T[m + 1][j] + (T[m + 1][j] - T[m][j]) * fac;
        }
    }
    for (int j = 0; j < 4; ++j) yout[j] = T[0][j];
}

/* returns 0 and (x, y, vx, vy) where the stream first reaches radius rad from
 * the white dwarf; 1 if it never does (the reference's roche.bspot raises) */
static inline int lfo_bspot(const lfo_roche *R, double rad, double out[4])
{
    if (!(rad > 0.0) || !(rad < R->xl1 - 2.0 * LFO_STREAM_EPS)) return 1;
    double A = R->omu / (R->xl1 * R->xl1 * R->xl1) + R->mu / (R->rs * R->rs * R->rs);
    double l2 = 0.5 * ((A - 2.0) + sqrt(A * (9.0 * A - 8.0)));
    double l1 = sqrt(l2);
    double m1 = (l2 - 2.0 * A - 1.0) / (2.0 * l1);
    double y[4] = {R->xl1 - LFO_STREAM_EPS, -m1 * LFO_STREAM_EPS, -l1 * LFO_STREAM_EPS,
                   -l1 * m1 * LFO_STREAM_EPS};
    for (int step = 0; step < LFO_STREAM_MAXSTEPS; ++step) {
        double x2 = y[0] - 1.0;
        double r1sq = y[0] * y[0] + y[1] * y[1], r2sq = x2 * x2 + y[1] * y[1];
        double w2 = R->omu / (r1sq * sqrt(r1sq)) + R->mu / (r2sq * sqrt(r2sq)) + 1.0;
        double H = LFO_STREAM_H0 / sqrt(w2);
        double yn[4];
        lfo_gbs_step(R, y, H, yn);
        double r0 = sqrt(r1sq), rn = sqrt(yn[0] * yn[0] + yn[1] * yn[1]);
        if (rn <= rad) {
            /* locate the crossing inside this step: Newton on the step length */
            double h = H * (r0 - rad) / (r0 - rn);
            for (int it = 0; it < 8; ++it) {
                lfo_gbs_step(R, y, h, yn);
                rn = sqrt(yn[0] * yn[0] + yn[1] * yn[1]);
                double rdot = (yn[0] * yn[2] + yn[1] * yn[3]) / rn;
                h -= (rn - rad) / rdot;
            }
            lfo_gbs_step(R, y, h, yn);
            for (int j = 0; j < 4; ++j) out[j] = yn[j];
            return 0;
        }
        if (yn[0] * yn[2] + yn[1] * yn[3] >= 0.0) return 1; /* past closest approach */
        for (int j = 0; j < 4; ++j) y[j] = yn[j];
    }
    return 1;
}

#endif
