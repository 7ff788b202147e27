/*
 * oracle/lfit_oracle.c -- CPU FP64 oracle: LFIT CV eclipse model + chi-squared + priors.
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lfit_oracle.h / roche_core.h).
 *
 * Follows the call sites and documented semantics of the reference:
 *   lfit.CV(pars).calcFlux(pars, phase, width)      /root/reference/CVModel.py:128,138
 *   parameter meaning                               /root/reference/README.md:24-43
 *   flux = sum of scaled unit components            /root/reference/testCV.py:59-65
 *   chisq / ln_like                                 /root/reference/CVModel.py:157-191
 *   validity priors                                 /root/reference/CVModel.py:193-324,440-491
 *   Prior.ln_prob, Node.ln_prior, Node.ln_prob      /root/reference/model.py:83-113,426-498
 * The algorithm is the direct one: every surface element gets an ingress and an
 * egress phase from the Roche line-of-sight solve, and every exposure sample sums
 * the elements visible at that phase.  No sorting, no tables, no shortcuts.
 */
#include "lfit_oracle.h"
#include "roche_core.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DEG (LFO_PI / 180.0)

void lfo_default_config(lfo_config *cfg)
{
    cfg->n_wd_rings = 10;
    cfg->n_disc_r = 25;
    cfg->n_disc_th = 40;
    cfg->n_bs = 200;
    cfg->n_donor_th = 18;
    cfg->n_quad = 3;
    cfg->donor_ulimb = 0.8;
    cfg->donor_gdexp = 0.32;
    cfg->solver = LFO_SOLVER_NEWTON;
}

int lfo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ roche API */

int lfo_roche_xl1(double q, double *out) { return lfo_xl1(q, out); }

static int ie_dispatch(int solver, const lfo_roche *R, double si, double ci, const lfo_point *T,
                       double *a, double *b)
{
    if (solver == LFO_SOLVER_ROBUST) return lfo_ingress_egress_robust(R, si, ci, T, a, b);
    return lfo_ingress_egress_newton(R, si, ci, T, a, b);
}

int lfo_roche_findphi(double q, double incl_deg, double *out)
{
    lfo_roche R;
    if (lfo_roche_init(&R, q)) return 1;
    if (!(incl_deg > 0.0) || !(incl_deg <= 90.0)) return 1;
    if (incl_deg == 90.0) { *out = lfo_findphi90(&R); return 0; }
    lfo_point T = {{0.0, 0.0, 0.0}, 0.0, 0.0};
    double a, b;
    if (!lfo_ingress_egress_robust(&R, sin(incl_deg * DEG), cos(incl_deg * DEG), &T, &a, &b)) return 2;
    *out = b - a;
    return 0;
}

int lfo_roche_findi(double q, double dphi, double *incl_deg)
{
    lfo_roche R;
    if (lfo_roche_init(&R, q)) return 1;
    double u;
    if (lfo_findi(&R, dphi, lfo_findphi90(&R), &u)) return 2;
    *incl_deg = asin(u) / DEG;
    return 0;
}

int lfo_roche_bspot(double q, double rad, double out[4])
{
    lfo_roche R;
    if (lfo_roche_init(&R, q)) return 1;
    return lfo_bspot(&R, rad, out) ? 2 : 0;
}

int lfo_roche_ingress_egress(double q, double incl_deg, const double p0[3], double xi, double eta,
                             int solver, double *ph_in, double *ph_out)
{
    lfo_roche R;
    if (lfo_roche_init(&R, q)) return -1;
    lfo_point T = {{p0[0], p0[1], p0[2]}, xi, eta};
    return ie_dispatch(solver, &R, sin(incl_deg * DEG), cos(incl_deg * DEG), &T, ph_in, ph_out);
}

/* ------------------------------------------------------------------ element grids */

typedef struct {
    int n;
    double *in, *out, *w;
    unsigned char *has;
    double total;
} tileset;

static int tileset_alloc(tileset *t, int n)
{
    t->n = n;
    t->in = (double *)malloc(sizeof(double) * n);
    t->out = (double *)malloc(sizeof(double) * n);
    t->w = (double *)malloc(sizeof(double) * n);
    t->has = (unsigned char *)malloc(n);
    t->total = 0.0;
    return !(t->in && t->out && t->w && t->has);
}
static void tileset_free(tileset *t)
{
    free(t->in); free(t->out); free(t->w); free(t->has);
    memset(t, 0, sizeof(*t));
}

/* Orbital phase is periodic and the solvers may return an interval that runs past +-0.5 (elements
 * on the far side of the donor, eclipsed around phase 0.5): those -- has == 2, set by
 * tileset_finish -- are tested against the neighbouring cycles too. */
static inline double tileset_visible(const tileset *t, double ph)
{
    double s = 0.0;
    for (int k = 0; k < t->n; ++k) {
        int ecl = 0;
        if (t->has[k]) {
            const double a = t->in[k], b = t->out[k];
            ecl = ph > a && ph < b;
            if (t->has[k] == 2 && !ecl) ecl = (ph + 1.0 > a && ph + 1.0 < b) || (ph - 1.0 > a && ph - 1.0 < b);
        }
        if (!ecl) s += t->w[k];
    }
    return s;
}

static void tileset_finish(tileset *t)
{
    for (int k = 0; k < t->n; ++k)
        if (t->has[k]) t->has[k] = (t->in[k] < -0.5 || t->out[k] > 0.5) ? 2 : 1;
}

/* white dwarf: limb-darkened disc on the sky, 4*n^2 equal-area tiles */
static int build_wd(const lfo_config *cfg, const lfo_roche *R, double si, double ci, double rwd_a,
                    double ulimb, tileset *t)
{
    int n = cfg->n_wd_rings;
    if (tileset_alloc(t, 4 * n * n)) return 1;
    int idx = 0;
    for (int k = 0; k < n; ++k) {
        double ra = (double)k / n, rb = (double)(k + 1) / n;
        double rho = sqrt(0.5 * (ra * ra + rb * rb));
        double mubar = (2.0 / 3.0) * (pow(1.0 - ra * ra, 1.5) - pow(1.0 - rb * rb, 1.5)) / (rb * rb - ra * ra);
        double w = (1.0 - ulimb) + ulimb * mubar;
        int nk = 4 * (2 * k + 1);
        for (int j = 0; j < nk; ++j, ++idx) {
            double al = (j + 0.5) * LFO_TWOPI / nk;
            lfo_point T = {{0.0, 0.0, 0.0}, rwd_a * rho * cos(al), rwd_a * rho * sin(al)};
            t->has[idx] = (unsigned char)ie_dispatch(cfg->solver, R, si, ci, &T, &t->in[idx], &t->out[idx]);
            t->w[idx] = w;
            t->total += w;
        }
    }
    return 0;
}

/* disc: flat annulus rwd..rdisc (units of a), brightness r^-dexp */
static int build_disc(const lfo_config *cfg, const lfo_roche *R, double si, double ci, double rin,
                      double rout, double dexp, tileset *t)
{
    int nr = cfg->n_disc_r, nt = cfg->n_disc_th;
    if (tileset_alloc(t, nr * nt)) return 1;
    int idx = 0;
    for (int m = 0; m < nr; ++m) {
        double r = rin + (m + 0.5) * (rout - rin) / nr;
        double w = pow(r, 1.0 - dexp);
        for (int j = 0; j < nt; ++j, ++idx) {
            double az = (j + 0.5) * LFO_TWOPI / nt;
            lfo_point T = {{r * cos(az), r * sin(az), 0.0}, 0.0, 0.0};
            t->has[idx] = (unsigned char)ie_dispatch(cfg->solver, R, si, ci, &T, &t->in[idx], &t->out[idx]);
            t->w[idx] = w;
            t->total += w;
        }
    }
    return 0;
}

/* bright spot: strip through the stream impact point along azimuth az */
static int build_bs(const lfo_config *cfg, const lfo_roche *R, double si, double ci, const double imp[4],
                    double len_a, double az_rad, double exp1, double exp2, tileset *t)
{
    int n = cfg->n_bs;
    if (tileset_alloc(t, n)) return 1;
    double smax = pow(exp1 / exp2, 1.0 / exp2);
    double smaxp = pow(smax, exp2);
    double shi = 20.0 + smax;
    double scut = pow(smaxp + 30.0, 1.0 / exp2);
    if (scut < shi) shi = scut;
    double tx = cos(az_rad), ty = sin(az_rad);
    for (int k = 0; k < n; ++k) {
        double s = shi * k / (n - 1);
        double b = (k == 0) ? 0.0 : pow(s / smax, exp1) * exp(smaxp - pow(s, exp2));
        lfo_point T = {{imp[0] + (s - smax) * len_a * tx, imp[1] + (s - smax) * len_a * ty, 0.0}, 0.0, 0.0};
        t->has[k] = (unsigned char)ie_dispatch(cfg->solver, R, si, ci, &T, &t->in[k], &t->out[k]);
        t->w[k] = b;
        t->total += b;
    }
    return 0;
}

/* donor: tiles on the critical lobe; a[k] cos th + b[k] sin th + d[k] = n.earth */
typedef struct {
    int n;
    double *a, *b, *d, *w;
    double norm;
} donorset;

static void donor_free(donorset *D)
{
    free(D->a); free(D->b); free(D->d); free(D->w);
    memset(D, 0, sizeof(*D));
}

static inline double donor_flux(const donorset *D, double ud, double c, double s)
{
    double f = 0.0;
    for (int k = 0; k < D->n; ++k) {
        double m = D->a[k] * c + D->b[k] * s + D->d[k];
        if (m > 0.0) f += D->w[k] * m * (1.0 - ud + ud * m);
    }
    return f;
}

static int donor_count(const lfo_config *cfg)
{
    int nth = cfg->n_donor_th, tot = 0;
    for (int k = 0; k < nth; ++k) {
        double th = (k + 0.5) * LFO_PI / nth;
        int nph = 4 * (int)fmax(1.0, floor(0.5 * nth * sin(th) + 0.5));
        tot += nph;
    }
    return tot;
}

static int build_donor(const lfo_config *cfg, const lfo_roche *R, double si, double ci, donorset *D)
{
    int nth = cfg->n_donor_th;
    int n = donor_count(cfg);
    D->n = n;
    D->a = (double *)malloc(sizeof(double) * n);
    D->b = (double *)malloc(sizeof(double) * n);
    D->d = (double *)malloc(sizeof(double) * n);
    D->w = (double *)malloc(sizeof(double) * n);
    if (!(D->a && D->b && D->d && D->w)) return 1;
    int idx = 0;
    for (int k = 0; k < nth; ++k) {
        double th = (k + 0.5) * LFO_PI / nth;
        int nph = 4 * (int)fmax(1.0, floor(0.5 * nth * sin(th) + 0.5));
        double dth = LFO_PI / nth, dphi = LFO_TWOPI / nph;
        for (int j = 0; j < nph; ++j, ++idx) {
            double ph = (j + 0.5) * dphi;
            double dx = -cos(th), dy = sin(th) * cos(ph), dz = sin(th) * sin(ph);
            /* radius of the critical surface along this direction: bisection */
            double lo = 0.02 * R->rs, hi = R->rs;
            for (int it = 0; it < 100; ++it) {
                double r = 0.5 * (lo + hi);
                if (lfo_pot(R, 1.0 + r * dx, r * dy, r * dz) < R->phic) lo = r; else hi = r;
            }
            double r = 0.5 * (lo + hi);
            double g[3];
            lfo_grad(R, 1.0 + r * dx, r * dy, r * dz, g);
            double gm = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
            double nx = g[0] / gm, ny = g[1] / gm, nz = g[2] / gm;
            double cosb = nx * dx + ny * dy + nz * dz;
            double area = r * r * sin(th) * dth * dphi / cosb;
            D->w[idx] = area * pow(gm, cfg->donor_gdexp);
            D->a[idx] = si * nx;
            D->b[idx] = -si * ny;
            D->d[idx] = ci * nz;
        }
    }
    /* "flux at maximum light": normalised at quadrature, phase 0.25 */
    D->norm = donor_flux(D, cfg->donor_ulimb, 0.0, 1.0);
    return 0;
}

/* ------------------------------------------------------------------ calcFlux */

static void quad_weights(int K, double *off, double *wq)
{
    if (K <= 1) { off[0] = 0.0; wq[0] = 1.0; return; }
    int nint = K - 1; /* even */
    for (int k = 0; k < K; ++k) {
        off[k] = -1.0 + 2.0 * k / nint;
        double c = (k == 0 || k == nint) ? 1.0 : ((k & 1) ? 4.0 : 2.0);
        wq[k] = c / (3.0 * nint);
    }
}

static void fill_nan(int n, double *a)
{
    if (a) for (int i = 0; i < n; ++i) a[i] = NAN;
}

int lfo_calc_flux(const lfo_config *cfg, const double *pars_in, int npars, int flags, int n_ph,
                  const double *phase, const double *width, double *out_total, double *out_wd,
                  double *out_d, double *out_s, double *out_rs)
{
    int status = 0;
    double pars[LFO_NPAR];
    tileset wd = {0}, disc = {0}, bs = {0};
    donorset don = {0};
    for (int i = 0; i < LFO_NPAR; ++i) pars[i] = 0.0;
    if (npars != 14 && npars != 18) { status = 1; goto fail; }
    for (int i = 0; i < npars; ++i) {
        pars[i] = pars_in[i];
        if (!isfinite(pars[i])) { status = 1; goto fail; }
    }
    if (npars == 14) { pars[LFO_EXP1] = 2.0; pars[LFO_EXP2] = 1.0; pars[LFO_TILT] = 90.0; pars[LFO_YAW] = 0.0; }
    if (cfg->n_quad < 1 || cfg->n_quad > 15 || !(cfg->n_quad & 1)) { status = 1; goto fail; }

    const int do_wd = !(flags & LFO_FLAG_SKIP_WD), do_disc = !(flags & LFO_FLAG_SKIP_DISC);
    const int do_bs = !(flags & LFO_FLAG_SKIP_BS), do_don = !(flags & LFO_FLAG_SKIP_DONOR);

    lfo_roche R;
    if (lfo_roche_init(&R, pars[LFO_Q])) { status = 2; goto fail; }
    double si, ci;
    if (flags & LFO_FLAG_INCL) {
        double inc = pars[LFO_DPHI];
        if (!(inc > 0.0) || !(inc <= 90.0)) { status = 3; goto fail; }
        si = sin(inc * DEG);
        ci = cos(inc * DEG);
    } else {
        double maxphi = lfo_findphi90(&R);
        if (lfo_findi(&R, pars[LFO_DPHI], maxphi, &si)) { status = 3; goto fail; }
        ci = sqrt(1.0 - si * si);
    }
    double rwd_a = pars[LFO_RWD] * R.xl1, rdisc_a = pars[LFO_RDISC] * R.xl1;
    if ((do_wd || do_disc) && !(pars[LFO_RWD] > 0.0)) { status = 4; goto fail; }
    if (do_disc && !(rdisc_a > rwd_a)) { status = 4; goto fail; }

    if (do_wd && build_wd(cfg, &R, si, ci, rwd_a, pars[LFO_ULIMB], &wd)) { status = 9; goto fail; }
    if (do_disc && build_disc(cfg, &R, si, ci, rwd_a, rdisc_a, pars[LFO_DEXP], &disc)) { status = 9; goto fail; }

    double beam_a = 0.0, beam_b = 0.0, beam_d = 0.0, beam_norm = 1.0, fis = pars[LFO_FIS];
    if (do_bs) {
        if (!(pars[LFO_SCALE] > 0.0) || !(pars[LFO_EXP1] > 0.0) || !(pars[LFO_EXP2] > 0.0)) { status = 5; goto fail; }
        double imp[4];
        if (lfo_bspot(&R, rdisc_a, imp)) { status = 6; goto fail; }
        if (build_bs(cfg, &R, si, ci, imp, pars[LFO_SCALE] * R.xl1, pars[LFO_AZ] * DEG, pars[LFO_EXP1],
                     pars[LFO_EXP2], &bs)) { status = 9; goto fail; }
        /* beamed fraction leaves along bhat: polar angle tilt from +z, azimuth az - 90 + yaw */
        double tilt = pars[LFO_TILT] * DEG, psib = (pars[LFO_AZ] - 90.0 + pars[LFO_YAW]) * DEG;
        double bx = sin(tilt) * cos(psib), by = sin(tilt) * sin(psib), bz = cos(tilt);
        beam_a = si * bx; beam_b = -si * by; beam_d = ci * bz;
        double cmax = si * sin(tilt) + ci * cos(tilt); /* = cos(i - tilt): best alignment over an orbit */
        beam_norm = fis + (1.0 - fis) * (cmax > 0.0 ? cmax : 0.0);
    }
    if (do_don && build_donor(cfg, &R, si, ci, &don)) { status = 9; goto fail; }
    if (do_wd) tileset_finish(&wd);
    if (do_disc) tileset_finish(&disc);
    if (do_bs) tileset_finish(&bs);

    int K = cfg->n_quad;
    double off[16], wq[16];
    quad_weights(K, off, wq);
    for (int j = 0; j < n_ph; ++j) {
        double ywd = 0.0, yd = 0.0, ys = 0.0, yrs = 0.0;
        for (int k = 0; k < K; ++k) {
            double ph = phase[j] + off[k] * (width ? width[j] : 0.0) - pars[LFO_PHI0];
            ph -= rint(ph); /* [-0.5, 0.5] */
            double th = LFO_TWOPI * ph, c = cos(th), s = sin(th);
            if (do_wd) ywd += wq[k] * tileset_visible(&wd, ph) / wd.total;
            if (do_disc) yd += wq[k] * tileset_visible(&disc, ph) / disc.total;
            if (do_bs) {
                double m = beam_a * c + beam_b * s + beam_d;
                double beam = fis + (1.0 - fis) * (m > 0.0 ? m : 0.0);
                double v = (beam_norm > 0.0 && bs.total > 0.0) ? beam / beam_norm * tileset_visible(&bs, ph) / bs.total : 0.0;
                ys += wq[k] * v;
            }
            if (do_don) yrs += wq[k] * donor_flux(&don, cfg->donor_ulimb, c, s) / don.norm;
        }
        double fwd = pars[LFO_WDFLUX] * ywd, fd = pars[LFO_DFLUX] * yd, fs = pars[LFO_SFLUX] * ys,
               frs = pars[LFO_RSFLUX] * yrs;
        if (out_wd) out_wd[j] = fwd;
        if (out_d) out_d[j] = fd;
        if (out_s) out_s[j] = fs;
        if (out_rs) out_rs[j] = frs;
        if (out_total) out_total[j] = fwd + fd + fs + frs;
    }
    tileset_free(&wd); tileset_free(&disc); tileset_free(&bs); donor_free(&don);
    return 0;
fail:
    tileset_free(&wd); tileset_free(&disc); tileset_free(&bs); donor_free(&don);
    fill_nan(n_ph, out_total); fill_nan(n_ph, out_wd); fill_nan(n_ph, out_d);
    fill_nan(n_ph, out_s); fill_nan(n_ph, out_rs);
    return status;
}

double lfo_chisq(const lfo_config *cfg, const double *pars, int npars, int n_ph, const double *phase,
                 const double *width, const double *y, const double *ye)
{
    double *f = (double *)malloc(sizeof(double) * (n_ph > 0 ? n_ph : 1));
    if (!f) return INFINITY;
    double chi = 0.0;
    if (lfo_calc_flux(cfg, pars, npars, 0, n_ph, phase, width, f, 0, 0, 0, 0)) {
        chi = INFINITY;
    } else {
        for (int j = 0; j < n_ph; ++j) {
            double r = (y[j] - f[j]) / ye[j];
            chi += r * r;
        }
        if (isnan(chi)) chi = INFINITY;
    }
    free(f);
    return chi;
}

/* ------------------------------------------------------------------ priors */

double lfo_prior_ln_prob(int type, double p1, double p2, double norm, double val)
{
    const double LN_SQRT_2PI = 0.91893853320467274178;
    const double LN_MIN_DENORMAL = -744.44007192138126;
    switch (type) {
    case LFO_PRIOR_GAUSSPOS:
        if (val <= 0.0) return -INFINITY;
        /* fall through */
    case LFO_PRIOR_GAUSS: {
        double z = (val - p1) / p2;
        double t = -0.5 * z * z - LN_SQRT_2PI;
        if (!(t >= LN_MIN_DENORMAL)) return -INFINITY; /* scipy's pdf underflows to 0 (model.py:85-89) */
        return t - log(p2);
    }
    case LFO_PRIOR_UNIFORM:
        if (val > p1 && val < p2) return log(1.0 / fabs(p1 - p2));
        return -INFINITY;
    case LFO_PRIOR_LOGUNIFORM:
        if (val > p1 && val < p2) return log(1.0 / norm / val);
        return -INFINITY;
    case LFO_PRIOR_MODJEFF:
        if (val > 0.0 && val < p2) return log(1.0 / norm / (val + p1));
        return -INFINITY;
    }
    return -INFINITY;
}

static inline double fetch(const lfo_layout *L, const double *theta, int src)
{
    return src >= 0 ? theta[src] : L->consts[-src - 1];
}

static double walker_ln_prior(const lfo_layout *L, const double *theta)
{
    /* LCModel.ln_prior (CVModel.py:440-491) */
    double q = fetch(L, theta, L->gather[LFO_Q]), dphi = fetch(L, theta, L->gather[LFO_DPHI]);
    lfo_roche R;
    if (lfo_roche_init(&R, q)) return -INFINITY;
    double maxphi = lfo_findphi90(&R);
    if (!(dphi <= maxphi - 1e-6)) return -INFINITY;
    /* Node.ln_prior over every Param (model.py:426-474) */
    double lnp = 0.0;
    for (int k = 0; k < L->n_prior; ++k) {
        double v = fetch(L, theta, L->prior_src[k]);
        double lp = lfo_prior_ln_prob(L->prior_type[k], L->prior_p1[k], L->prior_p2[k], L->prior_norm[k], v);
        if (!isfinite(lp)) return -INFINITY;
        if (L->prior_isvar[k]) lnp += lp;
    }
    /* SimpleEclipse.ln_prior (CVModel.py:193-324) */
    for (int e = 0; e < L->n_ecl; ++e) {
        const int *g = L->gather + e * LFO_NPAR;
        double rdisc = fetch(L, theta, g[LFO_RDISC]), rwd = fetch(L, theta, g[LFO_RWD]);
        double scale = fetch(L, theta, g[LFO_SCALE]), az = fetch(L, theta, g[LFO_AZ]);
        double rdisc_a = rdisc * R.xl1;
        if (!(rdisc_a <= 0.46)) return -INFINITY;
        if (!(scale <= rwd * 3.0) || !(scale >= rwd / 3.0)) return -INFINITY;
        double imp[4];
        if (lfo_bspot(&R, rdisc_a, imp)) return -INFINITY;
        double alpha = atan2(imp[1], imp[0]) / DEG;
        if (alpha < 0.0) alpha = 90.0 - alpha;
        double tangent = alpha + 90.0;
        double minaz = fmax(0.0, tangent - 80.0), maxaz = fmin(178.0, tangent + 80.0);
        if (!(az >= minaz) || !(az <= maxaz)) return -INFINITY;
    }
    return lnp;
}

static double walker_ln_like(const lfo_config *cfg, const lfo_layout *L, const double *theta, double *chis)
{
    double tot = 0.0;
    for (int e = 0; e < L->n_ecl; ++e) {
        double pars[LFO_NPAR];
        const int *g = L->gather + e * LFO_NPAR;
        for (int i = 0; i < L->npars; ++i) pars[i] = fetch(L, theta, g[i]);
        long long o = L->lc_off[e];
        int n_ph = (int)(L->lc_off[e + 1] - o);
        double chi = lfo_chisq(cfg, pars, L->npars, n_ph, L->lc_phase + o, L->lc_width + o, L->lc_y + o, L->lc_ye + o);
        if (chis) chis[e] = chi;
        tot += -0.5 * chi;
    }
    if (isnan(tot)) tot = -INFINITY;
    return tot;
}

int lfo_log_prob(const lfo_config *cfg, const lfo_layout *L, int what, long long n, const double *theta,
                 double *out, double *chisq_out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (long long w = 0; w < n; ++w) {
        const double *th = theta + w * L->ndim;
        double *chis = chisq_out ? chisq_out + w * L->n_ecl : 0;
        if (chis) for (int e = 0; e < L->n_ecl; ++e) chis[e] = NAN;
        double v;
        if (what == 0) {
            v = walker_ln_prior(L, th);
        } else if (what == 1) {
            v = walker_ln_like(cfg, L, th, chis);
        } else {
            v = walker_ln_prior(L, th);
            if (isfinite(v)) v += walker_ln_like(cfg, L, th, chis);
        }
        out[w] = v;
    }
    return 0;
}
