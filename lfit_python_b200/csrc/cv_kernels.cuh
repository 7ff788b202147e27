// cv_kernels.cuh -- sm_100a kernels of the LFIT CV eclipse model (FP64, no tensor cores:
// the work is root finding, interval scatter and prefix sums, not a contraction).
//
// Reference interfaces replaced (file:line under /root/reference):
//   lfit.CV(pars).calcFlux(pars, phase, width)         CVModel.py:128,138
//   SimpleEclipse.chisq / ln_like                       CVModel.py:157-191
//   LCModel.ln_prior / SimpleEclipse.ln_prior           CVModel.py:440-491,193-324
//   Node.ln_prior / Node.ln_prob, Prior.ln_prob         model.py:426-498,83-113
//   trm.roche.xl1 / findphi / findi / bspot / wdphases  CVModel.py:222,288,460,559,562
//   SimpleGPEclipse.ln_like (george GP)                 CVModel.py:603-696
//
// Pipeline of one log-probability call over n walkers x n_ecl eclipses ("jobs"):
//   walker_kernel    thread per walker: L1, Phi_c, inclination from (q, dphi), Param priors,
//                    scalar validity rules
//   jobcheck_kernel  thread per job: per-eclipse validity rules that need no stream
//   stream_kernel    thread per job, side stream: ballistic stream -> bright-spot impact point,
//                    azimuth rule, strip constants
//   elements_kernel  thread per surface element (white dwarf and donor per walker, disc and
//                    bright spot per job): ingress/egress phases from the Roche LOS solve /
//                    donor surface tiles -> HBM (16-32 B per element)
//   prep_kernel      warp per job: component totals, fixed-point element weights, constants
//   positions_kernel thread per element: the element's eclipse interval as events on the job's sorted
//                    exposure-sample axis (16 B record)
//   donor_table_kernel  CTA per walker: the donor's facing intervals sorted in phase, the five
//                    trigonometric moments of the facing tiles after every interval end
//   flux_kernel      CTA per job, in segments of the axis: tile events -> native 32-bit shared-memory
//                    atomics -> block scan; donor moments looked up in the walker's table; then
//                    component mix, exposure quadrature and the chi-squared reduction
//   gp_kernel        (lfb_set_gp) thread per job: Kalman-filter GP likelihood of the residuals
//   finish_kernel    thread per walker: ln_prior - chi^2/2 with the -inf rules
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lfit_b200.h"
#include "roche_device.cuh"
#include "gp_device.cuh"

namespace lfb {

enum { P_WDFLUX = 0, P_DFLUX, P_SFLUX, P_RSFLUX, P_Q, P_DPHI, P_RDISC, P_ULIMB, P_RWD, P_SCALE, P_AZ,
       P_FIS, P_DEXP, P_PHI0, P_EXP1, P_EXP2, P_TILT, P_YAW };

constexpr int kFluxThreads = 256;
constexpr int kElemThreads = 128;
#ifndef LFB_ELEM_BLOCKS
#define LFB_ELEM_BLOCKS 7
#endif
constexpr int kElemBlocks = LFB_ELEM_BLOCKS;  // resident CTAs per SM the element kernels are compiled for
constexpr int kMaxDonorRings = 128;
constexpr int kMaxQuad = 15;
constexpr int kNoEventPos = 0x7fffffff;
// Fixed-point scale of the normalised tile weights: 2^46, added into shared memory as a 30-bit high and a 16-bit
// low limb (native 32-bit atomics; up to 32767 events of one running sum may meet in one cell)
constexpr double kFix = 70368744177664.0;
constexpr double kInvFix = 1.0 / 70368744177664.0;
constexpr int kLimbBits = 16;
constexpr long long kLimbMask = 0xFFFF;

struct DevLayout {
    int ndim, n_ecl, npars, n_prior;
    const int* gather;
    const double* consts;
    const int *psrc, *ptype, *pisvar;
    const double *pp1, *pp2, *pnorm;
};

// Light curves as the flux kernel wants them: per eclipse, the K exposure samples of every
// point merged and sorted in phase (wrapped to [-0.5, 0.5)), with cos/sin of 2 pi phase and
// the map from (point, quadrature node) to sorted position.
struct DevSamples {
    const long long* lc_off;     // [n_ecl + 1] data-point offsets
    const double *y, *iye;       // [total] in phase order: flux, 1 / error
    const double *S, *cosS, *sinS;  // [K * total] sorted per eclipse
    const int* bins;             // [K * total + n_ecl] per eclipse M + 1 entries (see SampleAxis)
    const double4* axis;         // [n_ecl] first and last sample phase, bins per unit phase (see SampleAxis)
    const int* pos;              // [total * K] sorted position of each (point, node); points in phase order
    const int* pt_index;         // [total] original index of each phase-ordered point
    const double *gp_x, *gp_var; // [total] GP likelihood: raw phases ascending per eclipse, noise variances
    const int* gp_slot;          // [total] rank in that order of each phase-ordered point
    const double2* gp_span;      // [n_ecl] smallest and largest raw phase
    const long long* chunk_off;  // [n_ecl + 1] offsets into chunks
    const int4* chunks;          // per segment: first point, one past last point, first sample, last sample
    const double* seg_tr;        // per segment [3][RP][NT]: phase, cos, sin of its samples in the flux kernel's
                                 // thread-major order (sample t * RP + r of the segment at [r][t])
};

struct GridCfg {
    int n_wd_rings, n_wd, n_disc_r, n_disc_th, n_disc, n_bs, n_donor_th, n_donor_q, n_quad;
    int n_wd_half, n_disc_half;  // elements solved (the other half follows by the y -> -y mirror)
    double donor_ulimb, donor_gdexp;
    const int* donor_ring_off;  // [n_donor_th + 1] offsets of each ring's quarter tiles
    const unsigned short* rec_widx;  // [n_wd + n_disc + n_bs] index of each stored tile record's weight in the job's weight table
    const int* rec_slot;             // [n_wd + n_disc + n_bs] where tile record i is stored: neighbouring tiles far apart
    const int* disc_order;      // [n_disc_half] disc elements ordered along the line of centres, so that the
                                // threads of a warp solve elements of the same kind (deep / shallow / never eclipsed)
    const double4* disc_geo;    // [n_disc_half] in disc_order: cos, sin of the sector azimuth, (ring + 0.5) / n_disc_r, tile
    const double2* wd_geo;      // [n_wd_half] offset of a white-dwarf tile on the sky, in units of the star's radius
    double quad_off[kMaxQuad], quad_w[kMaxQuad];
};

struct WalkerScal {
    Roche R;
    double si, ci;
    double lnprior;
    Roots wd;    // grazing lines of sight of the white-dwarf centre: Newton starts for its tiles
    int status;  // 0: a model exists; else the parameters admit none
    int wd_ok;   // wd is set
};

struct JobScal {
    double xs, ys;             // stream impact point
    double smax, smaxp, shi;   // bright-spot strip: profile peak, peak^exp2, strip length (scale units)
    double lsmax, caz, saz;    // ln(smax); direction of the strip (cos, sin of az)
    int status;                // 0 ok, 1 walker invalid, 2 stream misses disc, 3 bad parameter, 4 not needed
    int veto;                  // what the stream ODE found (it runs beside other kernels and writes nothing they read):
                               // 0 nothing, 2 the stream misses the disc, 5 azimuth outside its window.  Merged into
                               // status / the walker's ln_prior after the join (prep_strip_kernel).
    int ev_lo, ev_hi;          // sample positions spanned by the eclipse events of the job's tiles
};

// walker of a job / unit (a batch holds far fewer than 2^31 jobs: 32-bit division, a fraction of the 64-bit one's cost)
__device__ __forceinline__ long long walker_of(long long job, int n_ecl)
{
    return (long long)((unsigned)job / (unsigned)n_ecl);
}

__device__ __forceinline__ double fetch(const DevLayout& L, const double* th, int src)
{
    return src >= 0 ? th[src] : L.consts[-src - 1];
}

// ---------------------------------------------------------------- walker_kernel
__global__ void walker_kernel(DevLayout L, int what, int flags, long long n, const double* __restrict__ theta,
                              WalkerScal* __restrict__ ws)
{
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const double* th = theta + w * L.ndim;
    WalkerScal W;
    W.status = 0;
    W.wd_ok = 0;
    W.lnprior = 0.0;
    W.si = 1.0;
    W.ci = 0.0;
    double q = fetch(L, th, L.gather[P_Q]), dphi = fetch(L, th, L.gather[P_DPHI]);
    double maxphi = 0.0;
    if (!isfinite(q) || !isfinite(dphi) || !roche_init(q, W.R)) {
        W.status = 1;
        W.R.mu = W.R.omu = W.R.xl1 = W.R.rs = W.R.phic = W.R.rin = 0.0;
    } else if (flags & LFB_FLAG_INCL) {
        if (!(dphi > 0.0) || !(dphi <= 90.0)) W.status = 1;
        else sincos_(dphi * kDeg, &W.si, &W.ci);
    } else {
        maxphi = findphi90(W.R);
        if (!findi(W.R, dphi, maxphi, W.si)) W.status = 1;
        else W.ci = sqrt(1.0 - W.si * W.si);
    }
    if (what != LFB_LN_LIKE) {
        double lnp = 0.0;
        // LCModel.ln_prior (CVModel.py:440-491): roche failure or dphi beyond the edge-on width
        if (!isfinite(q) || !(q > 0.0) || !(q < 1e6)) lnp = -INFINITY;
        else if (!(dphi <= maxphi - 1e-6)) lnp = -INFINITY;
        // Node.ln_prior (model.py:426-474): any invalid Param -> -inf, variable ones add up
        for (int k = 0; k < L.n_prior && lnp > -INFINITY; ++k) {
            double lp = prior_ln_prob(L.ptype[k], L.pp1[k], L.pp2[k], L.pnorm[k], fetch(L, th, L.psrc[k]));
            if (!isfinite(lp)) lnp = -INFINITY;
            else if (L.pisvar[k]) lnp += lp;
        }
        // SimpleEclipse.ln_prior (CVModel.py:217-276): disc radius and spot scale windows
        for (int e = 0; e < L.n_ecl && lnp > -INFINITY; ++e) {
            const int* g = L.gather + e * LFB_NPAR;
            double rdisc = fetch(L, th, g[P_RDISC]), rwd = fetch(L, th, g[P_RWD]), scale = fetch(L, th, g[P_SCALE]);
            if (!(rdisc * W.R.xl1 <= 0.46)) lnp = -INFINITY;
            if (!(scale <= rwd * 3.0) || !(scale >= rwd / 3.0)) lnp = -INFINITY;
        }
        W.lnprior = lnp;
    }
    ws[w] = W;
}

// ---------------------------------------------------------------- wdcentre_kernel
// Thread per walker, beside the disc solves: the grazing lines of sight of the white-dwarf centre,
// which the white-dwarf tiles (all within a few hundredths of it) take as their Newton starts.
__global__ void wdcentre_kernel(int what, long long n, WalkerScal* __restrict__ ws)
{
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    WalkerScal& W = ws[w];
    if (W.status != 0 || (what != LFB_LN_LIKE && !(W.lnprior > -INFINITY))) return;
    const Roche R = W.R;
    const Point T = {0.0, 0.0, 0.0, 0.0, 0.0};
    Roots r;
    r.lam[0] = NAN;
    double pin, pout;
    if (ingress_egress(R, W.si, W.ci, T, &pin, &pout, nullptr, &r) && r.lam[0] == r.lam[0]) {
        W.wd = r;
        W.wd_ok = 1;
    }
}

// ---------------------------------------------------------------- jobcheck_kernel / stream_kernel
// jobcheck: which jobs are worth evaluating, and the bright-spot strip constants.
__global__ void jobcheck_kernel(DevLayout L, int what, int flags, long long njobs, const double* __restrict__ theta,
                                const WalkerScal* __restrict__ ws, JobScal* __restrict__ js)
{
    long long job = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= njobs) return;
    long long w = walker_of(job, L.n_ecl);
    int e = (int)(job - w * L.n_ecl);
    JobScal J;
    J.xs = J.ys = 0.0;
    J.smax = J.smaxp = J.shi = 1.0;
    J.lsmax = 0.0;
    J.caz = 1.0;
    J.saz = 0.0;
    J.status = 0;
    J.veto = 0;
    J.ev_lo = kNoEventPos;
    J.ev_hi = -2;
    const WalkerScal& W = ws[w];
    if (W.status != 0) {
        J.status = 1;
    } else if (what != LFB_LN_LIKE && !(W.lnprior > -INFINITY)) {
        J.status = 4;  // the prior already vetoed this walker: nothing downstream is evaluated
    } else {
        const double* th = theta + w * L.ndim;
        const int* g = L.gather + e * LFB_NPAR;
        const bool do_wd = !(flags & LFB_FLAG_SKIP_WD), do_disc = !(flags & LFB_FLAG_SKIP_DISC);
        const bool do_bs = !(flags & LFB_FLAG_SKIP_BS);
        bool finite_all = true;
        for (int k = 0; k < L.npars; ++k) finite_all = finite_all && isfinite(fetch(L, th, g[k]));
        double rwd = fetch(L, th, g[P_RWD]), rdisc = fetch(L, th, g[P_RDISC]);
        double exp1 = L.npars > P_EXP1 ? fetch(L, th, g[P_EXP1]) : 2.0;
        double exp2 = L.npars > P_EXP2 ? fetch(L, th, g[P_EXP2]) : 1.0;
        double scale = fetch(L, th, g[P_SCALE]);
        if (!finite_all || ((do_wd || do_disc) && !(rwd > 0.0)) || (do_disc && !(rdisc > rwd)) ||
            (do_bs && (!(scale > 0.0) || !(exp1 > 0.0) || !(exp2 > 0.0)))) {
            J.status = 3;
        } else if (do_bs) {
            J.smax = pow(exp1 / exp2, 1.0 / exp2);
            J.smaxp = pow(J.smax, exp2);
            J.shi = fmin(20.0 + J.smax, pow(J.smaxp + 30.0, 1.0 / exp2));
            J.lsmax = log(J.smax);
            sincos_(fetch(L, th, g[P_AZ]) * kDeg, &J.saz, &J.caz);
        }
    }
    js[job] = J;
}

// stream: ballistic stream from L1 to the disc edge -> bright-spot impact point; the azimuth rule
// of SimpleEclipse.ln_prior.  A serial ODE per job: launched on a side stream so that it overlaps
// the element solves that do not need it.
// LANES: eight lanes per job (bspot_lanes, a third of the latency, three times the issue slots) -- for
// batches too small to hide the ODE behind the element solves; else one thread per job.
template <bool LANES>
__global__ void stream_kernel(DevLayout L, int what, int flags, long long njobs, const double* __restrict__ theta,
                              const WalkerScal* __restrict__ ws, JobScal* js)
{
    // (a group of lanes enters and leaves together)
    long long job = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> (LANES ? 3 : 0);
    if (job >= njobs || (flags & LFB_FLAG_SKIP_BS)) return;
    if (js[job].status != 0) return;
    long long w = walker_of(job, L.n_ecl);
    int e = (int)(job - w * L.n_ecl);
    const Roche R = ws[w].R;
    const double* th = theta + w * L.ndim;
    const int* g = L.gather + e * LFB_NPAR;
    double rdisc_a = fetch(L, th, g[P_RDISC]) * R.xl1;
    double imp[4];
    const bool hit = LANES ? bspot_lanes(R, rdisc_a, imp) : bspot(R, rdisc_a, imp);
    if (LANES && (threadIdx.x & 7) != 0) return;  // one lane of the group writes
    if (!hit) {
        js[job].veto = 2;  // the stream misses the disc (roche.bspot raises, CVModel.py:309-316)
        return;
    }
    js[job].xs = imp[0];
    js[job].ys = imp[1];
    if (what != LFB_LN_LIKE) {
        // azimuth window about the disc tangent at the impact point (CVModel.py:282-307)
        double az = fetch(L, th, g[P_AZ]);
        double alpha = atan2(imp[1], imp[0]) / kDeg;
        if (alpha < 0.0) alpha = 90.0 - alpha;
        double tangent = alpha + 90.0;
        double minaz = fmax(0.0, tangent - 80.0), maxaz = fmin(178.0, tangent + 80.0);
        if (!(az >= minaz) || !(az <= maxaz)) js[job].veto = 5;
    }
}

// veto_kernel: thread per walker, right after the join with the stream ODE: what the ODE found for the walker's
// eclipses becomes part of the walker's state -- a stream that misses the disc marks its job (no model:
// CVModel.py:309-316), and either finding vetoes the walker's prior (SimpleEclipse.ln_prior, CVModel.py:282-316).
// One writer per walker, and nothing else runs on these fields at this point.
__global__ void veto_kernel(int what, int n_ecl, long long n, WalkerScal* __restrict__ ws, JobScal* __restrict__ js)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    bool any = false;
    for (int e = 0; e < n_ecl; ++e) {
        JobScal& J = js[w * n_ecl + e];
        const int v = J.veto;
        if (v == 0) continue;
        any = true;
        if (v == 2 && J.status == 0) J.status = 2;
    }
    if (any && what != LFB_LN_LIKE) ws[w].lnprior = -INFINITY;
}

// ---------------------------------------------------------------- elements_kernel
struct ElemArgs {
    DevLayout L;
    GridCfg G;
    int what, flags;
    long long n, njobs;
    const double* theta;
    const WalkerScal* ws;
    const JobScal* js;
    double2* wd_io;    // [n][n_wd_half]       ingress, egress (cycles); ingress = kBig: never eclipsed
    double4* don;      // [n][n_donor_q]       outward normal and weight of a quarter tile
    double2* disc_io;  // [njobs][n_disc_half]
    double2* bs_io;    // [njobs][n_bs]
    double* bs_b;      // [njobs][n_bs]        strip brightness
};

// a walker's elements are worth computing if some job of it will be evaluated
__device__ __forceinline__ bool walker_live(const ElemArgs& A, const WalkerScal& W)
{
    if (W.status != 0) return false;
    if (A.what != LFB_LN_LIKE && !(W.lnprior > -INFINITY)) return false;
    return true;
}

// COMP 0: white dwarf (per walker), 1: disc (per job), 2: bright spot (per job), 3: donor (per walker)
template <int COMP>
__global__ void __launch_bounds__(kElemThreads, kElemBlocks) elements_kernel(const __grid_constant__ ElemArgs A)
{
    const GridCfg& G = A.G;
    const int per_unit = COMP == 0 ? G.n_wd_half : COMP == 1 ? G.n_disc_half : COMP == 2 ? G.n_bs : G.n_donor_q;
    // disc tiles keep whole warps per job (their order is tuned to warps, see disc_order); the other
    // components are dealt to threads back to back -- a warp may straddle two walkers, no idle lanes
    const int padded = COMP == 1 ? (per_unit + 31) & ~31 : per_unit;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long unit = gid < 0xffffffffLL ? (long long)((unsigned)gid / (unsigned)padded) : gid / padded;
    const int t = (int)(gid - unit * padded);
    const long long nunits = (COMP == 0 || COMP == 3) ? A.n : A.njobs;
    if (unit >= nunits || t >= per_unit) return;
    const long long w = (COMP == 0 || COMP == 3) ? unit : walker_of(unit, A.L.n_ecl);
    const int e = (COMP == 0 || COMP == 3) ? 0 : (int)(unit - w * A.L.n_ecl);
    const WalkerScal& W = A.ws[w];
    if (!walker_live(A, W)) return;
    const double* th = A.theta + w * A.L.ndim;
    const int* g = A.L.gather + e * LFB_NPAR;
    const Roche R = W.R;
    const double si = W.si, ci = W.ci;

    if (COMP == 3) {
        // donor: quarter (y > 0, z > 0) of the tiles on the critical surface
        int k = 0;
        while (G.donor_ring_off[k + 1] <= t) ++k;
        int j = t - G.donor_ring_off[k], mk = G.donor_ring_off[k + 1] - G.donor_ring_off[k];
        double sth, cth, sph, cph;
        double dth = kPi / G.n_donor_th, dph = kTwoPi / (4 * mk);
        sincos_((k + 0.5) * dth, &sth, &cth);
        sincos_((j + 0.5) * dph, &sph, &cph);
        double dx = -cth, dy = sth * cph, dz = sth * sph, gr[3];
        double r = donor_radius(R, dx, dy, dz, gr);
        double gm = sqrt(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
        double nx = gr[0] / gm, ny = gr[1] / gm, nz = gr[2] / gm;
        double area = r * r * sth * dth * dph / (nx * dx + ny * dy + nz * dz);
        A.don[unit * G.n_donor_q + t] = make_double4(nx, ny, nz, area * pow(gm, G.donor_gdexp));
        return;
    }

    Point T = {0.0, 0.0, 0.0, 0.0, 0.0};
    double wt = 0.0;
    int tile = t;  // where the result goes
    if (COMP == 0) {
        // white dwarf: limb-darkened disc on the sky, ring k, tiles with cos(alpha) > 0
        const double rwd_a = fetch(A.L, th, g[P_RWD]) * R.xl1;
        if (!(rwd_a > 0.0) || !isfinite(rwd_a)) return;
        const double2 og = __ldg(G.wd_geo + t);
        T.xi = rwd_a * og.x;
        T.eta = rwd_a * og.y;
    } else {
        const JobScal& J = A.js[unit];
        if (J.status != 0) return;
        if (COMP == 2 && J.veto != 0) return;  // (the strip is solved after the join with the stream ODE)
        if (COMP == 1) {
            // disc: ring m, sector j on the y > 0 side
            const double rwd_a = fetch(A.L, th, g[P_RWD]) * R.xl1, rdisc_a = fetch(A.L, th, g[P_RDISC]) * R.xl1;
            const double2 cs = __ldg((const double2*)(G.disc_geo + t)), ft = __ldg((const double2*)(G.disc_geo + t) + 1);
            tile = (int)ft.y;
            const double r = rwd_a + ft.x * (rdisc_a - rwd_a);
            T.x = r * cs.x;
            T.y = r * cs.y;
        } else {
            // bright spot: strip through the stream impact point along azimuth az
            const double exp1 = A.L.npars > P_EXP1 ? fetch(A.L, th, g[P_EXP1]) : 2.0;
            const double exp2 = A.L.npars > P_EXP2 ? fetch(A.L, th, g[P_EXP2]) : 1.0;
            const double s = J.shi * t / (G.n_bs - 1);
            // brightness (s / smax)^exp1 exp(smax^exp2 - s^exp2), through one logarithm
            const double ls = log(s);
            wt = t == 0 ? 0.0 : exp(exp1 * (ls - J.lsmax) + (J.smaxp - exp(exp2 * ls)));
            const double len = (s - J.smax) * fetch(A.L, th, g[P_SCALE]) * R.xl1;
            T.x = J.xs + len * J.caz;
            T.y = J.ys + len * J.saz;
        }
    }
    double pin, pout;
    // (disc and strip elements lie in the orbital plane: the solver's PLANAR form)
    if (!ingress_egress<COMP == 1 || COMP == 2>(R, si, ci, T, &pin, &pout, (COMP == 0 && W.wd_ok) ? &W.wd : nullptr)) { pin = kBig; pout = -kBig; }
    if (COMP == 0) A.wd_io[unit * G.n_wd_half + t] = make_double2(pin, pout);
    else if (COMP == 1) A.disc_io[unit * G.n_disc_half + tile] = make_double2(pin, pout);
    else {
        A.bs_io[unit * G.n_bs + t] = make_double2(pin, pout);
        A.bs_b[unit * G.n_bs + t] = wt;
    }
}

// ---------------------------------------------------------------- flux stage: prep / positions / donor table / flux
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA, result in every thread.  Fixed tree: deterministic.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red)
{
    v = warp_sum(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += red[i];
    return t;
}

constexpr int kNoEvent = kNoEventPos;  // event position: never happens
constexpr int kAtStart = -1;          // event position: before the first sample

// Per-job constants of the flux stage (written by prep_kernel)
struct JobConst {
    double f_wd, f_d, f_s, f_rs;        // flux scale of each component
    double F01;                         // |f_wd| + |f_d|: scale of the merged white-dwarf + disc running sum (chi-squared mode)
    double beam_a, beam_b, beam_d, fis; // bright-spot beaming: fis + (1-fis) max(0, a c + b s + d)
    double cphi, sphi, phi0w;           // phase offset wrapped to [-0.5, 0.5]
};

// A walker's donor curve as a table over orbital phase (donor_table_kernel): every tile image faces the observer
// on one phase interval, so the five trigonometric moments of the facing tiles are piecewise constant; the table
// holds the sorted break points and the moments between them.
struct DonorTable {
    unsigned short* first;  // [n][kDonorBins + 1] break points in the phase bins before bin g
    double* mom;            // [n][nb_max + 1][6]  row k: moments (1, c, s, c^2, c s) after k break points, normalised
                            //                     "at maximum light" (quadrature), then key[k], the sorted break point
                            //                     that ends the row (+inf for the last row).  An opening break point
                            //                     is nudged up by one ulp: break point k applies to phase x iff key[k] <= x
    int nb_max;             // 8 n_donor_q
};

struct FluxArgs {
    DevLayout L;
    GridCfg G;
    DevSamples smp;
    int what, flags, mode;  // mode 0: chi-squared, 1: flux curves
    int ni_total;           // event records per job: n_wd + n_disc + n_bs
    long long njobs;
    const double* theta;
    const WalkerScal* ws;
    JobScal* js;
    const double2* wd_io;
    const double4* don;
    const double2* disc_io;
    const double2* bs_io;
    const double* bs_b;
    JobConst* jc;           // [njobs]
    long long* wq;          // [njobs][n_wd_rings + n_disc_r + n_bs] fixed-point element weights
    ulonglong2* ivp;        // [njobs][ni_total] event records (EventRec) of every eclipse interval
    DonorTable dt;
    double* chisq_job;      // [njobs] chi-squared of each job (NaN: not evaluated)
    double* flux_tot;       // mode 1: [njobs][n_ph]
    double* flux_comp;      // mode 1 (optional): [4][njobs][n_ph]
    // Gaussian-process likelihood instead of chi-squared (lfb_set_gp): residuals y - model, one row per data
    // point (ascending raw phase, all eclipses back to back), one column per walker of the batch
    double* gp_resid;       // mode 0: [total points][n walkers] or null
    long long n_walkers;
    int gp_src[3];          // where ln_ampin_gp, ln_ampout_gp, ln_tau_gp come from (theta column / constant slot)
    const double* gp_dist;  // [n_ecl] distance of the change points from mid-eclipse
};

__device__ __forceinline__ bool job_live(const FluxArgs& A, const WalkerScal& W, const JobScal& J)
{
    if (J.status != 0) return false;
    if (A.what != LFB_LN_LIKE && !(W.lnprior > -INFINITY)) return false;  // prior said no: not evaluated
    return true;
}

// prep_kernel: one warp per job.  Component totals ("flux at maximum light", README.md:24-28),
// fixed-point (2^-46) element weights, beaming and phase-offset constants.  In chi-squared mode the white
// dwarf and the disc share one running sum: their weights carry f_wd / F01 and f_d / F01.
__global__ void __launch_bounds__(128) prep_kernel(const __grid_constant__ FluxArgs A)
{
    const GridCfg& G = A.G;
    const int lane = threadIdx.x & 31;
    const long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (job >= A.njobs) return;
    const long long w = walker_of(job, A.L.n_ecl);
    const int e = (int)(job - w * A.L.n_ecl);
    const WalkerScal& W = A.ws[w];
    const JobScal& J = A.js[job];
    if (!job_live(A, W, J)) return;
    const double* th = A.theta + w * A.L.ndim;
    const int* g = A.L.gather + e * LFB_NPAR;
    auto par = [&](int k, double dflt) { return k < A.L.npars ? fetch(A.L, th, g[k]) : dflt; };
    const bool do_wd = !(A.flags & LFB_FLAG_SKIP_WD), do_disc = !(A.flags & LFB_FLAG_SKIP_DISC);
    const bool do_bs = !(A.flags & LFB_FLAG_SKIP_BS), do_don = !(A.flags & LFB_FLAG_SKIP_DONOR);
    const double si = W.si, ci = W.ci, xl1 = W.R.xl1;
    const double rwd_a = par(P_RWD, 0.0) * xl1, rdisc_a = par(P_RDISC, 0.0) * xl1;
    const double ulimb = par(P_ULIMB, 0.0), dexp = par(P_DEXP, 0.0);
    const double f_wd = do_wd ? par(P_WDFLUX, 0.0) : 0.0, f_d = do_disc ? par(P_DFLUX, 0.0) : 0.0;
    const double F01 = fabs(f_wd) + fabs(f_d);
    // weight of a white-dwarf / disc tile relative to its own component (flux-curve mode) or to both (chi-squared mode)
    const double s_wd = A.mode ? (do_wd ? 1.0 : 0.0) : (F01 > 0.0 ? f_wd / F01 : 0.0);
    const double s_d = A.mode ? (do_disc ? 1.0 : 0.0) : (F01 > 0.0 ? f_d / F01 : 0.0);
    long long* wq_wd = A.wq + job * (G.n_wd_rings + G.n_disc_r + G.n_bs);
    long long* wq_disc = wq_wd + G.n_wd_rings;
    // white dwarf rings: equal-area tiles, weight (1 - u) + u <mu>_ring
    double p = 0.0;
    for (int k = lane; k < G.n_wd_rings; k += 32) {
        double inv = 1.0 / G.n_wd_rings, ra = k * inv, rb = (k + 1) * inv;
        double ua = 1.0 - ra * ra, ub = 1.0 - rb * rb;
        double mubar = (2.0 / 3.0) * (ua * sqrt(ua) - ub * sqrt(ub)) / (rb * rb - ra * ra);
        p += 4.0 * (2 * k + 1) * ((1.0 - ulimb) + ulimb * mubar);
    }
    const double tot_wd = warp_sum(p);
    for (int k = lane; k < G.n_wd_rings; k += 32) {
        double inv = 1.0 / G.n_wd_rings, ra = k * inv, rb = (k + 1) * inv;
        double ua = 1.0 - ra * ra, ub = 1.0 - rb * rb;
        double mubar = (2.0 / 3.0) * (ua * sqrt(ua) - ub * sqrt(ub)) / (rb * rb - ra * ra);
        wq_wd[k] = llrint(((1.0 - ulimb) + ulimb * mubar) / tot_wd * (s_wd * kFix));
    }
    // disc rings: brightness r^-dexp times area ~ r
    p = 0.0;
    for (int m = lane; m < G.n_disc_r; m += 32) {
        double r = rwd_a + (m + 0.5) * (rdisc_a - rwd_a) / G.n_disc_r;
        p += G.n_disc_th * pow(r, 1.0 - dexp);
    }
    const double tot_d = warp_sum(p);
    for (int m = lane; m < G.n_disc_r; m += 32) {
        double r = rwd_a + (m + 0.5) * (rdisc_a - rwd_a) / G.n_disc_r;
        wq_disc[m] = llrint(pow(r, 1.0 - dexp) / tot_d * (s_d * kFix));
    }
    // (the bright-spot strip's weights wait for the strip itself: prep_strip_kernel; the donor's normalisation
    // is in its table: donor_table_kernel)
    if (lane == 0) {
        JobConst C;
        // beamed part of the spot: polar angle tilt from +z, azimuth az - 90 + yaw
        const double fis = par(P_FIS, 0.0);
        double st, ct, sp, cp;
        sincos_(par(P_TILT, 90.0) * kDeg, &st, &ct);
        sincos_((par(P_AZ, 0.0) - 90.0 + par(P_YAW, 0.0)) * kDeg, &sp, &cp);
        C.beam_a = si * st * cp;
        C.beam_b = -si * st * sp;
        C.beam_d = ci * ct;
        C.fis = fis;
        double cmax = si * st + ci * ct;  // cos(i - tilt): best alignment over an orbit
        double beam_norm = fis + (1.0 - fis) * (cmax > 0.0 ? cmax : 0.0);
        C.f_wd = f_wd;
        C.f_d = f_d;
        C.F01 = F01;
        C.f_s = (do_bs && beam_norm > 0.0) ? par(P_SFLUX, 0.0) / beam_norm : 0.0;  // 0 if the strip is dark: prep_strip_kernel
        C.f_rs = do_don ? par(P_RSFLUX, 0.0) : 0.0;
        double phi0w = par(P_PHI0, 0.0);
        phi0w -= rint(phi0w);
        C.phi0w = phi0w;
        sincos_(kTwoPi * phi0w, &C.sphi, &C.cphi);
        A.jc[job] = C;
    }
}

// Sorted sample phases of one eclipse plus a bin table: bins[b] is the first sample whose phase
// falls in or after the b-th of M equal phase bins, so a search is a table look-up and a short
// walk instead of log2(M) dependent loads.
struct SampleAxis {
    const double* S;
    const int* bins;  // [M + 1]
    int M;
    double s_first, s_last, inv_binw;
};

__device__ __forceinline__ SampleAxis sample_axis(const DevSamples& smp, int e, int K)
{
    SampleAxis X;
    const long long lc0 = smp.lc_off[e];
    X.M = (int)(smp.lc_off[e + 1] - lc0) * K;
    X.S = smp.S + lc0 * K;
    X.bins = smp.bins + lc0 * K + e;
    const double2 fl = __ldg((const double2*)(smp.axis + e));
    X.s_first = fl.x;
    X.s_last = fl.y;
    X.inv_binw = __ldg((const double*)(smp.axis + e) + 2);
    return X;
}

// First index with S[idx] > v (strict) or S[idx] >= v (!strict); M if none.  The bin only seeds
// the walk, so host/device rounding of the bin index cannot change the answer.
__device__ __forceinline__ int sample_search(const SampleAxis& X, double v, bool strict)
{
    const double* __restrict__ S = X.S;
    const int M = X.M;
    double gf = (v - X.s_first) * X.inv_binw;
    int b = gf <= 0.0 ? 0 : (gf >= (double)(M - 1) ? M - 1 : (int)gf);
    int p = __ldg(X.bins + b);
    // the answer is almost always within a few samples of the seed: fetch a window at once instead of
    // walking load by load
    const double inf = 1e300;
    const double sm1 = p > 0 ? __ldg(S + p - 1) : -inf;
    const double s0 = p < M ? __ldg(S + p) : inf, s1 = p + 1 < M ? __ldg(S + p + 1) : inf;
    const double s2 = p + 2 < M ? __ldg(S + p + 2) : inf, s3 = p + 3 < M ? __ldg(S + p + 3) : inf;
    const bool b0 = strict ? (s0 <= v) : (s0 < v), b1 = strict ? (s1 <= v) : (s1 < v);
    const bool b2 = strict ? (s2 <= v) : (s2 < v), b3 = strict ? (s3 <= v) : (s3 < v);
    const bool bm = strict ? (sm1 <= v) : (sm1 < v);
    if (b0) {
        p += b1 ? (b2 ? (b3 ? 4 : 3) : 2) : 1;
        if (b1 && b2 && b3)
            while (p < M) {
                double s = __ldg(S + p);
                if (!(strict ? (s <= v) : (s < v))) break;
                ++p;
            }
    } else if (!bm && p > 0) {
        --p;
        while (p > 0) {
            double s = __ldg(S + p - 1);
            if (strict ? (s <= v) : (s < v)) break;
            --p;
        }
    }
    return p;
}

// Event record of one eclipse / facing interval: up to three pieces (open, close) of sample
// positions, six 21-bit fields in 128 bits.  Field value 0 = before the first sample
// (kAtStart), 0x1FFFFF = no event, else position + 1.  Positions ascend from field 0 to 5.
typedef ulonglong2 EventRec;
constexpr unsigned kFieldNone = 0x1FFFFFu;
constexpr int kMaxSamples = (1 << 21) - 3;

__device__ __forceinline__ unsigned long long enc_pos(int p)
{
    return p == kNoEvent ? (unsigned long long)kFieldNone : (unsigned long long)(unsigned)(p + 1);
}
__device__ __forceinline__ int dec_pos(unsigned long long word, int k)
{
    unsigned f = (unsigned)(word >> (21 * k)) & kFieldNone;
    return f == kFieldNone ? kNoEvent : (int)f - 1;
}
__device__ __forceinline__ EventRec no_events()
{
    const unsigned long long w = (unsigned long long)kFieldNone | ((unsigned long long)kFieldNone << 21) |
                                 ((unsigned long long)kFieldNone << 42);
    return make_ulonglong2(w, w);
}

// The samples with a < S + n < b for integer n.  Points are wrapped to [-0.5, 0.5) as a whole
// and exposures are shorter than an orbit, so the axis spans less than two cycles and b - a < 1:
// at most three shifts contribute.  Each piece is (first sample inside, first sample at or past
// the end); kAtStart = open before the first sample, kNoEvent = nothing.
__device__ __forceinline__ EventRec interval_pieces(const SampleAxis& X, double a, double b, int& lo, int& hi)
{
    // lo / hi: the span of the record's events on the axis as rec_first / rec_last_close would read them back
    // (kNoEvent / -2 for an empty record; hi = kNoEvent when the last piece stays open to the end)
    int ev[6] = {kNoEvent, kNoEvent, kNoEvent, kNoEvent, kNoEvent, kNoEvent};
    lo = kNoEvent;
    hi = -2;
    if ((a < b) && (a > -1e29) && (b < 1e29) && X.M > 0) {
        int n_lo = (int)ceil(X.s_first - b), n_hi = (int)floor(X.s_last - a);
        int np = 0;
        for (int n = n_lo; n <= n_hi && np < 3; ++n) {
            double an = a + n, bn = b + n;
            int po = an < X.s_first ? 0 : sample_search(X, an, true);
            int pc = bn > X.s_last ? X.M : sample_search(X, bn, false);
            if (pc <= po) continue;  // no sample inside
            if (po == 0) po = kAtStart;
            if (pc >= X.M) pc = kNoEvent;
            if (np == 0) { ev[0] = po; ev[1] = pc; lo = po; }
            else if (np == 1) { ev[2] = po; ev[3] = pc; }
            else { ev[4] = po; ev[5] = pc; }
            hi = pc;
            ++np;
        }
    }
    return make_ulonglong2(enc_pos(ev[0]) | (enc_pos(ev[1]) << 21) | (enc_pos(ev[2]) << 42),
                           enc_pos(ev[3]) | (enc_pos(ev[4]) << 21) | (enc_pos(ev[5]) << 42));
}

// prep_strip_kernel: warp per job, after the strip solve (which waits for the stream ODE): fixed-point
// weights of the strip elements; a strip without light switches its component off.
__global__ void __launch_bounds__(128) prep_strip_kernel(const __grid_constant__ FluxArgs A)
{
    const GridCfg& G = A.G;
    const int lane = threadIdx.x & 31;
    const long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (job >= A.njobs) return;
    const long long w = walker_of(job, A.L.n_ecl);
    if (!job_live(A, A.ws[w], A.js[job])) return;
    const bool do_bs = !(A.flags & LFB_FLAG_SKIP_BS);
    long long* wq_bs = A.wq + job * (G.n_wd_rings + G.n_disc_r + G.n_bs) + G.n_wd_rings + G.n_disc_r;
    const double* bsb = A.bs_b + job * G.n_bs;
    double p = 0.0;
    if (do_bs) for (int t = lane; t < G.n_bs; t += 32) p += bsb[t];
    const double tot_s = warp_sum(p);
    for (int t = lane; t < G.n_bs; t += 32) wq_bs[t] = (do_bs && tot_s > 0.0) ? llrint(bsb[t] / tot_s * kFix) : 0;
    if (lane == 0 && !(tot_s > 0.0)) A.jc[job].f_s = 0.0;
}

// positions_kernel: one thread per solved element of a job: where on the job's sorted sample axis its
// eclipse intervals open and close.
// PART 0: white dwarf and disc (nothing here needs the stream ODE); PART 1: the bright-spot strip.
template <int PART>
__global__ void __launch_bounds__(128) positions_kernel(const __grid_constant__ FluxArgs A)
{
    const GridCfg& G = A.G;
    const int per_job = PART == 0 ? G.n_wd_half + G.n_disc_half : G.n_bs;
    const int padded = (per_job + 31) & ~31;  // a warp never straddles two jobs
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long job = gid < 0xffffffffLL ? (long long)((unsigned)gid / (unsigned)padded) : gid / padded;
    const int t = (int)(gid - job * padded);
    if (job >= A.njobs) return;  // warp-uniform
    const long long w = walker_of(job, A.L.n_ecl);
    const int egather = (int)(job - w * A.L.n_ecl);
    const int e = A.mode ? 0 : egather;
    const WalkerScal& W = A.ws[w];
    if (!job_live(A, W, A.js[job])) return;  // warp-uniform
    int lo = kNoEvent, hi = -2;  // span of this thread's eclipse events on the sample axis
    if (t < per_job) {
        const SampleAxis X = sample_axis(A.smp, e, G.n_quad);
        const double phi0w = A.jc[job].phi0w;
        EventRec* ivp = A.ivp + job * A.ni_total;
        const EventRec none = no_events();
        double2 io;
        int i0;
        bool on, mirror = PART == 0;
        if (PART == 1) {
            on = !(A.flags & LFB_FLAG_SKIP_BS);
            io = on ? A.bs_io[job * G.n_bs + t] : make_double2(kBig, -kBig);
            i0 = G.n_wd + G.n_disc + t;
        } else if (t < G.n_wd_half) {
            on = !(A.flags & LFB_FLAG_SKIP_WD);
            io = on ? A.wd_io[w * G.n_wd_half + t] : make_double2(kBig, -kBig);
            i0 = 2 * t;
        } else {
            int h = t - G.n_wd_half;
            on = !(A.flags & LFB_FLAG_SKIP_DISC);
            io = on ? A.disc_io[job * G.n_disc_half + h] : make_double2(kBig, -kBig);
            i0 = G.n_wd + 2 * h;
        }
        const bool ecl = io.y > io.x;
        const EventRec r0 = ecl ? interval_pieces(X, io.x + phi0w, io.y + phi0w, lo, hi) : none;
        ivp[__ldg(G.rec_slot + i0)] = r0;
        // the y -> -y image is eclipsed from -egress to -ingress
        if (mirror) {
            int lo1 = kNoEvent, hi1 = -2;
            const EventRec r1 = ecl ? interval_pieces(X, -io.y + phi0w, -io.x + phi0w, lo1, hi1) : none;
            ivp[__ldg(G.rec_slot + i0 + 1)] = r1;
            lo = min(lo, lo1);
            hi = max(hi, hi1);
        }
    }
    // span of the job's eclipse events (lets segments far from the eclipse skip the tile records)
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && lo != kNoEvent) {
        atomicMin(&A.js[job].ev_lo, lo);
        atomicMax(&A.js[job].ev_hi, hi);
    }
}

// donor_table_kernel: one CTA per walker.  Tile image im of quarter tile t shows the observer the projected
// area W mu (1 - u + u mu), mu = A c + B s + D (c, s: cos, sin of the orbital angle), while mu > 0, i.e. on the
// phase interval |phase - centre| < half width.  The donor curve is therefore
//     yrs(phase) = m0 + m1 c + m2 s + m3 c^2 + m4 c s,   (m0..m4) = sum over the facing images,
// with moments that change only at the 2 x 4 x n_donor_q interval ends.  The kernel sorts the ends in phase
// (counting sort into kDonorBins uniform bins, then each short bin by insertion; ties by image, so the order is
// unique and the double-precision running sums below are reproducible) and stores the moments after every end,
// normalised "at maximum light" (quadrature, phase 0.25).  The table depends on (q, inclination) only: every
// eclipse of the walker reads it, through its own phase offset.
constexpr int kDonorThreads = 128;
constexpr int kDonorBins = 1024;

__host__ __device__ __forceinline__ int donor_bin(double phase)
{
    const double f = (phase + 0.5) * (double)kDonorBins;
    return f <= 0.0 ? 0 : (f >= (double)(kDonorBins - 1) ? kDonorBins - 1 : (int)f);
}

// moments (1, c, s, c^2, c s) of image im (bit 0: sign of B, bit 1: sign of D) from the eight parts of its
// quarter tile, m[part * stride]
__device__ __forceinline__ void donor_image_moments(const double* m, int stride, int im, double sgn, double* acc)
{
    const double sb = (im & 1) ? sgn : -sgn, sd = (im & 2) ? -1.0 : 1.0;
    acc[0] += sgn * (m[0] + sd * m[stride]);
    acc[1] += sgn * (m[2 * stride] + sd * m[3 * stride]);
    acc[2] += sb * (m[4 * stride] + sd * m[5 * stride]);
    acc[3] += sgn * m[6 * stride];
    acc[4] += sb * m[7 * stride];
}

__global__ void __launch_bounds__(kDonorThreads, 8) donor_table_kernel(const __grid_constant__ FluxArgs A)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const GridCfg& G = A.G;
    const int NDQ = G.n_donor_q, NB = 8 * NDQ;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = kDonorThreads / 32;
    double* qm = (double*)dsm;                         // [8][NDQ] moment parts of the quarter tiles (part-major: no bank conflicts)
    double* ukey = qm + 8 * NDQ;                       // [NB] break points as found (+inf: none)
    double* skey = ukey + NB;                          // [NB] sorted
    int* cnt = (int*)(skey + NB);                      // [kDonorBins + 1]
    unsigned short* uid = (unsigned short*)(cnt + kDonorBins + 1);  // [NB] image << 1 | (0: opens, 1: closes)
    unsigned short* sid = uid + NB;                    // [NB] sorted
    __shared__ double red[6][NW];
    __shared__ double s_base[6];
    __shared__ int s_itot[NW];
    const long long w = blockIdx.x;
    const WalkerScal& W = A.ws[w];
    // (one thread decides for the CTA: every thread must take the same way through the barriers below)
    __shared__ int s_live;
    if (tid == 0) s_live = !(W.status != 0 || (A.what != LFB_LN_LIKE && !(W.lnprior > -INFINITY)));
    __syncthreads();
    if (!s_live) return;
    const double ud = G.donor_ulimb;
    for (int q = tid; q < kDonorBins + 1; q += kDonorThreads) cnt[q] = 0;
    // ---- 1. intervals of the four images of every quarter tile ----
    double base[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // moments of the images facing at phase -0.5; [5]: normalisation
    for (int t = tid; t < NDQ; t += kDonorThreads) {
        const double4 q = A.don[w * NDQ + t];
        const double Aq = W.si * q.x, Bq = -W.si * q.y, Dq = W.ci * q.z, Bp = W.si * q.y;
        const double rho = sqrt(Aq * Aq + Bq * Bq);
        const double psi = atan2(Bq, Aq) * (1.0 / kTwoPi);
        const double ratio = rho > 0.0 ? Dq / rho : (Dq > 0.0 ? 2.0 : -2.0);
        // image with +D faces the observer iff cos(th - psi) > -D/rho
        const double hp = ratio >= 1.0 ? 0.5 : (ratio <= -1.0 ? -1.0 : acos(-ratio) * (1.0 / kTwoPi));
        const double hm = ratio <= -1.0 ? 0.5 : (ratio >= 1.0 ? -1.0 : acos(ratio) * (1.0 / kTwoPi));
        double* m = qm + t;
        m[0] = q.w * ud * (Dq * Dq + Bp * Bp);
        m[NDQ] = q.w * (1.0 - ud) * Dq;
        m[2 * NDQ] = q.w * (1.0 - ud) * Aq;
        m[3 * NDQ] = q.w * 2.0 * ud * Aq * Dq;
        m[4 * NDQ] = q.w * (1.0 - ud) * Bp;
        m[5 * NDQ] = q.w * 2.0 * ud * Bp * Dq;
        m[6 * NDQ] = q.w * ud * (Aq * Aq - Bp * Bp);
        m[7 * NDQ] = q.w * 2.0 * ud * Aq * Bp;
        // normalisation: the curve at quadrature (phase 0.25: c = 0, s = 1)
        {
            const double b = W.si * q.y, d = W.ci * q.z;
            double mu;
            mu = -b + d; if (mu > 0.0) base[5] += q.w * mu * (1.0 - ud + ud * mu);
            mu = b + d;  if (mu > 0.0) base[5] += q.w * mu * (1.0 - ud + ud * mu);
            mu = -b - d; if (mu > 0.0) base[5] += q.w * mu * (1.0 - ud + ud * mu);
            mu = b - d;  if (mu > 0.0) base[5] += q.w * mu * (1.0 - ud + ud * mu);
        }
#pragma unroll
        for (int im = 0; im < 4; ++im) {
            const double cen = (im & 1) ? -psi : psi, hw = (im & 2) ? hm : hp;
            double ko = INFINITY, kc = INFINITY;
            if (hw >= 0.5) {
                donor_image_moments(m, NDQ, im, 1.0, base);  // always faces the observer
            } else if (hw > 0.0) {
                double o = cen - hw, c2 = cen + hw;
                o -= rint(o);
                c2 -= rint(c2);
                if (o >= 0.5) o -= 1.0;
                if (c2 >= 0.5) c2 -= 1.0;
                if (o > c2) donor_image_moments(m, NDQ, im, 1.0, base);  // the interval holds phase -0.5: facing at the start
                if (o != c2) {
                    ko = nextafter(o, 1.0);  // opens just after o: applies to phase x iff o < x
                    kc = c2;                 // closed from c2 on:  applies iff c2 <= x
                }
            }
            const int slot = 2 * (4 * t + im);
            ukey[slot] = ko;
            ukey[slot + 1] = kc;
            uid[slot] = (unsigned short)((4 * t + im) << 1);
            uid[slot + 1] = (unsigned short)(((4 * t + im) << 1) | 1);
        }
    }
    // base moments and the normalisation: fixed-order sums over the CTA
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const double v = warp_sum(base[a]);
        if (lane == 0) red[a][wid] = v;
    }
    __syncthreads();
    if (tid < 6) {
        double v = 0.0;
        for (int i = 0; i < NW; ++i) v += red[tid][i];
        s_base[tid] = v;
    }
    // ---- 2. counting sort by phase bin ----
    for (int i = tid; i < NB; i += kDonorThreads) {
        const double k = ukey[i];
        if (k < 1e300) atomicAdd(&cnt[donor_bin(k) + 1], 1);
    }
    __syncthreads();
    {
        // exclusive scan of the bin counts: cnt[g + 1] <- break points in bins 0..g; every thread owns a run of bins
        constexpr int RB = kDonorBins / kDonorThreads;
        int c[RB], tot = 0;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            c[r] = cnt[1 + tid * RB + r];
            tot += c[r];
        }
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_itot[wid] = inc;
        __syncthreads();
        int run = inc - tot;
        for (int i = 0; i < wid; ++i) run += s_itot[i];
        unsigned short* first = A.dt.first + w * (kDonorBins + 1);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            first[tid * RB + r] = (unsigned short)run;  // break points in the bins before this one
            cnt[1 + tid * RB + r] = run;                // fill cursor of the bin
            run += c[r];
        }
        if (tid == kDonorThreads - 1) {
            first[kDonorBins] = (unsigned short)run;
            cnt[0] = run;  // the number of break points
        }
    }
    __syncthreads();
    const int nb = cnt[0];
    __syncthreads();
    for (int i = tid; i < NB; i += kDonorThreads) {
        const double k = ukey[i];
        if (k < 1e300) {
            const int at = atomicAdd(&cnt[donor_bin(k) + 1], 1);
            skey[at] = k;
            sid[at] = uid[i];
        }
    }
    __syncthreads();
    // afterwards cnt[g + 1] = end of bin g; its start is the end of bin g - 1 (0 for bin 0).  Order inside a bin
    // (a handful of break points at most): every break point counts those of its bin that come before it (ties
    // by image, so the order is unique) and takes that place -- one thread per break point, back into the
    // arrays the unsorted ones came from.
    for (int i = tid; i < nb; i += kDonorThreads) {
        const double k = skey[i];
        const unsigned short id = sid[i];
        const int g = donor_bin(k);
        const int b0 = g == 0 ? 0 : cnt[g], b1 = cnt[g + 1];
        int before = 0;
        for (int j = b0; j < b1; ++j) {
            const double kj = skey[j];
            before += (kj < k || (kj == k && sid[j] < id)) ? 1 : 0;
        }
        ukey[b0 + before] = k;
        uid[b0 + before] = id;
    }
    __syncthreads();
    skey = ukey;  // (sorted now)
    sid = uid;
    // ---- 3. moments after every break point: block scan in sorted order ----
    const double inv_norm = 1.0 / s_base[5];
    double* mom_out = A.dt.mom + w * (A.dt.nb_max + 1) * 6;
    // (row k: the moments after k break points, and in its sixth slot the break point that ends it)
    if (tid < 6) mom_out[tid] = tid < 5 ? s_base[tid] * inv_norm : (nb > 0 ? skey[0] : INFINITY);
    double carry[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) carry[a] = s_base[a];
    constexpr int EP = 8;  // break points per thread and round
    for (int r0 = 0; r0 < nb; r0 += EP * kDonorThreads) {
        // every thread owns EP consecutive break points: their sum first, the running sums in a second sweep
        double tot[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int q = 0; q < EP; ++q) {
            const int x = r0 + tid * EP + q;
            if (x < nb) {
                const int id = sid[x];
                donor_image_moments(qm + (id >> 3), NDQ, (id >> 1) & 3, (id & 1) ? -1.0 : 1.0, tot);
            }
        }
        double run[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            double v = tot[a];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            run[a] = v - tot[a];  // sum of the lanes before this one
            if (lane == 31) red[a][wid] = v;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            double before = carry[a], all = carry[a];
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                if (i < wid) before += red[a][i];
                all += red[a][i];
            }
            run[a] += before;
            carry[a] = all;
        }
#pragma unroll 1
        for (int q = 0; q < EP; ++q) {
            const int x = r0 + tid * EP + q;
            if (x < nb) {
                const int id = sid[x];
                donor_image_moments(qm + (id >> 3), NDQ, (id >> 1) & 3, (id & 1) ? -1.0 : 1.0, run);
                double2* row = (double2*)(mom_out + (size_t)(x + 1) * 6);
                row[0] = make_double2(run[0] * inv_norm, run[1] * inv_norm);
                row[1] = make_double2(run[2] * inv_norm, run[3] * inv_norm);
                row[2] = make_double2(run[4] * inv_norm, x + 1 < nb ? skey[x + 1] : INFINITY);
            }
        }
        __syncthreads();
    }
}

// flux_kernel: one CTA per job -- stages (2)-(4) of the model, one pass over the job's sorted
// exposure-sample axis in segments of at most Ms = NT * RP samples.
//
// Every eclipse interval of a tile is a few events (+w where it opens, -w where it closes) on the axis and the
// component curves are running sums of those events.  Weights are 2^-46 fixed point, split in a high and a 16-bit
// low limb, so that the events add into 32-bit shared-memory cells with native atomics and the sums do not depend
// on the order the events arrive in.  Per segment:
//   1. the cells are cleared; every tile record adds its events (chi-squared mode: white dwarf and disc share one
//      sum, their weights carry the component fluxes; flux-curve mode: three sums);
//   2. a block scan turns the cells into running sums: every thread owns RP consecutive samples, reads their
//      phase, cos and sin (stored per segment in thread-major order: coalesced), looks the donor moments up in the
//      walker's phase table (walking along it: consecutive samples mostly share a table row), evaluates the
//      components and leaves the flux in the sample's cell;
//   3. exposure quadrature, residuals, and the warp-shuffle + shared-memory chi-squared reduction.
// MODE 0: chi-squared (16 B cell per sample); MODE 1: the four component curves (32 B cell).
template <int MODE, int NT, int RP, int CTAS>
__global__ void __launch_bounds__(NT, CTAS) flux_kernel(const __grid_constant__ FluxArgs A)
{
    extern __shared__ __align__(16) int cells[];  // [Ms][CW]: limbs (hi, lo) of every running sum, then the flux
    const GridCfg& G = A.G;
    constexpr int Ms = NT * RP;
    constexpr int NW = NT / 32;
    constexpr int NA = MODE ? 3 : 2;  // running sums of tile events
    constexpr int CW = MODE ? 8 : 4;  // 32-bit words of a cell
    constexpr int NF = MODE ? 4 : 1;  // flux values left in a cell
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ double red[NW];
    __shared__ long long wtot[NA][NW];
    __shared__ long long s_carry[NA], s_next[NA];
    __shared__ JobConst C;  // (in shared memory: a dozen doubles every thread reads but need not hold in registers)

    const long long job = blockIdx.x;
    const long long w = walker_of(job, A.L.n_ecl);
    const int egather = (int)(job - w * A.L.n_ecl);
    const int e = MODE ? 0 : egather;
    const long long ch0 = A.smp.chunk_off[e];
    const int n_seg = (int)(A.smp.chunk_off[e + 1] - ch0);
    const long long lc0 = A.smp.lc_off[e];
    const int n_ph = (int)(A.smp.lc_off[e + 1] - lc0);
    const int K = G.n_quad;
    const WalkerScal& W = A.ws[w];
    const JobScal& J = A.js[job];
    if (!job_live(A, W, J)) {
        const bool skipped = J.status == 4 || J.status == 0;
        if (MODE == 0) {
            if (tid == 0) A.chisq_job[job] = skipped ? NAN : INFINITY;
        } else {
            for (int j = tid; j < n_ph; j += NT) {
                A.flux_tot[job * n_ph + j] = NAN;
                if (A.flux_comp)
                    for (int cidx = 0; cidx < 4; ++cidx) A.flux_comp[((long long)cidx * A.njobs + job) * n_ph + j] = NAN;
            }
        }
        return;
    }
    if (tid < (int)(sizeof(JobConst) / sizeof(double))) ((double*)&C)[tid] = ((const double*)(A.jc + job))[tid];
    __syncthreads();
    const long long* wq_tab = A.wq + job * (G.n_wd_rings + G.n_disc_r + G.n_bs);
    const EventRec* ivp = A.ivp + job * A.ni_total;
    const int n_tile_iv = A.ni_total;
    const bool do_don = C.f_rs != 0.0;
    // the walker's donor table
    const unsigned short* __restrict__ dfirst = A.dt.first + w * (kDonorBins + 1);
    const double* __restrict__ dmom = A.dt.mom + w * (A.dt.nb_max + 1) * 6;
    const int nb = do_don ? (int)__ldg(dfirst + kDonorBins) : 0;
    double chi = 0.0;

    for (int seg = 0; seg < n_seg; ++seg) {
        const int4 ch = __ldg(A.smp.chunks + ch0 + seg);  // first point, one past last point, first sample, last sample
        const int j0 = ch.x, j1 = ch.y, m0 = ch.z, m1 = ch.w + 1, len = m1 - m0;
        const bool tiles_matter = J.ev_lo < m1 && J.ev_hi >= m0;
        // the next segment starts at sample m0n of this one (segments may overlap, they leave no gap:
        // 1 <= m0n <= len); its running sums start from those after sample m0n - 1
        const int m0n = seg + 1 < n_seg ? __ldg(&A.smp.chunks[ch0 + seg + 1].z) - m0 : -1;

        // ---- 1. events of the segment into the cells; sums already open at sample 0 ----
        if (tiles_matter) {
            int4* c4 = (int4*)cells;
            for (int q = tid; q < Ms * CW / 4; q += NT) c4[q] = make_int4(0, 0, 0, 0);
        }
        __syncthreads();
        if (tiles_matter || seg == 0) {
            long long base[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) base[a] = 0;
            for (int i0 = tid; i0 < n_tile_iv; i0 += 4 * NT) {
                // (records are stored shuffled -- rec_slot -- so that the lanes of a warp hold tiles eclipsed at
                // different samples and their shared-memory atomics rarely meet)
                EventRec recs[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * NT;
                    recs[u] = i < n_tile_iv ? ivp[i] : no_events();
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * NT;
                    const EventRec rec = recs[u];
                    const int first = dec_pos(rec.x, 0);
                    if (first == kNoEvent || first >= m1) continue;
                    const int widx = __ldg(G.rec_widx + i);
                    const long long wq = __ldg(wq_tab + widx);  // the job's weight table: WD rings, disc rings, strip
                    const int arr = MODE ? (widx >= G.n_wd_rings) + (widx >= G.n_wd_rings + G.n_disc_r)
                                         : (widx >= G.n_wd_rings + G.n_disc_r);
                    const long long nq = -wq;
                    const int hi_p = (int)(wq >> kLimbBits), lo_p = (int)(wq & kLimbMask);
                    const int hi_n = (int)(nq >> kLimbBits), lo_n = (int)(nq & kLimbMask);
                    int* cell0 = cells + 2 * arr - m0 * CW;
                    // (nearly every record is one piece: it opens once and closes once)
                    const int nfld = dec_pos(rec.x, 2) == kNoEvent ? 2 : 6;
                    for (int k = 0; k < nfld; ++k) {
                        const int p = dec_pos(k < 3 ? rec.x : rec.y, k % 3);
                        if (p == kAtStart) {
                            if (seg == 0) base[arr] += wq;
                        } else if (p >= m0 && p < m1) {
                            int* cell = cell0 + p * CW;
                            atomicAdd(cell, (k & 1) ? hi_n : hi_p);
                            atomicAdd(cell + 1, (k & 1) ? lo_n : lo_p);
                        }
                    }
                }
            }
            if (seg == 0) {
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    long long v = base[a];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) wtot[a][wid] = v;
                }
            }
        }
        __syncthreads();
        if (seg == 0 && tid < NA) {
            long long v = 0;
            for (int i = 0; i < NW; ++i) v += wtot[tid][i];
            s_carry[tid] = v;  // running sums just before sample 0
        }
        __syncthreads();
        // ---- 2. cells -> running sums (block scan); every thread evaluates its RP consecutive samples ----
        long long run[NA];
        if (tiles_matter) {
            long long tot[NA], inc[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) tot[a] = 0;
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                const int* cell = cells + (tid * RP + r) * CW;
                if (MODE == 0) {
                    const int4 v = *(const int4*)cell;
                    tot[0] += ((long long)v.x << kLimbBits) + v.y;
                    tot[1] += ((long long)v.z << kLimbBits) + v.w;
                } else {
                    const int4 v = *(const int4*)cell;
                    const int2 u = *(const int2*)(cell + 4);
                    tot[0] += ((long long)v.x << kLimbBits) + v.y;
                    tot[1] += ((long long)v.z << kLimbBits) + v.w;
                    tot[2] += ((long long)u.x << kLimbBits) + u.y;
                }
            }
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                long long v = tot[a];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long u = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += u;
                }
                inc[a] = v;
                if (lane == 31) wtot[a][wid] = v;
            }
            __syncthreads();
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                long long before = s_carry[a] + inc[a] - tot[a];
#pragma unroll
                for (int i = 0; i < NW - 1; ++i) before += i < wid ? wtot[a][i] : 0;
                run[a] = before;
            }
        } else {
#pragma unroll
            for (int a = 0; a < NA; ++a) run[a] = s_carry[a];
            if (tid < NA) s_next[tid] = s_carry[tid];
        }
        {
            // phase, cos, sin of the segment's samples, thread-major: sample (tid, r) at [r][tid]
            const double* __restrict__ sp = A.smp.seg_tr + (size_t)(ch0 + seg) * 3 * Ms + tid;
            // Donor table state: dm = moments of row kd (the break points at or before the last sample's phase),
            // kend = the break point that ends the row.  (Fetching the next row ahead was measured: the extra
            // instructions cost more than the latency they hide -- other warps cover it.)
            int kd = -1;
            double kend = -INFINITY;
            double xprev = 2.0;
            double dm[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            double x0 = __ldg(sp), c0 = __ldg(sp + Ms), s0 = __ldg(sp + 2 * Ms);
            int* cell = cells + tid * RP * CW;
            int p = tid * RP;
            const int p_next = m0n - 1;
            // per-job constants of the mix (registers for the loop; the rest stays in shared memory)
            const double f01c = C.f_wd + C.f_d, f01k = C.F01 * kInvFix;
            const double b0 = C.f_s * C.fis, b1 = C.f_s * (1.0 - C.fis);
#pragma unroll 1
            for (int r = 0; r < RP; ++r, ++p, cell += CW) {
                // the next sample's phase, cos and sin travel while this one is evaluated
                const double* __restrict__ spn = r + 1 < RP ? sp + NT : sp;
                const double x0n = __ldg(spn), c0n = __ldg(spn + Ms), s0n = __ldg(spn + 2 * Ms);
                sp = spn;
                if (tiles_matter) {
                    const int4 v = *(const int4*)cell;
                    run[0] += ((long long)v.x << kLimbBits) + v.y;
                    run[1] += ((long long)v.z << kLimbBits) + v.w;
                    if (MODE) {
                        const int2 u = *(const int2*)(cell + 4);
                        run[2] += ((long long)u.x << kLimbBits) + u.y;
                    }
                    if (p == p_next) {
#pragma unroll
                        for (int a = 0; a < NA; ++a) s_next[a] = run[a];
                    }
                }
                if (p < len) {
                    const double cphi = C.cphi, sphi = C.sphi;
                    const double cc = c0 * cphi + s0 * sphi, ss = s0 * cphi - c0 * sphi;
                    double f3 = 0.0;
                    if (do_don) {
                        // the donor's moments at this sample: table row = number of break points at or before its phase
                        double x = x0 - C.phi0w;
                        x -= rint(x);
                        if (x >= 0.5) x -= 1.0;
                        if (x >= kend || x < xprev) {
                            // the row ended (or the phase wrapped, or this is the first sample): the next row,
                            // or -- after a jump in phase -- a look-up through the bin table, then a short walk
                            int k2 = kd + 1;
                            if (kd < 0 || x < xprev || x - xprev > 2.0 / kDonorBins) k2 = (int)__ldg(dfirst + donor_bin(x));
                            const double2* row = (const double2*)(dmom + (size_t)k2 * 6);
                            double2 a01 = __ldg(row), a23 = __ldg(row + 1), a45 = __ldg(row + 2);
                            while (a45.y <= x) {  // (the last row ends at +inf)
                                ++k2;
                                row += 3;
                                a01 = __ldg(row);
                                a23 = __ldg(row + 1);
                                a45 = __ldg(row + 2);
                            }
                            dm[0] = a01.x; dm[1] = a01.y; dm[2] = a23.x; dm[3] = a23.y; dm[4] = a45.x;
                            kend = a45.y;
                            kd = k2;
                        }
                        xprev = x;
                        f3 = C.f_rs * (dm[0] + dm[1] * cc + dm[2] * ss + dm[3] * (cc * cc) + dm[4] * (cc * ss));
                    }
                    const double bm = C.beam_a * cc + C.beam_b * ss + C.beam_d;
                    const double beam = b0 + b1 * (bm > 0.0 ? bm : 0.0);  // f_s (fis + (1 - fis) max(bm, 0))
                    if (MODE == 0) {
                        const double f01 = f01c - f01k * (double)run[0];
                        const double f2 = beam - (beam * kInvFix) * (double)run[1];
                        *(double*)cell = f01 + f2 + f3;
                    } else {
                        double* out = (double*)cell;
                        out[0] = C.f_wd * (1.0 - (double)run[0] * kInvFix);
                        out[1] = C.f_d * (1.0 - (double)run[1] * kInvFix);
                        out[2] = beam * (1.0 - (double)run[2] * kInvFix);
                        out[3] = f3;
                    }
                }
                x0 = x0n;
                c0 = c0n;
                s0 = s0n;
            }
        }
        __syncthreads();
        if (tid < NA && seg + 1 < n_seg) s_carry[tid] = s_next[tid];
        // ---- 3. exposure quadrature (Simpson over phase +- width), residuals / output ----
        for (int j = j0 + tid; j < j1; j += NT) {
            double acc[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] = 0.0;
            for (int k = 0; k < K; ++k) {
                const int q = __ldg(A.smp.pos + (lc0 + j) * K + k) - m0;
                const double qw = G.quad_w[k];
                const double* F = (const double*)(cells + q * CW);
#pragma unroll
                for (int f = 0; f < NF; ++f) acc[f] += qw * F[f];
            }
            if (MODE == 0) {
                const double dy = __ldg(A.smp.y + lc0 + j) - acc[0];
                const double r = dy * __ldg(A.smp.iye + lc0 + j);
                chi += r * r;
                if (A.gp_resid) A.gp_resid[(lc0 + __ldg(A.smp.gp_slot + lc0 + j)) * A.n_walkers + w] = dy;
            } else {
                const int jo = __ldg(A.smp.pt_index + lc0 + j);
                A.flux_tot[job * n_ph + jo] = acc[0] + acc[1 % NF] + acc[2 % NF] + acc[3 % NF];
                if (A.flux_comp)
                    for (int cidx = 0; cidx < 4; ++cidx)
                        A.flux_comp[((long long)cidx * A.njobs + job) * n_ph + jo] = acc[cidx % NF];
            }
        }
        __syncthreads();
    }
    if (MODE == 0) {
        chi = block_sum<NT>(chi, red);
        if (tid == 0) A.chisq_job[job] = chi;
    }
}

// ---------------------------------------------------------------- gp_kernel
// SimpleGPEclipse.ln_like (CVModel.py:650-696).  The filter is a serial recursion over the eclipse's
// points, so what matters for a thin batch is the latency of one job: two neighbouring lanes share a
// (walker, eclipse) -- one filters the first half of the points forwards, the other the second half
// backwards, and they meet in the middle (gp_merge) -- and the block (kGpWalkers walkers of one
// eclipse) stages tiles of kGpTile points per direction in shared memory: coalesced rows of the
// residual matrix, the shared times and variances.  Leaves -2 ln L where the chi-squared would be.
constexpr int kGpWalkers = 32;
constexpr int kGpThreads = 2 * kGpWalkers;
constexpr int kGpTile = 32;

__global__ void __launch_bounds__(kGpThreads) gp_kernel(const __grid_constant__ FluxArgs A)
{
    __shared__ double r_tile[2][kGpTile][kGpWalkers];
    __shared__ double x_tile[2][kGpTile], v_tile[2][kGpTile];
    const int e = blockIdx.y, tid = threadIdx.x;
    const int wl = tid >> 1, dir = tid & 1;  // dir 0: forwards over points [0, m); 1: backwards over [m, n)
    const long long w0 = (long long)blockIdx.x * kGpWalkers, w = w0 + wl;
    const long long lc0 = A.smp.lc_off[e];
    const int n_ph = (int)(A.smp.lc_off[e + 1] - lc0);
    const int m = n_ph < 4 ? n_ph : n_ph / 2;  // a handful of points: one filter does it all
    const int n_mine = dir ? n_ph - m : m, n_max = max(m, n_ph - m);
    const long long job = w * A.L.n_ecl + e;
    const bool live = w < A.n_walkers && job_live(A, A.ws[w], A.js[job]);  // else the flux kernel left NaN / +inf
    GpPars G;
    G.a_in = G.a_out = G.tau = 1.0;
    G.n_gaps = 0;
    bool run = false;
    if (live) {
        const double* th = A.theta + w * A.L.ndim;
        G.a_in = exp(fetch(A.L, th, A.gp_src[0]));
        G.a_out = exp(fetch(A.L, th, A.gp_src[1]));
        G.tau = exp(fetch(A.L, th, A.gp_src[2]));
        const double phi0 = fetch(A.L, th, A.L.gather[e * LFB_NPAR + P_PHI0]);
        const double2 span = A.smp.gp_span[e];
        const double dist = A.gp_dist[e];
        run = dist > 0.0;
        if (run) gp_changepoints(span.x, span.y, dist, phi0, G);
    }
    GpFilter F;
    F.init(G, dir ? -1 : 1);
    const int ncol = (int)min((long long)kGpWalkers, A.n_walkers - w0);
    for (int k0 = 0; k0 < n_max; k0 += kGpTile) {
        __syncthreads();
        for (int i = tid; i < 2 * kGpTile * kGpWalkers; i += kGpThreads) {
            const int d = i / (kGpTile * kGpWalkers), rem = i - d * (kGpTile * kGpWalkers);
            const int row = rem / kGpWalkers, col = rem - row * kGpWalkers;
            const int k = k0 + row;
            if (col < ncol && k < (d ? n_ph - m : m))
                r_tile[d][row][col] = A.gp_resid[(lc0 + (d ? n_ph - 1 - k : k)) * A.n_walkers + w0 + col];
        }
        {
            const int d = tid / kGpTile, row = tid - d * kGpTile, k = k0 + row;  // kGpThreads == 2 * kGpTile
            if (k < (d ? n_ph - m : m)) {
                const long long p = lc0 + (d ? n_ph - 1 - k : k);
                x_tile[d][row] = __ldg(A.smp.gp_x + p);
                v_tile[d][row] = __ldg(A.smp.gp_var + p);
            }
        }
        __syncthreads();
        if (run) {
            const int rows = min(kGpTile, n_mine - k0);
            for (int r = 0; r < rows; ++r) F.step(G, x_tile[dir][r], v_tile[dir][r], r_tile[dir][r][wl]);
        }
    }
    // the backward lane hands its half to the forward lane
    const double ll_mine = F.result();
    GpFilter B = F;
    const unsigned full = 0xffffffffu;
    B.m0 = __shfl_xor_sync(full, F.m0, 1); B.m1 = __shfl_xor_sync(full, F.m1, 1);
    B.m2 = __shfl_xor_sync(full, F.m2, 1); B.m3 = __shfl_xor_sync(full, F.m3, 1);
    B.p00 = __shfl_xor_sync(full, F.p00, 1); B.p01 = __shfl_xor_sync(full, F.p01, 1); B.p11 = __shfl_xor_sync(full, F.p11, 1);
    B.p02 = __shfl_xor_sync(full, F.p02, 1); B.p03 = __shfl_xor_sync(full, F.p03, 1);
    B.p12 = __shfl_xor_sync(full, F.p12, 1); B.p13 = __shfl_xor_sync(full, F.p13, 1);
    B.p22 = __shfl_xor_sync(full, F.p22, 1); B.p23 = __shfl_xor_sync(full, F.p23, 1); B.p33 = __shfl_xor_sync(full, F.p33, 1);
    B.last_gap = __shfl_xor_sync(full, F.last_gap, 1);
    B.bad = __shfl_xor_sync(full, (int)F.bad, 1) != 0;
    const double ll_other = __shfl_xor_sync(full, ll_mine, 1);
    if (live && dir == 0) {
        double ll = INFINITY;  // -> chi-squared + inf when there are no change points
        if (run) {
            ll = ll_mine;
            if (m < n_ph) {
                F.advance(G, __ldg(A.smp.gp_x + lc0 + m));
                ll += ll_other + gp_merge(F, B, G, F.last_gap >= 0 && F.last_gap == B.last_gap);
            }
        }
        A.chisq_job[job] = run ? -2.0 * ll : INFINITY;
    }
}

// lfb_gp_loglike: the same likelihood for caller-supplied residuals, thread per set
__global__ void gp_batch_kernel(long long n_sets, int n, const double* __restrict__ x, const double* __restrict__ ye,
                                const double* __restrict__ resid, const double* __restrict__ hyper, int n_gaps,
                                const double* __restrict__ gaps, double* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sets) return;
    GpPars G;
    G.a_in = hyper[3 * i];
    G.a_out = hyper[3 * i + 1];
    G.tau = hyper[3 * i + 2];
    G.n_gaps = n_gaps;
    for (int k = 0; k < n_gaps; ++k) {
        G.gap[k][0] = gaps[(i * n_gaps + k) * 2];
        G.gap[k][1] = gaps[(i * n_gaps + k) * 2 + 1];
    }
    const double* r = resid + i * n;
    out[i] = gp_loglike(
        n, [&](int k) { return x[k]; }, [&](int k) { return ye[k] * ye[k]; }, [&](int k) { return r[k]; }, G);
}

// lfb_wdphases: trm.roche.wdphases(q, iangle, r1, ntheta) (call site CVModel.py:562): earliest and
// latest egress over ntheta points on the limb of the white dwarf's disc on the sky
__global__ void wdphases_kernel(long long n, const double* __restrict__ q, const double* __restrict__ incl,
                                const double* __restrict__ r1, int ntheta, double* __restrict__ out, int* __restrict__ ok)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Roche R;
    int good = 0;
    double p3 = 1e30, p4 = -1e30;
    if (roche_init(q[i], R) && incl[i] > 0.0 && incl[i] <= 90.0 && r1[i] > 0.0 && ntheta > 0) {
        double si, ci;
        sincos_(incl[i] * kDeg, &si, &ci);
        good = 1;
        for (int k = 0; k < ntheta; ++k) {
            double sa, ca, pin, pout;
            sincos_(kTwoPi * k / ntheta, &sa, &ca);
            Point T = {0.0, 0.0, 0.0, r1[i] * ca, r1[i] * sa};
            if (!ingress_egress(R, si, ci, T, &pin, &pout)) {
                good = 0;
                break;
            }
            p3 = pout < p3 ? pout : p3;
            p4 = pout > p4 ? pout : p4;
        }
    }
    out[2 * i] = good ? p3 : NAN;
    out[2 * i + 1] = good ? p4 : NAN;
    ok[i] = good;
}

// lfb_ingress_egress: the element solve of stage (1) on its own, thread per element
__global__ void ingress_egress_kernel(long long n, const double* __restrict__ q, const double* __restrict__ incl,
                                      const double* __restrict__ pts, double* __restrict__ out, int* __restrict__ ok)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Roche R;
    int good = 0;
    double pin = NAN, pout = NAN;
    if (roche_init(q[i], R) && incl[i] > 0.0 && incl[i] <= 90.0) {
        double si, ci;
        sincos_(incl[i] * kDeg, &si, &ci);
        const Point T = {pts[5 * i], pts[5 * i + 1], pts[5 * i + 2], pts[5 * i + 3], pts[5 * i + 4]};
        // a tile of the white dwarf (offset on the sky only) is solved as the pipeline solves it: from the
        // grazing lines of sight of the centre (wdcentre_kernel), when those exist
        Roots hint;
        hint.lam[0] = NAN;
        bool have_hint = false;
        if (T.x == 0.0 && T.y == 0.0 && T.z == 0.0 && (T.xi != 0.0 || T.eta != 0.0)) {
            const Point T0 = {0.0, 0.0, 0.0, 0.0, 0.0};
            double a, b;
            have_hint = ingress_egress(R, si, ci, T0, &a, &b, nullptr, &hint) && hint.lam[0] == hint.lam[0];
        }
        good = ingress_egress(R, si, ci, T, &pin, &pout, have_hint ? &hint : nullptr);
    }
    out[2 * i] = good ? pin : NAN;
    out[2 * i + 1] = good ? pout : NAN;
    ok[i] = good;
}

// ---------------------------------------------------------------- finish_kernel
// Node.ln_prob = ln_prior + sum of -chi^2/2 with the -inf rules (model.py:476-498)
__global__ void finish_kernel(int what, int n_ecl, long long n, const WalkerScal* __restrict__ ws,
                              const double* __restrict__ chisq_job, double* __restrict__ chisq, double* __restrict__ out)
{
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    double lnp = ws[w].lnprior;
    double like = 0.0;
    if (what != LFB_LN_PRIOR) {
        for (int e = 0; e < n_ecl; ++e) {
            double chi = chisq_job[w * n_ecl + e];
            // NaN marks "not evaluated" (prior veto); a NaN model is +inf (CVModel.py:163-171)
            const bool skipped = what == LFB_LN_PROB && !(lnp > -INFINITY);
            if (isnan(chi) && !skipped) chi = INFINITY;
            if (chisq) chisq[w * n_ecl + e] = skipped ? NAN : chi;
            like += -0.5 * chi;
        }
    }
    if (!out) return;
    double v;
    if (what == LFB_LN_PRIOR) v = lnp;
    else if (what == LFB_LN_LIKE) v = like;
    else v = lnp > -INFINITY ? lnp + like : -INFINITY;
    if (isnan(v)) v = -INFINITY;  // never NaN towards the sampler (model.py:489-493)
    out[w] = v;
}

// ---------------------------------------------------------------- roche_kernel
__global__ void roche_kernel(int which, long long n, const double* __restrict__ a, const double* __restrict__ b,
                             double* __restrict__ out, int* __restrict__ ok)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o[4] = {0.0, 0.0, 0.0, 0.0};
    int good = 0;
    Roche R;
    if (which == LFB_ROCHE_ANGLE) {
        // diagnostic: the solver's arctangent of a unit vector, o[0] = angle_of(a, b) for a = cos, b = sin
        o[0] = angle_of(a[i], b[i]);
        good = 1;
    } else if (roche_init(a[i], R)) {
        if (which == LFB_ROCHE_XL1) {
            o[0] = R.xl1;
            good = 1;
        } else if (which == LFB_ROCHE_FINDPHI) {
            double inc = b[i];
            if (inc == 90.0) {
                o[0] = findphi90(R);
                good = 1;
            } else if (inc > 0.0 && inc < 90.0) {
                Point T = {0.0, 0.0, 0.0, 0.0, 0.0};
                double si, ci, pin, pout;
                sincos_(inc * kDeg, &si, &ci);
                if (ingress_egress(R, si, ci, T, &pin, &pout)) {
                    o[0] = pout - pin;
                    good = 1;
                }
            }
        } else if (which == LFB_ROCHE_FINDI) {
            double sini;
            if (findi(R, b[i], findphi90(R), sini)) {
                o[0] = asin(sini) / kDeg;
                good = 1;
            }
        } else if (which == LFB_ROCHE_BSPOT) {
            good = bspot(R, b[i], o) ? 1 : 0;
        }
    }
    for (int k = 0; k < 4; ++k) out[i * 4 + k] = good ? o[k] : NAN;
    ok[i] = good;
}

// ---------------------------------------------------------------- fp64_peak_kernel
// DFMA throughput probe: the FP64 roofline denominator is measured on the device the
// numbers are taken on (MEASURED_PEAKS.json has no FP64 vector figure).
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double x, double y, double* __restrict__ sink)
{
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456) sink[0] = s;
}

}  // namespace lfb
