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
//   positions_kernel thread per element: the element's eclipse / facing interval as events on the
//                    job's sorted exposure-sample axis (16 B record), donor moment parts
//   flux_kernel      CTA per job, in segments of the axis: tile events -> shared-memory atomics ->
//                    block scan; donor events -> counting sort -> block scan of five trigonometric
//                    moments; then component mix, exposure quadrature and the chi-squared reduction
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
constexpr int kNumArr = 8;  // event arrays: white dwarf, disc, bright spot, 5 donor moments
constexpr int kNoEventPos = 0x7fffffff;
constexpr double kFix = 72057594037927936.0;         // 2^56: fixed-point scale of normalised weights
constexpr double kInvFix = 1.0 / 72057594037927936.0;

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
    const double *y, *ye;        // [total] in phase order
    const double *S, *cosS, *sinS;  // [K * total] sorted per eclipse
    const int* bins;             // [K * total + n_ecl] per eclipse M + 1 entries (see SampleAxis)
    const int* pos;              // [total * K] sorted position of each (point, node); points in phase order
    const int* pt_index;         // [total] original index of each phase-ordered point
    const double *gp_x, *gp_var; // [total] GP likelihood: raw phases ascending per eclipse, noise variances
    const int* gp_slot;          // [total] rank in that order of each phase-ordered point
    const double2* gp_span;      // [n_ecl] smallest and largest raw phase
    const long long* chunk_off;  // [n_ecl + 1] offsets into chunks
    const int4* chunks;          // per segment: first point, one past last point, first sample, last sample
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
    int status;                // 0 ok, 1 walker invalid, 2 stream misses disc, 3 bad parameter, 4 not needed
    int ev_lo, ev_hi;          // sample positions spanned by the eclipse events of the job's tiles
};

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
    long long w = job / L.n_ecl;
    int e = (int)(job - w * L.n_ecl);
    JobScal J;
    J.xs = J.ys = 0.0;
    J.smax = J.smaxp = J.shi = 1.0;
    J.status = 0;
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
                              WalkerScal* ws, JobScal* js)
{
    // (a group of lanes enters and leaves together)
    long long job = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> (LANES ? 3 : 0);
    if (job >= njobs || (flags & LFB_FLAG_SKIP_BS)) return;
    if (js[job].status != 0) return;
    long long w = job / L.n_ecl;
    int e = (int)(job - w * L.n_ecl);
    const Roche R = ws[w].R;
    const double* th = theta + w * L.ndim;
    const int* g = L.gather + e * LFB_NPAR;
    double rdisc_a = fetch(L, th, g[P_RDISC]) * R.xl1;
    double imp[4];
    const bool hit = LANES ? bspot_lanes(R, rdisc_a, imp) : bspot(R, rdisc_a, imp);
    if (LANES && (threadIdx.x & 7) != 0) return;  // one lane of the group writes
    if (!hit) {
        js[job].status = 2;  // the stream misses the disc (roche.bspot raises, CVModel.py:309-316)
        if (what != LFB_LN_LIKE) ws[w].lnprior = -INFINITY;
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
        if (!(az >= minaz) || !(az <= maxaz)) ws[w].lnprior = -INFINITY;
    }
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
    const long long unit = gid / padded;
    const int t = (int)(gid - unit * padded);
    const long long nunits = (COMP == 0 || COMP == 3) ? A.n : A.njobs;
    if (unit >= nunits || t >= per_unit) return;
    const long long w = (COMP == 0 || COMP == 3) ? unit : unit / A.L.n_ecl;
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
        int k = (int)sqrt(0.5 * (double)t);
        while (2 * k * k > t) --k;
        while (2 * (k + 1) * (k + 1) <= t) ++k;
        int r = t - 2 * k * k, q1 = 2 * k + 1, nk = 4 * q1;
        int j = r < q1 ? r : r + 2 * q1;
        double inv = 1.0 / G.n_wd_rings;
        double ra = k * inv, rb = (k + 1) * inv;
        double rho = sqrt(0.5 * (ra * ra + rb * rb));
        double sa, ca;
        sincos_((j + 0.5) * kTwoPi / nk, &sa, &ca);
        T.xi = rwd_a * rho * ca;
        T.eta = rwd_a * rho * sa;
    } else {
        const JobScal& J = A.js[unit];
        if (J.status != 0) return;
        if (COMP == 1) {
            // disc: ring m, sector j on the y > 0 side
            const double rwd_a = fetch(A.L, th, g[P_RWD]) * R.xl1, rdisc_a = fetch(A.L, th, g[P_RDISC]) * R.xl1;
            int hth = G.n_disc_th / 2;
            tile = G.disc_order[t];
            int m = tile / hth, j = tile - m * hth;
            double r = rwd_a + (m + 0.5) * (rdisc_a - rwd_a) / G.n_disc_r;
            double sa, ca;
            sincos_((j + 0.5) * kTwoPi / G.n_disc_th, &sa, &ca);
            T.x = r * ca;
            T.y = r * sa;
        } else {
            // bright spot: strip through the stream impact point along azimuth az
            const double exp1 = A.L.npars > P_EXP1 ? fetch(A.L, th, g[P_EXP1]) : 2.0;
            const double exp2 = A.L.npars > P_EXP2 ? fetch(A.L, th, g[P_EXP2]) : 1.0;
            double s = J.shi * t / (G.n_bs - 1);
            wt = t == 0 ? 0.0 : pow(s / J.smax, exp1) * exp(J.smaxp - pow(s, exp2));
            double len = (s - J.smax) * fetch(A.L, th, g[P_SCALE]) * R.xl1;
            double tx, ty;
            sincos_(fetch(A.L, th, g[P_AZ]) * kDeg, &ty, &tx);
            T.x = J.xs + len * tx;
            T.y = J.ys + len * ty;
        }
    }
    double pin, pout;
    if (!ingress_egress(R, si, ci, T, &pin, &pout, (COMP == 0 && W.wd_ok) ? &W.wd : nullptr)) { pin = kBig; pout = -kBig; }
    if (COMP == 0) A.wd_io[unit * G.n_wd_half + t] = make_double2(pin, pout);
    else if (COMP == 1) A.disc_io[unit * G.n_disc_half + tile] = make_double2(pin, pout);
    else {
        A.bs_io[unit * G.n_bs + t] = make_double2(pin, pout);
        A.bs_b[unit * G.n_bs + t] = wt;
    }
}

// ---------------------------------------------------------------- flux stage: prep / positions / flux
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
    double f_wd, f_d, f_s, f_rs;        // flux scale of each component's running sum
    double beam_a, beam_b, beam_d, fis; // bright-spot beaming: fis + (1-fis) max(0, a c + b s + d)
    double cphi, sphi, phi0w;           // phase offset wrapped to [-0.5, 0.5]
    double don_sc;                      // 2^56 / sum of donor tile weights
};

struct FluxArgs {
    DevLayout L;
    GridCfg G;
    DevSamples smp;
    int what, flags, mode;  // mode 0: chi-squared, 1: flux curves
    int Ms;                 // capacity of a segment of the sample axis in samples
    int ni_total;           // event records per job: n_wd + n_disc + n_bs + 4 n_donor_q
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
    long long* qmom;        // [njobs][n_donor_q][8] fixed-point donor moment parts of each quarter tile
    ulonglong2* ivp;        // [njobs][ni_total] event records (EventRec) of every eclipse / facing interval
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
// fixed-point (2^-56) element weights, beaming and phase-offset constants.
__global__ void __launch_bounds__(128) prep_kernel(const __grid_constant__ FluxArgs A)
{
    const GridCfg& G = A.G;
    const int lane = threadIdx.x & 31;
    const long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (job >= A.njobs) return;
    const long long w = job / A.L.n_ecl;
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
        wq_wd[k] = do_wd ? llrint(((1.0 - ulimb) + ulimb * mubar) / tot_wd * kFix) : 0;
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
        wq_disc[m] = do_disc ? llrint(pow(r, 1.0 - dexp) / tot_d * kFix) : 0;
    }
    // (the bright-spot strip's weights wait for the strip itself: prep_strip_kernel)
    // donor: normalised at quadrature (phase 0.25: c = 0, s = 1)
    const double4* don = A.don + w * G.n_donor_q;
    const double ud = G.donor_ulimb;
    double p_rs = 0.0, p_rw = 0.0;
    if (do_don)
        for (int t = lane; t < G.n_donor_q; t += 32) {
            double4 q = don[t];
            double b = si * q.y, d = ci * q.z, m;
            m = -b + d; if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
            m = b + d;  if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
            m = -b - d; if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
            m = b - d;  if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
            p_rw += 4.0 * q.w;
        }
    const double tot_rs = warp_sum(p_rs), tot_rw = warp_sum(p_rw);
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
        C.f_wd = do_wd ? par(P_WDFLUX, 0.0) : 0.0;
        C.f_d = do_disc ? par(P_DFLUX, 0.0) : 0.0;
        C.f_s = (do_bs && beam_norm > 0.0) ? par(P_SFLUX, 0.0) / beam_norm : 0.0;  // 0 if the strip is dark: prep_strip_kernel
        C.f_rs = do_don ? par(P_RSFLUX, 0.0) / tot_rs * (tot_rw * kInvFix) : 0.0;
        double phi0w = par(P_PHI0, 0.0);
        phi0w -= rint(phi0w);
        C.phi0w = phi0w;
        sincos_(kTwoPi * phi0w, &C.sphi, &C.cphi);
        C.don_sc = do_don ? kFix / tot_rw : 0.0;
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
    X.s_first = X.M > 0 ? __ldg(X.S) : 0.0;
    X.s_last = X.M > 0 ? __ldg(X.S + X.M - 1) : 0.0;
    X.inv_binw = X.s_last > X.s_first ? (double)X.M / (X.s_last - X.s_first) : 0.0;
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
__device__ __forceinline__ EventRec interval_pieces(const SampleAxis& X, double a, double b)
{
    int ev[6] = {kNoEvent, kNoEvent, kNoEvent, kNoEvent, kNoEvent, kNoEvent};
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
            if (np == 0) { ev[0] = po; ev[1] = pc; }
            else if (np == 1) { ev[2] = po; ev[3] = pc; }
            else { ev[4] = po; ev[5] = pc; }
            ++np;
        }
    }
    return make_ulonglong2(enc_pos(ev[0]) | (enc_pos(ev[1]) << 21) | (enc_pos(ev[2]) << 42),
                           enc_pos(ev[3]) | (enc_pos(ev[4]) << 21) | (enc_pos(ev[5]) << 42));
}

// position of the last closing event of a record (kNoEvent if it stays open to the end; -2 if empty)
__device__ __forceinline__ int rec_last_close(const EventRec& rec)
{
    if (dec_pos(rec.x, 0) == kNoEvent) return -2;
    const int o1 = dec_pos(rec.x, 2), o2 = dec_pos(rec.y, 1);
    return o2 != kNoEvent ? dec_pos(rec.y, 2) : (o1 != kNoEvent ? dec_pos(rec.y, 0) : dec_pos(rec.x, 1));
}

// prep_strip_kernel: warp per job, after the strip solve (which waits for the stream ODE): fixed-point
// weights of the strip elements; a strip without light switches its component off.
__global__ void __launch_bounds__(128) prep_strip_kernel(const __grid_constant__ FluxArgs A)
{
    const GridCfg& G = A.G;
    const int lane = threadIdx.x & 31;
    const long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (job >= A.njobs) return;
    const long long w = job / A.L.n_ecl;
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

// positions_kernel: one thread per solved element (or donor quarter tile) of a job: where on the
// job's sorted sample axis its eclipse (facing) intervals open and close.
// PART 0: white dwarf, disc and donor (nothing here needs the stream ODE); PART 1: the bright-spot strip.
template <int PART>
__global__ void __launch_bounds__(128) positions_kernel(const __grid_constant__ FluxArgs A)
{
    const GridCfg& G = A.G;
    const int per_job = PART == 0 ? G.n_wd_half + G.n_disc_half + G.n_donor_q : G.n_bs;
    const int padded = (per_job + 31) & ~31;  // a warp never straddles two jobs
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long job = gid / padded;
    const int t = (int)(gid - job * padded);
    if (job >= A.njobs) return;  // warp-uniform
    const long long w = job / A.L.n_ecl;
    const int egather = (int)(job - w * A.L.n_ecl);
    const int e = A.mode ? 0 : egather;
    const WalkerScal& W = A.ws[w];
    if (!job_live(A, W, A.js[job])) return;  // warp-uniform
    int lo = kNoEvent, hi = -2;  // span of this thread's eclipse events on the sample axis
    if (t < per_job) {
        const SampleAxis X = sample_axis(A.smp, e, G.n_quad);
        const double phi0w = A.jc[job].phi0w;
        EventRec* ivp = A.ivp + job * A.ni_total;
        const int n_half = G.n_wd_half + G.n_disc_half;
        const EventRec none = no_events();
        if (PART == 1 || t < n_half) {
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
            const EventRec r0 = ecl ? interval_pieces(X, io.x + phi0w, io.y + phi0w) : none;
            ivp[__ldg(G.rec_slot + i0)] = r0;
            lo = dec_pos(r0.x, 0);
            hi = rec_last_close(r0);
            // the y -> -y image is eclipsed from -egress to -ingress
            if (mirror) {
                const EventRec r1 = ecl ? interval_pieces(X, -io.y + phi0w, -io.x + phi0w) : none;
                ivp[__ldg(G.rec_slot + i0 + 1)] = r1;
                lo = min(lo, dec_pos(r1.x, 0));
                hi = max(hi, rec_last_close(r1));
            }
        } else {
            // donor: every tile image faces the observer for |phase - centre| < half width
            const int h = t - n_half;
            const bool on = !(A.flags & LFB_FLAG_SKIP_DONOR);
            double4 q = on ? A.don[w * G.n_donor_q + h] : make_double4(1.0, 0.0, 0.0, 0.0);
            double Aq = W.si * q.x, Bq = -W.si * q.y, Dq = W.ci * q.z;
            double rho = sqrt(Aq * Aq + Bq * Bq);
            double psi = atan2(Bq, Aq) * (1.0 / kTwoPi);
            double ratio = rho > 0.0 ? Dq / rho : (Dq > 0.0 ? 2.0 : -2.0);
            // image with +D faces the observer iff cos(th - psi) > -D/rho
            double hp = ratio >= 1.0 ? 0.5 : (ratio <= -1.0 ? -1.0 : acos(-ratio) * (1.0 / kTwoPi));
            double hm = ratio <= -1.0 ? 0.5 : (ratio >= 1.0 ? -1.0 : acos(ratio) * (1.0 / kTwoPi));
            // W m (1 - u + u m), m = A c + B s + D, as moments of (1, c, s, c^2, c s): the four mirror
            // images differ by the signs of B and D, so eight integers serve all of them
            {
                const double ud = G.donor_ulimb, sc = on ? q.w * A.jc[job].don_sc : 0.0;
                const double Bp = W.si * q.y;
                long long* m = A.qmom + (job * G.n_donor_q + h) * 8;
                m[0] = llrint(sc * ud * (Dq * Dq + Bp * Bp));
                m[1] = llrint(sc * (1.0 - ud) * Dq);
                m[2] = llrint(sc * (1.0 - ud) * Aq);
                m[3] = llrint(sc * 2.0 * ud * Aq * Dq);
                m[4] = llrint(sc * (1.0 - ud) * Bp);
                m[5] = llrint(sc * 2.0 * ud * Bp * Dq);
                m[6] = llrint(sc * ud * (Aq * Aq - Bp * Bp));
                m[7] = llrint(sc * 2.0 * ud * Aq * Bp);
            }
            EventRec* dnp = ivp + G.n_wd + G.n_disc + G.n_bs + 4 * h;
#pragma unroll
            for (int im = 0; im < 4; ++im) {
                double cen = ((im & 1) ? -psi : psi) + phi0w, hw = (im & 2) ? hm : hp;
                EventRec p = none;
                if (on && hw >= 0.5) p.x = (p.x & ~(unsigned long long)kFieldNone) | enc_pos(kAtStart);  // always faces the observer
                else if (on && hw >= 0.0) p = interval_pieces(X, cen - hw, cen + hw);
                dnp[im] = p;
            }
        }
    }
    // span of the job's eclipse events (lets chunks far from the eclipse skip the tile records)
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && lo != kNoEvent) {
        atomicMin(&A.js[job].ev_lo, lo);
        atomicMax(&A.js[job].ev_hi, hi);
    }
}

// flux_kernel: one CTA per job -- stages (2)-(4) of the model, one pass over the job's sorted
// exposure-sample axis in segments of at most Ms samples.
//
// Every eclipse / facing interval is a few events (+w where it opens, -w where it closes) on the
// axis and the component curves are running sums of those events (2^-56 fixed point, so sums do
// not depend on the order the events arrive in).  Per segment:
//   * white-dwarf, disc and strip events are dense around the eclipse: shared-memory atomics add
//     them into three per-sample delta arrays, a block scan turns the deltas into running sums;
//   * donor events are sparse (one per ~5 samples): they are counting-sorted by sample, their
//     moment contributions block-scanned in event order, and the running sums kept per event;
//   * every sample then reads its three tile sums and the donor sums of the last event at or
//     before it, evaluates the four components, and the exposure quadrature, residuals and a
//     warp-shuffle + shared-memory chi-squared reduction follow.
template <int Ms, int EC, int CTAS>
__global__ void __launch_bounds__(kFluxThreads, CTAS) flux_kernel(const __grid_constant__ FluxArgs A)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const GridCfg& G = A.G;
    constexpr int NW = kFluxThreads / 32;
    constexpr int RP = Ms / kFluxThreads;  // samples per thread in the block scans
    constexpr int EP = EC / kFluxThreads;  // donor events per thread (EC = events whose running sums fit at once)
    constexpr int ND = kNumArr - 3;        // donor moment arrays
    static_assert(Ms % kFluxThreads == 0 && EC % kFluxThreads == 0, "capacities: multiples of the block size");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nF = A.mode ? 4 : 1;
    const int NDQ = G.n_donor_q;
    long long* qmom = (long long*)smraw;                              // [NDQ][8] donor moment parts of a quarter tile
    unsigned long long* Dt = (unsigned long long*)(qmom + 8 * NDQ);  // [3][Ms] tile deltas, then running sums (f64),
    double* Dsum = (double*)Dt;                                       // then, sample by sample, ...
    double* Fs = Dsum;                                                // [nF][Ms] flux per sample of the segment
    double* P = Dsum + (size_t)(nF > 3 ? nF : 3) * Ms;                // [ND][EC] donor running sums after each event
    int* S = (int*)(P + ND * EC);                                     // [Ms + 1] donor events per sample -> bucket ends
    unsigned short* ev = (unsigned short*)(S + Ms + 1);               // [6 * 4 NDQ] donor events sorted by sample
    __shared__ double red[NW];
    __shared__ long long wtot[kNumArr][NW];
    __shared__ long long s_carry[kNumArr], s_next[kNumArr];
    __shared__ int s_itot[NW];

    const long long job = blockIdx.x;
    const long long w = job / A.L.n_ecl;
    const int egather = (int)(job - w * A.L.n_ecl);
    const int e = A.mode ? 0 : egather;
    const long long ch0 = A.smp.chunk_off[e];
    const int n_seg = (int)(A.smp.chunk_off[e + 1] - ch0);
    const long long lc0 = A.smp.lc_off[e];
    const int n_ph = (int)(A.smp.lc_off[e + 1] - lc0);
    const int K = G.n_quad;
    const WalkerScal& W = A.ws[w];
    const JobScal& J = A.js[job];
    if (!job_live(A, W, J)) {
        const bool skipped = J.status == 4 || J.status == 0;
        if (A.mode == 0) {
            if (tid == 0) A.chisq_job[job] = skipped ? NAN : INFINITY;
        } else {
            for (int j = tid; j < n_ph; j += kFluxThreads) {
                A.flux_tot[job * n_ph + j] = NAN;
                if (A.flux_comp)
                    for (int cidx = 0; cidx < 4; ++cidx) A.flux_comp[((long long)cidx * A.njobs + job) * n_ph + j] = NAN;
            }
        }
        return;
    }
    const JobConst C = A.jc[job];
    const long long* wq_tab = A.wq + job * (G.n_wd_rings + G.n_disc_r + G.n_bs);
    const EventRec* ivp = A.ivp + job * A.ni_total;
    const int n_tile_iv = G.n_wd + G.n_disc + G.n_bs;
    const EventRec* dnp = ivp + n_tile_iv;
    const bool do_don = !(A.flags & LFB_FLAG_SKIP_DONOR);
    const int n_img = do_don ? 4 * NDQ : 0;

    if (do_don) {
        const longlong2* src = (const longlong2*)(A.qmom + job * NDQ * 8);
        for (int i = tid; i < 4 * NDQ; i += kFluxThreads) ((longlong2*)qmom)[i] = __ldg(src + i);
    }
    // moments (1, c, s, c^2, c s) of donor tile image im: the mirror images of a quarter tile differ by the
    // signs of B (bit 0 set: +) and D (bit 1 set: -)
    auto add_image = [&](int im, long long sgn, long long* acc) {
        const long long* m = qmom + 8 * (im >> 2);
        const long long sb = (im & 1) ? sgn : -sgn, sd = (im & 2) ? -1 : 1;
        acc[0] += sgn * (m[0] + sd * m[1]);
        acc[1] += sgn * (m[2] + sd * m[3]);
        acc[2] += sb * (m[4] + sd * m[5]);
        acc[3] += sgn * m[6];
        acc[4] += sb * m[7];
    };
    // inclusive warp scan of a 64-bit value
    auto warp_incl = [&](long long v) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        return v;
    };
    double chi = 0.0;
    __syncthreads();

    for (int seg = 0; seg < n_seg; ++seg) {
        const int4 ch = __ldg(A.smp.chunks + ch0 + seg);  // first point, one past last point, first sample, last sample
        const int j0 = ch.x, j1 = ch.y, m0 = ch.z, m1 = ch.w + 1, len = m1 - m0;
        const bool tiles_matter = J.ev_lo < m1 && J.ev_hi >= m0;
        // the next segment starts at sample m0n of this one (segments may overlap, they leave no gap:
        // 1 <= m0n <= len); its running sums start from those after sample m0n - 1
        const int m0n = seg + 1 < n_seg ? __ldg(&A.smp.chunks[ch0 + seg + 1].z) - m0 : -1;

        // ---- 1. events of the segment: tile deltas, donor counts, sums already open at sample 0 ----
        if (tiles_matter)
            for (int q = tid; q < 3 * Ms; q += kFluxThreads) Dt[q] = 0ull;
        for (int q = tid; q < Ms + 1; q += kFluxThreads) S[q] = 0;
        __syncthreads();
        long long base[kNumArr];
#pragma unroll
        for (int a = 0; a < kNumArr; ++a) base[a] = 0;
        if (tiles_matter || seg == 0)
            for (int i0 = tid; i0 < n_tile_iv; i0 += 4 * kFluxThreads) {
                // (records are stored shuffled -- rec_slot -- so that the lanes of a warp hold tiles eclipsed at
                // different samples and their shared-memory atomics rarely meet)
                EventRec recs[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kFluxThreads;
                    recs[u] = i < n_tile_iv ? ivp[i] : no_events();
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kFluxThreads;
                    const EventRec rec = recs[u];
                    const int first = dec_pos(rec.x, 0);
                    if (first == kNoEvent || first >= m1) continue;
                    const int widx = __ldg(G.rec_widx + i);
                    const long long wq = __ldg(wq_tab + widx);  // the job's weight table: WD rings, disc rings, strip
                    const int arr = (widx >= G.n_wd_rings) + (widx >= G.n_wd_rings + G.n_disc_r);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int p = dec_pos(k < 3 ? rec.x : rec.y, k % 3);
                        if (p == kAtStart) {
                            if (seg == 0) base[arr] += wq;
                        } else if (p >= m0 && p < m1) {
                            atomicAdd(Dt + arr * Ms + (p - m0), (unsigned long long)((k & 1) ? -wq : wq));
                        }
                    }
                }
            }
        for (int i = tid; i < n_img; i += kFluxThreads) {
            const EventRec rec = dnp[i];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int p = dec_pos(k < 3 ? rec.x : rec.y, k % 3);
                if (p == kNoEvent) break;  // positions ascend, empty fields come last
                if (p == kAtStart) {
                    if (seg == 0) add_image(i, 1, base + 3);
                } else if (p >= m0 && p < m1) {
                    atomicAdd(&S[p - m0 + 1], 1);
                }
            }
        }
        if (seg == 0) {
#pragma unroll
            for (int a = 0; a < kNumArr; ++a) {
                long long v = base[a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) wtot[a][wid] = v;
            }
        }
        __syncthreads();
        if (seg == 0 && tid < kNumArr) {
            long long v = 0;
            for (int i = 0; i < NW; ++i) v += wtot[tid][i];
            s_carry[tid] = v;  // running sums just before sample 0
        }
        // ---- 2. donor bucket offsets: exclusive scan of the counts (S[x + 1] = events before sample x) ----
        {
            int cnt[RP], tot = 0;
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                cnt[r] = S[1 + tid * RP + r];
                tot += cnt[r];
            }
            int inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) s_itot[wid] = inc;
            __syncthreads();
            int run = inc - tot;
#pragma unroll
            for (int i = 0; i < NW; ++i) run += i < wid ? s_itot[i] : 0;
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                const int c2 = cnt[r];
                S[1 + tid * RP + r] = run;  // start of sample (tid*RP + r)'s bucket, used as fill cursor
                run += c2;
            }
        }
        __syncthreads();
        // ---- 3. fill the donor buckets; afterwards S[x + 1] = number of events at or before sample x ----
        for (int i = tid; i < n_img; i += kFluxThreads) {
            const EventRec rec = dnp[i];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int p = dec_pos(k < 3 ? rec.x : rec.y, k % 3);
                if (p == kNoEvent) break;
                if (p >= m0 && p < m1) ev[atomicAdd(&S[p - m0 + 1], 1)] = (unsigned short)((i << 1) | (k & 1));
            }
        }
        // ---- 4. tile deltas -> running sums after every sample (block scan, in place, as f64) ----
        if (tiles_matter) {
            long long loc[3][RP], tot[3], inc[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                long long t = 0;
#pragma unroll
                for (int r = 0; r < RP; ++r) {
                    t += (long long)Dt[a * Ms + tid * RP + r];
                    loc[a][r] = t;
                }
                tot[a] = t;
                inc[a] = warp_incl(t);
                if (lane == 31) wtot[a][wid] = inc[a];
            }
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                long long before = s_carry[a] + inc[a] - tot[a];
#pragma unroll
                for (int i = 0; i < NW - 1; ++i) before += i < wid ? wtot[a][i] : 0;
#pragma unroll
                for (int r = 0; r < RP; ++r) {
                    const long long v = before + loc[a][r];
                    if (tid * RP + r == m0n - 1) s_next[a] = v;
                    Dsum[a * Ms + tid * RP + r] = (double)v;
                }
            }
        } else if (tid < 3) {
            s_next[tid] = s_carry[tid];
        }
        __syncthreads();
        // ---- 5. donor events -> running sums after every event, EC events at a time; 6. the samples ----
        const int n_ev = S[len];
        const int ev_next = m0n > 0 ? S[m0n] : 0;  // events at or before sample m0n - 1
        if (tid < ND && ev_next == 0) s_next[3 + tid] = s_carry[3 + tid];
        long long dcar[ND];  // donor sums before the first event of the round
#pragma unroll
        for (int a = 0; a < ND; ++a) dcar[a] = s_carry[3 + a];
        for (int r0 = 0; r0 == 0 || r0 < n_ev; r0 += EC) {
            if (r0 < n_ev) {
                long long loc[ND][EP], tot[ND], inc[ND];
#pragma unroll
                for (int a = 0; a < ND; ++a) tot[a] = 0;
#pragma unroll
                for (int q = 0; q < EP; ++q) {
                    const int x = r0 + tid * EP + q;
                    if (x < n_ev) {
                        const int word = ev[x];
                        add_image(word >> 1, (word & 1) ? -1 : 1, tot);
                    }
#pragma unroll
                    for (int a = 0; a < ND; ++a) loc[a][q] = tot[a];
                }
#pragma unroll
                for (int a = 0; a < ND; ++a) {
                    inc[a] = warp_incl(tot[a]);
                    if (lane == 31) wtot[3 + a][wid] = inc[a];
                }
                __syncthreads();
#pragma unroll
                for (int a = 0; a < ND; ++a) {
                    long long before = dcar[a] + inc[a] - tot[a], all = dcar[a];
#pragma unroll
                    for (int i = 0; i < NW; ++i) {
                        const long long t = wtot[3 + a][i];
                        before += i < wid ? t : 0;
                        all += t;
                    }
#pragma unroll
                    for (int q = 0; q < EP; ++q) {
                        const int x = r0 + tid * EP + q;
                        const long long v = before + loc[a][q];
                        if (x < n_ev) {
                            P[a * EC + tid * EP + q] = (double)v;
                            if (x == ev_next - 1) s_next[3 + a] = v;
                        }
                    }
                    dcar[a] = all;
                }
                __syncthreads();
            }
            // ---- 6. components at every sample whose last event lies in this round ----
            for (int p = tid; p < len; p += kFluxThreads) {
                const int idx = S[p + 1];
                if (!((idx > r0 && idx <= r0 + EC) || (idx == 0 && r0 == 0))) continue;
                double t3[3], dm[ND];
#pragma unroll
                for (int a = 0; a < 3; ++a) t3[a] = tiles_matter ? Dsum[a * Ms + p] : (double)s_carry[a];
#pragma unroll
                for (int a = 0; a < ND; ++a) dm[a] = idx == 0 ? (double)s_carry[3 + a] : P[a * EC + idx - r0 - 1];
                const int m = m0 + p;
                const double c0 = __ldg(A.smp.cosS + lc0 * K + m), s0 = __ldg(A.smp.sinS + lc0 * K + m);
                const double cc = c0 * C.cphi + s0 * C.sphi, ss = s0 * C.cphi - c0 * C.sphi;
                const double bm = C.beam_a * cc + C.beam_b * ss + C.beam_d;
                const double beam = C.fis + (1.0 - C.fis) * (bm > 0.0 ? bm : 0.0);
                const double dn = dm[0] + dm[1] * cc + dm[2] * ss + dm[3] * (cc * cc) + dm[4] * (cc * ss);
                const double f0 = C.f_wd * (1.0 - t3[0] * kInvFix);
                const double f1 = C.f_d * (1.0 - t3[1] * kInvFix);
                const double f2 = C.f_s * beam * (1.0 - t3[2] * kInvFix);
                const double f3 = C.f_rs * dn;
                if (A.mode == 0) {
                    Fs[p] = f0 + f1 + f2 + f3;
                } else {
                    Fs[p] = f0;
                    Fs[Ms + p] = f1;
                    Fs[2 * Ms + p] = f2;
                    Fs[3 * Ms + p] = f3;
                }
            }
            __syncthreads();
        }
        if (tid < kNumArr && seg + 1 < n_seg) s_carry[tid] = s_next[tid];
        // ---- 7. exposure quadrature (Simpson over phase +- width), residuals / output ----
        for (int j = j0 + tid; j < j1; j += kFluxThreads) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = 0; k < K; ++k) {
                const int q = __ldg(A.smp.pos + (lc0 + j) * K + k) - m0;
                const double qw = G.quad_w[k];
                acc[0] += qw * Fs[q];
                if (A.mode) {
                    acc[1] += qw * Fs[Ms + q];
                    acc[2] += qw * Fs[2 * Ms + q];
                    acc[3] += qw * Fs[3 * Ms + q];
                }
            }
            if (A.mode == 0) {
                const double dy = __ldg(A.smp.y + lc0 + j) - acc[0];
                const double r = dy / __ldg(A.smp.ye + lc0 + j);
                chi += r * r;
                if (A.gp_resid) A.gp_resid[(lc0 + __ldg(A.smp.gp_slot + lc0 + j)) * A.n_walkers + w] = dy;
            } else {
                const int jo = __ldg(A.smp.pt_index + lc0 + j);
                A.flux_tot[job * n_ph + jo] = acc[0] + acc[1] + acc[2] + acc[3];
                if (A.flux_comp)
                    for (int cidx = 0; cidx < 4; ++cidx)
                        A.flux_comp[((long long)cidx * A.njobs + job) * n_ph + jo] = acc[cidx];
            }
        }
        __syncthreads();
    }
    if (A.mode == 0) {
        chi = block_sum<kFluxThreads>(chi, red);
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
        good = ingress_egress(R, si, ci, T, &pin, &pout);
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
    if (roche_init(a[i], R)) {
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
