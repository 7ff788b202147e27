// cv_kernels.cuh -- sm_100a kernels of the LFIT CV eclipse model (FP64, no tensor cores:
// the work is root finding, interval scatter and prefix sums, not a contraction).
//
// Reference interfaces replaced (file:line under /root/reference):
//   lfit.CV(pars).calcFlux(pars, phase, width)         CVModel.py:128,138
//   SimpleEclipse.chisq / ln_like                       CVModel.py:157-191
//   LCModel.ln_prior / SimpleEclipse.ln_prior           CVModel.py:440-491,193-324
//   Node.ln_prior / Node.ln_prob, Prior.ln_prob         model.py:426-498,83-113
//   trm.roche.xl1 / findphi / findi / bspot             CVModel.py:222,288,460,561
//
// Pipeline of one log-probability call over n walkers x n_ecl eclipses ("jobs"):
//   walker_kernel    thread per walker: L1, Phi_c, inclination from (q, dphi), Param priors,
//                    scalar validity rules
//   stream_kernel    thread per job: ballistic stream -> bright-spot impact point, azimuth
//                    rule, strip constants, parameter validity
//   elements_kernel  thread per surface element (white dwarf and donor per walker, disc and
//                    bright spot per job): ingress/egress phases from the Roche LOS solve /
//                    donor surface tiles -> HBM (16-32 B per element)
//   flux_kernel      CTA per job: every element's eclipse interval becomes two events on the
//                    sorted exposure-sample axis (fixed-point shared-memory atomics), a block
//                    scan turns events into eclipsed flux per sample, the donor's facing
//                    intervals carry five trigonometric moments the same way; then component
//                    mix, exposure quadrature and the chi-squared reduction
//   finish_kernel    thread per walker: ln_prior - chi^2/2 with the -inf rules
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lfit_b200.h"
#include "roche_device.cuh"

namespace lfb {

enum { P_WDFLUX = 0, P_DFLUX, P_SFLUX, P_RSFLUX, P_Q, P_DPHI, P_RDISC, P_ULIMB, P_RWD, P_SCALE, P_AZ,
       P_FIS, P_DEXP, P_PHI0, P_EXP1, P_EXP2, P_TILT, P_YAW };

constexpr int kFluxThreads = 256;
constexpr int kElemThreads = 128;
constexpr int kMaxDonorRings = 128;
constexpr int kMaxQuad = 15;
constexpr int kNumArr = 8;  // event arrays: white dwarf, disc, bright spot, 5 donor moments
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
    const double *y, *ye;        // [total]
    const double *S, *cosS, *sinS;  // [K * total] sorted per eclipse
    const int* pos;              // [total * K]
    const long long* chunk_off;  // [n_ecl + 1] offsets into chunk_j
    const int* chunk_j;          // [2 * n_chunks] first / last data point touching each chunk
};

struct GridCfg {
    int n_wd_rings, n_wd, n_disc_r, n_disc_th, n_disc, n_bs, n_donor_th, n_donor_q, n_quad;
    int n_wd_half, n_disc_half;  // elements solved (the other half follows by the y -> -y mirror)
    double donor_ulimb, donor_gdexp;
    const int* donor_ring_off;  // [n_donor_th + 1] offsets of each ring's quarter tiles
    double quad_off[kMaxQuad], quad_w[kMaxQuad];
};

struct WalkerScal {
    Roche R;
    double si, ci;
    double lnprior;
    int status;  // 0: a model exists; else the parameters admit none
};

struct JobScal {
    double xs, ys;             // stream impact point
    double smax, smaxp, shi;   // bright-spot strip: profile peak, peak^exp2, strip length (scale units)
    int status;                // 0 ok, 1 walker invalid, 2 stream misses disc, 3 bad parameter, 4 not needed
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

// ---------------------------------------------------------------- stream_kernel
__global__ void stream_kernel(DevLayout L, int what, int flags, long long njobs, const double* __restrict__ theta,
                              WalkerScal* ws, JobScal* __restrict__ js)
{
    long long job = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= njobs) return;
    long long w = job / L.n_ecl;
    int e = (int)(job - w * L.n_ecl);
    JobScal J;
    J.xs = J.ys = 0.0;
    J.smax = J.smaxp = J.shi = 1.0;
    J.status = 0;
    const WalkerScal W = ws[w];
    if (W.status != 0) {
        J.status = 1;
        js[job] = J;
        return;
    }
    if (what != LFB_LN_LIKE && !(W.lnprior > -INFINITY)) {
        J.status = 4;  // the prior already vetoed this walker: nothing downstream is evaluated
        js[job] = J;
        return;
    }
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
        js[job] = J;
        return;
    }
    if (do_bs) {
        double rdisc_a = rdisc * W.R.xl1;
        double imp[4];
        if (!bspot(W.R, rdisc_a, imp)) {
            J.status = 2;  // the stream misses the disc (roche.bspot raises, CVModel.py:309-316)
            if (what != LFB_LN_LIKE) ws[w].lnprior = -INFINITY;
        } else {
            J.xs = imp[0];
            J.ys = imp[1];
            J.smax = pow(exp1 / exp2, 1.0 / exp2);
            J.smaxp = pow(J.smax, exp2);
            J.shi = fmin(20.0 + J.smax, pow(J.smaxp + 30.0, 1.0 / exp2));
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
    }
    js[job] = J;
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
__global__ void __launch_bounds__(kElemThreads) elements_kernel(const __grid_constant__ ElemArgs A)
{
    const GridCfg& G = A.G;
    const int per_unit = COMP == 0 ? G.n_wd_half : COMP == 1 ? G.n_disc_half : COMP == 2 ? G.n_bs : G.n_donor_q;
    const int padded = (per_unit + 31) & ~31;
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
            int m = t / hth, j = t - m * hth;
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
    if (!ingress_egress(R, si, ci, T, &pin, &pout)) { pin = kBig; pout = -kBig; }
    if (COMP == 0) A.wd_io[unit * G.n_wd_half + t] = make_double2(pin, pout);
    else if (COMP == 1) A.disc_io[unit * G.n_disc_half + t] = make_double2(pin, pout);
    else {
        A.bs_io[unit * G.n_bs + t] = make_double2(pin, pout);
        A.bs_b[unit * G.n_bs + t] = wt;
    }
}

// ---------------------------------------------------------------- flux_kernel
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA, result in every thread.  Fixed tree: deterministic.
__device__ __forceinline__ double block_sum(double v, double* red)
{
    v = warp_sum(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kFluxThreads / 32; ++i) t += red[i];
    return t;
}

struct FluxArgs {
    DevLayout L;
    GridCfg G;
    DevSamples smp;
    int what, flags, mode;  // mode 0: chi-squared, 1: flux curves
    int Mc;                 // samples per chunk (multiple of kFluxThreads)
    int model_in_smem;      // per-point partial sums live in shared memory (else in model_scratch)
    int max_nph;
    long long njobs;
    const double* theta;
    const WalkerScal* ws;
    const JobScal* js;
    const double2* wd_io;
    const double4* don;
    const double2* disc_io;
    const double2* bs_io;
    const double* bs_b;
    double* chisq;          // [njobs]
    double* flux_tot;       // mode 1: [njobs][n_ph]
    double* flux_comp;      // mode 1 (optional): [4][njobs][n_ph]
    double* model_scratch;  // [gridDim.x][nF * max_nph] when !model_in_smem
};

// Sorted sample phases of one eclipse, with what the interpolating search needs.
struct SampleAxis {
    const double* S;
    int M;
    double s_first, s_last, scale;  // scale = (M - 1) / (s_last - s_first)
};

// First index with S[idx] > v (strict) or S[idx] >= v (!strict); M if none.  Light curves are
// close to uniformly sampled, so start from the interpolated position and gallop: a couple of
// loads instead of log2(M).
__device__ __forceinline__ int sample_search(const SampleAxis& X, double v, bool strict)
{
    const double* __restrict__ S = X.S;
    const int M = X.M;
    double gf = (v - X.s_first) * X.scale;
    int g = gf <= 0.0 ? 0 : (gf >= (double)(M - 1) ? M - 1 : (int)gf);
    int lo, hi;
    double sg = __ldg(S + g);
    if (strict ? (sg <= v) : (sg < v)) {
        lo = g + 1;
        hi = M;
        for (int step = 1; lo < M; step <<= 1) {
            int j = min(M - 1, g + step);
            double s = __ldg(S + j);
            if (strict ? (s <= v) : (s < v)) {
                lo = j + 1;
                if (j == M - 1) break;
            } else {
                hi = j;
                break;
            }
        }
    } else {
        hi = g;
        lo = 0;
        for (int step = 1; hi > 0; step <<= 1) {
            int j = max(0, g - step);
            double s = __ldg(S + j);
            if (strict ? (s <= v) : (s < v)) {
                lo = j + 1;
                break;
            } else {
                hi = j;
                if (j == 0) break;
            }
        }
    }
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        double s = __ldg(S + mid);
        if (strict ? (s <= v) : (s < v)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The samples with a < S + n < b for integer n: since S lies in [-0.5, 0.5] and b - a < 1 at
// most two shifts contribute.  Each piece is (first sample inside, first sample at or past the
// end); -1 marks "no event": an opening at sample 0 is returned through *open_at_start instead
// (the caller adds the weight to the running sum's start value), a closing past the last sample
// never happens.
__device__ __forceinline__ int4 interval_pieces(const SampleAxis& X, double a, double b, int* open_at_start)
{
    int4 p = make_int4(-1, -1, -1, -1);
    *open_at_start = 0;
    if (!(a < b) || !(a > -1e29) || !(b < 1e29)) return p;
    int n_lo = (int)ceil(X.s_first - b), n_hi = (int)floor(X.s_last - a);
    int np = 0;
    for (int n = n_lo; n <= n_hi && np < 2; ++n) {
        double an = a + n, bn = b + n;
        int po = an < X.s_first ? 0 : sample_search(X, an, true);
        int pc = bn > X.s_last ? X.M : sample_search(X, bn, false);
        if (pc <= po) continue;  // no sample inside
        if (po == 0) { *open_at_start += 1; po = -1; }
        if (pc >= X.M) pc = -1;
        if (np == 0) { p.x = po; p.y = pc; } else { p.z = po; p.w = pc; }
        ++np;
    }
    return p;
}

__device__ __forceinline__ void add_event(unsigned long long* Darr, int p, int m0, int m1, long long w)
{
    if (p >= m0 && p < m1) atomicAdd(Darr + (p - m0), (unsigned long long)w);
}

// five moments (1, c, s, c^2, c s) of W m (1 - u + u m), m = A c + B s + D, in 2^-56 fixed point
__device__ __forceinline__ void donor_moments(double sc, double ud, double Aq, double Bi, double Di, long long mo[5])
{
    mo[0] = llrint(sc * ((1.0 - ud) * Di + ud * (Di * Di + Bi * Bi)));
    mo[1] = llrint(sc * ((1.0 - ud) * Aq + 2.0 * ud * Aq * Di));
    mo[2] = llrint(sc * ((1.0 - ud) * Bi + 2.0 * ud * Bi * Di));
    mo[3] = llrint(sc * (ud * (Aq * Aq - Bi * Bi)));
    mo[4] = llrint(sc * (2.0 * ud * Aq * Bi));
}

__global__ void __launch_bounds__(kFluxThreads) flux_kernel(const __grid_constant__ FluxArgs A)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const GridCfg& G = A.G;
    constexpr int NW = kFluxThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NI = G.n_wd + G.n_disc + G.n_bs;   // eclipse intervals (mirrors included)
    const int NDQ = G.n_donor_q;
    const int Mc = A.Mc, R = Mc / kFluxThreads;  // samples per thread per chunk
    const int nF = A.mode ? 4 : 1;
    // shared-memory carve-up (16-byte aligned pieces first)
    int4* iv_p = (int4*)smraw;                            // [NI]      sample positions of each interval's events
    int4* dn_p = iv_p + NI;                               // [4 NDQ]   same for the donor tile images
    double4* dn_q = (double4*)(dn_p + 4 * NDQ);           // [NDQ]     (A, B, D, scaled weight) of a quarter tile
    unsigned long long* D = (unsigned long long*)(dn_q + NDQ);  // [kNumArr][Mc] events, then nothing else
    long long* part = (long long*)(D + kNumArr * Mc);     // [kNumArr][kFluxThreads] scan partials
    double* Fs = (double*)(part + kNumArr * kFluxThreads);  // [nF][Mc] flux per sample of the chunk
    long long* wq_bs = (long long*)(Fs + nF * Mc);        // [n_bs]    fixed-point strip weights
    long long* wq_ring = wq_bs + G.n_bs;                  // [n_disc_r + n_wd_rings]
    double* ringw = (double*)(wq_ring + G.n_disc_r + G.n_wd_rings);  // [n_disc_r + n_wd_rings]
    double* model_sm = ringw + G.n_disc_r + G.n_wd_rings; // [nF * n_ph] when model_in_smem
    __shared__ double s_par[LFB_NPAR];
    __shared__ double red[NW];
    __shared__ long long wtot[kNumArr][NW];
    __shared__ long long s_tot[kNumArr];

    const bool do_wd = !(A.flags & LFB_FLAG_SKIP_WD), do_disc = !(A.flags & LFB_FLAG_SKIP_DISC);
    const bool do_bs = !(A.flags & LFB_FLAG_SKIP_BS), do_don = !(A.flags & LFB_FLAG_SKIP_DONOR);

    for (long long job = blockIdx.x; job < A.njobs; job += gridDim.x) {
        const long long w = job / A.L.n_ecl;
        const int egather = (int)(job - w * A.L.n_ecl);
        const int e = A.mode ? 0 : egather;
        const long long lc0 = A.smp.lc_off[e];
        const int n_ph = (int)(A.smp.lc_off[e + 1] - lc0);
        const int K = G.n_quad;
        const int M = n_ph * K;
        const double* cosS = A.smp.cosS + lc0 * K;
        const double* sinS = A.smp.sinS + lc0 * K;
        const int* pos = A.smp.pos + lc0 * K;
        const int* chunk_j = A.smp.chunk_j + 2 * A.smp.chunk_off[e];
        double* model = A.model_in_smem ? model_sm : A.model_scratch + (size_t)blockIdx.x * nF * A.max_nph;
        __syncthreads();
        if (tid < LFB_NPAR) {
            double v = 0.0;
            if (tid < A.L.npars) v = fetch(A.L, A.theta + w * A.L.ndim, A.L.gather[egather * LFB_NPAR + tid]);
            else if (tid == P_EXP1) v = 2.0;
            else if (tid == P_EXP2) v = 1.0;
            else if (tid == P_TILT) v = 90.0;
            s_par[tid] = v;
        }
        __syncthreads();
        const WalkerScal W = A.ws[w];
        const JobScal J = A.js[job];
        const bool vetoed = A.what != LFB_LN_LIKE && !(W.lnprior > -INFINITY);  // prior said no: not evaluated
        if (J.status != 0 || n_ph == 0 || vetoed) {
            if (A.mode == 0) {
                if (tid == 0) A.chisq[job] = (J.status == 4 || vetoed) ? NAN : (J.status != 0 ? INFINITY : 0.0);
            } else {
                for (int j = tid; j < n_ph; j += kFluxThreads) {
                    A.flux_tot[job * n_ph + j] = NAN;
                    if (A.flux_comp)
                        for (int cidx = 0; cidx < 4; ++cidx) A.flux_comp[((long long)cidx * A.njobs + job) * n_ph + j] = NAN;
                }
            }
            continue;
        }
        SampleAxis X;
        X.S = A.smp.S + lc0 * K;
        X.M = M;
        X.s_first = __ldg(X.S);
        X.s_last = __ldg(X.S + M - 1);
        X.scale = X.s_last > X.s_first ? (double)(M - 1) / (X.s_last - X.s_first) : 0.0;
        const Roche Rr = W.R;
        const double si = W.si, ci = W.ci;
        const double rwd_a = s_par[P_RWD] * Rr.xl1, rdisc_a = s_par[P_RDISC] * Rr.xl1;
        double phi0w = s_par[P_PHI0];
        phi0w -= rint(phi0w);
        double sphi, cphi;
        sincos_(kTwoPi * phi0w, &sphi, &cphi);

        // ---- ring weights and component totals ("flux at maximum light", README.md:24-28) ----
        double* wdw = ringw + G.n_disc_r;
        if (do_disc)
            for (int m = tid; m < G.n_disc_r; m += kFluxThreads) {
                double r = rwd_a + (m + 0.5) * (rdisc_a - rwd_a) / G.n_disc_r;
                ringw[m] = pow(r, 1.0 - s_par[P_DEXP]);
            }
        if (do_wd)
            for (int k = tid; k < G.n_wd_rings; k += kFluxThreads) {
                double inv = 1.0 / G.n_wd_rings, ra = k * inv, rb = (k + 1) * inv;
                double ua = 1.0 - ra * ra, ub = 1.0 - rb * rb;
                double mubar = (2.0 / 3.0) * (ua * sqrt(ua) - ub * sqrt(ub)) / (rb * rb - ra * ra);
                wdw[k] = (1.0 - s_par[P_ULIMB]) + s_par[P_ULIMB] * mubar;
            }
        const double* bsb = A.bs_b + job * G.n_bs;
        const double4* don = A.don + w * NDQ;
        double p_s = 0.0, p_rs = 0.0, p_rw = 0.0;
        if (do_bs) for (int t = tid; t < G.n_bs; t += kFluxThreads) p_s += bsb[t];
        const double ud = G.donor_ulimb;
        if (do_don)
            for (int t = tid; t < NDQ; t += kFluxThreads) {
                // donor at quadrature (phase 0.25): c = 0, s = 1
                double4 q = don[t];
                double b = si * q.y, d = ci * q.z, m;
                m = -b + d; if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
                m = b + d;  if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
                m = -b - d; if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
                m = b - d;  if (m > 0.0) p_rs += q.w * m * (1.0 - ud + ud * m);
                p_rw += 4.0 * q.w;
            }
        const double tot_s = block_sum(p_s, red), tot_rs = block_sum(p_rs, red), tot_rw = block_sum(p_rw, red);
        double tot_wd = 0.0, tot_d = 0.0;  // after block_sum's barriers the ring weights are visible
        for (int k = 0; k < G.n_wd_rings; ++k) tot_wd += 4.0 * (2 * k + 1) * wdw[k];
        for (int m = 0; m < G.n_disc_r; ++m) tot_d += G.n_disc_th * ringw[m];
        // beamed part of the spot: polar angle tilt from +z, azimuth az - 90 + yaw
        double beam_a = 0.0, beam_b = 0.0, beam_d = 0.0, beam_norm = 1.0;
        const double fis = s_par[P_FIS];
        if (do_bs) {
            double st, ct, sp, cp;
            sincos_(s_par[P_TILT] * kDeg, &st, &ct);
            sincos_((s_par[P_AZ] - 90.0 + s_par[P_YAW]) * kDeg, &sp, &cp);
            beam_a = si * st * cp;
            beam_b = -si * st * sp;
            beam_d = ci * ct;
            double cmax = si * st + ci * ct;
            beam_norm = fis + (1.0 - fis) * (cmax > 0.0 ? cmax : 0.0);
        }
        const double f_wd = do_wd ? s_par[P_WDFLUX] : 0.0, f_d = do_disc ? s_par[P_DFLUX] : 0.0;
        const double f_s = (do_bs && beam_norm > 0.0 && tot_s > 0.0) ? s_par[P_SFLUX] / beam_norm : 0.0;
        const double f_rs = do_don ? s_par[P_RSFLUX] / tot_rs * (tot_rw * kInvFix) : 0.0;
        // fixed-point (2^-56) weights: per ring for the white dwarf and the disc, per element for the strip
        long long* wq_disc = wq_ring;
        long long* wq_wd = wq_ring + G.n_disc_r;
        for (int m = tid; m < G.n_disc_r; m += kFluxThreads) wq_disc[m] = do_disc ? llrint(ringw[m] / tot_d * kFix) : 0;
        for (int k = tid; k < G.n_wd_rings; k += kFluxThreads) wq_wd[k] = do_wd ? llrint(wdw[k] / tot_wd * kFix) : 0;
        for (int t = tid; t < G.n_bs; t += kFluxThreads) wq_bs[t] = (do_bs && tot_s > 0.0) ? llrint(bsb[t] / tot_s * kFix) : 0;
        for (int j = tid; j < nF * n_ph; j += kFluxThreads) model[j] = 0.0;
        __syncthreads();

        // ---- pass 1: every eclipse / facing interval -> positions of its events on the sample axis ----
        long long base[kNumArr];
#pragma unroll
        for (int a = 0; a < kNumArr; ++a) base[a] = 0;
        const double2* wdio = A.wd_io + w * G.n_wd_half;
        const double2* dio = A.disc_io + job * G.n_disc_half;
        const double2* bio = A.bs_io + job * G.n_bs;
        const int n_half = G.n_wd_half + G.n_disc_half;
        for (int t = tid; t < n_half + G.n_bs; t += kFluxThreads) {
            double2 io;
            int i0, arr;
            long long wq;
            bool mirror = t < n_half;
            if (t < G.n_wd_half) {
                int k = (int)sqrt(0.5 * (double)t);
                while (2 * k * k > t) --k;
                while (2 * (k + 1) * (k + 1) <= t) ++k;
                io = do_wd ? wdio[t] : make_double2(kBig, -kBig);
                wq = wq_wd[k];
                i0 = 2 * t;
                arr = 0;
            } else if (t < n_half) {
                int h = t - G.n_wd_half;
                io = do_disc ? dio[h] : make_double2(kBig, -kBig);
                wq = wq_disc[h / (G.n_disc_th / 2)];
                i0 = G.n_wd + 2 * h;
                arr = 1;
            } else {
                int h = t - n_half;
                io = do_bs ? bio[h] : make_double2(kBig, -kBig);
                wq = wq_bs[h];
                i0 = G.n_wd + G.n_disc + h;
                arr = 2;
            }
            const bool ecl = io.y > io.x;
            int nopen = 0;
            int4 p = ecl ? interval_pieces(X, io.x + phi0w, io.y + phi0w, &nopen) : make_int4(-1, -1, -1, -1);
            iv_p[i0] = p;
            long long b0 = nopen * wq;
            if (mirror) {  // the y -> -y image is eclipsed from -egress to -ingress
                p = ecl ? interval_pieces(X, -io.y + phi0w, -io.x + phi0w, &nopen) : make_int4(-1, -1, -1, -1);
                iv_p[i0 + 1] = p;
                b0 += nopen * wq;
            }
            base[0] += arr == 0 ? b0 : 0;
            base[1] += arr == 1 ? b0 : 0;
            base[2] += arr == 2 ? b0 : 0;
        }
        // donor: every tile image faces the observer for |phase - centre| < half width; while it does it
        // adds W m (1 - u + u m), m = A c + B s + D
        for (int t = tid; t < NDQ; t += kFluxThreads) {
            double4 q = do_don ? don[t] : make_double4(1.0, 0.0, 0.0, 0.0);
            double Aq = si * q.x, Bq = -si * q.y, Dq = ci * q.z;
            double rho = sqrt(Aq * Aq + Bq * Bq);
            double psi = atan2(Bq, Aq) * (1.0 / kTwoPi);
            double ratio = rho > 0.0 ? Dq / rho : (Dq > 0.0 ? 2.0 : -2.0);
            // image with +D faces the observer iff cos(th - psi) > -D/rho
            double hp = ratio >= 1.0 ? 0.5 : (ratio <= -1.0 ? -1.0 : acos(-ratio) * (1.0 / kTwoPi));
            double hm = ratio <= -1.0 ? 0.5 : (ratio >= 1.0 ? -1.0 : acos(ratio) * (1.0 / kTwoPi));
            double sc = do_don ? q.w / tot_rw * kFix : 0.0;
            dn_q[t] = make_double4(Aq, Bq, Dq, sc);
#pragma unroll
            for (int im = 0; im < 4; ++im) {
                double cen = ((im & 1) ? -psi : psi) + phi0w, hw = (im & 2) ? hm : hp;
                int4 p = make_int4(-1, -1, -1, -1);
                int nopen = 0;
                if (!do_don || hw < 0.0) {
                    // never faces the observer
                } else if (hw >= 0.5) {
                    nopen = 1;  // always does
                } else {
                    p = interval_pieces(X, cen - hw, cen + hw, &nopen);
                }
                dn_p[4 * t + im] = p;
                if (nopen) {
                    long long mo[5];
                    donor_moments(sc, ud, Aq, (im & 1) ? -Bq : Bq, (im & 2) ? -Dq : Dq, mo);
#pragma unroll
                    for (int k = 0; k < 5; ++k) base[3 + k] += nopen * mo[k];
                }
            }
        }
        // running sums start from the intervals already open at the first sample (exact integer reduction)
        long long carry[kNumArr];
#pragma unroll
        for (int a = 0; a < kNumArr; ++a) {
            long long v = base[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) wtot[a][wid] = v;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < kNumArr; ++a) {
            long long v = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) v += wtot[a][i];
            carry[a] = v;
        }

        // ---- pass 2: chunks of the sorted sample axis ----
        const int n_chunks = (M + Mc - 1) / Mc;
        for (int c = 0; c < n_chunks; ++c) {
            const int m0 = c * Mc, m1 = min(M, m0 + Mc);
            for (int i = tid; i < kNumArr * Mc; i += kFluxThreads) D[i] = 0ull;
            __syncthreads();
            for (int i = tid; i < NI; i += kFluxThreads) {
                const int4 p = iv_p[i];
                if (!((p.x >= m0 && p.x < m1) || (p.y >= m0 && p.y < m1) || (p.z >= m0 && p.z < m1) ||
                      (p.w >= m0 && p.w < m1)))
                    continue;
                int arr;
                long long wq;
                if (i < G.n_wd) {
                    int t = i >> 1;
                    int k = (int)sqrt(0.5 * (double)t);
                    while (2 * k * k > t) --k;
                    while (2 * (k + 1) * (k + 1) <= t) ++k;
                    wq = wq_wd[k];
                    arr = 0;
                } else if (i < G.n_wd + G.n_disc) {
                    wq = wq_disc[((i - G.n_wd) >> 1) / (G.n_disc_th / 2)];
                    arr = 1;
                } else {
                    wq = wq_bs[i - G.n_wd - G.n_disc];
                    arr = 2;
                }
                unsigned long long* Da = D + arr * Mc;
                add_event(Da, p.x, m0, m1, wq);
                add_event(Da, p.y, m0, m1, -wq);
                add_event(Da, p.z, m0, m1, wq);
                add_event(Da, p.w, m0, m1, -wq);
            }
            if (do_don)
                for (int i = tid; i < 4 * NDQ; i += kFluxThreads) {
                    const int4 p = dn_p[i];
                    const bool hx = p.x >= m0 && p.x < m1, hy = p.y >= m0 && p.y < m1;
                    const bool hz = p.z >= m0 && p.z < m1, hw_ = p.w >= m0 && p.w < m1;
                    if (!(hx || hy || hz || hw_)) continue;
                    const double4 q = dn_q[i >> 2];
                    long long mo[5];
                    donor_moments(q.w, ud, q.x, (i & 1) ? -q.y : q.y, (i & 2) ? -q.z : q.z, mo);
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        unsigned long long* Da = D + (3 + k) * Mc;
                        if (hx) atomicAdd(Da + (p.x - m0), (unsigned long long)mo[k]);
                        if (hy) atomicAdd(Da + (p.y - m0), (unsigned long long)(-mo[k]));
                        if (hz) atomicAdd(Da + (p.z - m0), (unsigned long long)mo[k]);
                        if (hw_) atomicAdd(Da + (p.w - m0), (unsigned long long)(-mo[k]));
                    }
                }
            __syncthreads();
            // block scan: thread t owns samples m0 + t*R .. +R-1; warp a scans the partials of array a
#pragma unroll
            for (int a = 0; a < kNumArr; ++a) {
                long long v = 0;
                for (int r = 0; r < R; ++r) v += (long long)D[a * Mc + tid * R + r];
                part[a * kFluxThreads + tid] = v;
            }
            __syncthreads();
            for (int a = wid; a < kNumArr; a += NW) {
                long long* pa = part + a * kFluxThreads + lane * NW;
                long long loc[NW], t = 0;
#pragma unroll
                for (int j = 0; j < NW; ++j) { loc[j] = t; t += pa[j]; }
                long long inc = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    long long u = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += u;
                }
                long long exc = inc - t;
#pragma unroll
                for (int j = 0; j < NW; ++j) pa[j] = exc + loc[j];
                if (lane == 31) s_tot[a] = inc;
            }
            __syncthreads();
            long long pre[kNumArr];
#pragma unroll
            for (int a = 0; a < kNumArr; ++a) {
                pre[a] = carry[a] + part[a * kFluxThreads + tid];
                carry[a] += s_tot[a];
            }
            // per-sample flux
            for (int r = 0; r < R; ++r) {
                int ml = tid * R + r, m = m0 + ml;
#pragma unroll
                for (int a = 0; a < kNumArr; ++a) pre[a] += (long long)D[a * Mc + ml];
                if (m >= m1) break;
                double c0 = __ldg(cosS + m), s0 = __ldg(sinS + m);
                double cc = c0 * cphi + s0 * sphi, ss = s0 * cphi - c0 * sphi;
                double v_wd = 1.0 - (double)pre[0] * kInvFix;
                double v_d = 1.0 - (double)pre[1] * kInvFix;
                double v_s = 1.0 - (double)pre[2] * kInvFix;
                double bm = beam_a * cc + beam_b * ss + beam_d;
                double beam = fis + (1.0 - fis) * (bm > 0.0 ? bm : 0.0);
                double dn = (double)pre[3] + (double)pre[4] * cc + (double)pre[5] * ss + (double)pre[6] * (cc * cc) +
                            (double)pre[7] * (cc * ss);
                double fwd = f_wd * v_wd, fd = f_d * v_d, fs = f_s * beam * v_s, frs = f_rs * dn;
                if (A.mode == 0) {
                    Fs[ml] = fwd + fd + fs + frs;
                } else {
                    Fs[ml] = fwd;
                    Fs[Mc + ml] = fd;
                    Fs[2 * Mc + ml] = fs;
                    Fs[3 * Mc + ml] = frs;
                }
            }
            __syncthreads();
            // exposure quadrature: every point touching this chunk collects its samples in it
            const int jlo = chunk_j[2 * c], jhi = chunk_j[2 * c + 1];
            for (int j = jlo + tid; j <= jhi; j += kFluxThreads) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                bool any = false;
                for (int k = 0; k < K; ++k) {
                    int p = __ldg(pos + j * K + k);
                    if (p >= m0 && p < m1) {
                        any = true;
                        double qw = G.quad_w[k];
                        acc[0] += qw * Fs[p - m0];
                        if (A.mode) {
                            acc[1] += qw * Fs[Mc + p - m0];
                            acc[2] += qw * Fs[2 * Mc + p - m0];
                            acc[3] += qw * Fs[3 * Mc + p - m0];
                        }
                    }
                }
                if (any) {
                    model[j] += acc[0];
                    if (A.mode) {
                        model[n_ph + j] += acc[1];
                        model[2 * n_ph + j] += acc[2];
                        model[3 * n_ph + j] += acc[3];
                    }
                }
            }
            __syncthreads();
        }

        // ---- residuals / output ----
        if (A.mode == 0) {
            double chi = 0.0;
            for (int j = tid; j < n_ph; j += kFluxThreads) {
                double r = (__ldg(A.smp.y + lc0 + j) - model[j]) / __ldg(A.smp.ye + lc0 + j);
                chi += r * r;
            }
            chi = block_sum(chi, red);
            if (tid == 0) A.chisq[job] = isnan(chi) ? INFINITY : chi;  // NaN model -> +inf (CVModel.py:163-171)
        } else {
            for (int j = tid; j < n_ph; j += kFluxThreads) {
                double fwd = model[j], fd = model[n_ph + j], fs = model[2 * n_ph + j], frs = model[3 * n_ph + j];
                A.flux_tot[job * n_ph + j] = fwd + fd + fs + frs;
                if (A.flux_comp) {
                    A.flux_comp[((long long)0 * A.njobs + job) * n_ph + j] = fwd;
                    A.flux_comp[((long long)1 * A.njobs + job) * n_ph + j] = fd;
                    A.flux_comp[((long long)2 * A.njobs + job) * n_ph + j] = fs;
                    A.flux_comp[((long long)3 * A.njobs + job) * n_ph + j] = frs;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- finish_kernel
__global__ void finish_kernel(int what, int n_ecl, long long n, const WalkerScal* __restrict__ ws,
                              const double* __restrict__ chisq, double* __restrict__ out)
{
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    double lnp = ws[w].lnprior;
    double v;
    if (what == LFB_LN_PRIOR) {
        v = lnp;
    } else {
        double like = 0.0;
        if (what == LFB_LN_LIKE || lnp > -INFINITY)
            for (int e = 0; e < n_ecl; ++e) like += -0.5 * chisq[w * n_ecl + e];
        v = what == LFB_LN_LIKE ? like : (lnp > -INFINITY ? lnp + like : -INFINITY);
    }
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
