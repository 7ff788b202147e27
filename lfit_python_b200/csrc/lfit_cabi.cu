// lfit_cabi.cu -- C ABI (include/lfit_b200.h) over the sm_100a kernels in cv_kernels.cuh.
//
// Host side only: buffer management, the one-off preprocessing of light curves into
// sorted exposure samples, and kernel launches.  There is no CPU implementation of the
// model here: without a CUDA device lfb_create fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "cv_kernels.cuh"
#include "peer.cuh"

using namespace lfb;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned_host = false;
    // Generation counter of the owning handle (null: a buffer no captured CUDA graph points into).  Bumped
    // whenever the buffer is (re)allocated or freed: a graph is only replayed while the buffers it was
    // captured with are still the ones in use.
    unsigned long long* gen = nullptr;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (gen) ++*gen;
        release();
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = pinned_host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release()
    {
        if (p) {
            if (gen) ++*gen;
            if (pinned_host) cudaFreeHost(p); else cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() { return (T*)p; }
};

// Sorted exposure samples of a set of light curves (device copies)
struct SampleSet {
    DevBuf lc_off, y, ye, S, cosS, sinS, bins, pos, pt_index, chunk_off, chunks, seg_tr, axis;
    DevBuf gp_x, gp_var, gp_slot, gp_span;  // GP likelihood: points in ascending raw phase
    int max_chunks = 1;
    int max_nph = 0;
    int max_gaps = 0;  // GP likelihood: the most change-point gaps any of the light curves spans
    long long total = 0;
    void release()
    {
        DevBuf* b[] = {&lc_off, &y, &ye, &S, &cosS, &sinS, &bins, &pos, &pt_index, &chunk_off, &chunks, &seg_tr, &axis,
                       &gp_x, &gp_var, &gp_slot, &gp_span};
        for (DevBuf* x : b) x->release();
    }
    DevSamples view()
    {
        DevSamples v;
        v.lc_off = lc_off.as<long long>();
        v.y = y.as<double>();
        v.iye = ye.as<double>();
        v.S = S.as<double>();
        v.cosS = cosS.as<double>();
        v.sinS = sinS.as<double>();
        v.bins = bins.as<int>();
        v.axis = axis.as<double4>();
        v.pos = pos.as<int>();
        v.pt_index = pt_index.as<int>();
        v.gp_x = gp_x.as<double>();
        v.gp_var = gp_var.as<double>();
        v.gp_slot = gp_slot.as<int>();
        v.gp_span = gp_span.as<double2>();
        v.chunk_off = chunk_off.as<long long>();
        v.chunks = chunks.as<int4>();
        v.seg_tr = seg_tr.as<double>();
        return v;
    }
};

enum { ST_WALKER = 0, ST_STREAM, ST_ELEMENTS, ST_FLUX, ST_FINISH, ST_COUNT };

// A lane runs one batch of walkers through the whole pipeline on its own streams and buffers.
// Two lanes work on alternate batches so that one batch's FP64-bound element solves overlap
// the other's shared-memory-bound flux stage.
struct Lane {
    cudaStream_t st = nullptr, side = nullptr, side2 = nullptr;  // side: the serial stream ODE beside the element
                                                                 // solves; side2: the white-dwarf centre's LOS
    cudaEvent_t fork_ev = nullptr, fork2_ev = nullptr, join_ev = nullptr, done_ev = nullptr, wd_ev = nullptr, don_ev = nullptr;
    cudaEvent_t dev[3] = {};  // trace of the donor chain on the second side stream
    cudaEvent_t ev[ST_COUNT + 1] = {};
    cudaEvent_t kev[LFB_K_COUNT + 1] = {}, sev[2] = {};  // per-kernel trace (lfb_set_trace), side-stream pair
    bool kev_set[LFB_K_COUNT + 1] = {};
    DevBuf ws, js, wd_io, don, disc_io, bs_io, bs_b, jc, wq, ivp, chi_part, gp_resid, dt_first, dt_mom;
    void track(unsigned long long* gen)
    {
        DevBuf* b[] = {&ws, &js, &wd_io, &don, &disc_io, &bs_io, &bs_b, &jc, &wq, &ivp, &chi_part, &gp_resid, &dt_first, &dt_mom};
        for (DevBuf* x : b) x->gen = gen;
    }
    cudaError_t create()
    {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&done_ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&wd_ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&don_ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&fork2_ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        for (int i = 0; i < 3; ++i)
            if ((e = cudaEventCreate(&dev[i])) != cudaSuccess) return e;
        for (int i = 0; i <= ST_COUNT; ++i)
            if ((e = cudaEventCreate(&ev[i])) != cudaSuccess) return e;
        for (int i = 0; i <= LFB_K_COUNT; ++i)
            if ((e = cudaEventCreate(&kev[i])) != cudaSuccess) return e;
        for (int i = 0; i < 2; ++i)
            if ((e = cudaEventCreate(&sev[i])) != cudaSuccess) return e;
        return cudaSuccess;
    }
    void destroy()
    {
        DevBuf* b[] = {&ws, &js, &wd_io, &don, &disc_io, &bs_io, &bs_b, &jc, &wq, &ivp, &chi_part, &gp_resid, &dt_first, &dt_mom};
        for (DevBuf* x : b) x->release();
        for (int i = 0; i <= ST_COUNT; ++i)
            if (ev[i]) cudaEventDestroy(ev[i]);
        for (int i = 0; i <= LFB_K_COUNT; ++i)
            if (kev[i]) cudaEventDestroy(kev[i]);
        for (int i = 0; i < 2; ++i)
            if (sev[i]) cudaEventDestroy(sev[i]);
        if (fork_ev) cudaEventDestroy(fork_ev);
        if (join_ev) cudaEventDestroy(join_ev);
        if (done_ev) cudaEventDestroy(done_ev);
        if (wd_ev) cudaEventDestroy(wd_ev);
        if (don_ev) cudaEventDestroy(don_ev);
        if (fork2_ev) cudaEventDestroy(fork2_ev);
        for (int i = 0; i < 3; ++i)
            if (dev[i]) cudaEventDestroy(dev[i]);
        if (side) cudaStreamDestroy(side);
        if (side2) cudaStreamDestroy(side2);
        if (st) cudaStreamDestroy(st);
    }
};
constexpr int kLanes = 2;

}  // namespace

struct lfb_handle {
    int device = 0;
    lfb_config cfg{};
    GridCfg grid{};
    cudaStream_t stream = nullptr;
    Lane lanes[kLanes];
    cudaEvent_t enter_ev = nullptr, t0_ev = nullptr, t1_ev = nullptr;
    bool ev_valid = false;
    bool trace = false;  // lfb_set_trace
    bool debug_sync = false;  // LFB_DEBUG_SYNC=1: synchronise after every kernel launch and name the one that faults
    // Gaussian-process likelihood (lfb_set_gp)
    bool gp_on = false;
    int gp_src[3] = {0, 0, 0};
    DevBuf gp_dist;
    std::string err;
    long long launches = 0;
    int sm_count = 148;
    int max_smem = 0;
    int flux_variant = 0;  // shape of the chi-squared flux kernel (LFB_FLUX_VARIANT: 0 = 256 x 13, 1 = 512 x 13, 2 = 128 x 13)
    long long max_jobs_per_batch = 131072;
    // CUDA graphs for small, repeated calls (launch-bound): one captured pass per (what, n, pointers)
    struct GraphEntry {
        int what = -1;
        long long n = 0;
        const void *theta = nullptr, *out = nullptr, *chi = nullptr;
        int seen = 0;            // identical calls so far (the second one is captured)
        long long launches = 0;  // kernels inside
        unsigned long long generation = 0;  // alloc_generation at capture
        cudaGraphExec_t exec = nullptr;
    };
    GraphEntry graphs[4];
    int graph_next = 0;
    bool graphs_on = true, stages_valid = true;
    long long graph_max_jobs = 131072;  // = max_jobs_per_batch: every one-batch call (LFB_GRAPH_MAX_JOBS)
    long long stream_lanes_below = 1536;  // batches smaller than this spread each stream ODE over eight lanes
    size_t flux_smem_pad = 0, donor_smem_pad = 0;  // tuning: extra dynamic shared memory = fewer resident CTAs per SM
    int n_lanes = 2;       // LFB_LANES=1 serialises the batches (clean per-stage timings for profiling)
    // layout
    bool have_layout = false, have_lc = false;
    int ndim = 0, n_ecl = 0, npars = 0, n_prior = 0, n_consts = 0;
    DevBuf gather, consts, psrc, ptype, pisvar, pp1, pp2, pnorm, donor_off, disc_order, rec_widx, rec_slot, tile_geo;
    SampleSet lc, cf_lc;
    // calc_flux scratch
    DevBuf cf_gather, cf_pars, cf_tot, cf_comp;
    // work
    DevBuf theta, out, chisq;
    DevBuf h_in, h_out, h_chisq;
    DevBuf scratch[6];  // device staging of the scalar helpers (lfb_roche, lfb_wdphases, ...): kept between calls
    bool h2d_pending = false;  // the last call staged theta through h_in and returned without synchronising
    // bumped when a buffer a captured graph may point into moves (the lanes' buffers, theta / out / chisq staging)
    unsigned long long alloc_generation = 0;
    // exchange over NVLink peer memory (lfb_peer_*): this rank's window and the peers' windows as mapped here
    struct Peer {
        int rank = -1, world = 0;
        long long slot_bytes = 0;
        size_t bytes = 0;
        unsigned char* win[kMaxPeers] = {};
        unsigned int* ticket = nullptr;
        unsigned long long step = 0;
        bool connected = false;
    } peer;
    lfb_handle()
    {
        h_in.pinned_host = h_out.pinned_host = h_chisq.pinned_host = true;
        for (int i = 0; i < kLanes; ++i) lanes[i].track(&alloc_generation);
        theta.gen = out.gen = chisq.gen = h_in.gen = &alloc_generation;
    }
};

static std::string g_create_error;

// The flux kernel's shape per mode: threads, consecutive samples per thread (segment capacity = NT * RP samples),
// resident CTAs per SM.  Chi-squared mode has a few variants for tuning (LFB_FLUX_VARIANT).
struct FluxShape {
    int NT, RP, CTAS;
};
static FluxShape flux_shape(const lfb_handle* h, int mode)
{
    if (mode) return {256, 6, 4};
    if (h->flux_variant == 1) return {512, 13, 2};
    if (h->flux_variant == 2) return {128, 13, 8};
    return {256, 13, 4};
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return LFB_ECUDA;                                                                      \
        }                                                                                          \
    } while (0)

static int fail(lfb_handle* h, int code, const std::string& msg)
{
    h->err = msg;
    return code;
}

static void drop_graphs(lfb_handle* h)
{
    for (auto& g : h->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = lfb_handle::GraphEntry();
    }
}

static bool is_device_ptr(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor): copies to and from it are
// asynchronous as they stand, so it is not staged through the handle's own pinned buffers
static bool is_pinned_host_ptr(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static int upload(lfb_handle* h, DevBuf& b, const void* src, size_t bytes)
{
    CK(b.reserve(bytes ? bytes : 8));
    if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyDefault, h->stream));
    return LFB_OK;
}

static int donor_ring_count(int nth, int k)
{
    double th = (k + 0.5) * lfb::kPi / nth;
    return (int)fmax(1.0, floor(0.5 * nth * sin(th) + 0.5));
}

// Light-curve preprocessing, done once per set_lightcurves / per calc_flux phase grid and shared
// by all walkers: order the points in (wrapped) phase, merge their K exposure samples, wrap to
// [-0.5, 0.5], sort, record where each (point, node) landed, build the bin table of the sorted
// axis, and cut the points into chunks whose samples span at most Mc consecutive sorted samples.
static int build_samples(lfb_handle* h, SampleSet& ss, int NT, int RP, int n_ecl, const long long* off,
                         const double* phase, const double* width, const double* y, const double* ye)
{
    const int Mc = NT * RP;  // capacity of a flux-kernel segment in samples
    const GridCfg& G = h->grid;
    const int K = G.n_quad;
    const long long total = off[n_ecl];
    std::vector<double> S((size_t)total * K), cS((size_t)total * K), sS((size_t)total * K);
    std::vector<double> ys((size_t)total), yes((size_t)total);
    std::vector<int> pos((size_t)total * K), bins((size_t)total * K + n_ecl, 0), pt_index((size_t)total);
    std::vector<double> gp_x((size_t)total), gp_var((size_t)total);
    std::vector<int> gp_slot((size_t)total), gp_rank;
    std::vector<double2> gp_span((size_t)std::max(n_ecl, 1));
    std::vector<double4> axis((size_t)std::max(n_ecl, 1), make_double4(0.0, 0.0, 0.0, 0.0));
    std::vector<long long> chunk_off(n_ecl + 1, 0);
    std::vector<int4> chunks;
    int max_nph = 0, max_chunks = 1, max_gaps = 0;
    std::vector<int> order, pt;
    std::vector<double> raw, wph;
    for (int e = 0; e < n_ecl; ++e) {
        const long long o = off[e];
        const int n_ph = (int)(off[e + 1] - o), M = n_ph * K;
        max_nph = std::max(max_nph, n_ph);
        if (M > kMaxSamples) {
            h->err = "set_lightcurves: more than 2^21 exposure samples in one light curve";
            return LFB_EINVAL;
        }
        for (int j = 0; j < n_ph; ++j)
            if (width && !(fabs(width[o + j]) < 0.25)) {
                h->err = "set_lightcurves: exposure half-width must be below a quarter of the orbit";
                return LFB_EINVAL;
            }
        // points in wrapped-phase order
        wph.resize(n_ph);
        pt.resize(n_ph);
        for (int j = 0; j < n_ph; ++j) wph[j] = phase[o + j] - rint(phase[o + j]);
        std::iota(pt.begin(), pt.end(), 0);
        std::stable_sort(pt.begin(), pt.end(), [&](int a, int b) { return wph[a] < wph[b]; });
        raw.resize(M);
        order.resize(M);
        for (int jj = 0; jj < n_ph; ++jj) {
            const int j = pt[jj];
            pt_index[o + jj] = j;
            ys[o + jj] = y ? y[o + j] : 0.0;
            yes[o + jj] = ye ? 1.0 / ye[o + j] : 1.0;  // the kernel multiplies
            // the point is wrapped as a whole: its samples stay together on the axis
            for (int k = 0; k < K; ++k) raw[jj * K + k] = wph[j] + G.quad_off[k] * (width ? width[o + j] : 0.0);
        }
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return raw[a] < raw[b]; });
        for (int r = 0; r < M; ++r) {
            int src = order[r];
            S[o * K + r] = raw[src];
            cS[o * K + r] = cos(kTwoPi * raw[src]);
            sS[o * K + r] = sin(kTwoPi * raw[src]);
            pos[o * K + src] = r;
        }
        // bin table: first sample at or after the start of each of M equal phase bins
        if (M > 0) {
            int* bt = bins.data() + o * K + e;
            const double s0 = S[o * K], s1 = S[o * K + M - 1];
            const double inv_binw = s1 > s0 ? (double)M / (s1 - s0) : 0.0;
            axis[e] = make_double4(s0, s1, inv_binw, 0.0);
            int r = 0;
            for (int b = 0; b <= M; ++b) {
                while (r < M) {
                    double gf = (S[o * K + r] - s0) * inv_binw;
                    int br = gf <= 0.0 ? 0 : (gf >= (double)(M - 1) ? M - 1 : (int)gf);
                    if (br >= b) break;
                    ++r;
                }
                bt[b] = r;
            }
        }
        // chunks of consecutive points: all their samples inside [pmin, pmax], pmax - pmin < Mc
        int nch = 0;
        for (int j0 = 0; j0 < n_ph;) {
            int pmin = M, pmax = -1, j1 = j0;
            while (j1 < n_ph) {
                int lo = pmin, hi = pmax;
                for (int k = 0; k < K; ++k) {
                    lo = std::min(lo, pos[o * K + j1 * K + k]);
                    hi = std::max(hi, pos[o * K + j1 * K + k]);
                }
                if (hi - lo + 1 > Mc) break;
                pmin = lo;
                pmax = hi;
                ++j1;
            }
            if (j1 == j0) {
                h->err = "set_lightcurves: one exposure spans more sorted samples than a chunk holds "
                         "(width far larger than the sampling); not supported";
                return LFB_EINVAL;
            }
            chunks.push_back(make_int4(j0, j1, pmin, pmax));
            ++nch;
            j0 = j1;
        }
        // the flux kernel carries running sums from one segment to the next: the sample ranges must
        // start at sample 0 and leave no gap (they may overlap by a few samples)
        for (int k = 0; k < nch; ++k) {
            int4& cur = chunks[chunks.size() - nch + k];
            const int want = k == 0 ? 0 : chunks[chunks.size() - nch + k - 1].w + 1;
            if (cur.z > want) cur.z = want;
            if (cur.w - cur.z + 1 > Mc) {
                h->err = "set_lightcurves: exposures overlap too irregularly for the segmented sample axis";
                return LFB_EINVAL;
            }
        }
        // GP likelihood (CVModel.py:650-696): the points in ascending raw (unwrapped) phase, their noise
        // variances, and where each wrapped-phase-ordered point sits in that order
        gp_rank.resize(n_ph);
        std::iota(gp_rank.begin(), gp_rank.end(), 0);
        std::stable_sort(gp_rank.begin(), gp_rank.end(), [&](int a, int b) { return phase[o + a] < phase[o + b]; });
        for (int r = 0; r < n_ph; ++r) {
            const int j = gp_rank[r];
            gp_x[o + r] = phase[o + j];
            gp_var[o + r] = ye ? ye[o + j] * ye[o + j] : 1.0;
            wph[j] = (double)r;  // reuse: rank of original point j
        }
        for (int jj = 0; jj < n_ph; ++jj) gp_slot[o + jj] = (int)wph[pt_index[o + jj]];
        gp_span[e] = n_ph ? make_double2(gp_x[o], gp_x[o + n_ph - 1]) : make_double2(0.0, 0.0);
        {
            // one gap per cycle number c with x_min < c < 1 + x_max (gp_changepoints)
            int ngap = 0;
            for (int c = (int)floor(gp_span[e].x); n_ph && c <= (int)ceil(gp_span[e].y); ++c)
                if ((double)c > gp_span[e].x && (double)c < 1.0 + gp_span[e].y) ++ngap;
            max_gaps = std::max(max_gaps, ngap);
        }
        chunk_off[e + 1] = chunk_off[e] + nch;
        max_chunks = std::max(max_chunks, nch);
    }
    if (chunks.empty()) chunks.push_back(make_int4(0, 0, 0, -1));
    // phase, cos, sin of every segment's samples in the flux kernel's thread-major order: thread t evaluates
    // samples t * RP .. t * RP + RP - 1 of the segment and reads [r][t], coalesced
    std::vector<double> seg_tr((size_t)chunks.size() * 3 * Mc, 0.0);
    {
        size_t ci = 0;
        for (int e = 0; e < n_ecl; ++e) {
            const size_t so = (size_t)off[e] * K;
            for (long long c = chunk_off[e]; c < chunk_off[e + 1]; ++c, ++ci) {
                const int4 ch = chunks[(size_t)c];
                double* dst = seg_tr.data() + (size_t)c * 3 * Mc;
                for (int p = 0; p <= ch.w - ch.z; ++p) {
                    const size_t at = (size_t)(p % RP) * NT + (size_t)(p / RP);
                    dst[at] = S[so + ch.z + p];
                    dst[(size_t)Mc + at] = cS[so + ch.z + p];
                    dst[2 * (size_t)Mc + at] = sS[so + ch.z + p];
                }
            }
        }
    }
    int rc;
    if ((rc = upload(h, ss.seg_tr, seg_tr.data(), sizeof(double) * seg_tr.size()))) return rc;
    if ((rc = upload(h, ss.lc_off, off, sizeof(long long) * (size_t)(n_ecl + 1)))) return rc;
    if ((rc = upload(h, ss.y, ys.data(), sizeof(double) * (size_t)total))) return rc;
    if ((rc = upload(h, ss.ye, yes.data(), sizeof(double) * (size_t)total))) return rc;
    if ((rc = upload(h, ss.S, S.data(), sizeof(double) * S.size()))) return rc;
    if ((rc = upload(h, ss.cosS, cS.data(), sizeof(double) * cS.size()))) return rc;
    if ((rc = upload(h, ss.sinS, sS.data(), sizeof(double) * sS.size()))) return rc;
    if ((rc = upload(h, ss.bins, bins.data(), sizeof(int) * bins.size()))) return rc;
    if ((rc = upload(h, ss.axis, axis.data(), sizeof(double4) * axis.size()))) return rc;
    if ((rc = upload(h, ss.pos, pos.data(), sizeof(int) * pos.size()))) return rc;
    if ((rc = upload(h, ss.pt_index, pt_index.data(), sizeof(int) * pt_index.size()))) return rc;
    if ((rc = upload(h, ss.gp_x, gp_x.data(), sizeof(double) * gp_x.size()))) return rc;
    if ((rc = upload(h, ss.gp_var, gp_var.data(), sizeof(double) * gp_var.size()))) return rc;
    if ((rc = upload(h, ss.gp_slot, gp_slot.data(), sizeof(int) * gp_slot.size()))) return rc;
    if ((rc = upload(h, ss.gp_span, gp_span.data(), sizeof(double2) * gp_span.size()))) return rc;
    if ((rc = upload(h, ss.chunk_off, chunk_off.data(), sizeof(long long) * chunk_off.size()))) return rc;
    if ((rc = upload(h, ss.chunks, chunks.data(), sizeof(int4) * chunks.size()))) return rc;
    CK(cudaStreamSynchronize(h->stream));  // the host vectors die here
    ss.max_nph = max_nph;
    ss.max_chunks = max_chunks;
    ss.max_gaps = max_gaps;
    ss.total = total;
    return LFB_OK;
}

// One pass of the pipeline over walkers [0, n) (device pointers, one batch).
static int run_batch(lfb_handle* h, Lane& ln, const DevLayout& L, SampleSet& ss, int what, int flags, int mode,
                     long long n, const double* d_theta, double* d_out, double* d_chi, double* d_tot, double* d_comp,
                     bool record)
{
    const GridCfg& G = h->grid;
    const long long njobs = n * L.n_ecl;
    cudaStream_t st = ln.st;
    CK(ln.ws.reserve(sizeof(WalkerScal) * (size_t)n));
    CK(ln.js.reserve(sizeof(JobScal) * (size_t)njobs));
    CK(ln.wd_io.reserve(sizeof(double2) * (size_t)n * G.n_wd_half));
    CK(ln.don.reserve(sizeof(double4) * (size_t)n * G.n_donor_q));
    CK(ln.disc_io.reserve(sizeof(double2) * (size_t)njobs * G.n_disc_half));
    CK(ln.bs_io.reserve(sizeof(double2) * (size_t)njobs * G.n_bs));
    CK(ln.bs_b.reserve(sizeof(double) * (size_t)njobs * G.n_bs));
    const bool trace = record && h->trace;
    if (trace)
        for (int i = 0; i <= LFB_K_COUNT; ++i) ln.kev_set[i] = false;
    // LFB_DEBUG_SYNC=1: synchronise the device after every launch and name the kernel that faulted
#define KSYNC(name)                                                                                  \
    do {                                                                                             \
        if (h->debug_sync) {                                                                         \
            cudaError_t e_ = cudaDeviceSynchronize();                                                \
            if (e_ != cudaSuccess) return fail(h, LFB_ECUDA, std::string(name) + ": " + cudaGetErrorString(e_)); \
        }                                                                                            \
    } while (0)
#define KREC(i)                                      \
    do {                                             \
        if (trace) {                                 \
            CK(cudaEventRecord(ln.kev[i], st));      \
            ln.kev_set[i] = true;                    \
        }                                            \
    } while (0)
    // argument blocks of the element and flux stages (buffers sized before anything is launched)
    ElemArgs E;
    FluxArgs A;
    const bool gp = h->gp_on && mode == 0;
    if (what != LFB_LN_PRIOR) {
        E.L = L;
        E.G = G;
        E.what = what;
        E.flags = flags;
        E.n = n;
        E.njobs = njobs;
        E.theta = d_theta;
        E.ws = ln.ws.as<WalkerScal>();
        E.js = ln.js.as<JobScal>();
        E.wd_io = ln.wd_io.as<double2>();
        E.don = ln.don.as<double4>();
        E.disc_io = ln.disc_io.as<double2>();
        E.bs_io = ln.bs_io.as<double2>();
        E.bs_b = ln.bs_b.as<double>();
        A.L = L;
        A.G = G;
        A.smp = ss.view();
        A.what = what;
        A.flags = flags;
        A.mode = mode;
        A.ni_total = G.n_wd + G.n_disc + G.n_bs;
        A.njobs = njobs;
        A.theta = d_theta;
        A.ws = E.ws;
        A.js = ln.js.as<JobScal>();
        A.wd_io = E.wd_io;
        A.don = E.don;
        A.disc_io = E.disc_io;
        A.bs_io = E.bs_io;
        A.bs_b = E.bs_b;
        const size_t nwq = (size_t)G.n_wd_rings + G.n_disc_r + G.n_bs;
        const size_t nb_max = 8 * (size_t)G.n_donor_q;
        CK(ln.jc.reserve(sizeof(JobConst) * (size_t)njobs));
        CK(ln.wq.reserve(sizeof(long long) * (size_t)njobs * nwq));
        CK(ln.ivp.reserve(sizeof(EventRec) * (size_t)njobs * A.ni_total));
        CK(ln.chi_part.reserve(sizeof(double) * (size_t)njobs));
        CK(ln.dt_first.reserve(sizeof(unsigned short) * (size_t)n * (kDonorBins + 1)));
        CK(ln.dt_mom.reserve(sizeof(double) * (size_t)n * (nb_max + 1) * 6));
        A.jc = ln.jc.as<JobConst>();
        A.wq = ln.wq.as<long long>();
        A.ivp = ln.ivp.as<EventRec>();
        A.dt.first = ln.dt_first.as<unsigned short>();
        A.dt.mom = ln.dt_mom.as<double>();
        A.dt.nb_max = (int)nb_max;
        A.chisq_job = ln.chi_part.as<double>();
        A.flux_tot = d_tot;
        A.flux_comp = d_comp;
        A.gp_resid = nullptr;
        A.n_walkers = n;
        A.gp_dist = h->gp_dist.as<double>();
        for (int i = 0; i < 3; ++i) A.gp_src[i] = h->gp_src[i];
        if (gp) {
            CK(ln.gp_resid.reserve(sizeof(double) * (size_t)ss.total * (size_t)n));
            A.gp_resid = ln.gp_resid.as<double>();
        }
        if (8 * G.n_donor_q > 65535) return fail(h, LFB_EINVAL, "donor grid too dense (16-bit break-point index)");
        if (G.n_wd + G.n_disc > 32767 || G.n_bs > 32767)
            return fail(h, LFB_EINVAL, "surface grid too dense: at most 32767 tiles per running sum (16-bit low limb)");
    }
    auto blocks = [&](long long units, int per_unit, bool whole_warps = false) {
        const long long padded = whole_warps ? (per_unit + 31) & ~31 : per_unit;
        return (unsigned)((units * padded + kElemThreads - 1) / kElemThreads);
    };
    if (record) CK(cudaEventRecord(ln.ev[ST_WALKER], st));
    KREC(LFB_K_WALKER);
    walker_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(L, what, flags, n, d_theta, ln.ws.as<WalkerScal>());
    KSYNC("walker_kernel");
    // fork 1: what needs the walker scalars only runs beside everything else on a second side stream -- the
    // white-dwarf centre's lines of sight (Newton starts of its tiles), then the donor: its tiles and its phase table
    CK(cudaEventRecord(ln.fork2_ev, st));
    CK(cudaStreamWaitEvent(ln.side2, ln.fork2_ev, 0));
    if (what != LFB_LN_PRIOR && !(flags & LFB_FLAG_SKIP_WD)) {
        wdcentre_kernel<<<(unsigned)((n + 63) / 64), 64, 0, ln.side2>>>(what, n, ln.ws.as<WalkerScal>());
        KSYNC("wdcentre_kernel");
        h->launches++;
    }
    CK(cudaEventRecord(ln.wd_ev, ln.side2));
    if (what != LFB_LN_PRIOR && !(flags & LFB_FLAG_SKIP_DONOR)) {
        const size_t dsm = 224 * (size_t)G.n_donor_q + 4 * (kDonorBins + 1) + 64 + h->donor_smem_pad;
        if (dsm > (size_t)h->max_smem - 2048) return fail(h, LFB_EINVAL, "donor grid too dense for the table kernel's shared memory");
        if (trace) CK(cudaEventRecord(ln.dev[0], ln.side2));
        elements_kernel<3><<<blocks(n, G.n_donor_q), kElemThreads, 0, ln.side2>>>(E);
        KSYNC("elements_kernel");
        if (trace) CK(cudaEventRecord(ln.dev[1], ln.side2));
        // the donor's curve as a table over phase, once per walker (every eclipse of the walker reads it)
        CK(cudaFuncSetAttribute(donor_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
        donor_table_kernel<<<(unsigned)n, kDonorThreads, dsm, ln.side2>>>(A);
        KSYNC("donor_table_kernel");
        if (trace) CK(cudaEventRecord(ln.dev[2], ln.side2));
        h->launches += 2;
    }
    CK(cudaEventRecord(ln.don_ev, ln.side2));
    if (record) CK(cudaEventRecord(ln.ev[ST_STREAM], st));
    KREC(LFB_K_JOBCHECK);
    jobcheck_kernel<<<(unsigned)((njobs + 127) / 128), 128, 0, st>>>(L, what, flags, njobs, d_theta, ln.ws.as<WalkerScal>(),
                                                                      ln.js.as<JobScal>());
    KSYNC("jobcheck_kernel");
    // fork 2: the ballistic-stream ODE (one serial integration per job) runs beside the element solves
    CK(cudaEventRecord(ln.fork_ev, st));
    CK(cudaStreamWaitEvent(ln.side, ln.fork_ev, 0));
    if (trace) CK(cudaEventRecord(ln.sev[0], ln.side));
    if (njobs < h->stream_lanes_below)
        stream_kernel<true><<<(unsigned)((njobs * 8 + 63) / 64), 64, 0, ln.side>>>(L, what, flags, njobs, d_theta, ln.ws.as<WalkerScal>(),
                                                                      ln.js.as<JobScal>());
    else
        stream_kernel<false><<<(unsigned)((njobs + 63) / 64), 64, 0, ln.side>>>(L, what, flags, njobs, d_theta, ln.ws.as<WalkerScal>(),
                                                                      ln.js.as<JobScal>());
        KSYNC("stream_kernel");
    if (trace) CK(cudaEventRecord(ln.sev[1], ln.side));
    CK(cudaEventRecord(ln.join_ev, ln.side));
    h->launches += 3;
    if (record) CK(cudaEventRecord(ln.ev[ST_ELEMENTS], st));
    if (what != LFB_LN_PRIOR) {
        if (!(flags & LFB_FLAG_SKIP_DISC)) {
            KREC(LFB_K_ELEM_DISC);
            elements_kernel<1><<<blocks(njobs, G.n_disc_half, true), kElemThreads, 0, st>>>(E);
            KSYNC("elements_kernel");
            h->launches++;
        }
        if (!(flags & LFB_FLAG_SKIP_WD)) {
            CK(cudaStreamWaitEvent(st, ln.wd_ev, 0));  // the white-dwarf centre's lines of sight (side stream)
            KREC(LFB_K_ELEM_WD);
            elements_kernel<0><<<blocks(n, G.n_wd_half), kElemThreads, 0, st>>>(E);
            KSYNC("elements_kernel");
            h->launches++;
        }
        if (record) CK(cudaEventRecord(ln.ev[ST_FLUX], st));
        // everything of the flux preparation that does not need the strip goes before the join with the
        // stream ODE, so that the main stream has work while the ODE finishes
        KREC(LFB_K_PREP);
        prep_kernel<<<(unsigned)((njobs + 3) / 4), 128, 0, st>>>(A);
        KSYNC("prep_kernel");
        const long long per_job0 = ((G.n_wd_half + G.n_disc_half) + 31) & ~31;
        const long long per_job1 = (G.n_bs + 31) & ~31;
        KREC(LFB_K_POSITIONS);
        positions_kernel<0><<<(unsigned)((njobs * per_job0 + 127) / 128), 128, 0, st>>>(A);
        KSYNC("positions_kernel");
        CK(cudaStreamWaitEvent(st, ln.join_ev, 0));  // join: the strip needs the impact point
        KREC(LFB_K_ELEM_BS);  // includes any wait for the stream ODE on the side stream
        veto_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(what, L.n_ecl, n, ln.ws.as<WalkerScal>(), ln.js.as<JobScal>());
        KSYNC("veto_kernel");
        h->launches++;
        if (!(flags & LFB_FLAG_SKIP_BS)) {
            elements_kernel<2><<<blocks(njobs, G.n_bs), kElemThreads, 0, st>>>(E);
            KSYNC("elements_kernel");
            h->launches++;
        }
        KREC(LFB_K_PREP_BS);
        prep_strip_kernel<<<(unsigned)((njobs + 3) / 4), 128, 0, st>>>(A);
        KSYNC("prep_strip_kernel");
        positions_kernel<1><<<(unsigned)((njobs * per_job1 + 127) / 128), 128, 0, st>>>(A);
        KSYNC("positions_kernel");
        h->launches += 2;
        const FluxShape fs = flux_shape(h, mode);
        const size_t smem = (size_t)fs.NT * fs.RP * (mode ? 32 : 16) + h->flux_smem_pad;
        const dim3 fgrid((unsigned)njobs);
        CK(cudaStreamWaitEvent(st, ln.don_ev, 0));  // the donor tables (second side stream)
        KREC(LFB_K_FLUX);
#define LFB_LAUNCH_FLUX(MODE, NT, RP, CTAS)                                                                               \
    do {                                                                                                                  \
        CK(cudaFuncSetAttribute(flux_kernel<MODE, NT, RP, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        flux_kernel<MODE, NT, RP, CTAS><<<fgrid, NT, smem, st>>>(A);                                                      \
    } while (0)
        if (mode) LFB_LAUNCH_FLUX(1, 256, 6, 4);
        else if (fs.NT == 512) LFB_LAUNCH_FLUX(0, 512, 13, 2);
        else if (fs.NT == 128) LFB_LAUNCH_FLUX(0, 128, 13, 8);
        else LFB_LAUNCH_FLUX(0, 256, 13, 4);
#undef LFB_LAUNCH_FLUX
        KSYNC("flux_kernel");
        h->launches += 3;
        if (gp) {
            KREC(LFB_K_GP);
            gp_kernel<<<dim3((unsigned)((n + kGpWalkers - 1) / kGpWalkers), (unsigned)L.n_ecl), kGpThreads, 0, st>>>(A);
            KSYNC("gp_kernel");
            h->launches++;
        }
    } else {
        CK(cudaStreamWaitEvent(st, ln.join_ev, 0));
        veto_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(what, L.n_ecl, n, ln.ws.as<WalkerScal>(), ln.js.as<JobScal>());
        KSYNC("veto_kernel");
        h->launches++;
        if (record) CK(cudaEventRecord(ln.ev[ST_FLUX], st));
    }
    CK(cudaStreamWaitEvent(st, ln.don_ev, 0));  // the second side stream always rejoins (stream capture needs it)
    if (record) CK(cudaEventRecord(ln.ev[ST_FINISH], st));
    KREC(LFB_K_FINISH);
    if (d_out || d_chi) {
        finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(what, L.n_ecl, n, ln.ws.as<WalkerScal>(),
                                                                  ln.chi_part.as<double>(), d_chi, d_out);
        KSYNC("finish_kernel");
        h->launches++;
    }
    if (record) {
        CK(cudaEventRecord(ln.ev[ST_COUNT], st));
    }
    KREC(LFB_K_COUNT);
#undef KREC
#undef KSYNC
    CK(cudaGetLastError());
    return LFB_OK;
}

extern "C" {

int lfb_create(int device, const lfb_config* cfg_in, lfb_handle** out)
{
    if (!out) return LFB_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                         " (this engine has no CPU fallback)";
        cudaGetLastError();
        return LFB_ECUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "device index out of range";
        return LFB_EINVAL;
    }
    lfb_config c{};
    if (cfg_in) c = *cfg_in;
    if (c.n_wd_rings <= 0) c.n_wd_rings = 10;
    if (c.n_disc_r <= 0) c.n_disc_r = 25;
    if (c.n_disc_th <= 0) c.n_disc_th = 40;
    if (c.n_bs <= 0) c.n_bs = 200;
    if (c.n_donor_th <= 0) c.n_donor_th = 18;
    if (c.n_quad <= 0) c.n_quad = 3;
    if (!(c.donor_ulimb != 0.0)) c.donor_ulimb = 0.8;
    if (!(c.donor_gdexp != 0.0)) c.donor_gdexp = 0.32;
    if ((c.n_disc_th & 1) || !(c.n_quad & 1) || c.n_quad > kMaxQuad || c.n_donor_th > kMaxDonorRings ||
        c.n_bs < 2 || c.n_wd_rings > 256 || c.n_disc_r > 4096) {
        g_create_error = "bad grid configuration (n_disc_th even, n_quad odd <= 15, n_donor_th <= 128, n_bs >= 2)";
        return LFB_EINVAL;
    }
    lfb_handle* h = new lfb_handle();
    h->device = device;
    h->cfg = c;
    auto bail = [&](const char* what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        delete h;
        return LFB_ECUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (int i = 0; i < kLanes; ++i)
        if ((e = h->lanes[i].create()) != cudaSuccess) return bail("lane streams/events", e);
    if ((e = cudaEventCreateWithFlags(&h->enter_ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->t0_ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->t1_ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    GridCfg& G = h->grid;
    G.n_wd_rings = c.n_wd_rings;
    G.n_wd = 4 * c.n_wd_rings * c.n_wd_rings;
    G.n_wd_half = G.n_wd / 2;
    G.n_disc_r = c.n_disc_r;
    G.n_disc_th = c.n_disc_th;
    G.n_disc = c.n_disc_r * c.n_disc_th;
    G.n_disc_half = G.n_disc / 2;
    G.n_bs = c.n_bs;
    G.n_donor_th = c.n_donor_th;
    G.n_quad = c.n_quad;
    G.donor_ulimb = c.donor_ulimb;
    G.donor_gdexp = c.donor_gdexp;
    std::vector<int> off(c.n_donor_th + 1, 0);
    for (int k = 0; k < c.n_donor_th; ++k) off[k + 1] = off[k] + donor_ring_count(c.n_donor_th, k);
    G.n_donor_q = off[c.n_donor_th];
    if (h->donor_off.reserve(off.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(h->donor_off.p, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
        return bail("donor ring table", cudaGetLastError());
    G.donor_ring_off = h->donor_off.as<int>();
    // disc elements in order of their distance from the donor (for a disc of typical size, 0.35 a):
    // elements next to each other in this order are eclipsed in the same way (deeply / barely /
    // never), which keeps the 32 threads of a warp on the same branch of the solver
    {
        const int hth = c.n_disc_th / 2;
        std::vector<int> order(G.n_disc_half);
        std::iota(order.begin(), order.end(), 0);
        const double rtyp = 0.35;
        auto key = [&](int tile) {
            int m = tile / hth, j = tile % hth;
            const double rho = rtyp * (m + 0.5) / c.n_disc_r, ca = cos((j + 0.5) * kTwoPi / c.n_disc_th);
            return 2.0 * rho * ca - rho * rho;  // 1 - (distance to the donor)^2
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) > key(b); });
        if (h->disc_order.reserve(order.size() * sizeof(int)) != cudaSuccess ||
            cudaMemcpy(h->disc_order.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail("disc order table", cudaGetLastError());
        G.disc_order = h->disc_order.as<int>();
        // what an element kernel thread needs to place its tile, looked up instead of worked out per thread:
        // disc (in disc_order) cos and sin of the sector's azimuth, the ring's place between rwd and rdisc,
        // the tile's number; white dwarf (half disc) the tile's offset on the sky in units of its radius
        std::vector<double> geo((size_t)G.n_disc_half * 4 + (size_t)G.n_wd_half * 2);
        for (int t = 0; t < G.n_disc_half; ++t) {
            const int tile = order[t], m = tile / hth, j = tile % hth;
            const double a = (j + 0.5) * kTwoPi / c.n_disc_th;
            geo[4 * (size_t)t] = cos(a);
            geo[4 * (size_t)t + 1] = sin(a);
            geo[4 * (size_t)t + 2] = (m + 0.5) / c.n_disc_r;
            geo[4 * (size_t)t + 3] = (double)tile;
        }
        double* wdg = geo.data() + (size_t)G.n_disc_half * 4;
        for (int t = 0; t < G.n_wd_half; ++t) {
            int k = (int)sqrt(0.5 * (double)t);
            while (2 * k * k > t) --k;
            while (2 * (k + 1) * (k + 1) <= t) ++k;
            const int r = t - 2 * k * k, q1 = 2 * k + 1, nk = 4 * q1, j = r < q1 ? r : r + 2 * q1;
            const double inv = 1.0 / c.n_wd_rings, ra = k * inv, rb = (k + 1) * inv;
            const double rho = sqrt(0.5 * (ra * ra + rb * rb)), a = (j + 0.5) * kTwoPi / nk;
            wdg[2 * (size_t)t] = rho * cos(a);
            wdg[2 * (size_t)t + 1] = rho * sin(a);
        }
        if (h->tile_geo.reserve(geo.size() * sizeof(double)) != cudaSuccess ||
            cudaMemcpy(h->tile_geo.p, geo.data(), geo.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail("tile geometry table", cudaGetLastError());
        G.disc_geo = (const double4*)h->tile_geo.p;
        G.wd_geo = (const double2*)(h->tile_geo.as<double>() + (size_t)G.n_disc_half * 4);
    }
    // where each tile record (mirrors included) finds its weight in a job's weight table
    // [white-dwarf rings | disc rings | strip elements]
    {
        std::vector<unsigned short> widx((size_t)G.n_wd + G.n_disc + G.n_bs);
        for (int i = 0; i < G.n_wd; ++i) {
            int t = i >> 1, k = (int)sqrt(0.5 * (double)t);
            while (2 * k * k > t) --k;
            while (2 * (k + 1) * (k + 1) <= t) ++k;
            widx[i] = (unsigned short)k;
        }
        for (int i = 0; i < G.n_disc; ++i) widx[G.n_wd + i] = (unsigned short)(G.n_wd_rings + (i >> 1) / (G.n_disc_th / 2));
        for (int i = 0; i < G.n_bs; ++i) widx[G.n_wd + G.n_disc + i] = (unsigned short)(G.n_wd_rings + G.n_disc_r + i);
        if (G.n_wd_rings + G.n_disc_r + G.n_bs > 65535) {
            g_create_error = "surface grid too dense (weight table index)";
            delete h;
            return LFB_EINVAL;
        }
        // Records are stored shuffled: slot s holds tile record (s * stride) mod n, stride co-prime with n.
        // Neighbouring tiles are eclipsed at neighbouring samples; this way the lanes of a flux-kernel
        // warp hold distant tiles and their shared-memory atomics rarely hit the same sample.
        const long long nrec = (long long)widx.size();
        long long stride = nrec > 4096 ? 521 : 67;
        if (const char* env = getenv("LFB_REC_STRIDE")) stride = std::max(1, atoi(env));
        auto gcd = [](long long a, long long b) {
            while (b) {
                long long r = a % b;
                a = b;
                b = r;
            }
            return a;
        };
        while (nrec > 1 && gcd(stride, nrec) != 1) ++stride;
        std::vector<int> slot(widx.size());
        std::vector<unsigned short> widx_slot(widx.size());
        for (long long s = 0; s < nrec; ++s) {
            const long long i = (s * stride) % nrec;
            slot[(size_t)i] = (int)s;
            widx_slot[(size_t)s] = widx[(size_t)i];
        }
        if (h->rec_widx.reserve(widx.size() * sizeof(unsigned short)) != cudaSuccess ||
            cudaMemcpy(h->rec_widx.p, widx_slot.data(), widx.size() * sizeof(unsigned short), cudaMemcpyHostToDevice) != cudaSuccess ||
            h->rec_slot.reserve(slot.size() * sizeof(int)) != cudaSuccess ||
            cudaMemcpy(h->rec_slot.p, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail("record weight index table", cudaGetLastError());
        G.rec_widx = h->rec_widx.as<unsigned short>();
        G.rec_slot = h->rec_slot.as<int>();
    }
    // composite Simpson nodes on [-1, 1] (exposure = phase +- width, CVModel.py:64)
    if (c.n_quad == 1) {
        G.quad_off[0] = 0.0;
        G.quad_w[0] = 1.0;
    } else {
        int nint = c.n_quad - 1;
        for (int k = 0; k < c.n_quad; ++k) {
            G.quad_off[k] = -1.0 + 2.0 * k / nint;
            double cw = (k == 0 || k == nint) ? 1.0 : ((k & 1) ? 4.0 : 2.0);
            G.quad_w[k] = cw / (3.0 * nint);
        }
    }
    if (const char* env = getenv("LFB_FLUX_VARIANT")) {
        int v = atoi(env);
        if (v >= 0 && v <= 2) h->flux_variant = v;
    }
    if (const char* env = getenv("LFB_FLUX_SMEM_PAD")) h->flux_smem_pad = (size_t)std::max(0, atoi(env));
    if (const char* env = getenv("LFB_DONOR_SMEM_PAD")) h->donor_smem_pad = (size_t)std::max(0, atoi(env));
    if (const char* env = getenv("LFB_LANES")) {
        int v = atoi(env);
        if (v >= 1 && v <= kLanes) h->n_lanes = v;
    }
    if (const char* env = getenv("LFB_STREAM_LANES_BELOW")) h->stream_lanes_below = atoll(env);
    if (const char* env = getenv("LFB_GRAPHS")) h->graphs_on = atoi(env) != 0;
    if (const char* env = getenv("LFB_GRAPH_MAX_JOBS")) h->graph_max_jobs = atoll(env);
    if (const char* env = getenv("LFB_DEBUG_SYNC")) {
        h->debug_sync = atoi(env) != 0;
        if (h->debug_sync) h->graphs_on = false;
    }
    *out = h;
    return LFB_OK;
}

// ---- exchange over NVLink peer memory (peer.cuh) ----
void lfb_peer_destroy(lfb_handle* h)
{
    if (!h || h->peer.world == 0) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < h->peer.world; ++p)
        if (p != h->peer.rank && h->peer.win[p]) cudaIpcCloseMemHandle(h->peer.win[p]);
    if (h->peer.rank >= 0 && h->peer.win[h->peer.rank]) cudaFree(h->peer.win[h->peer.rank]);
    if (h->peer.ticket) cudaFree(h->peer.ticket);
    h->peer = lfb_handle::Peer();
}

int lfb_peer_create(lfb_handle* h, int rank, int world, long long slot_bytes, unsigned char handle_out[64])
{
    if (!h || !handle_out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || slot_bytes <= 0 || (slot_bytes & 7))
        return h ? fail(h, LFB_EINVAL, "peer_create: need 1 <= world <= 16, 0 <= rank < world, slot_bytes a positive multiple of 8") : LFB_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the IPC handle is what the header promises");
    CK(cudaSetDevice(h->device));
    lfb_peer_destroy(h);
    lfb_handle::Peer& P = h->peer;
    P.bytes = (size_t)kPeerHeader + 2 * (size_t)world * (size_t)slot_bytes;
    void* w = nullptr;
    CK(cudaMalloc(&w, P.bytes));
    cudaError_t e = cudaMemset(w, 0, P.bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&P.ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(P.ticket, 0, sizeof(unsigned int));
    cudaIpcMemHandle_t ipc;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&ipc, w);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(w);
        if (P.ticket) cudaFree(P.ticket);
        P = lfb_handle::Peer();
        return fail(h, LFB_ECUDA, std::string("peer_create: ") + cudaGetErrorString(e));
    }
    P.rank = rank;
    P.world = world;
    P.slot_bytes = slot_bytes;
    P.win[rank] = (unsigned char*)w;
    memcpy(handle_out, &ipc, 64);
    return LFB_OK;
}

int lfb_peer_connect(lfb_handle* h, const unsigned char* handles)
{
    if (!h || !handles || h->peer.world == 0) return h ? fail(h, LFB_ESTATE, "peer_connect: call peer_create first") : LFB_EINVAL;
    CK(cudaSetDevice(h->device));
    lfb_handle::Peer& P = h->peer;
    for (int p = 0; p < P.world; ++p) {
        if (p == P.rank || P.win[p]) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + 64 * (size_t)p, 64);
        void* w = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&w, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(h, LFB_ECUDA, std::string("peer_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        }
        P.win[p] = (unsigned char*)w;
    }
    P.connected = true;
    return LFB_OK;
}

int lfb_peer_allgather(lfb_handle* h, const double* a, int ca, const double* b, int cb, long long rows,
                       const double** gathered, void* stream_v)
{
    if (!h) return LFB_EINVAL;
    lfb_handle::Peer& P = h->peer;
    if (!P.connected) return fail(h, LFB_ESTATE, "peer_allgather: call peer_create and peer_connect first");
    if (!a || ca < 1 || cb < 0 || (cb && !b) || rows < 0 || rows * (long long)(ca + cb) * 8 > P.slot_bytes)
        return fail(h, LFB_EINVAL, "peer_allgather: rows do not fit the slot");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : h->stream;
    PeerArgs A;
    A.a = a;
    A.b = b;
    A.ca = ca;
    A.cb = cb;
    A.rows = rows;
    for (int p = 0; p < kMaxPeers; ++p) A.win[p] = P.win[p];
    A.rank = P.rank;
    A.world = P.world;
    A.step = ++P.step;
    A.buf = (int)(A.step & 1);
    A.slot_bytes = P.slot_bytes;
    A.ticket = P.ticket;
    A.spin_limit = 60000000000LL;  // ~30 s of SM clocks: a rank that died must not hang its peers for ever
    const long long total = rows * (ca + cb);
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count);
    if (blocks < 1) blocks = 1;
    peer_allgather_kernel<<<blocks, 256, 0, st>>>(A);
    CK(cudaGetLastError());
    h->launches++;
    if (gathered) *gathered = (const double*)(P.win[P.rank] + kPeerHeader + (size_t)A.buf * P.world * P.slot_bytes);
    return LFB_OK;
}

int lfb_peer_status(lfb_handle* h, int* timed_out)
{
    if (!h || !timed_out || h->peer.world == 0) return h ? fail(h, LFB_ESTATE, "peer_status: no window") : LFB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    unsigned int v = 0;
    CK(cudaMemcpy(&v, h->peer.win[h->peer.rank] + kMaxPeers * 8, sizeof(v), cudaMemcpyDeviceToHost));
    *timed_out = (int)v;
    return LFB_OK;
}

void lfb_destroy(lfb_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graphs(h);
    DevBuf* bufs[] = {&h->gather, &h->consts, &h->psrc, &h->ptype, &h->pisvar, &h->pp1, &h->pp2, &h->pnorm,
                      &h->donor_off, &h->disc_order, &h->rec_widx, &h->rec_slot, &h->tile_geo, &h->cf_gather, &h->cf_pars, &h->cf_tot, &h->cf_comp, &h->theta, &h->out,
                      &h->chisq, &h->h_in, &h->h_out, &h->h_chisq};
    for (DevBuf* b : bufs) b->release();
    for (DevBuf& b : h->scratch) b.release();
    h->gp_dist.release();
    h->lc.release();
    h->cf_lc.release();
    lfb_peer_destroy(h);
    for (int i = 0; i < kLanes; ++i) {
        if (h->lanes[i].st) cudaStreamSynchronize(h->lanes[i].st);
        h->lanes[i].destroy();
    }
    if (h->enter_ev) cudaEventDestroy(h->enter_ev);
    if (h->t0_ev) cudaEventDestroy(h->t0_ev);
    if (h->t1_ev) cudaEventDestroy(h->t1_ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* lfb_last_error(const lfb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lfb_get_config(const lfb_handle* h, lfb_config* out)
{
    if (!h || !out) return LFB_EINVAL;
    *out = h->cfg;
    return LFB_OK;
}

long long lfb_launch_count(const lfb_handle* h) { return h ? h->launches : 0; }

int lfb_last_stage_ms(lfb_handle* h, float out[6])
{
    if (!h || !out) return LFB_EINVAL;
    for (int i = 0; i < 6; ++i) out[i] = -1.0f;
    if (!h->ev_valid) return LFB_ESTATE;
    // stages of the last batch that ran on lane 0 (they overlap the other lane's work);
    // total = the whole call on the caller's stream
    Lane& ln = h->lanes[0];
    for (int i = 0; i < ST_COUNT && h->stages_valid; ++i)  // a replayed CUDA graph has no stage events: total only
        if (cudaEventElapsedTime(&out[i], ln.ev[i], ln.ev[i + 1]) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, LFB_ECUDA, "stage events not complete: synchronise the stream first");
        }
    if (cudaEventElapsedTime(&out[5], h->t0_ev, h->t1_ev) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, LFB_ECUDA, "stage events not complete: synchronise the stream first");
    }
    return LFB_OK;
}

int lfb_set_trace(lfb_handle* h, int on)
{
    if (!h) return LFB_EINVAL;
    drop_graphs(h);  // captured passes hold the old tables / buffers
    h->trace = on != 0;
    return LFB_OK;
}

int lfb_last_trace_ms(lfb_handle* h, float out[LFB_K_COUNT + 1])
{
    if (!h || !out) return LFB_EINVAL;
    for (int i = 0; i <= LFB_K_COUNT; ++i) out[i] = -1.0f;
    if (!h->ev_valid || !h->trace) return fail(h, LFB_ESTATE, "no trace: lfb_set_trace(h, 1), then lfb_log_prob");
    Lane& ln = h->lanes[0];
    // a kernel's slot runs from its own record to the next one that was recorded
    for (int i = 0; i < LFB_K_COUNT; ++i) {
        if (!ln.kev_set[i]) continue;
        int nx = i + 1;
        while (nx <= LFB_K_COUNT && !ln.kev_set[nx]) ++nx;
        if (nx > LFB_K_COUNT || cudaEventElapsedTime(&out[i], ln.kev[i], ln.kev[nx]) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, LFB_ECUDA, "trace events not complete: synchronise the stream first");
        }
    }
    if (cudaEventElapsedTime(&out[LFB_K_COUNT], ln.sev[0], ln.sev[1]) != cudaSuccess) {
        cudaGetLastError();
        out[LFB_K_COUNT] = -1.0f;
    }
    // the donor chain runs on the second side stream, beside the kernels above
    if (cudaEventElapsedTime(&out[LFB_K_ELEM_DONOR], ln.dev[0], ln.dev[1]) != cudaSuccess ||
        cudaEventElapsedTime(&out[LFB_K_DONOR_TABLE], ln.dev[1], ln.dev[2]) != cudaSuccess) {
        cudaGetLastError();
        out[LFB_K_ELEM_DONOR] = out[LFB_K_DONOR_TABLE] = -1.0f;
    }
    return LFB_OK;
}

float lfb_last_kernel_ms(lfb_handle* h)
{
    float t[6];
    if (lfb_last_stage_ms(h, t) != LFB_OK) return -1.0f;
    return t[5];
}

int lfb_set_layout(lfb_handle* h, int ndim, int n_ecl, int npars, const int* gather, int n_consts,
                   const double* consts)
{
    if (!h) return LFB_EINVAL;
    drop_graphs(h);  // captured passes hold the old tables / buffers
    if (ndim < 0 || n_ecl < 1 || (npars != 14 && npars != 18) || !gather || n_consts < 0 || (n_consts && !consts))
        return fail(h, LFB_EINVAL, "set_layout: need n_ecl >= 1, npars in {14, 18}, gather");
    for (int e = 0; e < n_ecl; ++e)
        for (int k = 0; k < npars; ++k) {
            int g = gather[e * LFB_NPAR + k];
            if (g >= ndim || (g < 0 && -g - 1 >= n_consts)) return fail(h, LFB_EINVAL, "set_layout: gather index out of range");
        }
    // q, dphi and rwd live on the root of the tree (LCModel.node_par_names, CVModel.py:434)
    for (int e = 1; e < n_ecl; ++e)
        if (gather[e * LFB_NPAR + P_Q] != gather[P_Q] || gather[e * LFB_NPAR + P_DPHI] != gather[P_DPHI] ||
            gather[e * LFB_NPAR + P_RWD] != gather[P_RWD])
            return fail(h, LFB_EINVAL, "set_layout: q, dphi and rwd must be shared by every eclipse");
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = upload(h, h->gather, gather, sizeof(int) * (size_t)n_ecl * LFB_NPAR))) return rc;
    if ((rc = upload(h, h->consts, consts, sizeof(double) * (size_t)n_consts))) return rc;
    h->n_consts = n_consts;
    h->gp_on = false;  // a new layout: the GP sources must be set again
    CK(cudaStreamSynchronize(h->stream));
    h->ndim = ndim;
    h->n_ecl = n_ecl;
    h->npars = npars;
    h->have_layout = true;
    h->have_lc = false;
    h->n_prior = 0;
    return LFB_OK;
}

int lfb_set_priors(lfb_handle* h, int n_prior, const int* src, const int* type, const double* p1, const double* p2,
                   const double* norm, const int* isvar)
{
    if (!h) return LFB_EINVAL;
    drop_graphs(h);  // captured passes hold the old tables / buffers
    if (!h->have_layout) return fail(h, LFB_ESTATE, "set_priors: call set_layout first");
    if (n_prior < 0 || (n_prior && (!src || !type || !p1 || !p2 || !norm || !isvar)))
        return fail(h, LFB_EINVAL, "set_priors: NULL array");
    for (int k = 0; k < n_prior; ++k) {
        if (src[k] >= h->ndim || src[k] < -h->n_consts)
            return fail(h, LFB_EINVAL, "set_priors: source column / constant slot out of range");
        if (type[k] < 0 || type[k] > LFB_PRIOR_MODJEFF) return fail(h, LFB_EINVAL, "set_priors: unknown prior type");
    }
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = upload(h, h->psrc, src, sizeof(int) * (size_t)n_prior))) return rc;
    if ((rc = upload(h, h->ptype, type, sizeof(int) * (size_t)n_prior))) return rc;
    if ((rc = upload(h, h->pisvar, isvar, sizeof(int) * (size_t)n_prior))) return rc;
    if ((rc = upload(h, h->pp1, p1, sizeof(double) * (size_t)n_prior))) return rc;
    if ((rc = upload(h, h->pp2, p2, sizeof(double) * (size_t)n_prior))) return rc;
    if ((rc = upload(h, h->pnorm, norm, sizeof(double) * (size_t)n_prior))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->n_prior = n_prior;
    return LFB_OK;
}

int lfb_set_lightcurves(lfb_handle* h, int n_ecl, const long long* off, const double* phase, const double* width,
                        const double* y, const double* ye)
{
    if (!h) return LFB_EINVAL;
    drop_graphs(h);  // captured passes hold the old tables / buffers
    if (!h->have_layout) return fail(h, LFB_ESTATE, "set_lightcurves: call set_layout first");
    if (n_ecl != h->n_ecl || !off || !phase || !width || !y || !ye)
        return fail(h, LFB_EINVAL, "set_lightcurves: n_ecl must match the layout; arrays must be non-NULL");
    if (off[0] != 0) return fail(h, LFB_EINVAL, "set_lightcurves: off[0] must be 0");
    for (int e = 0; e < n_ecl; ++e)
        if (off[e + 1] < off[e] || off[e + 1] - off[e] > 100000000LL) return fail(h, LFB_EINVAL, "set_lightcurves: offsets must ascend");
    CK(cudaSetDevice(h->device));
    const FluxShape fs = flux_shape(h, 0);
    int rc = build_samples(h, h->lc, fs.NT, fs.RP, n_ecl, off, phase, width, y, ye);
    if (rc) return rc;
    h->have_lc = true;
    if (h->gp_on && h->lc.max_gaps > kMaxGaps) {
        h->gp_on = false;
        return fail(h, LFB_EINVAL, "set_lightcurves: a light curve spans more than 8 orbital cycles; the GP likelihood "
                                   "holds at most 8 change-point gaps (GP switched off)");
    }
    return LFB_OK;
}

static DevLayout make_layout(lfb_handle* h)
{
    DevLayout L;
    L.ndim = h->ndim;
    L.n_ecl = h->n_ecl;
    L.npars = h->npars;
    L.n_prior = h->n_prior;
    L.gather = h->gather.as<int>();
    L.consts = h->consts.as<double>();
    L.psrc = h->psrc.as<int>();
    L.ptype = h->ptype.as<int>();
    L.pisvar = h->pisvar.as<int>();
    L.pp1 = h->pp1.as<double>();
    L.pp2 = h->pp2.as<double>();
    L.pnorm = h->pnorm.as<double>();
    return L;
}

}  // extern "C"

// One pass over n walkers (device pointers) enqueued on st with plain launches on lane 0: no CUDA graph of its
// own, no timing events, nothing allocated once the buffers have their sizes -- safe inside a stream capture.
static int enqueue_pass_inline(lfb_handle* h, int what, long long n, const double* d_theta, double* d_out, cudaStream_t st)
{
    if (what == LFB_LN_PRIOR) CK(h->lanes[0].chi_part.reserve(8));
    DevLayout L = make_layout(h);
    Lane& ln = h->lanes[0];
    cudaStream_t keep = ln.st;
    ln.st = st;
    int rc = LFB_OK;
    const long long per = std::max(1LL, h->max_jobs_per_batch / h->n_ecl);
    for (long long w0 = 0; w0 < n && rc == LFB_OK; w0 += per)
        rc = run_batch(h, ln, L, h->lc, what, 0, 0, std::min(per, n - w0), d_theta + w0 * h->ndim, d_out + w0, nullptr,
                       nullptr, nullptr, false);
    ln.st = keep;
    return rc;
}

extern "C" {

int lfb_log_prob(lfb_handle* h, int what, long long n, const double* theta, double* out, double* chisq_out,
                 void* stream_v)
{
    if (!h) return LFB_EINVAL;
    if (!h->have_layout) return fail(h, LFB_ESTATE, "log_prob: call set_layout first");
    if (what != LFB_LN_PRIOR && !h->have_lc) return fail(h, LFB_ESTATE, "log_prob: call set_lightcurves first");
    if (what < LFB_LN_PRIOR || what > LFB_LN_PROB || n < 0 || (n && (!theta || !out)))
        return fail(h, LFB_EINVAL, "log_prob: bad arguments");
    if (n == 0) return LFB_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : h->stream;
    const long long njobs = n * h->n_ecl;
    const size_t th_bytes = sizeof(double) * (size_t)n * (size_t)(h->ndim > 0 ? h->ndim : 1);
    const bool th_dev = is_device_ptr(theta), out_dev = is_device_ptr(out);
    const bool chi_dev = chisq_out && is_device_ptr(chisq_out);
    const bool th_pinned = !th_dev && is_pinned_host_ptr(theta), out_pinned = !out_dev && is_pinned_host_ptr(out);
    const double* d_theta = theta;
    if (!th_dev) {
        // the previous call may have returned with its copy out of h_in still in flight (host theta, device
        // outputs): t1_ev was recorded behind it
        if (h->h2d_pending) CK(cudaEventSynchronize(h->t1_ev));
        h->h2d_pending = false;
        CK(h->theta.reserve(th_bytes));
        if (!th_pinned) CK(h->h_in.reserve(th_bytes));
        d_theta = h->theta.as<double>();  // staged batch by batch below, so that a copy overlaps the batch before
    }
    double* d_out = out;
    if (!out_dev) {
        CK(h->out.reserve(sizeof(double) * (size_t)n));
        d_out = h->out.as<double>();
    }
    double* d_chi = chi_dev ? chisq_out : nullptr;
    if (!chi_dev) {
        CK(h->chisq.reserve(sizeof(double) * (size_t)njobs));
        d_chi = h->chisq.as<double>();
    }
    if (what == LFB_LN_PRIOR && chisq_out) CK(cudaMemsetAsync(d_chi, 0xff, sizeof(double) * (size_t)njobs, st));
    if (what == LFB_LN_PRIOR)
        for (int i = 0; i < kLanes; ++i) CK(h->lanes[i].chi_part.reserve(8));
    DevLayout L = make_layout(h);
    // Batches of walkers alternate between two lanes (own streams and buffers): one batch's
    // FP64-bound element solves overlap the other's flux stage, and the element / event buffers
    // stay bounded (~60 KB per job in flight).
    CK(cudaEventRecord(h->enter_ev, st));
    CK(cudaEventRecord(h->t0_ev, st));
    const long long njobs_all = n * h->n_ecl;
    long long nbatch = (njobs_all + h->max_jobs_per_batch - 1) / h->max_jobs_per_batch;
    // A call that fits one batch runs as ONE batch on one lane and, from its second identical occurrence on, as a
    // replayed CUDA graph: measured against two plain lanes of half the size each, 8 % faster at 2048 light curves,
    // 4.6 % at 4096, 2 % at 16 384 (C2; 4.5 % for C5's grid at 4096) -- the gaps between the pass's dependent
    // kernels, not the overlap of two lanes, are what is left to win.  Without graphs (LFB_GRAPHS=0, tracing) a
    // call of 1024 jobs or more is split over the lanes.
    const bool graph_eligible = h->graphs_on && !h->trace && nbatch == 1 && njobs_all <= h->graph_max_jobs;
    if (!graph_eligible && nbatch < h->n_lanes && njobs_all >= 1024) nbatch = h->n_lanes;
    if (h->n_lanes > 1 && nbatch > 1) nbatch = (nbatch + h->n_lanes - 1) / h->n_lanes * h->n_lanes;  // the lanes get equal shares
    const long long per = (n + nbatch - 1) / nbatch;
    int used = 0;
    long long b = 0;
    h->stages_valid = true;
    // A small ensemble is launch bound (a dozen dependent kernels of 10-30 us): the second identical call
    // is captured into a CUDA graph and later ones replay it.
    if (graph_eligible) {
        lfb_handle::GraphEntry* ge = nullptr;
        for (auto& g : h->graphs)
            if (g.what == what && g.n == n && g.theta == (const void*)d_theta && g.out == (const void*)d_out &&
                g.chi == (const void*)d_chi)
                ge = &g;
        if (!ge) {
            ge = &h->graphs[h->graph_next];
            h->graph_next = (h->graph_next + 1) % 4;
            if (ge->exec) cudaGraphExecDestroy(ge->exec);
            *ge = lfb_handle::GraphEntry();
            ge->what = what;
            ge->n = n;
            ge->theta = d_theta;
            ge->out = d_out;
            ge->chi = d_chi;
        }
        if (ge->exec && ge->generation != h->alloc_generation) {  // some buffer moved since the capture
            cudaGraphExecDestroy(ge->exec);
            ge->exec = nullptr;
            ge->seen = 0;  // this call runs plainly (and re-sizes the buffers), the next one captures again
        }
        ++ge->seen;
        if (ge->seen >= 2) {
            Lane& ln = h->lanes[0];
            CK(cudaStreamWaitEvent(ln.st, h->enter_ev, 0));
            if (!th_dev) {
                const void* src = theta;
                if (!th_pinned) {
                    memcpy(h->h_in.p, theta, sizeof(double) * (size_t)n * h->ndim);
                    src = h->h_in.p;
                }
                CK(cudaMemcpyAsync(h->theta.p, src, sizeof(double) * (size_t)n * h->ndim, cudaMemcpyHostToDevice, ln.st));
            }
            bool ok = ge->exec != nullptr;
            if (!ok) {
                // buffers have their sizes from the first call: nothing inside allocates while capturing
                const long long l0 = h->launches;
                cudaGraph_t graph = nullptr;
                if (cudaStreamBeginCapture(ln.st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    const int rc = run_batch(h, ln, L, h->lc, what, 0, 0, n, d_theta, d_out, d_chi, nullptr, nullptr, false);
                    const cudaError_t e = cudaStreamEndCapture(ln.st, &graph);
                    if (rc == LFB_OK && e == cudaSuccess && graph &&
                        cudaGraphInstantiate(&ge->exec, graph, 0) == cudaSuccess) {
                        ge->launches = h->launches - l0;
                        ge->generation = h->alloc_generation;
                        ok = true;
                    }
                    if (graph) cudaGraphDestroy(graph);
                }
                h->launches = l0;
                if (!ok) {
                    cudaGetLastError();
                    h->graphs_on = false;  // this driver / configuration cannot capture the pass: plain launches from now on
                    ge->exec = nullptr;
                }
            }
            if (ok) {
                CK(cudaGraphLaunch(ge->exec, ln.st));
                h->launches += ge->launches;
                h->stages_valid = false;
                CK(cudaEventRecord(ln.done_ev, ln.st));
                CK(cudaStreamWaitEvent(st, ln.done_ev, 0));
                goto finished;
            }
        }
    }
    {
    // one lane: run on the caller's own stream (no cross-stream hand-off)
    const bool inline_lane = h->n_lanes == 1;
    cudaStream_t lane0_st = h->lanes[0].st;
    if (inline_lane) h->lanes[0].st = st;
    for (long long w0 = 0; w0 < n; w0 += per, ++b) {
        Lane& ln = h->lanes[b % h->n_lanes];
        if (b < h->n_lanes && !inline_lane) {
            CK(cudaStreamWaitEvent(ln.st, h->enter_ev, 0));
            used = (int)b + 1;
        }
        const long long nb = std::min(per, n - w0);
        if (!th_dev) {
            const size_t off = (size_t)w0 * h->ndim, cnt = sizeof(double) * (size_t)nb * h->ndim;
            const double* src = theta + off;  // page-locked memory of the caller's: copied from as it stands
            if (!th_pinned) {
                memcpy((double*)h->h_in.p + off, theta + off, cnt);
                src = (const double*)h->h_in.p + off;
            }
            CK(cudaMemcpyAsync(h->theta.as<double>() + off, src, cnt, cudaMemcpyHostToDevice, ln.st));
        }
        const bool last_on_lane0 = (b % h->n_lanes) == 0 && w0 + (long long)h->n_lanes * per >= n;
        int rc = run_batch(h, ln, L, h->lc, what, 0, 0, nb, d_theta + w0 * h->ndim, d_out + w0, d_chi + w0 * h->n_ecl,
                           nullptr, nullptr, last_on_lane0);
        if (rc) {
            h->lanes[0].st = lane0_st;
            return rc;
        }
    }
    h->lanes[0].st = lane0_st;
    for (int i = 0; i < used; ++i) {
        CK(cudaEventRecord(h->lanes[i].done_ev, h->lanes[i].st));
        CK(cudaStreamWaitEvent(st, h->lanes[i].done_ev, 0));
    }
    }
finished:
    CK(cudaEventRecord(h->t1_ev, st));
    h->ev_valid = true;
    if (!out_dev || (chisq_out && !chi_dev)) {
        if (!out_pinned) CK(h->h_out.reserve(sizeof(double) * (size_t)n));
        if (!out_dev) CK(cudaMemcpyAsync(out_pinned ? (void*)out : h->h_out.p, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (chisq_out && !chi_dev) {
            CK(h->h_chisq.reserve(sizeof(double) * (size_t)njobs));
            CK(cudaMemcpyAsync(h->h_chisq.p, d_chi, sizeof(double) * (size_t)njobs, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        if (!out_dev && !out_pinned) memcpy(out, h->h_out.p, sizeof(double) * (size_t)n);
        if (chisq_out && !chi_dev) memcpy(chisq_out, h->h_chisq.p, sizeof(double) * (size_t)njobs);
    } else if (!th_dev) {
        // host theta, device outputs: the call returns with work in flight.  A copy out of the caller's own
        // page-locked buffer must be over before the caller may write there again: wait for it here
        if (th_pinned) CK(cudaEventSynchronize(h->t1_ev));
        else h->h2d_pending = true;
    }
    return LFB_OK;
}

int lfb_calc_flux(lfb_handle* h, long long n_sets, const double* pars, int npars, int flags, int n_ph,
                  const double* phase, const double* width, double* out_total, double* out_comp, void* stream_v)
{
    if (!h) return LFB_EINVAL;
    if (n_sets < 0 || (npars != 14 && npars != 18) || n_ph < 0 || (n_sets && !pars) || (n_ph && (!phase || !out_total)))
        return fail(h, LFB_EINVAL, "calc_flux: need npars in {14, 18} and non-NULL arrays");
    if (n_sets == 0 || n_ph == 0) return LFB_OK;
    if (is_device_ptr(phase) || (width && is_device_ptr(width)))
        return fail(h, LFB_EINVAL, "calc_flux: phase and width are host arrays");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : h->stream;
    std::vector<int> gather(LFB_NPAR, 0);
    for (int k = 0; k < LFB_NPAR; ++k) gather[k] = k < npars ? k : 0;
    long long off[2] = {0, n_ph};
    const FluxShape fs = flux_shape(h, 1);
    int rc = build_samples(h, h->cf_lc, fs.NT, fs.RP, 1, off, phase, width, nullptr, nullptr);
    if (rc) return rc;
    CK(h->cf_gather.reserve(sizeof(int) * LFB_NPAR));
    CK(cudaMemcpyAsync(h->cf_gather.p, gather.data(), sizeof(int) * LFB_NPAR, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    const double* d_pars = pars;
    if (!is_device_ptr(pars)) {
        CK(h->cf_pars.reserve(sizeof(double) * (size_t)n_sets * npars));
        CK(cudaMemcpyAsync(h->cf_pars.p, pars, sizeof(double) * (size_t)n_sets * npars, cudaMemcpyHostToDevice, st));
        d_pars = h->cf_pars.as<double>();
    }
    const size_t cur = sizeof(double) * (size_t)n_sets * n_ph;
    const bool tot_dev = is_device_ptr(out_total), comp_dev = out_comp && is_device_ptr(out_comp);
    double* d_tot = out_total;
    if (!tot_dev) {
        CK(h->cf_tot.reserve(cur));
        d_tot = h->cf_tot.as<double>();
    }
    double* d_comp = comp_dev ? out_comp : nullptr;
    if (out_comp && !comp_dev) {
        CK(h->cf_comp.reserve(4 * cur));
        d_comp = h->cf_comp.as<double>();
    }
    DevLayout L;
    memset(&L, 0, sizeof(L));
    L.ndim = npars;
    L.n_ecl = 1;
    L.npars = npars;
    L.n_prior = 0;
    L.gather = h->cf_gather.as<int>();
    Lane& ln = h->lanes[0];
    CK(cudaEventRecord(h->enter_ev, st));
    CK(cudaStreamWaitEvent(ln.st, h->enter_ev, 0));
    rc = run_batch(h, ln, L, h->cf_lc, LFB_LN_LIKE, flags, 1, n_sets, d_pars, nullptr, nullptr, d_tot, d_comp, false);
    if (rc) return rc;
    CK(cudaEventRecord(ln.done_ev, ln.st));
    CK(cudaStreamWaitEvent(st, ln.done_ev, 0));
    if (!tot_dev) CK(cudaMemcpyAsync(out_total, d_tot, cur, cudaMemcpyDeviceToHost, st));
    if (out_comp && !comp_dev) CK(cudaMemcpyAsync(out_comp, d_comp, 4 * cur, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LFB_OK;
}

int lfb_set_gp(lfb_handle* h, int enabled, const int gp_src[3], const double* dist_cp)
{
    if (!h) return LFB_EINVAL;
    drop_graphs(h);  // captured passes hold the old tables / buffers
    if (!enabled) {
        h->gp_on = false;
        return LFB_OK;
    }
    if (!h->have_layout || !gp_src || !dist_cp) return fail(h, LFB_ESTATE, "set_gp: set_layout first; need gp_src[3] and dist_cp[n_ecl]");
    for (int i = 0; i < 3; ++i)
        if (gp_src[i] >= h->ndim || gp_src[i] < -h->n_consts)
            return fail(h, LFB_EINVAL, "set_gp: hyper-parameter source out of range");
    for (int e = 0; e < h->n_ecl; ++e)
        if (!(dist_cp[e] > 0.0) || !(dist_cp[e] < 0.5))
            return fail(h, LFB_EINVAL, "set_gp: dist_cp must lie in (0, 0.5): no change points otherwise");
    if (h->have_lc && h->lc.max_gaps > kMaxGaps)
        return fail(h, LFB_EINVAL, "set_gp: a light curve spans more than 8 orbital cycles; the GP likelihood holds at "
                                   "most 8 change-point gaps");
    for (int i = 0; i < 3; ++i) h->gp_src[i] = gp_src[i];
    CK(cudaSetDevice(h->device));
    int rc = upload(h, h->gp_dist, dist_cp, sizeof(double) * (size_t)h->n_ecl);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->gp_on = true;
    return LFB_OK;
}

int lfb_gp_loglike(lfb_handle* h, long long n_sets, int n, const double* x, const double* ye, const double* resid,
                   const double* hyper, int n_gaps, const double* gaps, double* out)
{
    if (!h) return LFB_EINVAL;
    if (n_sets < 0 || n < 0 || n_gaps < 0 || n_gaps > kMaxGaps || (n_sets && (!hyper || !out)) ||
        (n_sets && n && (!x || !ye || !resid)) || (n_sets && n_gaps && !gaps))
        return fail(h, LFB_EINVAL, "gp_loglike: bad arguments (at most 8 gaps)");
    if (n_sets == 0) return LFB_OK;
    for (int k = 1; k < n; ++k)
        if (!(x[k] >= x[k - 1])) return fail(h, LFB_EINVAL, "gp_loglike: x must ascend");
    CK(cudaSetDevice(h->device));
    DevBuf &dx = h->scratch[0], &dye = h->scratch[1], &dr = h->scratch[2], &dh = h->scratch[3], &dg = h->scratch[4],
           &dout = h->scratch[5];
    const size_t nn = (size_t)std::max(n, 1), ng = (size_t)std::max(n_gaps, 1);
    CK(dx.reserve(8 * nn));
    CK(dye.reserve(8 * nn));
    CK(dr.reserve(8 * nn * (size_t)n_sets));
    CK(dh.reserve(24 * (size_t)n_sets));
    CK(dg.reserve(16 * ng * (size_t)n_sets));
    CK(dout.reserve(8 * (size_t)n_sets));
    cudaStream_t st = h->stream;
    if (n) {
        CK(cudaMemcpyAsync(dx.p, x, 8 * (size_t)n, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(dye.p, ye, 8 * (size_t)n, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(dr.p, resid, 8 * (size_t)n * (size_t)n_sets, cudaMemcpyDefault, st));
    }
    CK(cudaMemcpyAsync(dh.p, hyper, 24 * (size_t)n_sets, cudaMemcpyDefault, st));
    if (n_gaps) CK(cudaMemcpyAsync(dg.p, gaps, 16 * (size_t)n_gaps * (size_t)n_sets, cudaMemcpyDefault, st));
    gp_batch_kernel<<<(unsigned)((n_sets + 63) / 64), 64, 0, st>>>(n_sets, n, dx.as<double>(), dye.as<double>(), dr.as<double>(),
                                                                    dh.as<double>(), n_gaps, dg.as<double>(), dout.as<double>());
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaMemcpyAsync(out, dout.p, 8 * (size_t)n_sets, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    return LFB_OK;
}

int lfb_wdphases(lfb_handle* h, long long n, const double* q, const double* incl_deg, const double* r1, int ntheta,
                 double* out, int* ok)
{
    if (!h) return LFB_EINVAL;
    if (n < 0 || ntheta < 1 || (n && (!q || !incl_deg || !r1 || !out || !ok)))
        return fail(h, LFB_EINVAL, "wdphases: bad arguments");
    if (n == 0) return LFB_OK;
    CK(cudaSetDevice(h->device));
    DevBuf &dq = h->scratch[0], &di = h->scratch[1], &dr = h->scratch[2], &dout = h->scratch[3], &dok = h->scratch[4];
    CK(dq.reserve(8 * (size_t)n));
    CK(di.reserve(8 * (size_t)n));
    CK(dr.reserve(8 * (size_t)n));
    CK(dout.reserve(16 * (size_t)n));
    CK(dok.reserve(4 * (size_t)n));
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(dq.p, q, 8 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(di.p, incl_deg, 8 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(dr.p, r1, 8 * (size_t)n, cudaMemcpyDefault, st));
    wdphases_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(n, dq.as<double>(), di.as<double>(), dr.as<double>(), ntheta,
                                                              dout.as<double>(), dok.as<int>());
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaMemcpyAsync(out, dout.p, 16 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(ok, dok.p, 4 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    return LFB_OK;
}

int lfb_ingress_egress(lfb_handle* h, long long n, const double* q, const double* incl_deg, const double* pts, double* out,
                       int* ok)
{
    if (!h) return LFB_EINVAL;
    if (n < 0 || (n && (!q || !incl_deg || !pts || !out || !ok))) return fail(h, LFB_EINVAL, "ingress_egress: bad arguments");
    if (n == 0) return LFB_OK;
    CK(cudaSetDevice(h->device));
    DevBuf &dq = h->scratch[0], &di = h->scratch[1], &dp = h->scratch[2], &dout = h->scratch[3], &dok = h->scratch[4];
    CK(dq.reserve(8 * (size_t)n));
    CK(di.reserve(8 * (size_t)n));
    CK(dp.reserve(40 * (size_t)n));
    CK(dout.reserve(16 * (size_t)n));
    CK(dok.reserve(4 * (size_t)n));
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(dq.p, q, 8 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(di.p, incl_deg, 8 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(dp.p, pts, 40 * (size_t)n, cudaMemcpyDefault, st));
    ingress_egress_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, dq.as<double>(), di.as<double>(), dp.as<double>(),
                                                                      dout.as<double>(), dok.as<int>());
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaMemcpyAsync(out, dout.p, 16 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(ok, dok.p, 4 * (size_t)n, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    return LFB_OK;
}

int lfb_roche(lfb_handle* h, int which, long long n, const double* a, const double* b, double* out, int* ok)
{
    if (!h) return LFB_EINVAL;
    if (which < LFB_ROCHE_XL1 || which > LFB_ROCHE_ANGLE || n < 0 || (n && (!a || !out || !ok)) ||
        (n && which != LFB_ROCHE_XL1 && !b))
        return fail(h, LFB_EINVAL, "roche: bad arguments");
    if (n == 0) return LFB_OK;
    CK(cudaSetDevice(h->device));
    DevBuf &da = h->scratch[0], &db = h->scratch[1], &dout = h->scratch[2], &dok = h->scratch[3];
    CK(da.reserve(sizeof(double) * n));
    CK(db.reserve(sizeof(double) * n));
    CK(dout.reserve(sizeof(double) * 4 * n));
    CK(dok.reserve(sizeof(int) * n));
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(da.p, a, sizeof(double) * n, cudaMemcpyDefault, st));
    if (b) CK(cudaMemcpyAsync(db.p, b, sizeof(double) * n, cudaMemcpyDefault, st));
    else CK(cudaMemsetAsync(db.p, 0, sizeof(double) * n, st));
    roche_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(which, n, da.as<double>(), db.as<double>(), dout.as<double>(),
                                                           dok.as<int>());
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaMemcpyAsync(out, dout.p, sizeof(double) * 4 * n, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(ok, dok.p, sizeof(int) * n, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    return LFB_OK;
}

// Diagnostics: element solves that fell through to the scan + golden-section + bisection solver since the
// library was loaded (all handles of this device).
long long lfb_robust_calls(lfb_handle* h)
{
    if (!h) return -1;
    unsigned long long v = 0;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpyFromSymbol(&v, g_robust_calls, sizeof(v)) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (long long)v;
}

int lfb_measure_fp64_peak(lfb_handle* h, int iters, double* tflops)
{
    if (!h || !tflops || iters <= 0) return LFB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(h->out.reserve(64));
    const int blocks = h->sm_count * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, h->stream);
        fp64_peak_kernel<<<blocks, 256, 0, h->stream>>>(iters, 0.999999, 1e-7, h->out.as<double>());
        cudaEventRecord(e1, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return fail(h, LFB_ECUDA, cudaGetErrorString(e));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 16.0 * iters * 256.0 * blocks / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->launches += 5;
    *tflops = best;
    return LFB_OK;
}

}  // extern "C"

#include "sampler.cuh"
