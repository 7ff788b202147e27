// sampler.cuh -- the affine-invariant stretch move with the ensemble resident in HBM, and the chain writer.
//
// Reference callers replaced (file:line under /root/reference):
//   emcee.EnsembleSampler(nwalkers, npars, ln_prob, args=(model,), pool=pool)     mcmcfit.py:283-288
//   sampler.sample(...) loops of run_burnin / run_mcmc_save                        mcmc_utils.py:114-183
//   the chain file, "{k:4d} {pos...} {lnprob:f}" per walker per step               mcmc_utils.py:157-164
// emcee itself is third-party and not vendored; the move is the published one (Goodman & Weare 2010;
// emcee's StretchMove): for walker k of one half S with a partner j drawn from the other half C,
//   z = ((a - 1) u + 1)^2 / a,  Y = X_j - z (X_j - X_k),  accept if (ndim - 1) ln z + lnp(Y) - lnp(X_k) > ln u'.
// Halves are the first and the second half of the ensemble (emcee 2.x; requirements.txt:10 asks for
// emcee >= 2.2.1), updated one after the other.
//
// Random numbers are Philox4x32-10 keyed by the seed and counted by (step, half, walker): any rank can
// draw any walker's numbers, so an ensemble sharded over N GPUs follows the 1-GPU chain bit for bit.
// Included at the end of lfit_cabi.cu (it needs the handle's internals).
#pragma once

namespace lfb {

struct SamplerState {
    unsigned long long seed;
    unsigned long long step;  // full steps done: the Philox counter
    unsigned long long rec;   // steps recorded into the chain buffer since it was last flushed
};

__host__ __device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                                        unsigned k1, unsigned out[4])
{
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform on [0, 1) with 53 random bits
__host__ __device__ __forceinline__ double u01(unsigned hi, unsigned lo)
{
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

// The draws of walker `row` of half `hsel` (temperature `temp`) at step `step`: z, the partner's row in the
// other half, ln u' of the acceptance test.
struct StretchDraw {
    double z, lnu;
    long long partner;
};
__device__ __forceinline__ StretchDraw stretch_draw(unsigned long long seed, unsigned long long step, int hsel, int temp,
                                                    long long row, long long half_n, double a)
{
    unsigned r[4], q[4];
    const unsigned tag = (unsigned)hsel | ((unsigned)temp << 1);
    philox4x32_10((unsigned)row, (unsigned)step, (unsigned)(step >> 32), tag, (unsigned)seed, (unsigned)(seed >> 32), r);
    philox4x32_10((unsigned)row, (unsigned)step, (unsigned)(step >> 32), tag | 0x80000000u, (unsigned)seed,
                  (unsigned)(seed >> 32), q);
    StretchDraw d;
    // (no fused multiply-add: the numbers are those of emcee's numpy expression, bit for bit)
    const double t = __dadd_rn(__dmul_rn(a - 1.0, u01(r[0], r[1])), 1.0);
    d.z = __dmul_rn(t, t) / a;
    d.partner = (long long)(((unsigned long long)r[2] * (unsigned long long)half_n) >> 32);
    d.lnu = log(u01(q[0], q[1]));
    return d;
}

// Proposals of rows [lo, lo + cnt) of half hsel: thread per (row, dimension).
__global__ void stretch_propose_kernel(const SamplerState* __restrict__ S, long long half_n, int ndim, int hsel, long long lo,
                                       long long cnt, double a, const double* __restrict__ pos, double* __restrict__ prop,
                                       double* __restrict__ zf, double* __restrict__ lnu)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cnt * ndim) return;
    const long long row = idx / ndim;
    const int d = (int)(idx - row * ndim);
    const long long r = lo + row;
    const StretchDraw D = stretch_draw(S->seed, S->step, hsel, 0, r, half_n, a);
    const long long k = (long long)hsel * half_n + r, j = (long long)(1 - hsel) * half_n + D.partner;
    const double xs = pos[k * ndim + d], xc = pos[j * ndim + d];
    prop[idx] = __dsub_rn(xc, __dmul_rn(__dsub_rn(xc, xs), D.z));  // c[rint] - (c[rint] - s) * zz, unfused
    if (d == 0) {
        zf[row] = (ndim - 1.0) * log(D.z);
        lnu[row] = D.lnu;
    }
}

// Accept / reject rows [lo, lo + cnt) of half hsel: warp per row (lane 0 decides before any lane writes).
// PACKED: the outcome goes to packed[row][ndim + 2] = (position, ln_prob, accepted) for the all-gather of a
// sharded ensemble; else the ensemble is updated in place.
template <bool PACKED>
__global__ void stretch_accept_kernel(long long half_n, int ndim, int hsel, long long lo, long long cnt,
                                      const double* __restrict__ prop, const double* __restrict__ new_lnp,
                                      const double* __restrict__ zf, const double* __restrict__ lnu, double* pos, double* lnp,
                                      long long* nacc, double* __restrict__ packed)
{
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= cnt) return;
    const long long k = (long long)hsel * half_n + lo + row;
    int acc = 0;
    double old = 0.0, nw = 0.0;
    if (lane == 0) {
        old = lnp[k];
        nw = new_lnp[row];
        acc = (zf[row] + nw - old) > lnu[row] ? 1 : 0;  // a NaN anywhere rejects
    }
    acc = __shfl_sync(0xffffffffu, acc, 0);
    if (PACKED) {
        double* out = packed + row * (ndim + 2);
        for (int d = lane; d < ndim; d += 32) out[d] = acc ? prop[row * ndim + d] : pos[k * ndim + d];
        if (lane == 0) {
            out[ndim] = acc ? nw : old;
            out[ndim + 1] = (double)acc;
        }
    } else if (acc) {
        for (int d = lane; d < ndim; d += 32) pos[k * ndim + d] = prop[row * ndim + d];
        if (lane == 0) {
            lnp[k] = nw;
            nacc[k] += 1;
        }
    }
}

// After the all-gather: gathered[world][slot][ndim + 2] holds every rank's packed rows (balanced contiguous
// shards of the half; rank r's rows start at its slot 0); write them into the replicated ensemble.
__global__ void stretch_update_kernel(long long half_n, int ndim, int hsel, int world, long long slot,
                                      const double* __restrict__ gathered, double* __restrict__ pos, double* __restrict__ lnp,
                                      long long* __restrict__ nacc)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = ndim + 2;
    if (idx >= half_n * w) return;
    const long long r = idx / w;
    const int d = (int)(idx - r * w);
    const long long base = half_n / world, extra = half_n % world;
    long long rank, lo;
    if (r < (base + 1) * extra) {
        rank = r / (base + 1);
        lo = rank * (base + 1);
    } else {
        rank = extra + (r - (base + 1) * extra) / base;
        lo = (base + 1) * extra + (rank - extra) * base;
    }
    const double v = gathered[(rank * slot + (r - lo)) * w + d];
    const long long k = (long long)hsel * half_n + r;
    if (d < ndim) pos[k * ndim + d] = v;
    else if (d == ndim) lnp[k] = v;
    else nacc[k] += (long long)v;
}

// End of a full step: count it, and (optionally) append the ensemble to the chain buffer
// chain[rec][n][ndim + 1] = (position, ln_prob).
__global__ void stretch_record_kernel(SamplerState* S, long long n, int ndim, const double* __restrict__ pos,
                                      const double* __restrict__ lnp, double* __restrict__ chain, long long chain_cap)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = ndim + 1;
    const unsigned long long rec = S->rec;
    if (chain && rec < (unsigned long long)chain_cap && idx < n * w) {
        const long long k = idx / w;
        const int d = (int)(idx - k * w);
        chain[(long long)rec * n * w + idx] = d < ndim ? pos[k * ndim + d] : lnp[k];
    }
}
__global__ void stretch_bump_kernel(SamplerState* S, int recorded)
{
    S->step += 1;
    if (recorded) S->rec += 1;
}

}  // namespace lfb

struct lfb_sampler {
    lfb_handle* h = nullptr;
    long long n = 0, half = 0;
    int ndim = 0, what = LFB_LN_PROB;
    double a = 2.0;
    DevBuf pos, lnp, nacc, prop, new_lnp, zf, lnu, state, packed, chain;
    long long chain_cap = 0;          // steps the chain buffer holds
    long long iterations = 0;         // full steps done (host mirror of state.step)
    long long recorded = 0;           // host mirror of state.rec
    bool have_state = false;
    cudaGraphExec_t graph = nullptr;  // one full step (small ensembles)
    bool graph_records = false;
    unsigned long long graph_generation = 0;
    long long graph_launches = 0;
    bool graphs_ok = true;
};

static void sampler_drop_graph(lfb_sampler* s)
{
    if (s->graph) cudaGraphExecDestroy(s->graph);
    s->graph = nullptr;
}

// propose + ln_prob + accept of rows [lo, hi) of one half, enqueued on st.  packed: null = update in place.
static int sampler_half(lfb_sampler* s, int hsel, long long lo, long long hi, double* packed, cudaStream_t st, bool inline_pass)
{
    lfb_handle* h = s->h;
    const long long cnt = hi - lo;
    if (cnt <= 0) return LFB_OK;
    const int nd = s->ndim;
    stretch_propose_kernel<<<(unsigned)((cnt * nd + 255) / 256), 256, 0, st>>>(
        s->state.as<SamplerState>(), s->half, nd, hsel, lo, cnt, s->a, s->pos.as<double>(), s->prop.as<double>(),
        s->zf.as<double>(), s->lnu.as<double>());
    h->launches++;
    int rc;
    if (inline_pass) rc = enqueue_pass_inline(h, s->what, cnt, s->prop.as<double>(), s->new_lnp.as<double>(), st);
    else rc = lfb_log_prob(h, s->what, cnt, s->prop.as<double>(), s->new_lnp.as<double>(), nullptr, (void*)st);
    if (rc) return rc;
    const unsigned blocks = (unsigned)((cnt * 32 + 255) / 256);
    if (packed)
        stretch_accept_kernel<true><<<blocks, 256, 0, st>>>(s->half, nd, hsel, lo, cnt, s->prop.as<double>(),
                                                             s->new_lnp.as<double>(), s->zf.as<double>(), s->lnu.as<double>(),
                                                             s->pos.as<double>(), s->lnp.as<double>(), s->nacc.as<long long>(),
                                                             packed);
    else
        stretch_accept_kernel<false><<<blocks, 256, 0, st>>>(s->half, nd, hsel, lo, cnt, s->prop.as<double>(),
                                                              s->new_lnp.as<double>(), s->zf.as<double>(), s->lnu.as<double>(),
                                                              s->pos.as<double>(), s->lnp.as<double>(), s->nacc.as<long long>(),
                                                              nullptr);
    h->launches++;
    CK(cudaGetLastError());
    return LFB_OK;
}

static int sampler_end_step(lfb_sampler* s, bool record, cudaStream_t st)
{
    lfb_handle* h = s->h;
    if (record) {
        const long long tot = s->n * (s->ndim + 1);
        stretch_record_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(s->state.as<SamplerState>(), s->n, s->ndim,
                                                                              s->pos.as<double>(), s->lnp.as<double>(),
                                                                              s->chain.as<double>(), s->chain_cap);
        h->launches++;
    }
    stretch_bump_kernel<<<1, 1, 0, st>>>(s->state.as<SamplerState>(), record ? 1 : 0);
    h->launches++;
    CK(cudaGetLastError());
    return LFB_OK;
}

extern "C" {

void lfb_sampler_destroy(lfb_sampler* s);

int lfb_sampler_create(lfb_handle* h, long long nwalkers, double a, unsigned long long seed, int what, lfb_sampler** out)
{
    if (!h || !out) return LFB_EINVAL;
    *out = nullptr;
    if (!h->have_layout) return fail(h, LFB_ESTATE, "sampler_create: call set_layout first");
    if (nwalkers < 2 || (nwalkers & 1) || nwalkers < 2LL * h->ndim)
        return fail(h, LFB_EINVAL, "sampler_create: need an even number of walkers, at least twice the number of dimensions");
    if (!(a > 1.0) || what < LFB_LN_PRIOR || what > LFB_LN_PROB) return fail(h, LFB_EINVAL, "sampler_create: need a > 1");
    CK(cudaSetDevice(h->device));
    lfb_sampler* s = new lfb_sampler();
    s->h = h;
    s->n = nwalkers;
    s->half = nwalkers / 2;
    s->ndim = h->ndim;
    s->what = what;
    s->a = a;
    const size_t nd = (size_t)s->ndim;
    cudaError_t e = cudaSuccess;
    auto need = [&](DevBuf& b, size_t bytes) {
        if (e == cudaSuccess) e = b.reserve(bytes);
    };
    need(s->pos, 8 * (size_t)s->n * nd);
    need(s->lnp, 8 * (size_t)s->n);
    need(s->nacc, 8 * (size_t)s->n);
    need(s->prop, 8 * (size_t)s->half * nd);
    need(s->new_lnp, 8 * (size_t)s->half);
    need(s->zf, 8 * (size_t)s->half);
    need(s->lnu, 8 * (size_t)s->half);
    need(s->state, sizeof(SamplerState));
    need(s->packed, 8 * (size_t)s->half * (nd + 2));
    SamplerState S0 = {seed, 0ull, 0ull};
    if (e == cudaSuccess) e = cudaMemcpy(s->state.p, &S0, sizeof(S0), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(s->nacc.p, 0, 8 * (size_t)s->n);
    if (e != cudaSuccess) {
        h->err = std::string("sampler_create: ") + cudaGetErrorString(e);
        lfb_sampler_destroy(s);
        return LFB_ECUDA;
    }
    *out = s;
    return LFB_OK;
}

void lfb_sampler_destroy(lfb_sampler* s)
{
    if (!s) return;
    cudaSetDevice(s->h->device);
    cudaDeviceSynchronize();
    sampler_drop_graph(s);
    DevBuf* b[] = {&s->pos, &s->lnp, &s->nacc, &s->prop, &s->new_lnp, &s->zf, &s->lnu, &s->state, &s->packed, &s->chain};
    for (DevBuf* x : b) x->release();
    delete s;
}

// Positions (host or device, [n][ndim]) and, optionally, their log-probabilities (else they are evaluated).
int lfb_sampler_set_state(lfb_sampler* s, const double* pos, const double* lnp, void* stream_v)
{
    if (!s || !pos) return LFB_EINVAL;
    lfb_handle* h = s->h;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : h->stream;
    CK(cudaMemcpyAsync(s->pos.p, pos, 8 * (size_t)s->n * s->ndim, cudaMemcpyDefault, st));
    if (lnp) CK(cudaMemcpyAsync(s->lnp.p, lnp, 8 * (size_t)s->n, cudaMemcpyDefault, st));
    else {
        int rc = lfb_log_prob(h, s->what, s->n, s->pos.as<double>(), s->lnp.as<double>(), nullptr, (void*)st);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(st));
    s->have_state = true;
    return LFB_OK;
}

// Record the ensemble after every step of lfb_sampler_run into a device buffer of `steps` steps
// ([steps][n][ndim + 1]: position, ln_prob); 0 switches recording off.  Read it back with lfb_sampler_read_chain.
int lfb_sampler_set_chain(lfb_sampler* s, long long steps)
{
    if (!s || steps < 0) return LFB_EINVAL;
    lfb_handle* h = s->h;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    sampler_drop_graph(s);
    if (steps > 0) CK(s->chain.reserve(8 * (size_t)steps * (size_t)s->n * (size_t)(s->ndim + 1)));
    s->chain_cap = steps;
    s->recorded = 0;
    unsigned long long zero = 0;
    CK(cudaMemcpy(&s->state.as<SamplerState>()->rec, &zero, sizeof(zero), cudaMemcpyHostToDevice));
    return LFB_OK;
}

// Copy the recorded steps to out (host, [steps recorded][n][ndim + 1]) and empty the buffer.
int lfb_sampler_read_chain(lfb_sampler* s, double* out, long long* n_steps)
{
    if (!s || !n_steps) return LFB_EINVAL;
    lfb_handle* h = s->h;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    *n_steps = s->recorded;
    if (s->recorded && out)
        CK(cudaMemcpy(out, s->chain.p, 8 * (size_t)s->recorded * (size_t)s->n * (size_t)(s->ndim + 1), cudaMemcpyDeviceToHost));
    s->recorded = 0;
    unsigned long long zero = 0;
    CK(cudaMemcpy(&s->state.as<SamplerState>()->rec, &zero, sizeof(zero), cudaMemcpyHostToDevice));
    return LFB_OK;
}

// nsteps full steps (two half-steps each) of the whole ensemble on this GPU, nothing crossing PCIe.  Small
// ensembles replay one captured CUDA graph per step (the pass is launch bound).  Returns without synchronising.
int lfb_sampler_run(lfb_sampler* s, long long nsteps, void* stream_v)
{
    if (!s || nsteps < 0) return LFB_EINVAL;
    lfb_handle* h = s->h;
    if (!s->have_state) return fail(h, LFB_ESTATE, "sampler_run: call sampler_set_state first");
    if (s->what != LFB_LN_PRIOR && !h->have_lc) return fail(h, LFB_ESTATE, "sampler_run: call set_lightcurves first");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : h->stream;
    const bool rec = s->chain_cap > 0;
    if (rec && s->recorded + nsteps > s->chain_cap) return fail(h, LFB_EINVAL, "sampler_run: the chain buffer is full (read it first)");
    const bool small = h->graphs_on && s->graphs_ok && !h->trace && s->half * h->n_ecl <= h->graph_max_jobs;
    long long done = 0;
    if (small) {
        if (s->graph && (s->graph_generation != h->alloc_generation || s->graph_records != rec)) sampler_drop_graph(s);
        if (!s->graph) {
            // one plain step sizes every buffer, then the same step is captured
            if (nsteps == 0) return LFB_OK;
            int rc;
            for (int hs = 0; hs < 2; ++hs)
                if ((rc = sampler_half(s, hs, 0, s->half, nullptr, st, true))) return rc;
            if ((rc = sampler_end_step(s, rec, st))) return rc;
            ++done;
            CK(cudaStreamSynchronize(st));
            const long long l0 = h->launches;
            cudaGraph_t g = nullptr;
            bool ok = false;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                rc = LFB_OK;
                for (int hs = 0; hs < 2 && rc == LFB_OK; ++hs) rc = sampler_half(s, hs, 0, s->half, nullptr, st, true);
                if (rc == LFB_OK) rc = sampler_end_step(s, rec, st);
                const cudaError_t e = cudaStreamEndCapture(st, &g);
                if (rc == LFB_OK && e == cudaSuccess && g && cudaGraphInstantiate(&s->graph, g, 0) == cudaSuccess) {
                    s->graph_launches = h->launches - l0;
                    s->graph_generation = h->alloc_generation;
                    s->graph_records = rec;
                    ok = true;
                }
                if (g) cudaGraphDestroy(g);
            }
            h->launches = l0;
            if (!ok) {
                cudaGetLastError();
                s->graph = nullptr;
                s->graphs_ok = false;  // plain launches from now on
            }
        }
        if (s->graph) {
            for (; done < nsteps; ++done) {
                CK(cudaGraphLaunch(s->graph, st));
                h->launches += s->graph_launches;
            }
        }
    }
    for (; done < nsteps; ++done) {
        int rc;
        for (int hs = 0; hs < 2; ++hs)
            if ((rc = sampler_half(s, hs, 0, s->half, nullptr, st, false))) return rc;
        if ((rc = sampler_end_step(s, rec, st))) return rc;
    }
    s->iterations += nsteps;
    if (rec) s->recorded += nsteps;
    return LFB_OK;
}

// Sharded ensemble, one rank's share of a half-step: propose, evaluate and accept rows [lo, hi) of half
// `half` (0 / 1); the outcome lands in packed[hi - lo][ndim + 2] (device; position, ln_prob, accepted) for
// the all-gather.  packed = NULL uses the sampler's own buffer (lfb_sampler_packed).
int lfb_sampler_half_begin(lfb_sampler* s, int half, long long lo, long long hi, double* packed, void* stream_v)
{
    if (!s || half < 0 || half > 1 || lo < 0 || hi < lo || hi > s->half) return LFB_EINVAL;
    lfb_handle* h = s->h;
    if (!s->have_state) return fail(h, LFB_ESTATE, "sampler_half_begin: call sampler_set_state first");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : h->stream;
    return sampler_half(s, half, lo, hi, packed ? packed : s->packed.as<double>(), st, false);
}

// ... and after the all-gather: gathered[world][slot][ndim + 2] (device) -> the replicated ensemble.  The
// second half's call ends the step (counter, chain record).
int lfb_sampler_half_end(lfb_sampler* s, int half, const double* gathered, int world, long long slot, void* stream_v)
{
    if (!s || half < 0 || half > 1 || !gathered || world < 1 || slot * world < s->half) return LFB_EINVAL;
    lfb_handle* h = s->h;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : h->stream;
    const long long tot = s->half * (s->ndim + 2);
    stretch_update_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(s->half, s->ndim, half, world, slot, gathered,
                                                                          s->pos.as<double>(), s->lnp.as<double>(),
                                                                          s->nacc.as<long long>());
    h->launches++;
    CK(cudaGetLastError());
    if (half == 1) {
        const bool rec = s->chain_cap > 0;
        if (rec && s->recorded + 1 > s->chain_cap) return fail(h, LFB_EINVAL, "sampler_half_end: the chain buffer is full");
        int rc = sampler_end_step(s, rec, st);
        if (rc) return rc;
        s->iterations += 1;
        if (rec) s->recorded += 1;
    }
    return LFB_OK;
}

double* lfb_sampler_packed(lfb_sampler* s) { return s ? s->packed.as<double>() : nullptr; }
double* lfb_sampler_positions(lfb_sampler* s) { return s ? s->pos.as<double>() : nullptr; }
double* lfb_sampler_log_prob(lfb_sampler* s) { return s ? s->lnp.as<double>() : nullptr; }

// Ensemble, log-probabilities and per-walker acceptance counts to the host (any may be NULL).
int lfb_sampler_get_state(lfb_sampler* s, double* pos, double* lnp, long long* naccepted, long long* iterations)
{
    if (!s) return LFB_EINVAL;
    lfb_handle* h = s->h;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    if (pos) CK(cudaMemcpy(pos, s->pos.p, 8 * (size_t)s->n * s->ndim, cudaMemcpyDeviceToHost));
    if (lnp) CK(cudaMemcpy(lnp, s->lnp.p, 8 * (size_t)s->n, cudaMemcpyDeviceToHost));
    if (naccepted) CK(cudaMemcpy(naccepted, s->nacc.p, 8 * (size_t)s->n, cudaMemcpyDeviceToHost));
    if (iterations) *iterations = s->iterations;
    return LFB_OK;
}

// The draws of the stretch move for rows [0, cnt) of one half at one step, on the host (tests: the kernels'
// random stream against an independent Philox): out[cnt][3] = (z, partner row, ln u').
int lfb_stretch_draws(unsigned long long seed, unsigned long long step, int half, long long half_n, double a, long long cnt,
                      double* out)
{
    if (!out || cnt < 0 || half_n < 1) return LFB_EINVAL;
    for (long long r = 0; r < cnt; ++r) {
        unsigned w[4], q[4];
        const unsigned tag = (unsigned)half;
        philox4x32_10((unsigned)r, (unsigned)step, (unsigned)(step >> 32), tag, (unsigned)seed, (unsigned)(seed >> 32), w);
        philox4x32_10((unsigned)r, (unsigned)step, (unsigned)(step >> 32), tag | 0x80000000u, (unsigned)seed,
                      (unsigned)(seed >> 32), q);
        const double t = (a - 1.0) * u01(w[0], w[1]) + 1.0;
        out[3 * r] = t * t / a;
        out[3 * r + 1] = (double)(((unsigned long long)w[2] * (unsigned long long)half_n) >> 32);
        out[3 * r + 2] = log(u01(q[0], q[1]));
    }
    return LFB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- chain writer (host)
// The production chain file of the reference: after the header line, one line per walker per step,
//     "{0:4d} {1:s} {2:f}\n".format(k, " ".join(map(str, pos[k])), prob[k])          mcmc_utils.py:157-164
// str() of a numpy float64 is the shortest decimal string that reads back to the same double, positional for
// 1e-4 <= |x| < 1e16 and scientific otherwise.  The reference opens the file once per walker per step; here a
// whole block of steps is formatted (in parallel over walkers) and appended with one write -- the same bytes.
#include <charconv>
#include <thread>

namespace lfb {

// str(numpy.float64(v)) appended to out
static void append_py_float(std::string& out, double v)
{
    if (v != v) { out += "nan"; return; }
    if (v == INFINITY) { out += "inf"; return; }
    if (v == -INFINITY) { out += "-inf"; return; }
    if (v == 0.0) { out += std::signbit(v) ? "-0.0" : "0.0"; return; }
    char buf[48];
    const auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);  // shortest round trip
    const char *p = buf, *end = res.ptr;
    if (*p == '-') { out += '-'; ++p; }
    // d[.ddd]e[+-]XX
    char digits[24];
    int nd = 0;
    while (p < end && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        ++p;
    }
    ++p;  // 'e'
    const bool eneg = *p == '-';
    ++p;
    int ex = 0;
    while (p < end) ex = ex * 10 + (*p++ - '0');
    if (eneg) ex = -ex;
    const int decpt = ex + 1;
    if (ex >= 16 || ex < -4) {
        out += digits[0];
        if (nd > 1) {
            out += '.';
            out.append(digits + 1, nd - 1);
        }
        out += 'e';
        out += ex < 0 ? '-' : '+';
        const int ae = ex < 0 ? -ex : ex;
        if (ae < 10) out += '0';
        out += std::to_string(ae);
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out += '.';
        out.append(digits + decpt, nd - decpt);
    }
}

// "{:f}".format(v)
static void append_py_fixed6(std::string& out, double v)
{
    if (v != v) { out += "nan"; return; }
    if (v == INFINITY) { out += "inf"; return; }
    if (v == -INFINITY) { out += "-inf"; return; }
    char buf[400];
    const int len = snprintf(buf, sizeof(buf), "%f", v);
    out.append(buf, (size_t)(len < (int)sizeof(buf) ? len : (int)sizeof(buf) - 1));
}

// rows[n_steps][n][ndim + 1] (position, ln_prob) -> text
static void format_chain(long long n_steps, long long n, int ndim, const double* rows, std::string& text)
{
    const long long total = n_steps * n;
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1 : (nt > 16 ? 16 : nt);
    if (total * (ndim + 1) < 20000) nt = 1;
    std::vector<std::string> parts(nt);
    auto work = [&](unsigned t) {
        std::string& s = parts[t];
        const long long a = total * t / nt, b = total * (t + 1) / nt;
        s.reserve((size_t)(b - a) * (size_t)(20 * (ndim + 1) + 8));
        char head[32];
        for (long long i = a; i < b; ++i) {
            const long long k = i % n;
            const double* r = rows + i * (ndim + 1);
            const int hl = snprintf(head, sizeof(head), "%4lld ", k);
            s.append(head, (size_t)hl);
            for (int d = 0; d < ndim; ++d) {
                if (d) s += ' ';
                append_py_float(s, r[d]);
            }
            s += ' ';
            append_py_fixed6(s, r[ndim]);
            s += '\n';
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    size_t len = 0;
    for (auto& p : parts) len += p.size();
    text.clear();
    text.reserve(len);
    for (auto& p : parts) text += p;
}

}  // namespace lfb

extern "C" {

// Format n_steps steps of n walkers (rows[n_steps][n][ndim + 1]: position, ln_prob; host) as the reference's
// chain lines.  Returns the number of bytes the text takes; it is copied to out when cap holds it.
long long lfb_chain_format(long long n_steps, long long n, int ndim, const double* rows, char* out, long long cap)
{
    if (n_steps < 0 || n < 0 || ndim < 0 || (n_steps * n && !rows)) return -1;
    std::string text;
    format_chain(n_steps, n, ndim, rows, text);
    if (out && cap >= (long long)text.size()) memcpy(out, text.data(), text.size());
    return (long long)text.size();
}

// ... and append them to `path` with a single write.  0 on success.
int lfb_chain_append(const char* path, long long n_steps, long long n, int ndim, const double* rows)
{
    if (!path || n_steps < 0 || n < 0 || ndim < 0 || (n_steps * n && !rows)) return LFB_EINVAL;
    std::string text;
    format_chain(n_steps, n, ndim, rows, text);
    FILE* f = fopen(path, "ab");
    if (!f) return LFB_EINVAL;
    const size_t wr = fwrite(text.data(), 1, text.size(), f);
    const int rc = fclose(f);
    return (wr == text.size() && rc == 0) ? LFB_OK : LFB_EINVAL;
}

}  // extern "C"
