/*
 * include/lfit_b200.h -- C ABI of the B200 (sm_100a) LFIT CV eclipse-model engine.
 *
 * This is the drop-in boundary for the one hot path of wildjames/lfit_python:
 * evaluating lfit.CV(pars).calcFlux(pars, phase, width) and its chi-squared /
 * log-probability for every emcee walker.  Plain pointers and sizes only; no
 * torch types.  Every entry point names the reference interface it replaces
 * (file:line under /root/reference).  The reference binds this path through
 * Cython (`import lfit`, CVModel.py:13) and a CPython extension (`from trm
 * import roche`, CVModel.py:15); the ctypes stub that replaces both is
 * lfit_python_b200/_cabi.py and is shown in INTEGRATION.md.
 *
 * Conventions
 *  - all functions return 0 on success, a negative LFB_E* code on failure;
 *    lfb_last_error() gives the message.  An INVALID MODEL IS DATA, not an
 *    error: -inf log-probability / +inf chi-squared / NaN flux, as the reference
 *    produces at CVModel.py:139-144,163-171 and model.py:489-493.
 *  - there is no CPU fallback: every call needs a CUDA device.
 *  - pointers marked "host or device" are classified with
 *    cudaPointerGetAttributes; host buffers are staged through pinned memory.
 *  - `stream` is a cudaStream_t passed as void* (NULL = the handle's own stream).
 *  - a handle is not thread-safe; use one per GPU / rank.
 */
#ifndef LFIT_B200_H
#define LFIT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LFB_NPAR 18 /* CV parameter slots, order of ComplexEclipse.cv_parnames (CVModel.py:383-390) */

enum {
    LFB_OK = 0,
    LFB_EINVAL = -1,   /* bad argument (shape, count, NULL) */
    LFB_ECUDA = -2,    /* CUDA runtime error */
    LFB_ESTATE = -3,   /* call order: layout / lightcurves not set */
    LFB_ENOMEM = -4
};

/* lfb_calc_flux flags */
enum {
    LFB_FLAG_INCL = 1,      /* slot 5 (dphi) holds the inclination in degrees: the lfit.Py* component API (testCV.py:44-49) */
    LFB_FLAG_SKIP_WD = 2,
    LFB_FLAG_SKIP_DISC = 4,
    LFB_FLAG_SKIP_BS = 8,
    LFB_FLAG_SKIP_DONOR = 16
};

/* prior codes = Prior.type (model.py:70) */
enum { LFB_PRIOR_GAUSS = 0, LFB_PRIOR_GAUSSPOS, LFB_PRIOR_UNIFORM, LFB_PRIOR_LOGUNIFORM, LFB_PRIOR_MODJEFF };

/* lfb_log_prob `what` = the three wrappers of mcmcfit.py:30-48 */
enum { LFB_LN_PRIOR = 0, LFB_LN_LIKE = 1, LFB_LN_PROB = 2 };

/* lfb_roche `which` = trm.roche calls on the path (CVModel.py:222,288,460,561) */
enum { LFB_ROCHE_XL1 = 0, LFB_ROCHE_FINDPHI = 1, LFB_ROCHE_FINDI = 2, LFB_ROCHE_BSPOT = 3,
       LFB_ROCHE_ANGLE = 4 /* test aid, no reference counterpart: out = the solver's arctangent of the unit vector (a, b) */ };

/* Surface-grid density (the element counts the lfit component constructors take,
 * testCV.py:31,43).  Zero / NULL selects the defaults in brackets. */
typedef struct {
    int n_wd_rings;      /* [10]  white dwarf: 4*n^2 sky tiles */
    int n_disc_r;        /* [25]  disc rings  */
    int n_disc_th;       /* [40]  disc sectors (even): 25*40 = 1000 = testCV.py:31 */
    int n_bs;            /* [200] bright-spot strip elements */
    int n_donor_th;      /* [18]  donor rings -> 412 tiles (testCV.py:43 asks for 400) */
    int n_quad;          /* [3]   exposure quadrature points (odd; Simpson) */
    double donor_ulimb;  /* [0.8] */
    double donor_gdexp;  /* [0.32] */
} lfb_config;

typedef struct lfb_handle lfb_handle;

/* lfit.CV.__init__ (CVModel.py:128): allocate an engine on CUDA device `device`. */
int lfb_create(int device, const lfb_config *cfg, lfb_handle **out);
void lfb_destroy(lfb_handle *h);
const char *lfb_last_error(const lfb_handle *h); /* h may be NULL: last create() error */
int lfb_get_config(const lfb_handle *h, lfb_config *out);

/* The flattened tree: replaces the per-walker traversal of
 * Node.__set_parameter_vector__ (model.py:586-603) + SimpleEclipse.cv_parlist
 * (CVModel.py:335-354).  gather[e*18+k] >= 0: column of theta; < 0: consts[-g-1].
 * npars is 14 (SimpleEclipse) or 18 (ComplexEclipse). */
int lfb_set_layout(lfb_handle *h, int ndim, int n_ecl, int npars, const int *gather, int n_consts,
                   const double *consts);

/* Every Param of the tree with its Prior (model.py:40-141, Node.ln_prior model.py:426-474). */
int lfb_set_priors(lfb_handle *h, int n_prior, const int *src, const int *type, const double *p1,
                   const double *p2, const double *norm, const int *isvar);

/* Lightcurve.x/.w/.y/.ye of every leaf (CVModel.py:20-83), concatenated; copied once to HBM. */
int lfb_set_lightcurves(lfb_handle *h, int n_ecl, const long long *off, const double *phase,
                        const double *width, const double *y, const double *ye);

/* mcmcfit.ln_prior / ln_like / ln_prob (mcmcfit.py:30-48) for n walkers at once.
 * theta: row-major f64[n][ndim], host or device.  out: f64[n], host or device.
 * chisq_out (optional): f64[n][n_ecl] per-leaf SimpleEclipse.chisq (CVModel.py:157-178). */
int lfb_log_prob(lfb_handle *h, int what, long long n, const double *theta, double *out,
                 double *chisq_out, void *stream);

/* lfit.CV.calcFlux(pars, phase, width) (CVModel.py:138,154; plot_lc_model.py:134) for n_sets
 * parameter sets sharing one phase grid.  pars: f64[n_sets][npars] (host or device), phase/width:
 * f64[n_ph] host.  out_total: f64[n_sets][n_ph]; out_comp (optional): f64[4][n_sets][n_ph] in
 * the order ywd, yd, ys, yrs (the scaled contributions, plot_lc_model.py:135-138). */
int lfb_calc_flux(lfb_handle *h, long long n_sets, const double *pars, int npars, int flags, int n_ph,
                  const double *phase, const double *width, double *out_total, double *out_comp,
                  void *stream);

/* trm.roche.{xl1(q), findphi(q,i), findi(q,dphi), bspot(q,rad)} batched on the device.
 * a: first argument (q), b: second argument (ignored for xl1); out: f64[n][4]
 * (xl1 / dphi / incl_deg in out[.][0]; bspot -> x, y, vx, vy); ok[n]: 1 = value, 0 = the
 * reference would raise (CVModel.py:223,309). */
int lfb_roche(lfb_handle *h, int which, long long n, const double *a, const double *b, double *out,
              int *ok);

/* ---- Gaussian-process likelihood (GPLCModel / SimpleGPEclipse / ComplexGPEclipse, CVModel.py:494-711) ----
 * With the GP on, lfb_log_prob's likelihood of an eclipse is george's
 *     GP(ampin * Matern32(tau) + sum over gaps of ampout * Matern32(tau, block=gap)).compute(x, ye)
 *       .log_likelihood(y - model, quiet=True)                       (CVModel.py:603-696)
 * instead of -chi^2/2, evaluated exactly (4-state Kalman filter over the points in ascending x)
 * rather than with george's approximate HODLR solver.  The gaps are the reference's change points
 * [(e - 1) + dist_cp + phi0, e - dist_cp + phi0] (CVModel.py:580-599).
 * gp_src[3]: where ln_ampin_gp, ln_ampout_gp, ln_tau_gp come from -- a column of theta (>= 0) or
 * -(k + 1) for constant k of set_layout, like the gather table; dist_cp[n_ecl]: (dphi + dpwd) / 2
 * of each eclipse (CVModel.py:558-568; the reference computes it once and caches it).
 * chisq_out of lfb_log_prob then holds -2 ln L per eclipse.  enabled = 0 switches back. */
int lfb_set_gp(lfb_handle *h, int enabled, const int gp_src[3], const double *dist_cp);
/* The same likelihood for caller-supplied residuals (the scalar tree path and tests):
 * x[n] ascending, ye[n], resid[n_sets][n], hyper[n_sets][3] = (ampin, ampout, tau) -- not
 * logarithms --, gaps[n_sets][n_gaps][2] (n_gaps <= 8, disjoint, ascending); out[n_sets],
 * -inf where george would return -inf (quiet=True). Host or device pointers. */
int lfb_gp_loglike(lfb_handle *h, long long n_sets, int n, const double *x, const double *ye, const double *resid,
                   const double *hyper, int n_gaps, const double *gaps, double *out);
/* trm.roche.wdphases(q, iangle, r1, ntheta=...) (call site CVModel.py:562): third and fourth
 * contact of the white dwarf, out[n][2] = (phi3, phi4); ok[n] = 0 where trm.roche would raise. */
int lfb_wdphases(lfb_handle *h, long long n, const double *q, const double *incl_deg, const double *r1, int ntheta,
                 double *out, int *ok);

/* Stage (1) on its own: ingress / egress phases (cycles, ingress < egress) of surface elements
 * seen at inclination incl_deg in a binary of mass ratio q.  pts[n][5] = (x, y, z, xi, eta): a
 * point fixed in the rotating frame (units of a, origin at the white dwarf) plus an offset fixed
 * on the sky (xi along the orbital motion at phase 0, eta towards the projected pole).  ok[n] = 0:
 * never eclipsed (out = NaN).  No reference counterpart (lfit does this internally); parity tests. */
int lfb_ingress_egress(lfb_handle *h, long long n, const double *q, const double *incl_deg, const double *pts,
                       double *out, int *ok);

/* ---- Stretch-move sampler with the ensemble resident in HBM ----
 * Replaces emcee.EnsembleSampler(nwalkers, npars, ln_prob, args=(model,), pool=pool) (mcmcfit.py:283-288)
 * and the sampler.sample() loops of run_burnin / run_mcmc_save (mcmc_utils.py:114-183) for the model set on
 * the handle.  emcee is third-party and not vendored; the move is the published stretch move (a = 2 by
 * default), first half of the ensemble then second half (emcee 2.x).  Random numbers are Philox4x32-10
 * counted by (step, half, walker): an ensemble sharded over N GPUs follows the 1-GPU chain bit for bit.
 * what = LFB_LN_PROB normally.  nwalkers even and >= 2 ndim (mcmcfit.py:195-196). */
typedef struct lfb_sampler lfb_sampler;
int lfb_sampler_create(lfb_handle *h, long long nwalkers, double a, unsigned long long seed, int what,
                       lfb_sampler **out);
void lfb_sampler_destroy(lfb_sampler *s);
/* Start positions pos[n][ndim] (host or device) and optionally their log-probabilities (NULL: evaluated). */
int lfb_sampler_set_state(lfb_sampler *s, const double *pos, const double *lnp, void *stream);
/* nsteps full steps on this GPU (nothing crosses PCIe; ensembles whose half fits one batch replay one CUDA graph per step).
 * Asynchronous: lfb_sampler_get_state synchronises. */
int lfb_sampler_run(lfb_sampler *s, long long nsteps, void *stream);
/* Sharded ensemble (one process per GPU, the ensemble replicated): this rank proposes, evaluates and
 * accepts rows [lo, hi) of half `half` (0 or 1) into packed[hi - lo][ndim + 2] = (position, ln_prob,
 * accepted) on the device (NULL: the sampler's own buffer, lfb_sampler_packed); the caller all-gathers
 * the packed rows of all ranks (lfb_peer_allgather below, or NCCL), then lfb_sampler_half_end writes gathered[world][slot][ndim + 2]
 * (rank r's rows at gathered[r][0..]; balanced contiguous shards, lower ranks one row longer) into the
 * ensemble.  The second half's half_end ends the step. */
int lfb_sampler_half_begin(lfb_sampler *s, int half, long long lo, long long hi, double *packed, void *stream);
int lfb_sampler_half_end(lfb_sampler *s, int half, const double *gathered, int world, long long slot, void *stream);
double *lfb_sampler_packed(lfb_sampler *s);    /* device, [nwalkers / 2][ndim + 2] */
double *lfb_sampler_positions(lfb_sampler *s); /* device, [nwalkers][ndim] */
double *lfb_sampler_log_prob(lfb_sampler *s);  /* device, [nwalkers] */
/* Ensemble, log-probabilities, per-walker acceptance counts and the step count to the host (any NULL). */
int lfb_sampler_get_state(lfb_sampler *s, double *pos, double *lnp, long long *naccepted, long long *iterations);
/* Record the ensemble after every step into a device buffer of `steps` steps (0: off); read_chain copies
 * the steps recorded since the last read to out[steps][n][ndim + 1] (position, ln_prob; host) and empties it. */
int lfb_sampler_set_chain(lfb_sampler *s, long long steps);
int lfb_sampler_read_chain(lfb_sampler *s, double *out, long long *n_steps);
/* The move's draws on the host (tests): out[cnt][3] = (z, partner row, ln u') of rows [0, cnt) of one half. */
int lfb_stretch_draws(unsigned long long seed, unsigned long long step, int half, long long half_n, double a,
                      long long cnt, double *out);

/* ---- Chain file (mcmc_utils.py:157-164; reader mcmc_utils.py:252-272) ----
 * rows[n_steps][n][ndim + 1] (position, ln_prob; host) as the reference's lines
 * "{0:4d} {1:s} {2:f}\n".format(k, " ".join(map(str, pos[k])), prob[k]) -- byte for byte, formatted in
 * parallel and appended with ONE write per call instead of one open() per walker per step.
 * lfb_chain_format returns the size of the text and copies it to out when cap holds it. */
long long lfb_chain_format(long long n_steps, long long n, int ndim, const double *rows, char *out, long long cap);
int lfb_chain_append(const char *path, long long n_steps, long long n, int ndim, const double *rows);

/* counters for bench.py: kernels launched by this handle since creation */
long long lfb_launch_count(const lfb_handle *h);
/* Diagnostics: element solves that needed the last-resort (scan + golden section + bisection) solver since
 * the library was loaded, on this handle's device; -1 on error.  Synchronises the device. */
long long lfb_robust_calls(lfb_handle *h);
/* Device time (ms) of the stages of the last lfb_log_prob (its last batch), from CUDA events
 * recorded on the stream the kernels ran on; valid once that stream is synchronised.
 * out = {walker, stream, elements (4 launches), flux, finish, total}; the stages are -1 when the
 * call replayed a CUDA graph (a call that fits one batch does, from its third identical occurrence on). */
int lfb_last_stage_ms(lfb_handle *h, float out[6]);
/* elements + flux of the same call, <0 if none */
float lfb_last_kernel_ms(lfb_handle *h);
/* Per-kernel trace (measurement aid, no reference counterpart).  lfb_set_trace(h, 1) brackets
 * every kernel of the last batch on lane 0 with CUDA events; lfb_last_trace_ms returns the
 * device time of each in LFB_K_* order (-1 where the kernel did not run) and, in the last
 * slot, the ballistic-stream kernel on its side stream.  Off by default: the extra event
 * records cost a few microseconds per batch. */
enum {
    LFB_K_WALKER = 0,  /* walker_kernel */
    LFB_K_JOBCHECK,    /* jobcheck_kernel */
    LFB_K_ELEM_DISC,   /* elements_kernel<1> */
    LFB_K_ELEM_WD,     /* elements_kernel<0> */
    LFB_K_ELEM_DONOR,  /* elements_kernel<3> (second side stream, beside the kernels of the main stream) */
    LFB_K_DONOR_TABLE, /* donor_table_kernel: the donor curve's phase table, per walker (second side stream) */
    LFB_K_PREP,        /* prep_kernel */
    LFB_K_POSITIONS,   /* positions_kernel<0>: white dwarf, disc */
    LFB_K_ELEM_BS,     /* elements_kernel<2>, with any wait for the stream ODE */
    LFB_K_PREP_BS,     /* prep_strip_kernel, positions_kernel<1>: the strip's share of the flux preparation */
    LFB_K_FLUX,        /* flux_kernel */
    LFB_K_GP,          /* gp_kernel (only with lfb_set_gp) */
    LFB_K_FINISH,      /* finish_kernel */
    LFB_K_COUNT        /* last_trace_ms slot of stream_kernel (side stream) */
};
int lfb_set_trace(lfb_handle *h, int on);
int lfb_last_trace_ms(lfb_handle *h, float out[LFB_K_COUNT + 1]);

/* ---- Exchange of positions and log-probabilities between the GPUs of one box over NVLink peer memory ----
 * Replaces the hand-back of the workers' results to the parent of the reference's multiprocessing.Pool
 * (mcmcfit.py:273-288); north_star's all-gather per stretch-move half-step.  One process per GPU:
 *   lfb_peer_create   allocates this rank's window (2 buffers of world slots of slot_bytes) and returns its CUDA
 *                     IPC handle; the caller exchanges the 64-byte handles of all ranks (any transport),
 *   lfb_peer_connect  maps the peers' windows, handles[world][64] in rank order,
 *   lfb_peer_allgather packs rows [a row | b row] (a: rows x ca, b: rows x cb doubles on the device; cb may be 0),
 *                     stores them into slot `rank` of every rank's window with one kernel, raises this rank's flag
 *                     everywhere and waits -- on the device, on `stream` -- for every rank's flag.  *gathered =
 *                     this step's buffer in the local window, [world][slot_bytes], valid to work queued on `stream`
 *                     until the next-but-one exchange.  All ranks must call it the same number of times.
 *   lfb_peer_status   synchronises and reports whether an exchange gave up waiting for a peer (~30 s). */
int lfb_peer_create(lfb_handle *h, int rank, int world, long long slot_bytes, unsigned char handle_out[64]);
int lfb_peer_connect(lfb_handle *h, const unsigned char *handles);
int lfb_peer_allgather(lfb_handle *h, const double *a, int ca, const double *b, int cb, long long rows,
                       const double **gathered, void *stream);
int lfb_peer_status(lfb_handle *h, int *timed_out);
void lfb_peer_destroy(lfb_handle *h);

/* Measurement aid (no reference counterpart): sustained FP64 FMA rate of the device in
 * TFLOP/s (FMA = 2), the roofline denominator bench.py reports against. */
int lfb_measure_fp64_peak(lfb_handle *h, int iters, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
