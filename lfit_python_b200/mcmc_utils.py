"""Sampler glue: walker initialisation, burn-in / production loops, chain file I/O, and a
minimal affine-invariant ensemble sampler.

Mirror of the parts of /root/reference/mcmc_utils.py either side of the hot path
(initialise_walkers :46-72, run_burnin :114-132, run_mcmc_save :135-183, flatchain :242-249,
readchain :252-272), vectorised: ln_prior is called once per resampling round on the whole
walker matrix, and the chain file is appended once per step instead of once per walker.
`emcee` is not installed here; EnsembleSampler below implements the same stretch move
(Goodman & Weare 2010, a = 2, two half-ensembles per step) with emcee's call shape, so the
same driver runs with either.
"""
import numpy as np

TINY = -np.inf


def sample_ball(p0, std, size, rng):
    """emcee.utils.sample_ball: Gaussian ball about p0 with per-dimension std."""
    p0, std = np.asarray(p0, dtype=np.float64), np.asarray(std, dtype=np.float64)
    return p0 + std * rng.standard_normal((size, p0.shape[0]))


def initialise_walkers(p, scatter, nwalkers, ln_prior, model, rng=None, max_rounds=200, verbose=True):
    """Ball of walkers about p; walkers violating the priors are redrawn from the valid ones with
    half the scatter until all are valid (mcmc_utils.py:46-72).  ln_prior(matrix, model) is called
    once per round on all walkers."""
    rng = np.random.default_rng() if rng is None else rng
    p = np.asarray(p, dtype=np.float64)
    scatter = np.asarray(scatter, dtype=np.float64) * np.ones_like(p)
    p0 = sample_ball(p, scatter * p, nwalkers, rng)
    if verbose:
        print('Initialising walkers...')
        print('Number of walkers currently invalid:')
    for _ in range(max_rounds):
        ok = np.isfinite(np.asarray(ln_prior(p0, model)))
        nbad = int((~ok).sum())
        if verbose:
            print(nbad)
        if nbad == 0:
            return p0
        good = p0[ok]
        if good.shape[0] == 0:
            raise RuntimeError("no walker satisfies the priors; check the starting position")
        repl = good[rng.integers(good.shape[0], size=nbad)]
        repl = repl + 0.5 * repl * scatter * rng.standard_normal(repl.shape)
        p0[~ok] = repl
    raise RuntimeError("walker initialisation did not converge")


class EnsembleSampler:
    """Affine-invariant ensemble sampler with the stretch move, emcee call shape.

    log_prob_fn(theta, *args): with vectorize=True theta is (n, ndim) and n values come back
    (one call per half-step); otherwise it is called per walker."""

    def __init__(self, nwalkers, ndim, log_prob_fn, args=(), kwargs=None, a=2.0, vectorize=False, pool=None, rng=None):
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("need an even number of walkers, at least twice the number of dimensions")
        self.nwalkers, self.ndim, self.a = nwalkers, ndim, float(a)
        self.log_prob_fn, self.args, self.kwargs = log_prob_fn, tuple(args), dict(kwargs or {})
        self.vectorize, self.pool = vectorize, pool
        self.rng = np.random.default_rng() if rng is None else rng
        self.reset()

    def reset(self):
        self._chain, self._lnprob = [], []
        self.naccepted = np.zeros(self.nwalkers)
        self.iterations = 0

    def compute_log_prob(self, coords):
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(coords, *self.args, **self.kwargs), dtype=np.float64)
        else:
            mapper = self.pool.map if self.pool is not None else map
            lp = np.asarray(list(mapper(lambda x: self.log_prob_fn(x, *self.args, **self.kwargs), coords)),
                            dtype=np.float64)
        if np.isnan(lp).any():
            raise ValueError("Probability function returned NaN")
        return lp

    def sample(self, initial_state, iterations=1, store=True, storechain=None, log_prob0=None, rstate0=None,
               skip_initial_state_check=True, **_):
        if storechain is not None:
            store = storechain
        pos = np.array(initial_state, dtype=np.float64, copy=True)
        if pos.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        lnp = self.compute_log_prob(pos) if log_prob0 is None else np.array(log_prob0, dtype=np.float64)
        half = self.nwalkers // 2
        for _ in range(iterations):
            perm = self.rng.permutation(self.nwalkers)
            for first, second in ((perm[:half], perm[half:]), (perm[half:], perm[:half])):
                s, c = pos[first], pos[second]
                zz = ((self.a - 1.0) * self.rng.random(half) + 1.0) ** 2 / self.a
                partner = c[self.rng.integers(half, size=half)]
                prop = partner - (partner - s) * zz[:, None]
                new_lnp = self.compute_log_prob(prop)
                lnpdiff = (self.ndim - 1.0) * np.log(zz) + new_lnp - lnp[first]
                accept = lnpdiff > np.log(self.rng.random(half))
                idx = first[accept]
                pos[idx] = prop[accept]
                lnp[idx] = new_lnp[accept]
                self.naccepted[idx] += 1
            self.iterations += 1
            if store:
                self._chain.append(pos.copy())
                self._lnprob.append(lnp.copy())
            yield pos, lnp, self.rng

    def run_mcmc(self, initial_state, nsteps, **kw):
        out = None
        for out in self.sample(initial_state, iterations=nsteps, **kw):
            pass
        return out

    @property
    def chain(self):
        """(nwalkers, nsteps, ndim)"""
        return np.swapaxes(np.asarray(self._chain), 0, 1) if self._chain else np.empty((self.nwalkers, 0, self.ndim))

    @property
    def lnprobability(self):
        return np.asarray(self._lnprob).T if self._lnprob else np.empty((self.nwalkers, 0))

    @property
    def flatchain(self):
        return self.chain.reshape(-1, self.ndim)

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iterations, 1)


def _sample_iter(sampler, startPos, nSteps, store, **kwargs):
    """sampler.sample(...) with whichever keyword the sampler knows for storing the chain: emcee 3 and the
    samplers here take store=, emcee 2 and ptemcee take storechain= (the reference tries one, then the other:
    mcmc_utils.py:119-129,199-239)."""
    if nSteps <= 0:
        return
    try:
        it = sampler.sample(startPos, iterations=nSteps, store=store, **kwargs)
        first = next(it)
    except TypeError:
        it = sampler.sample(startPos, iterations=nSteps, storechain=store, **kwargs)
        first = next(it)
    yield first
    yield from it


def run_burnin(sampler, startPos, nSteps, storechain=False, progress=False):
    """Advance nSteps without storing; returns (pos, prob, state) (mcmc_utils.py:114-132).  A parallel-tempered
    sampler returns (pos, logpost, logl) with a leading temperature axis."""
    out = None
    for out in _sample_iter(sampler, startPos, nSteps, storechain):
        pass
    return out[0], out[1], out[2]


def format_step(pos, prob):
    """One step of every walker in the reference's chain format: "{k:4d} {pos...} {lnprob:f}"
    (mcmc_utils.py:163-164), formatted natively (lfb_chain_format) in one call."""
    from . import _cabi
    rows = np.concatenate([np.asarray(pos, dtype=np.float64), np.asarray(prob, dtype=np.float64)[:, None]], axis=1)
    return _cabi.chain_text(rows[None]).decode()


def format_step_python(pos, prob):
    """The same lines with the reference's own expression (the checker of the native formatter)."""
    return "".join("{0:4d} {1:s} {2:f}\n".format(k, " ".join(map(str, pos[k])), prob[k]) for k in range(pos.shape[0]))


def run_mcmc_save(sampler, startPos, nSteps, rState, file, col_names='', progress=False, flush_every=16, **kwargs):
    """Run nSteps storing the chain and append it to `file` (mcmc_utils.py:135-183): the steps are
    formatted natively and appended `flush_every` steps at a time with one write, instead of the
    reference's one open() per walker per step (mcmc_utils.py:157-164); same bytes.  A DeviceSampler
    keeps the chain on the GPU between flushes (run_mcmc_save_device)."""
    from . import _cabi
    if hasattr(sampler, "run_block"):
        run_mcmc_save_device(sampler, startPos, nSteps, file, col_names=col_names, block=max(flush_every, 1), keep=False)
        return sampler
    if file:
        with open(file, "w") as f:
            f.write(col_names)
            if col_names:
                f.write("\n")
    pending = []

    def flush():
        if file and pending:
            _cabi.chain_append(file, np.asarray(pending))
        pending.clear()

    for pos, prob, state in _sample_iter(sampler, startPos, nSteps, True, **kwargs):
        pending.append(np.concatenate([pos, np.asarray(prob)[:, None]], axis=1))
        if len(pending) >= flush_every:
            flush()
    flush()
    return sampler


def flatchain(chain, npars=None, nskip=0, thin=1):
    """(nwalkers, nsteps, npars) -> (nwalkers * nsteps', npars), skipping and thinning steps."""
    if npars is None:
        npars = chain.shape[2]
    return chain[:, nskip::thin, :].reshape((-1, npars))


def readchain(file, nskip=0, thin=1):
    """Read a chain_prod.txt back as (nwalkers, nsteps, npars + 1) (mcmc_utils.py:252-272)."""
    with open(file) as f:
        first = f.readline().split()
    skip = 1 if first and first[0] == 'walker_no' else 0
    data = np.loadtxt(file, skiprows=skip, ndmin=2)
    nwalkers = int(data[:, 0].max()) + 1
    nprod = data.shape[0] // nwalkers
    npars = data.shape[1] - 1
    chain = data[:nprod * nwalkers, 1:].reshape((nprod, nwalkers, npars))
    return np.swapaxes(chain, 0, 1)[:, nskip::thin, :]


def initialise_walkers_pt(p, scatter, nwalkers, ntemps, ln_prior, model, rng=None, max_rounds=200, verbose=True):
    """A ball of walkers for every temperature, (ntemps, nwalkers, ndim); walkers violating the priors are
    redrawn from the valid ones of the whole set (mcmc_utils.py:75-111).  ln_prior(matrix, model) is called
    once per round on all ntemps x nwalkers rows."""
    p = np.asarray(p, dtype=np.float64)
    flat = initialise_walkers(p, scatter, nwalkers * ntemps, ln_prior, model, rng=rng, max_rounds=max_rounds,
                              verbose=verbose)
    return flat.reshape(ntemps, nwalkers, p.shape[0])


def default_beta_ladder(ndim, ntemps, Tmax=None):
    """Geometric ladder of inverse temperatures, beta_i = tstep ** -i.  ptemcee (not vendored in the reference
    tree, call site mcmcfit.py:264-270) takes tstep from a table tuned for 25 % swap acceptance between
    neighbouring temperatures of a Gaussian posterior and, beyond the table, from
    1 + 2 sqrt(ln 4) / sqrt(ndim); that expression is used here for every ndim [the table itself is not
    reproduced: pass `betas` for a pinned ladder].  With Tmax the ladder spans 1 .. Tmax in ntemps steps."""
    if ntemps < 1:
        raise ValueError("need at least one temperature")
    if Tmax is not None and ntemps > 1:
        tstep = float(Tmax) ** (1.0 / (ntemps - 1))
    else:
        tstep = 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)
    return tstep ** -np.arange(ntemps, dtype=np.float64)


class PTSampler:
    """Parallel-tempered affine-invariant ensemble sampler with ptemcee's call shape
    (ptemcee.sampler.Sampler(nwalkers, dim, logl, logp, loglargs=, logpargs=, ntemps=), mcmcfit.py:264-270;
    ptemcee is third-party and not vendored -- this is its published algorithm, Vousden, Farr & Mandel 2016,
    restated): every temperature runs the stretch move on the tempered posterior beta * logl + logp (walkers
    of even index proposed from those of odd index, then the reverse; z log-uniform on [1/a, a], acceptance
    z^dim), then neighbouring temperatures try to swap walkers, hottest pair first.  Fixed ladder
    (ptemcee's default when sample() is not asked to adapt).

    vectorize=True: logl and logp take a (n, dim) matrix and return n values -- all ntemps x nwalkers / 2
    proposals of a half-step are ONE call each (one CUDA pass); `joint`, if given, returns (logl, logp) of a
    matrix in one call instead.  Walkers outside the prior (logp = -inf) are given logl = 0, as ptemcee does."""

    def __init__(self, nwalkers, dim, logl, logp, ntemps=None, Tmax=None, betas=None, loglargs=(), logpargs=(),
                 loglkwargs=None, logpkwargs=None, a=2.0, vectorize=False, joint=None, rng=None, pool=None):
        if nwalkers % 2:
            raise ValueError("The number of walkers must be even.")
        if nwalkers < 2 * dim:
            raise ValueError("The number of walkers must be greater than 2*dimension.")
        self.nwalkers, self.dim, self.a = int(nwalkers), int(dim), float(a)
        self.logl, self.logp, self.joint = logl, logp, joint
        self.loglargs, self.logpargs = tuple(loglargs), tuple(logpargs)
        self.loglkwargs, self.logpkwargs = dict(loglkwargs or {}), dict(logpkwargs or {})
        self.vectorize, self.pool = vectorize, pool
        if betas is None:
            betas = default_beta_ladder(dim, ntemps if ntemps is not None else 1, Tmax)
        self._betas = np.array(betas, dtype=np.float64)
        self.ntemps = self._betas.shape[0]
        self._random = np.random.default_rng() if rng is None else rng
        self.reset()

    def reset(self):
        self._chain = []
        self._logposterior, self._loglikelihood = [], []
        self.nswap = np.zeros(self.ntemps)
        self.nswap_accepted = np.zeros(self.ntemps)
        self.nprop = np.zeros((self.ntemps, self.nwalkers))
        self.nprop_accepted = np.zeros((self.ntemps, self.nwalkers))
        self._p0 = self._logposterior0 = self._loglikelihood0 = None
        self.time = 0

    @property
    def betas(self):
        return self._betas

    def _tempered_likelihood(self, logl, betas=None):
        betas = self._betas if betas is None else betas
        with np.errstate(invalid='ignore'):
            out = logl * betas[:, None]
        out[np.isnan(out)] = -np.inf
        return out

    def _evaluate(self, ps):
        """(logl, logp) of positions ps (..., dim); walkers outside the prior are not asked for their likelihood."""
        shape = ps.shape[:-1]
        flat = np.ascontiguousarray(ps.reshape(-1, self.dim))
        if self.joint is not None:
            ll, lp = self.joint(flat)
            ll, lp = np.array(ll, dtype=np.float64), np.array(lp, dtype=np.float64)
        elif self.vectorize:
            lp = np.array(self.logp(flat, *self.logpargs, **self.logpkwargs), dtype=np.float64)
            ll = np.array(self.logl(flat, *self.loglargs, **self.loglkwargs), dtype=np.float64)
        else:
            mapper = self.pool.map if self.pool is not None else map
            lp = np.array(list(mapper(lambda x: self.logp(x, *self.logpargs, **self.logpkwargs), flat)), dtype=np.float64)
            ll = np.zeros_like(lp)
            ok = lp > -np.inf
            ll[ok] = list(mapper(lambda x: self.logl(x, *self.loglargs, **self.loglkwargs), flat[ok]))
        ll = np.where(lp == -np.inf, 0.0, ll)
        if np.isnan(lp).any():
            raise ValueError('Prior function returned NaN.')
        if np.isnan(ll).any():
            raise ValueError('Log likelihood function returned NaN.')
        return ll.reshape(shape), lp.reshape(shape)

    def sample(self, p0=None, iterations=1, thin=1, storechain=True, adapt=False, **kwargs):
        """Advance the chains; yields (p, logpost, logl) after every iteration, shapes (ntemps, nwalkers, dim),
        (ntemps, nwalkers), (ntemps, nwalkers)."""
        if 'store' in kwargs:
            # the reference calls sample(..., store=True) first and falls back to storechain= on any exception
            # (mcmc_utils.py:199,220): ptemcee raises here, so does this
            raise TypeError("sample() got an unexpected keyword argument 'store'")
        if adapt:
            raise NotImplementedError("adaptive temperature ladders are not implemented (ptemcee's default is a fixed ladder)")
        if p0 is None:
            if self._p0 is None:
                raise ValueError('Initial walker positions not specified.')
            p, logpost, logl = self._p0, self._logposterior0, self._loglikelihood0
        else:
            p = np.array(p0, dtype=np.float64, copy=True)
            if p.shape != (self.ntemps, self.nwalkers, self.dim):
                raise ValueError("incompatible input dimensions: expected (ntemps, nwalkers, dim)")
            logl, logp = self._evaluate(p)
            logpost = self._tempered_likelihood(logl) + logp
            if (logp == -np.inf).any():
                raise ValueError('Attempting to start with samples outside posterior support.')
        for i in range(iterations):
            for j in (0, 1):
                jupdate, jsample = j, (j + 1) % 2
                pupdate, psample = p[:, jupdate::2, :], p[:, jsample::2, :]
                half = self.nwalkers // 2
                zs = np.exp(self._random.uniform(low=-np.log(self.a), high=np.log(self.a), size=(self.ntemps, half)))
                qs = np.empty((self.ntemps, half, self.dim))
                for k in range(self.ntemps):
                    js = self._random.integers(0, half, size=half)
                    qs[k] = psample[k, js, :] + zs[k, :, None] * (pupdate[k] - psample[k, js, :])
                qslogl, qslogp = self._evaluate(qs)
                qslogpost = self._tempered_likelihood(qslogl) + qslogp
                with np.errstate(invalid='ignore'):
                    logpaccept = self.dim * np.log(zs) + qslogpost - logpost[:, jupdate::2]
                logr = np.log(self._random.uniform(low=0.0, high=1.0, size=(self.ntemps, half)))
                accepts = logr < logpaccept
                pupdate[accepts] = qs[accepts]          # (views: p is updated in place)
                logpost[:, jupdate::2][accepts] = qslogpost[accepts]
                logl[:, jupdate::2][accepts] = qslogl[accepts]
                self.nprop[:, jupdate::2] += 1.0
                self.nprop_accepted[:, jupdate::2] += accepts
            self._temperature_swaps(p, logpost, logl)
            self.time += 1
            if storechain and (i + 1) % thin == 0:
                self._chain.append(p.copy())
                self._logposterior.append(logpost.copy())
                self._loglikelihood.append(logl.copy())
            self._p0, self._logposterior0, self._loglikelihood0 = p, logpost, logl
            yield p, logpost, logl

    def _temperature_swaps(self, p, logpost, logl):
        """Neighbouring temperatures exchange walkers, hottest pair first: walker a of temperature i and walker b
        of temperature i - 1 swap with probability min(1, exp((beta_{i-1} - beta_i) (logl_a - logl_b)))."""
        for i in range(self.ntemps - 1, 0, -1):
            dbeta = self._betas[i - 1] - self._betas[i]
            iperm = self._random.permutation(self.nwalkers)
            i1perm = self._random.permutation(self.nwalkers)
            raccept = np.log(self._random.uniform(size=self.nwalkers))
            with np.errstate(invalid='ignore'):
                paccept = dbeta * (logl[i, iperm] - logl[i - 1, i1perm])
            self.nswap[i] += self.nwalkers
            self.nswap[i - 1] += self.nwalkers
            asel = paccept > raccept
            nacc = int(asel.sum())
            self.nswap_accepted[i] += nacc
            self.nswap_accepted[i - 1] += nacc
            a, b = iperm[asel], i1perm[asel]
            ptemp, ltemp, prtemp = p[i, a, :].copy(), logl[i, a].copy(), logpost[i, a].copy()
            p[i, a, :] = p[i - 1, b, :]
            logl[i, a] = logl[i - 1, b]
            logpost[i, a] = logpost[i - 1, b] - dbeta * logl[i - 1, b]
            p[i - 1, b, :] = ptemp
            logl[i - 1, b] = ltemp
            logpost[i - 1, b] = prtemp + dbeta * ltemp

    def run_mcmc(self, p0=None, iterations=1, **kw):
        out = None
        for out in self.sample(p0, iterations=iterations, **kw):
            pass
        return out

    @property
    def chain(self):
        """(ntemps, nwalkers, nsteps, dim)"""
        if not self._chain:
            return np.empty((self.ntemps, self.nwalkers, 0, self.dim))
        return np.moveaxis(np.asarray(self._chain), 0, 2)

    @property
    def flatchain(self):
        """(ntemps, nwalkers * nsteps, dim); the reference takes flatchain[0] (mcmcfit.py:328)."""
        c = self.chain
        return c.reshape(self.ntemps, -1, self.dim)

    @property
    def logprobability(self):
        return np.moveaxis(np.asarray(self._logposterior), 0, 2) if self._logposterior else np.empty((self.ntemps, self.nwalkers, 0))

    @property
    def loglikelihood(self):
        return np.moveaxis(np.asarray(self._loglikelihood), 0, 2) if self._loglikelihood else np.empty((self.ntemps, self.nwalkers, 0))

    @property
    def tswap_acceptance_fraction(self):
        return self.nswap_accepted / np.maximum(self.nswap, 1)

    @property
    def acceptance_fraction(self):
        return self.nprop_accepted / np.maximum(self.nprop, 1)


def run_ptmcmc_save(sampler, startPos, nSteps, file, progress=False, col_names='', flush_every=16, **kwargs):
    """Run a parallel-tempered chain and save the chain of the first temperature (beta = 1: the only one that
    samples the posterior) in the reference's format (mcmc_utils.py:186-239), `flush_every` steps per write."""
    from . import _cabi
    if file:
        with open(file, "w") as f:
            f.write(col_names)
            if col_names:
                f.write("\n")
    pending = []

    def flush():
        if file and pending:
            _cabi.chain_append(file, np.asarray(pending))
        pending.clear()

    for pos, prob, like in _sample_iter(sampler, startPos, nSteps, True, **kwargs):
        pending.append(np.concatenate([pos[0], np.asarray(prob[0])[:, None]], axis=1))
        if len(pending) >= flush_every:
            flush()
    flush()
    return sampler


class DeviceSampler:
    """emcee's stretch move with the ensemble resident in GPU memory:
    hand-written kernels (csrc/sampler.cuh: Philox draws + proposals, accept / update, chain record) around
    the engine's log-probability pass, through the C ABI's lfb_sampler_*.  A step moves nothing across PCIe;
    a small ensemble replays one captured CUDA graph per step.  Replaces emcee.EnsembleSampler + pool
    (/root/reference/mcmcfit.py:283-288) for the model set on `engine` (an _cabi.Engine or anything with
    `.engine`, e.g. a flatten.VectorModel)."""

    def __init__(self, engine, nwalkers, a=2.0, seed=0, what=2):
        import ctypes as C
        from . import _cabi
        self._C, self._cabi = C, _cabi
        self.engine = getattr(engine, "engine", engine)
        self._lib = _cabi.load()
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(self.engine.ndim), float(a)
        h = C.c_void_p()
        rc = self._lib.lfb_sampler_create(self.engine._h, self.nwalkers, self.a, int(seed) & (2 ** 64 - 1), int(what),
                                          C.byref(h))
        self.engine._check(rc, "lfb_sampler_create")
        self._s = h
        self._chain_cap = 0

    def close(self):
        if getattr(self, "_s", None):
            self._lib.lfb_sampler_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, pos, lnp=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        if pos.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        lnp = None if lnp is None else np.ascontiguousarray(lnp, dtype=np.float64)
        rc = self._lib.lfb_sampler_set_state(self._s, pos.ctypes.data, None if lnp is None else lnp.ctypes.data, None)
        self.engine._check(rc, "lfb_sampler_set_state")

    def run(self, nsteps, stream=None):
        """Enqueue nsteps full steps (asynchronous; get_state / read_chain synchronise)."""
        self.engine._check(self._lib.lfb_sampler_run(self._s, int(nsteps), stream), "lfb_sampler_run")

    def _state(self, want_pos=True):
        C = self._C
        pos = np.empty((self.nwalkers, self.ndim)) if want_pos else None
        lnp = np.empty(self.nwalkers)
        acc = np.empty(self.nwalkers, dtype=np.int64)
        it = C.c_longlong()
        rc = self._lib.lfb_sampler_get_state(self._s, pos.ctypes.data if want_pos else None, lnp.ctypes.data,
                                             acc.ctypes.data, C.byref(it))
        self.engine._check(rc, "lfb_sampler_get_state")
        return pos, lnp, acc, int(it.value)

    def get_state(self):
        pos, lnp, _, _ = self._state()
        return pos, lnp

    @property
    def naccepted(self):
        return self._state(False)[2]

    @property
    def iterations(self):
        return self._state(False)[3]

    @property
    def acceptance_fraction(self):
        _, _, acc, it = self._state(False)
        return acc / max(it, 1)

    def set_chain(self, steps):
        """Record the ensemble after every step in a device buffer holding `steps` steps (0: off)."""
        self.engine._check(self._lib.lfb_sampler_set_chain(self._s, int(steps)), "lfb_sampler_set_chain")
        self._chain_cap = int(steps)

    def read_chain(self):
        """The steps recorded since the last read, (steps, nwalkers, ndim + 1); empties the buffer."""
        C = self._C
        out = np.empty((max(self._chain_cap, 1), self.nwalkers, self.ndim + 1))
        n = C.c_longlong()
        self.engine._check(self._lib.lfb_sampler_read_chain(self._s, out.ctypes.data, C.byref(n)), "lfb_sampler_read_chain")
        return out[: int(n.value)]

    # -- the emcee call shape the reference's loops use (mcmc_utils.py:114-183) --
    def reset(self):
        pass

    def run_mcmc(self, initial_state, nsteps, log_prob0=None, **_):
        if initial_state is not None:
            self.set_state(initial_state, log_prob0)
        self.run(nsteps)
        pos, lnp = self.get_state()
        return pos, lnp, None

    def sample(self, initial_state, iterations=1, log_prob0=None, **_):
        """Step by step (one device -> host copy per step); run_block / run_mcmc_save are the fast paths."""
        if initial_state is not None:
            self.set_state(initial_state, log_prob0)
        for _ in range(iterations):
            self.run(1)
            pos, lnp = self.get_state()
            yield pos, lnp, None

    def run_block(self, nsteps, block=64):
        """nsteps recorded steps, read back `block` at a time: yields (steps, nwalkers, ndim + 1) arrays."""
        block = max(1, min(int(block), int(nsteps)))
        self.set_chain(block)
        done = 0
        while done < nsteps:
            k = min(block, nsteps - done)
            self.run(k)
            yield self.read_chain()
            done += k
        self.set_chain(0)


def run_mcmc_save_device(sampler, startPos, nSteps, file, col_names='', block=64, keep=True):
    """run_mcmc_save (mcmc_utils.py:135-183) for a DeviceSampler: the chain is recorded on the device,
    read back `block` steps at a time and appended with one native write per block -- the same bytes as
    the reference's one open() per walker per step.  Returns the chain (nwalkers, nsteps, ndim + 1) if keep."""
    from . import _cabi
    if file:
        with open(file, "w") as f:
            f.write(col_names)
            if col_names:
                f.write("\n")
    if startPos is not None:
        sampler.set_state(startPos)
    blocks = []
    for rows in sampler.run_block(nSteps, block):
        if file:
            _cabi.chain_append(file, rows)
        if keep:
            blocks.append(rows.copy())
    if keep and blocks:
        return np.swapaxes(np.concatenate(blocks, axis=0), 0, 1)
    return None
