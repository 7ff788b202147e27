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


def run_burnin(sampler, startPos, nSteps, storechain=False, progress=False):
    """Advance nSteps without storing; returns (pos, prob, state) (mcmc_utils.py:114-132)."""
    out = None
    try:
        it = sampler.sample(startPos, iterations=nSteps, store=storechain)
    except TypeError:
        it = sampler.sample(startPos, iterations=nSteps, storechain=storechain)
    for out in it:
        pass
    return out[0], out[1], out[2]


def format_step(pos, prob):
    """One step of every walker in the reference's chain format: "{k:4d} {pos...} {lnprob:f}"
    (mcmc_utils.py:163-164)."""
    return "".join("{0:4d} {1:s} {2:f}\n".format(k, " ".join(map(str, pos[k])), prob[k]) for k in range(pos.shape[0]))


def run_mcmc_save(sampler, startPos, nSteps, rState, file, col_names='', progress=False, **kwargs):
    """Run nSteps storing the chain; append every step to `file` (one write per step instead of
    the reference's one open() per walker per step, mcmc_utils.py:157-164; same bytes)."""
    if file:
        with open(file, "w") as f:
            f.write(col_names)
            if col_names:
                f.write("\n")
    for pos, prob, state in sampler.sample(startPos, iterations=nSteps, store=True, **kwargs):
        if file:
            with open(file, 'a') as f:
                f.write(format_step(pos, prob))
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


class DeviceEnsembleSampler:
    """The same stretch move with the ensemble resident in GPU memory.

    Positions, log-probabilities, proposals and the accept/reject step live in torch CUDA tensors
    (plumbing only); the log-probability is the CUDA engine called through raw device pointers, so
    a step moves nothing across PCIe.  `vec` is a flatten.VectorModel (or anything with `.engine`
    and `.ndim`)."""

    def __init__(self, nwalkers, vec, a=2.0, seed=None, what=2):
        import torch
        if nwalkers % 2 or nwalkers < 2 * vec.ndim:
            raise ValueError("need an even number of walkers, at least twice the number of dimensions")
        self.torch = torch
        self.engine, self.ndim, self.nwalkers, self.a, self.what = vec.engine, vec.ndim, nwalkers, float(a), what
        self.device = torch.device("cuda", self.engine.device)
        self.gen = torch.Generator(device=self.device)
        if seed is not None:
            self.gen.manual_seed(int(seed))
        self.stream = torch.cuda.Stream(device=self.device)
        self.pos = self.lnp = None
        self.naccepted = torch.zeros(nwalkers, dtype=torch.int64, device=self.device)
        self.iterations = 0

    def _log_prob(self, theta, out=None):
        torch = self.torch
        theta = theta.contiguous()
        if out is None:
            out = torch.empty(theta.shape[0], dtype=torch.float64, device=self.device)
        self.engine.log_prob_device(theta.data_ptr(), theta.shape[0], out.data_ptr(), what=self.what,
                                    stream=self.stream.cuda_stream)
        return out

    def run_mcmc(self, initial_state, nsteps):
        """Advance nsteps; returns (positions, log-probabilities) as host arrays."""
        torch = self.torch
        half = self.nwalkers // 2
        with torch.cuda.stream(self.stream):
            if initial_state is not None:
                self.pos = torch.as_tensor(np.asarray(initial_state, dtype=np.float64)).to(self.device)
                self.lnp = self._log_prob(self.pos)
            pos, lnp = self.pos, self.lnp
            # proposals and their log-probabilities live in fixed buffers: identical calls, which the engine
            # replays as a CUDA graph when the ensemble is small
            if getattr(self, "_prop", None) is None:
                self._prop = torch.empty((half, self.ndim), dtype=torch.float64, device=self.device)
                self._new_lnp = torch.empty(half, dtype=torch.float64, device=self.device)
            prop, new_lnp = self._prop, self._new_lnp
            for _ in range(nsteps):
                perm = torch.randperm(self.nwalkers, device=self.device, generator=self.gen)
                for first, second in ((perm[:half], perm[half:]), (perm[half:], perm[:half])):
                    s, c = pos[first], pos[second]
                    u = torch.rand(half, dtype=torch.float64, device=self.device, generator=self.gen)
                    zz = ((self.a - 1.0) * u + 1.0) ** 2 / self.a
                    partner = c[torch.randint(half, (half,), device=self.device, generator=self.gen)]
                    torch.sub(partner, (partner - s) * zz[:, None], out=prop)
                    self._log_prob(prop, out=new_lnp)
                    lnpdiff = (self.ndim - 1.0) * torch.log(zz) + new_lnp - lnp[first]
                    accept = lnpdiff > torch.log(torch.rand(half, dtype=torch.float64, device=self.device,
                                                            generator=self.gen))
                    # masked updates with fixed shapes: nothing here makes the host wait for the GPU
                    pos[first] = torch.where(accept[:, None], prop, s)
                    lnp[first] = torch.where(accept, new_lnp, lnp[first])
                    self.naccepted[first] += accept.to(torch.int64)
                self.iterations += 1
            self.pos, self.lnp = pos, lnp
        self.stream.synchronize()
        return pos.cpu().numpy(), lnp.cpu().numpy()

    @property
    def acceptance_fraction(self):
        return self.naccepted.cpu().numpy() / max(self.iterations, 1)
