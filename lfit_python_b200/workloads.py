"""Synthetic workloads of the shapes named in BASELINE.json.

Everything is derived from the example shipped with the reference,
/root/reference/test_data/mcmc_input.dat: its six eclipse blocks (lines 70-178)
are the "truth" parameter sets, its priors are the priors, `first_scatter = 0.10`
(line 13) is the walker scatter.  The values are restated here as data so that
nothing reads /root/reference at run time.

A workload is a flattened model tree (the arrays the C ABI's set_layout /
set_priors / set_lightcurves take) plus a walker generator.  The light-curve
model that produces the synthetic fluxes is passed in as a callable, so the same
generator serves the CUDA engine and the CPU oracle legs of bench.py.
"""
import numpy as np

NPAR = 18
CV_NAMES = ["wdFlux", "dFlux", "sFlux", "rsFlux", "q", "dphi", "rdisc", "ulimb", "rwd", "scale", "az",
            "fis", "dexp", "phi0", "exp1", "exp2", "tilt", "yaw"]
CORE_NAMES = ["q", "dphi", "rwd"]
BAND_NAMES = ["wdFlux", "rsFlux", "ulimb"]
ECL_SIMPLE = ["dFlux", "sFlux", "rdisc", "scale", "az", "fis", "dexp", "phi0"]
ECL_COMPLEX = ECL_SIMPLE + ["exp1", "exp2", "yaw", "tilt"]  # node order (CVModel.py:376-380)
PRIOR_CODES = {"gauss": 0, "gaussPos": 1, "uniform": 2, "log_uniform": 3, "mod_jeff": 4}

# (value, prior, p1, p2) -- test_data/mcmc_input.dat:48-50
CORE = {"q": (0.1037, "uniform", 0.03, 0.5), "dphi": (0.0392, "uniform", 0.01, 0.1),
        "rwd": (0.0187, "uniform", 0.001, 0.1)}
# test_data/mcmc_input.dat:53-63
BANDS = [
    {"wdFlux": (0.0528, "uniform", 0.001, 0.2), "rsFlux": (0.0131, "uniform", 0.001, 0.2), "ulimb": (0.284, "gauss", 0.284, 0.001)},
    {"wdFlux": (0.0508, "uniform", 0.001, 0.2), "rsFlux": (0.0262, "uniform", 0.001, 0.2), "ulimb": (0.284, "gauss", 0.284, 0.001)},
    {"wdFlux": (0.0324, "uniform", 0.001, 0.2), "rsFlux": (0.0262, "uniform", 0.001, 0.2), "ulimb": (0.284, "gauss", 0.284, 0.001)},
]
_ECL_PRIORS = {"dFlux": ("uniform", 0.001, 0.2), "sFlux": ("uniform", 0.001, 0.2), "rdisc": ("uniform", 0.2, 0.7),
               "scale": ("log_uniform", 0.001, 0.2), "az": ("uniform", 50.0, 175.0), "fis": ("uniform", 0.001, 1.0),
               "dexp": ("log_uniform", 0.001, 2.0), "phi0": ("uniform", -0.2, 0.2), "exp1": ("uniform", 0.001, 5.0),
               "exp2": ("uniform", 0.5, 5.0), "yaw": ("uniform", -90.0, 90.0), "tilt": ("uniform", 0.001, 180.0)}
# test_data/mcmc_input.dat:75-178 (dFlux sFlux rdisc scale az fis dexp phi0 exp1 exp2 yaw tilt)
_ECL_VALUES = [
    (0.0707, 0.0613, 0.2953, 0.0430, 120.0000, 0.0480, 0.5000, 0.0010, 1.1342, 4.5971, 5.4000, 72.0006),
    (0.1238, 0.1518, 0.5214, 0.0497, 122.0724, 0.1684, 1.9539, -0.0013, 3.4876, 1.4429, 15.6635, 52.4720),
    (0.0496, 0.0631, 0.5487, 0.0410, 125.1563, 0.0467, 0.7073, -0.0004, 2.7938, 1.2241, -1.3957, 49.0896),
    (0.0938, 0.0618, 0.4212, 0.0430, 125.1298, 0.1099, 1.0794, -0.0002, 0.2859, 0.9702, -2.4394, 52.2358),
    (0.1267, 0.1084, 0.5954, 0.0487, 99.1185, 0.0317, 1.4333, 0.0004, 2.9879, 1.2802, 22.7125, 138.7462),
    (0.0845, 0.0697, 0.5702, 0.0191, 120.9071, 0.0506, 1.7123, -0.0010, 3.0050, 1.3828, 8.6757, 111.5294),
]
ECLIPSES = [{n: (v,) + _ECL_PRIORS[n] for n, v in zip(ECL_COMPLEX, vals)} for vals in _ECL_VALUES]


def log_uniform_norm(p1, p2):
    """Prior.normalise for log_uniform as the reference computes it (model.py:77-79):
    |integral of ln(1/x) over (p1, p2)| -- the integral of ln_prob, not of the pdf."""
    f = lambda x: x - x * np.log(x)
    return abs(f(p2) - f(p1))


def prior_norm(ptype, p1, p2):
    if ptype == "log_uniform":
        return log_uniform_norm(max(p1, 1.0e-30), p2)
    if ptype == "mod_jeff":
        return float(np.log((p1 + p2) / p1))
    return 1.0


class Workload:
    """Flattened tree + light curves + walker generator."""

    def __init__(self, name, n_bands, ecl_per_band, n_ph, complex_bs=True, phase_range=(-0.5, 0.5), sigma=0.004):
        self.name = name
        self.complex_bs = bool(complex_bs)
        self.npars = 18 if complex_bs else 14
        self.n_ph = int(n_ph)
        self.sigma = float(sigma)
        ecl_names = ECL_COMPLEX if complex_bs else ECL_SIMPLE
        if isinstance(ecl_per_band, int):
            ecl_per_band = [ecl_per_band] * n_bands
        names, p0, pri = [], [], []

        def add(par, label, spec):
            names.append("%s_%s" % (par, label))
            p0.append(spec[0])
            pri.append(spec[1:])
            return len(names) - 1

        col = {}
        for par in CORE_NAMES:
            col[par] = add(par, "core", CORE[par])
        gather = []
        band_of = []
        k = 0
        for b in range(n_bands):
            bcol = {par: add(par, "b%d" % b, BANDS[b % len(BANDS)][par]) for par in BAND_NAMES}
            for _ in range(ecl_per_band[b]):
                spec = ECLIPSES[k % len(ECLIPSES)]
                ecol = {par: add(par, "e%d" % k, spec[par]) for par in ecl_names}
                allc = dict(col)
                allc.update(bcol)
                allc.update(ecol)
                gather.append([allc.get(nm, 0) for nm in CV_NAMES])  # unused slots (simple BS) -> 0
                band_of.append(b)
                k += 1
        self.names = names
        self.ndim = len(names)
        self.n_ecl = len(gather)
        self.gather = np.asarray(gather, dtype=np.int32)
        self.band_of = np.asarray(band_of, dtype=np.int32)
        self.consts = np.zeros(0)
        self.p0 = np.asarray(p0, dtype=np.float64)
        self.prior_src = np.arange(self.ndim, dtype=np.int32)
        self.prior_type = np.asarray([PRIOR_CODES[p[0]] for p in pri], dtype=np.int32)
        self.prior_p1 = np.asarray([p[1] for p in pri], dtype=np.float64)
        self.prior_p2 = np.asarray([p[2] for p in pri], dtype=np.float64)
        self.prior_norm = np.asarray([prior_norm(*p) for p in pri], dtype=np.float64)
        self.prior_isvar = np.ones(self.ndim, dtype=np.int32)
        # phases: uniform grid; width = mean(diff(phase))/2 for every point (CVModel.py:64)
        x = np.linspace(phase_range[0], phase_range[1], self.n_ph)
        w = np.mean(np.diff(x)) * np.ones_like(x) / 2.0
        self.lc_off = np.arange(self.n_ecl + 1, dtype=np.int64) * self.n_ph
        self.lc_phase = np.tile(x, self.n_ecl)
        self.lc_width = np.tile(w, self.n_ecl)
        self.lc_y = None
        self.lc_ye = np.full(self.n_ecl * self.n_ph, self.sigma)
        # per-parameter scatter multipliers (mcmcfit.py:208-246)
        self.scatter_mult = np.ones(self.ndim)
        for i, nm in enumerate(names):
            par = nm.split("_")[0]
            if par == "dphi":
                self.scatter_mult[i] = 0.2
            elif par == "ulimb":
                self.scatter_mult[i] = 1e-6

    def cv_pars(self, theta, e):
        """The CV parameter list of eclipse e for one parameter vector (CVModel.py:335-354)."""
        return np.asarray(theta)[self.gather[e][: self.npars]]

    def make_data(self, model_fn, seed=12345):
        """y = model(truth) + N(0, sigma); model_fn(pars, phase, width) -> flux."""
        rng = np.random.default_rng(seed)
        y = np.empty(self.n_ecl * self.n_ph)
        for e in range(self.n_ecl):
            sl = slice(self.lc_off[e], self.lc_off[e + 1])
            f = np.asarray(model_fn(self.cv_pars(self.p0, e), self.lc_phase[sl], self.lc_width[sl]))
            if not np.all(np.isfinite(f)):
                raise RuntimeError("truth parameters of eclipse %d give no model" % e)
            y[sl] = f + rng.normal(0.0, self.sigma, self.n_ph)
        self.lc_y = y
        return y

    def make_noise_only_data(self, seed=12345):
        """Cheap stand-in fluxes (timing legs that must not call any model): 0.2 + noise."""
        rng = np.random.default_rng(seed)
        self.lc_y = 0.2 + rng.normal(0.0, self.sigma, self.n_ecl * self.n_ph)
        return self.lc_y

    def walkers(self, n, ln_prior_fn=None, scatter=0.10, seed=2024, max_rounds=2000):
        """emcee.utils.sample_ball(p0, scatter*p0) then the resampling loop of
        mcmc_utils.initialise_walkers (mcmc_utils.py:46-72), with ln_prior_fn(theta[n, ndim]) -> (n,)."""
        rng = np.random.default_rng(seed)
        std = scatter * self.scatter_mult * self.p0
        p = self.p0 + std * rng.standard_normal((n, self.ndim))
        if ln_prior_fn is None:
            return p
        for _ in range(max_rounds):
            ok = np.isfinite(ln_prior_fn(p))
            nbad = int((~ok).sum())
            if nbad == 0:
                return p
            good = p[ok]
            if good.shape[0] == 0:
                raise RuntimeError("no walker satisfies the priors")
            rows = rng.integers(good.shape[0], size=nbad)
            repl = good[rows]
            repl = repl + 0.5 * repl * (scatter * self.scatter_mult) * rng.standard_normal(repl.shape)
            p[~ok] = repl
        raise RuntimeError("walker initialisation did not converge")

    def apply(self, engine):
        """Push the layout into a lfit_python_b200._cabi.Engine."""
        engine.set_layout(self.ndim, self.npars, self.gather, self.consts)
        engine.set_priors(self.prior_src, self.prior_type, self.prior_p1, self.prior_p2, self.prior_norm,
                          self.prior_isvar)
        if self.lc_y is not None:
            engine.set_lightcurves(self.lc_off, self.lc_phase, self.lc_width, self.lc_y, self.lc_ye)


    def apply_gp(self, engine, ln_hyper=(-9.5118, -9.0953, -2.2131)):
        """The same layout judged by the Gaussian process (GPLCModel, CVModel.py:494-711): the three
        ln hyper-parameters as fixed Params (defaults: test_data/mcmc_input.dat:43-45), change points
        from the truth parameters."""
        consts = np.concatenate([self.consts, np.asarray(ln_hyper, dtype=np.float64)])
        nc = self.consts.shape[0]
        engine.set_layout(self.ndim, self.npars, self.gather, consts)
        engine.set_priors(self.prior_src, self.prior_type, self.prior_p1, self.prior_p2, self.prior_norm,
                          self.prior_isvar)
        engine.set_lightcurves(self.lc_off, self.lc_phase, self.lc_width, self.lc_y, self.lc_ye)
        q, dphi, rwd = (self.p0[self.names.index(n + "_core")] for n in ("q", "dphi", "rwd"))
        from . import _cabi
        inc, ok = engine.roche(_cabi.ROCHE_FINDI, q, dphi)
        ph, ok2 = engine.wdphases(q, inc[0, 0], rwd, 10)
        if not (ok.all() and ok2.all()):
            raise RuntimeError("no change points for the truth parameters")
        dist = (dphi + (ph[0, 1] - ph[0, 0])) / 2.0
        engine.set_gp(np.asarray([-(nc + 1), -(nc + 2), -(nc + 3)], dtype=np.int32), np.full(self.n_ecl, dist))
        return dist


def config(idx, **over):
    """The five BASELINE.json configurations (index 0..4)."""
    specs = [
        dict(name="C1: 1 eclipse, simple BS, 300 pts", n_bands=1, ecl_per_band=1, n_ph=300, complex_bs=False,
             phase_range=(-0.2, 0.3), walkers=50),
        dict(name="C2: 1 eclipse, complex BS, 2000 pts", n_bands=1, ecl_per_band=1, n_ph=2000, complex_bs=True,
             walkers=4096),
        dict(name="C3: 3 bands x 8 eclipses, 1000 pts", n_bands=3, ecl_per_band=8, n_ph=1000, complex_bs=True,
             walkers=8192),
        dict(name="C4: 20 eclipses, 1000 pts", n_bands=3, ecl_per_band=[7, 7, 6], n_ph=1000, complex_bs=True,
             walkers=65536),
        dict(name="C5: 1 eclipse, 5000 pts, 4x disc/BS density", n_bands=1, ecl_per_band=1, n_ph=5000,
             complex_bs=True, walkers=16384, grid=dict(n_disc_r=50, n_disc_th=80, n_bs=800)),
    ]
    s = dict(specs[idx])
    s.update(over)
    nwalk = s.pop("walkers")
    grid = s.pop("grid", {})
    wl = Workload(**s)
    wl.n_walkers = nwalk
    wl.grid = grid
    return wl
