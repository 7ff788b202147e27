"""Gaussian-process likelihood row (CVModel.py:494-711).

CPU part: the oracle's dense statement of the kernel, the change-point rule, the device Kalman
filter compiled for the host (tests/host_gp_harness.cpp) against the dense Cholesky, and the
GP tree classes.  GPU part: the C ABI (lfb_gp_loglike, lfb_wdphases, lfb_set_gp + lfb_log_prob)
against the oracle, and the scalar tree walk against the batched call.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from lfit_python_b200.CVModel import (ComplexGPEclipse, GPLCModel, SimpleGPEclipse, construct_model)

from test_tree import write_input

HERE = os.path.dirname(os.path.abspath(__file__))
GP_EXTRA = """ln_ampin_gp = -9.99 uniform -25.0 -1.0 1
ln_ampout_gp = -9.5 uniform -25.0 -1.0 1
ln_tau_gp = -5.5 uniform -10.0 -1.0 1
"""


def gp_input(tmp_path, complex=1):
    path = write_input(tmp_path, complex=complex, extra=GP_EXTRA)
    text = open(path).read().replace("useGP = 0", "useGP = 1")
    open(path, "w").write(text)
    return path


@pytest.fixture(scope="module")
def hostgp(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostgp") / "libhostgp.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "host_gp_harness.cpp")],
                   check=True)
    lib = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    for fn in (lib.host_gp_loglike, lib.host_gp_loglike_two_sided):
        fn.restype = C.c_double
        fn.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, dp]
    lib.host_gp_changepoints.argtypes = [C.c_double] * 4 + [dp]

    def loglike(x, ye, r, ampin, ampout, tau, gaps, two_sided=False):
        x, ye, r = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, ye, r))
        g = np.ascontiguousarray(np.asarray(gaps, dtype=np.float64).ravel())
        if g.size == 0:
            g = np.zeros(2)
        fn = lib.host_gp_loglike_two_sided if two_sided else lib.host_gp_loglike
        return fn(len(x), x.ctypes.data_as(dp), ye.ctypes.data_as(dp), r.ctypes.data_as(dp), ampin, ampout, tau, len(gaps),
                  g.ctypes.data_as(dp))

    def changepoints(xmin, xmax, dist, phi0):
        g = np.zeros(16)
        n = lib.host_gp_changepoints(xmin, xmax, dist, phi0, g.ctypes.data_as(dp))
        return g[:2 * n].reshape(n, 2)

    return loglike, changepoints


def random_case(rng, n=None):
    n = int(rng.integers(2, 300)) if n is None else n
    lo, hi = ((-0.5, 0.5), (-0.2, 0.3), (0.05, 0.45), (-1.2, 1.3))[int(rng.integers(4))]
    x = np.sort(rng.uniform(lo, hi, n))
    ye = rng.uniform(0.002, 0.006, n)
    hyper = np.exp([rng.uniform(-12, -7), rng.uniform(-12, -7), rng.uniform(-8, -3)])
    gaps = O.gp_changepoints(x, rng.uniform(0.02, 0.08), rng.normal(0, 0.002))
    K = O.gp_kernel_matrix(x, ye, *hyper, gaps)
    r = np.linalg.cholesky(K) @ rng.standard_normal(n)
    return x, ye, r, hyper, gaps


def test_kernel_matrix_entries():
    x = np.array([-0.3, -0.1, 0.0, 0.2, 0.4])
    ye = np.full(5, 0.01)
    gaps = [[-0.96, -0.04], [0.04, 0.96]]
    K = O.gp_kernel_matrix(x, ye, 2.0, 3.0, 0.01, gaps)
    m = lambda d: (1 + np.sqrt(3 * d * d / 0.01)) * np.exp(-np.sqrt(3 * d * d / 0.01))
    assert K[0, 1] == pytest.approx((2.0 + 3.0) * m(0.2))      # same gap: both amplitudes
    assert K[0, 3] == pytest.approx(2.0 * m(0.5))              # different gaps: global only
    assert K[1, 2] == pytest.approx(2.0 * m(0.1))              # in eclipse: global only
    assert K[2, 2] == pytest.approx(2.0 + 1e-4 + O.GP_WHITE_NOISE)
    assert K[3, 3] == pytest.approx(5.0 + 1e-4 + O.GP_WHITE_NOISE)
    # one point: a plain Gaussian
    ll = O.gp_log_like([0.0], [0.1], [0.3], 2.0, 3.0, 0.01, gaps)
    v = 2.0 + 0.01 + O.GP_WHITE_NOISE
    assert ll == pytest.approx(-0.5 * (0.09 / v + np.log(2 * np.pi * v)))
    assert O.gp_log_like(x, ye, [0, 0, np.nan, 0, 0], 2.0, 3.0, 0.01, gaps) == -np.inf
    assert O.gp_log_like(x, ye, np.zeros(5), 2.0, -3.0, 0.01, gaps) == -np.inf


def test_changepoints_rule(hostgp):
    _, cp = hostgp
    x = np.linspace(-0.5, 0.5, 11)
    g = O.gp_changepoints(x, 0.03, 0.001)             # cycles 0 and 1 (CVModel.py:580-599)
    assert np.allclose(g, [[-1 + 0.031, -0.029], [0.031, 1 - 0.029]])
    assert O.gp_changepoints(np.array([0.1, 0.4]), 0.03, 0.0) == [[0.03, 0.97]]
    assert len(O.gp_changepoints(np.array([-1.2, 1.3]), 0.03, 0.0)) == 4
    for lo, hi in ((-0.5, 0.5), (0.1, 0.4), (-1.2, 1.3), (-0.2, 0.3), (0.0, 1.0), (-1.0, 0.0)):
        want = np.asarray(O.gp_changepoints(np.array([lo, hi]), 0.04, -0.002)).reshape(-1, 2)
        assert np.array_equal(cp(lo, hi, 0.04, -0.002), want)


def test_kalman_filter_equals_dense_cholesky(hostgp):
    loglike, _ = hostgp
    rng = np.random.default_rng(7)
    for trial in range(40):
        x, ye, r, hyper, gaps = random_case(rng)
        if trial % 5 == 0:
            x[len(x) // 2:] = np.maximum(x[len(x) // 2:], x[len(x) // 2])  # keep ascending, make duplicates
            x[1] = x[0]
        if trial % 7 == 0:
            gaps = []
        if trial % 9 == 0 and len(x) > 4 and gaps:
            x[3] = gaps[0][1]                                              # a point exactly on a gap edge
            x.sort()
        a = O.gp_log_like(x, ye, r, *hyper, gaps)
        b = loglike(x, ye, r, *hyper, gaps)
        assert np.isfinite(a) and b == pytest.approx(a, rel=1e-10, abs=1e-9)
        # two filters meeting in the middle (what the kernel runs on two lanes): in a gap or in eclipse
        assert loglike(x, ye, r, *hyper, gaps, two_sided=True) == pytest.approx(a, rel=1e-10, abs=1e-9)
    x, ye, r, hyper, gaps = random_case(rng, n=1)
    assert loglike(x, ye, r, *hyper, gaps) == pytest.approx(O.gp_log_like(x, ye, r, *hyper, gaps), rel=1e-13)
    r[0] = np.inf
    assert loglike(x, ye, r, *hyper, gaps) == -np.inf
    assert loglike(x, ye, np.zeros(1), hyper[0], hyper[1], 0.0, gaps) == -np.inf


def test_gp_tree_structure(tmp_path):
    m = construct_model(gp_input(tmp_path))
    assert isinstance(m, GPLCModel) and m.node_par_names[-3:] == ('ln_ampin_gp', 'ln_ampout_gp', 'ln_tau_gp')
    ecl = list(m.search_node_type("Eclipse"))
    assert len(ecl) == 2 and all(isinstance(e, ComplexGPEclipse) for e in ecl)
    assert ecl[0].cv_parnames[-2:] == ['tilt', 'yaw'] and len(ecl[0].cv_parlist) == 18
    names = [p.name for p in m.__get_descendant_params__()[0]]
    assert names[:6] == ['q', 'dphi', 'rwd', 'ln_ampin_gp', 'ln_ampout_gp', 'ln_tau_gp']
    m2 = construct_model(gp_input(tmp_path, complex=0))
    assert all(type(e) is SimpleGPEclipse for e in m2.search_node_type("Eclipse"))


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_gp_loglike_matches_dense(engine):
    rng = np.random.default_rng(11)
    for trial in range(6):
        x, ye, r, hyper, gaps = random_case(rng, n=int(rng.integers(50, 500)))
        n_sets = 5
        R = np.stack([r * s for s in (1.0, 0.5, 2.0, -1.0, 0.0)])
        H = np.stack([hyper * f for f in (1.0, 1.3, 0.7, 1.0, 2.0)])
        Gp = np.stack([np.asarray(gaps) + d for d in (0.0, 0.001, -0.002, 0.0, 0.003)])
        got = engine.gp_loglike(x, ye, R, H, Gp)
        want = [O.gp_log_like(x, ye, R[i], *H[i], Gp[i].tolist()) for i in range(n_sets)]
        assert np.allclose(got, want, rtol=1e-10, atol=1e-8)
    R[1, 3] = np.nan
    H[2, 2] = -1.0
    got = engine.gp_loglike(x, ye, R, H, Gp)
    assert got[1] == -np.inf and got[2] == -np.inf and np.isfinite(got[0])
    with pytest.raises(Exception):
        engine.gp_loglike(x[::-1].copy(), ye, R, H, Gp)


@pytest.mark.gpu
def test_gpu_wdphases_matches_oracle(engine):
    from lfit_python_b200 import roche
    for q, dphi, rwd in ((0.1037, 0.0392, 0.0187), (0.3, 0.06, 0.012), (0.05, 0.03, 0.02)):
        inc = O.findi(q, dphi)
        want = O.wdphases(q, inc, rwd, 10)
        got = roche.wdphases(q, inc, rwd, ntheta=10)
        assert np.allclose(got, want, rtol=0, atol=1e-10)
        assert 0 < got[0] < dphi / 2 < got[1]                      # egress straddles the centre's egress
    with pytest.raises(roche.RocheError):
        roche.wdphases(0.1, 60.0, 0.02, ntheta=10)                  # no eclipse at this inclination


@pytest.mark.gpu
def test_gp_tree_vector_path_matches_oracle_and_scalar_path(tmp_path, engine):
    from lfit_python_b200 import mcmcfit
    from lfit_python_b200.flatten import VectorModel
    m = construct_model(gp_input(tmp_path))
    vec = VectorModel(m)
    L = vec.layout
    assert L.gp and L.ndim == len(m.dynasty_par_vals)
    ecl = L.eclipses
    dist = [O.gp_dist_cp(0.1037, 0.0392, 0.0187, 10)] * 2
    assert np.allclose(L.gp_dist, dist, rtol=0, atol=1e-10)
    p0 = np.asarray(m.dynasty_par_vals)
    rng = np.random.default_rng(3)
    theta = p0 * (1 + 0.01 * rng.standard_normal((12, L.ndim)))
    theta[0] = p0
    theta[3, L.names.index("q_core")] = 0.9                         # invalid walker
    got = vec.ln_like(theta)
    names = L.names
    for k in range(12):
        m.dynasty_par_vals = theta[k]
        want = 0.0
        for e in ecl:
            pars = np.asarray(e.cv_parlist)
            st, flx = O.calc_flux(pars, e.lc.x, e.lc.w)
            pd = e.ancestor_param_dict
            hyper = [np.exp(pd[n].currVal) for n in ('ln_ampin_gp', 'ln_ampout_gp', 'ln_tau_gp')]
            if st != 0 or np.any(np.isnan(flx)):
                want = -np.inf
                break
            order = np.argsort(e.lc.x, kind='stable')
            gaps = O.gp_changepoints(e.lc.x, dist[0], pd['phi0'].currVal)
            want += O.gp_log_like(e.lc.x[order], e.lc.ye[order], (e.lc.y - flx)[order], *hyper, gaps)
        if np.isfinite(want):
            assert got[k] == pytest.approx(want, rel=1e-8), k
            assert mcmcfit.ln_like(theta[k], m) == pytest.approx(want, rel=1e-8)   # scalar tree walk
        else:
            assert got[k] == -np.inf
    m.dynasty_par_vals = p0
    lp = vec.ln_prob(theta)
    assert np.array_equal(np.isfinite(lp), np.isfinite(vec.ln_prior(theta)) & np.isfinite(got))
    fin = np.isfinite(lp)
    assert np.allclose(lp[fin], (vec.ln_prior(theta) + got)[fin], rtol=1e-12)
    assert np.allclose(-0.5 * vec.chisq(theta[:3]).sum(axis=1), got[:3], rtol=1e-12)
    # back to chi-squared on the same engine
    vec.engine.set_gp()
    chi = vec.ln_like(theta[:2])
    assert np.all(chi != got[:2])


@pytest.mark.gpu
@pytest.mark.parametrize("use_gp", [0, 1])
def test_driver_runs_end_to_end(tmp_path, monkeypatch, capsys, use_gp):
    """python -m lfit_python_b200.mcmcfit input.dat: burn-in, production chain, chain file (mcmcfit.py:162-330)."""
    from lfit_python_b200 import mcmcfit, mcmc_utils
    path = gp_input(tmp_path) if use_gp else write_input(tmp_path)
    txt = open(path).read().replace("fit = 0", "fit = 1").replace("nburn = 10", "nburn = 3")
    txt = txt.replace("nprod = 10", "nprod = 4").replace("nwalkers = 40", "nwalkers = 96")
    open(path, "w").write(txt)
    monkeypatch.chdir(tmp_path)
    sampler = mcmcfit.main([path, "--seed", "3"])
    out = capsys.readouterr().out
    assert "Initial guess has a chisq of" in out and "Starting the main MCMC chain." in out
    ndim = sampler.chain.shape[-1]
    assert sampler.chain.shape == (96, 4, ndim)
    assert np.all(np.isfinite(sampler.lnprobability))
    chain = mcmc_utils.readchain(str(tmp_path / "chain_prod.txt"))
    assert chain.shape == (96, 4, ndim + 1)
    assert np.allclose(chain[:, -1, :ndim], sampler.chain[:, -1, :], rtol=1e-5, atol=1e-6)   # "%f"-style file precision


@pytest.mark.gpu
def test_driver_runs_parallel_tempering(tmp_path, monkeypatch, capsys):
    """usePT = 1 (mcmcfit.py:251-270,319-328): ntemps x nwalkers rows per half-step in one CUDA pass; the chain
    file holds the first temperature."""
    from lfit_python_b200 import mcmcfit, mcmc_utils
    path = write_input(tmp_path)
    txt = open(path).read().replace("fit = 0", "fit = 1").replace("nburn = 10", "nburn = 3")
    txt = txt.replace("nprod = 10", "nprod = 5").replace("nwalkers = 40", "nwalkers = 96")
    txt = txt.replace("usePT = 0", "usePT = 1").replace("ntemps = 1", "ntemps = 3")
    assert "usePT = 1" in txt and "ntemps = 3" in txt
    open(path, "w").write(txt)
    monkeypatch.chdir(tmp_path)
    sampler = mcmcfit.main([path, "--seed", "3"])
    out = capsys.readouterr().out
    assert "MCMC using parallel tempering at 3 levels, for 288 total walkers." in out
    ndim = sampler.chain.shape[-1]
    assert sampler.chain.shape == (3, 96, 5, ndim) and sampler.ntemps == 3
    assert np.all(np.isfinite(sampler.logprobability)) and np.all(sampler.tswap_acceptance_fraction >= 0)
    # the stored posterior of the first temperature is ln_like + ln_prob of the stored positions (the reference's
    # wiring: ptemcee's "prior" is ln_prob, mcmcfit.py:266-268)
    model = mcmcfit.construct_model(path)
    last = np.ascontiguousarray(sampler.chain[0, :, -1, :])
    want = mcmcfit.ln_like(last, model) + mcmcfit.ln_prob(last, model)
    assert np.allclose(sampler.logprobability[0, :, -1], want, rtol=1e-12)
    chain = mcmc_utils.readchain(str(tmp_path / "chain_prod.txt"))
    assert chain.shape == (96, 5, ndim + 1)
    assert np.allclose(chain[:, -1, :ndim], sampler.chain[0, :, -1, :], rtol=1e-5, atol=1e-6)


def _gp_golden():
    g = np.load(os.path.join(HERE, "golden", "gp.npz"))
    return g, [dict((k, g["%s_%d" % (k, i)]) for k in ("x", "ye", "resid", "hyper", "gaps", "lnl")) for i in range(int(g["n_cases"]))]


def test_gp_golden_vectors(hostgp):
    """The committed numbers (tests/golden/make_golden.py): oracle today == oracle then, host build of the
    device filter == both, one- and two-sided."""
    loglike, _ = hostgp
    g, cases = _gp_golden()
    for c in cases:
        want = float(c["lnl"])
        assert O.gp_log_like(c["x"], c["ye"], c["resid"], *c["hyper"], c["gaps"].tolist()) == pytest.approx(want, rel=1e-12)
        for two in (False, True):
            assert loglike(c["x"], c["ye"], c["resid"], *c["hyper"], c["gaps"].tolist(), two_sided=two) == pytest.approx(want, rel=1e-10)
    for (q, dphi, rwd), (inc, p3, p4, dist) in zip(g["wd_in"], g["wd_out"]):
        assert O.findi(q, dphi) == pytest.approx(inc, rel=1e-12)
        assert np.allclose(O.wdphases(q, inc, rwd, 10), (p3, p4), rtol=0, atol=1e-12)
        assert O.gp_dist_cp(q, dphi, rwd, 10) == pytest.approx(dist, rel=1e-11)


@pytest.mark.gpu
def test_gpu_gp_golden_vectors(engine):
    g, cases = _gp_golden()
    for c in cases:
        got = engine.gp_loglike(c["x"], c["ye"], c["resid"], c["hyper"], c["gaps"])
        assert got[0] == pytest.approx(float(c["lnl"]), rel=1e-10)
    for (q, dphi, rwd), (inc, p3, p4, dist) in zip(g["wd_in"], g["wd_out"]):
        out, ok = engine.wdphases(q, inc, rwd, 10)
        assert ok.all() and np.allclose(out[0], (p3, p4), rtol=0, atol=1e-10)


@pytest.mark.gpu
def test_gp_batched_path_with_unsorted_and_multi_cycle_phases(engine):
    """lfb_set_gp + lfb_log_prob on light curves whose phases are shuffled / span several cycles: equal to
    the per-eclipse lfb_gp_loglike of the residuals sorted by phase, with the reference's change points."""
    from lfit_python_b200 import _cabi, workloads
    wl = workloads.config(2, n_bands=1, ecl_per_band=3, n_ph=210)
    rng = np.random.default_rng(17)
    shapes = (np.linspace(-0.45, 0.48, 210), rng.permutation(np.linspace(-0.3, 0.4, 210)), np.linspace(-1.2, 1.3, 210))
    for e, x in enumerate(shapes):
        sl = slice(wl.lc_off[e], wl.lc_off[e + 1])
        wl.lc_phase[sl] = x
        wl.lc_width[sl] = 0.5 * (x.max() - x.min()) / 210
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    ln_hyper = (-9.5118, -9.0953, -5.0)
    dist = wl.apply_gp(engine, ln_hyper)
    theta = wl.walkers(24, scatter=0.01)
    theta[0] = wl.p0
    like, m2ll = engine.log_prob(theta, what=_cabi.LN_LIKE, return_chisq=True)
    hyper = np.exp(ln_hyper)
    for k in (0, 5, 23):
        total = 0.0
        for e in range(3):
            sl = slice(wl.lc_off[e], wl.lc_off[e + 1])
            x, ye = wl.lc_phase[sl], wl.lc_ye[sl]
            pars = wl.cv_pars(theta[k], e)
            resid = wl.lc_y[sl] - engine.calc_flux(pars, x, wl.lc_width[sl])
            order = np.argsort(x, kind="stable")
            gaps = O.gp_changepoints(x, dist, pars[13])
            one = engine.gp_loglike(x[order], ye[order], resid[order], hyper, gaps)[0]
            assert one == pytest.approx(O.gp_log_like(x[order], ye[order], resid[order], *hyper, gaps), rel=1e-9)
            assert m2ll[k, e] == pytest.approx(-2.0 * one, rel=1e-9)
            total += one
        assert like[k] == pytest.approx(total, rel=1e-9)
    wl.apply(engine)      # leave the shared engine in chi-squared mode
    engine.set_gp()
