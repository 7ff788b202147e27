"""Sampler glue on the CPU: the ensemble sampler, walker initialisation, chain file format."""
import numpy as np
import pytest

from lfit_python_b200 import mcmc_utils as utils


def gauss_lnprob(theta, mu, isig):
    d = (np.atleast_2d(theta) - mu) * isig
    return -0.5 * np.sum(d * d, axis=1)


def test_stretch_move_samples_a_gaussian():
    rng = np.random.default_rng(3)
    ndim, nw = 4, 64
    mu, sig = np.array([1.0, -2.0, 0.5, 3.0]), np.array([0.5, 2.0, 1.0, 0.1])
    calls = []

    def fn(theta, *a):
        calls.append(theta.shape)
        return gauss_lnprob(theta, mu, 1 / sig)

    s = utils.EnsembleSampler(nw, ndim, fn, vectorize=True, rng=rng)
    p0 = mu + 0.1 * rng.standard_normal((nw, ndim))
    pos, prob, _ = utils.run_burnin(s, p0, 300)
    assert s.chain.shape == (nw, 0, ndim)          # burn-in is not stored (mcmc_utils.py:114)
    s.reset()
    s.run_mcmc(pos, 1500)
    # one vectorised call per half-step, plus one for the starting ensemble of each run
    assert sorted(set(calls)) == [(nw // 2, ndim), (nw, ndim)] and calls.count((nw, ndim)) == 2
    assert len(calls) == 2 + 2 * (300 + 1500)
    flat = utils.flatchain(s.chain, ndim, thin=5)
    assert np.allclose(flat.mean(axis=0), mu, atol=4 * sig / np.sqrt(300))
    assert np.allclose(flat.std(axis=0), sig, rtol=0.15)
    assert 0.2 < s.acceptance_fraction.mean() < 0.9
    assert s.lnprobability.shape == (nw, 1500)


def test_sampler_argument_checks():
    with pytest.raises(ValueError):
        utils.EnsembleSampler(6, 4, gauss_lnprob)   # fewer than 2 * ndim walkers (mcmcfit.py:195)
    s = utils.EnsembleSampler(8, 2, lambda t: np.full(len(t), np.nan), vectorize=True)
    with pytest.raises(ValueError, match="NaN"):
        s.run_mcmc(np.zeros((8, 2)), 1)


def test_nonvectorised_and_vectorised_agree():
    mu, isig = np.zeros(3), np.ones(3)
    p0 = np.random.default_rng(1).standard_normal((12, 3))
    a = utils.EnsembleSampler(12, 3, gauss_lnprob, args=(mu, isig), vectorize=True, rng=np.random.default_rng(9))
    b = utils.EnsembleSampler(12, 3, lambda t, m, i: float(gauss_lnprob(t, m, i)[0]), args=(mu, isig),
                              rng=np.random.default_rng(9))
    pa, la, _ = a.run_mcmc(p0, 20)
    pb, lb, _ = b.run_mcmc(p0, 20)
    assert np.array_equal(pa, pb) and np.allclose(la, lb)


def test_initialise_walkers_resamples_invalid_ones():
    rng = np.random.default_rng(0)
    p = np.array([0.5, 2.0, 10.0])

    def ln_prior(theta, model):
        theta = np.atleast_2d(theta)
        ok = (theta[:, 0] > 0.45) & (theta[:, 0] < 0.6) & (theta[:, 1] > 0)
        return np.where(ok, 0.0, -np.inf)

    p0 = utils.initialise_walkers(p, np.array([0.2, 0.1, 0.1]), 200, ln_prior, None, rng=rng, verbose=False)
    assert p0.shape == (200, 3) and np.isfinite(ln_prior(p0, None)).all()
    assert abs(p0[:, 2].mean() - 10.0) < 0.5


def test_chain_file_format_and_reader(tmp_path):
    rng = np.random.default_rng(2)
    s = utils.EnsembleSampler(8, 2, gauss_lnprob, args=(np.zeros(2), np.ones(2)), vectorize=True, rng=rng)
    path = tmp_path / "chain_prod.txt"
    utils.run_mcmc_save(s, rng.standard_normal((8, 2)), 5, None, str(path), col_names="walker_no a_core b_core ln_prob")
    lines = path.read_text().splitlines()
    assert lines[0] == "walker_no a_core b_core ln_prob" and len(lines) == 1 + 5 * 8
    k, a, b, lp = lines[1].split()
    assert lines[1].startswith("   0 ") and int(k) == 0
    assert float(a) == s.chain[0, 0, 0] and float(b) == s.chain[0, 0, 1]       # positions written with str(): exact
    assert lp == "{:f}".format(s.lnprobability[0, 0])                           # ln_prob with %f, as the reference
    chain = utils.readchain(str(path))
    assert chain.shape == (8, 5, 3)
    assert np.array_equal(chain[:, :, :2], s.chain)
