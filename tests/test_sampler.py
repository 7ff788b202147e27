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


# ---- the stretch move of the device sampler: random stream, algebra, statistics (numpy restatement) ----
def test_philox_known_answers():
    """Random123's known-answer vectors for Philox4x32-10."""
    from oracle.stretch import philox4x32_10
    z = np.zeros(1, dtype=np.uint32)
    out = [int(v[0]) for v in philox4x32_10(z, z, z, z, 0, 0)]
    assert out == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.full(1, 0xffffffff, dtype=np.uint32)
    out = [int(v[0]) for v in philox4x32_10(f, f, f, f, 0xffffffff, 0xffffffff)]
    assert out == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    c = [np.full(1, v, dtype=np.uint32) for v in (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)]
    out = [int(v[0]) for v in philox4x32_10(*c, 0xa4093822, 0x299f31d0)]
    assert out == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_native_draws_equal_the_numpy_restatement():
    """lfb_stretch_draws (the kernels' Philox + z + partner, compiled for the host) against oracle/stretch.py."""
    from lfit_python_b200 import _cabi
    from oracle import stretch as S
    for seed, step, half, hn in [(0, 0, 0, 25), (12345678901234567, 7, 1, 2048), (2 ** 63 + 5, 2 ** 33 + 1, 0, 32768)]:
        d = _cabi.stretch_draws(seed, step, half, hn, 2.0, 500)
        z, partner, lnu = S.draws(seed, step, half, hn, 2.0, np.arange(500))
        assert np.array_equal(d[:, 0], z) and np.array_equal(d[:, 1], partner)
        assert np.allclose(d[:, 2], lnu, rtol=1e-15, atol=0)
        assert (z >= 0.5).all() and (z <= 2.0).all() and partner.min() >= 0 and partner.max() < hn
    # g(z) ~ 1/sqrt(z) on [1/a, a]: E[z] = (a + 1 + 1/a) / 3
    z, partner, lnu = S.draws(5, 3, 0, 1000, 2.0, np.arange(200000) % 1000 + 1000 * (np.arange(200000) // 1000))
    assert abs(z.mean() - 3.5 / 3.0) < 3e-3 and abs(np.exp(lnu).mean() - 0.5) < 3e-3


def test_stretch_oracle_samples_a_gaussian():
    from oracle.stretch import StretchOracle
    mu, sig = np.array([1.0, -2.0, 0.5, 3.0]), np.array([0.5, 2.0, 1.0, 0.1])
    s = StretchOracle(lambda t: gauss_lnprob(t, mu, 1 / sig), 64, 4, seed=8)
    s.set_state(mu + 0.1 * np.random.default_rng(3).standard_normal((64, 4)))
    s.run(300)
    s.run(1500, record=True)
    flat = np.asarray(s.chain)[::5, :, :4].reshape(-1, 4)
    assert np.allclose(flat.mean(axis=0), mu, atol=4 * sig / np.sqrt(300))
    assert np.allclose(flat.std(axis=0), sig, rtol=0.15)
    assert 0.2 < (s.naccepted / 1800).mean() < 0.9


def test_native_chain_text_equals_the_reference_expression():
    """lfb_chain_format against "{0:4d} {1:s} {2:f}".format(k, " ".join(map(str, pos)), prob) (mcmc_utils.py:163-164)."""
    from lfit_python_b200 import _cabi
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.standard_normal(40000) * 10.0 ** rng.integers(-30, 30, 40000),
                           rng.standard_normal(20000), rng.uniform(0, 200, 20000),
                           [0.0, -0.0, 1e16, 1e-4, 9.999e-5, 1e15, 123456789012345678.0, 0.1037, 120.0, 1e-5, 1.5e-5, 1e100,
                            1e-100, 5e-324, 1.7976931348623157e308, np.inf, -np.inf, np.nan, 1e22, 1e21,
                            9007199254740993.0, 0.1 + 0.2, 1 / 3, 2.5e-5]])
    vals = vals[: len(vals) // 4 * 4]
    rows = vals.reshape(2, -1, 4)
    txt = _cabi.chain_text(rows).decode()
    ref = "".join(utils.format_step_python(r[:, :3], r[:, 3]) for r in rows)
    assert txt == ref
    assert utils.format_step(rows[0, :, :3], rows[0, :, 3]) == utils.format_step_python(rows[0, :, :3], rows[0, :, 3])
    with pytest.raises(ValueError):
        _cabi.chain_text(np.zeros((3, 4)))


# ---- parallel tempering (ptemcee's algorithm; /root/reference/mcmcfit.py:251-270, mcmc_utils.py:75-111,186-239) ----
def test_parallel_tempering_samples_the_tempered_gaussians(tmp_path):
    rng = np.random.default_rng(5)
    mu, sig = np.array([1.0, -2.0, 0.5]), np.array([0.5, 2.0, 1.0])
    calls = []

    def logl(t):
        calls.append(np.shape(t))
        return gauss_lnprob(t, mu, 1 / sig)

    def logp(t):
        return np.where(np.all(np.abs(np.atleast_2d(t)) < 50, axis=1), 0.0, -np.inf)

    s = utils.PTSampler(20, 3, logl, logp, ntemps=4, vectorize=True, rng=rng)
    assert s.betas[0] == 1.0 and np.all(np.diff(s.betas) < 0) and s.ntemps == 4
    p0 = utils.initialise_walkers_pt(mu, np.full(3, 0.1), 20, 4, lambda t, m: logp(t), None, rng=rng, verbose=False)
    assert p0.shape == (4, 20, 3)
    pos, prob, like = utils.run_burnin(s, p0, 400)
    assert s.chain.shape == (4, 20, 0, 3)
    # one vectorised call per half-step for ALL temperatures: 4 x 10 rows
    assert set(calls[1:]) == {(40, 3)} and len(calls) == 1 + 2 * 400
    s.reset()
    path = tmp_path / "chain_prod.txt"
    utils.run_ptmcmc_save(s, pos, 2500, str(path), col_names="walker_no a b c ln_prob")
    c = s.chain
    assert c.shape == (4, 20, 2500, 3) and s.flatchain.shape == (4, 20 * 2500, 3)
    for t in range(4):   # temperature t samples N(mu, sig^2 / beta_t)
        f = c[t, :, ::5, :].reshape(-1, 3)
        assert np.allclose(f.mean(axis=0), mu, atol=5 * sig / np.sqrt(s.betas[t]) / np.sqrt(250))
        assert np.allclose(f.std(axis=0) * np.sqrt(s.betas[t]), sig, rtol=0.12)
    assert np.all(s.tswap_acceptance_fraction > 0.2) and np.all(s.tswap_acceptance_fraction < 0.8)
    assert np.all(s.acceptance_fraction.mean(axis=1) > 0.3)
    # the file holds the first temperature only, in the reference's format, and reads back
    chain = utils.readchain(str(path))
    assert chain.shape == (20, 2500, 4) and np.array_equal(chain[:, :, :3], c[0])
    assert np.allclose(chain[:, :, 3], s.logprobability[0], atol=1e-6)


def test_parallel_tempering_argument_checks():
    with pytest.raises(ValueError):
        utils.PTSampler(7, 2, gauss_lnprob, gauss_lnprob, ntemps=2)
    s = utils.PTSampler(8, 2, lambda t: -0.5 * np.sum(t * t, axis=1), lambda t: np.where(t[:, 0] > 0, 0.0, -np.inf), ntemps=2,
                        vectorize=True, rng=np.random.default_rng(1))
    with pytest.raises(ValueError, match="outside posterior support"):
        s.run_mcmc(np.full((2, 8, 2), -1.0), 1)
    out = s.run_mcmc(np.abs(np.random.default_rng(2).standard_normal((2, 8, 2))) + 0.1, 50)
    assert (out[0][..., 0] > 0).all()                 # walkers never leave the prior support
    one = utils.PTSampler(8, 2, lambda t: -0.5 * np.sum(t * t, axis=1), lambda t: np.zeros(len(t)), vectorize=True)
    assert one.ntemps == 1 and one.betas[0] == 1.0
