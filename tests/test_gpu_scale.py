"""GPU: BASELINE.json-sized workloads checked through size-independent properties (the oracle
needs seconds per hundred light curves, so full sizes are not compared point by point)."""
import numpy as np
import pytest

from oracle import oracle as O
from lfit_python_b200 import _cabi, workloads

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties(engine):
    wl = workloads.config(1)                     # 1 eclipse, complex BS, 2000 points, 4096 walkers
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    theta = wl.walkers(wl.n_walkers, ln_prior_fn=lambda t: engine.log_prob(t, what=_cabi.LN_PRIOR))
    lnp, chi = engine.log_prob(theta, return_chisq=True)
    prior = engine.log_prob(theta, what=_cabi.LN_PRIOR)
    like = engine.log_prob(theta, what=_cabi.LN_LIKE)
    assert np.isfinite(lnp).all() and not np.isnan(lnp).any()
    assert np.allclose(lnp, prior + like, rtol=1e-12)                 # ln_prob = ln_prior + ln_like
    assert np.allclose(like, -0.5 * chi[:, 0], rtol=1e-14)            # ln_like = -chi^2/2
    # a spot check of 32 walkers against the oracle
    lay = O.FlatLayout(wl.ndim, wl.npars, wl.gather, wl.consts, wl.prior_src, wl.prior_type, wl.prior_p1,
                       wl.prior_p2, wl.prior_norm, wl.prior_isvar, wl.lc_off, wl.lc_phase, wl.lc_width, wl.lc_y,
                       wl.lc_ye)
    sel = np.linspace(0, wl.n_walkers - 1, 32).astype(int)
    ref = O.log_prob(lay, theta[sel])
    assert np.allclose(lnp[sel], ref, rtol=1e-7)
    # chi-squared is a quadratic form in the data: check it through the flux API for one walker
    f = engine.calc_flux(wl.cv_pars(theta[7], 0), wl.lc_phase, wl.lc_width)
    assert np.sum(((wl.lc_y - f) / wl.lc_ye) ** 2) == pytest.approx(chi[7, 0], rel=1e-10)
    # flux scales linearly with the four component fluxes
    p = wl.cv_pars(theta[7], 0).copy()
    p2 = p.copy()
    p2[:4] *= 3.0
    assert np.allclose(engine.calc_flux(p2, wl.lc_phase, wl.lc_width), 3.0 * f, rtol=1e-13)
    # phi0 is a pure shift of the phase axis
    p3 = p.copy()
    p3[13] += 0.01
    assert np.allclose(engine.calc_flux(p3, wl.lc_phase + 0.01, wl.lc_width), f, rtol=1e-9)


def test_hierarchical_tree_shares_core_parameters(engine):
    wl = workloads.config(2, n_ph=400)           # 3 bands x 8 eclipses
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    theta = wl.walkers(256, scatter=0.02, ln_prior_fn=lambda t: engine.log_prob(t, what=_cabi.LN_PRIOR))
    lnp, chi = engine.log_prob(theta, return_chisq=True)
    assert chi.shape == (256, 24) and np.isfinite(lnp).all()
    # each leaf's chi-squared equals a one-eclipse evaluation with the gathered CV parameters
    for e in (0, 9, 23):
        sl = slice(wl.lc_off[e], wl.lc_off[e + 1])
        f = engine.calc_flux(np.array([wl.cv_pars(t, e) for t in theta[:8]]), wl.lc_phase[sl], wl.lc_width[sl])
        ref = np.sum(((wl.lc_y[sl] - f) / wl.lc_ye[sl]) ** 2, axis=1)
        assert np.allclose(chi[:8, e], ref, rtol=1e-10)
    assert np.allclose(lnp, engine.log_prob(theta, what=_cabi.LN_PRIOR) - 0.5 * chi.sum(axis=1), rtol=1e-12)


def test_batching_of_large_ensembles_is_invisible(engine):
    wl = workloads.config(3, n_ph=60)            # 20 eclipses; > 262144 jobs forces several batches
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    theta = wl.walkers(14000, scatter=0.01)
    lnp = engine.log_prob(theta)
    assert lnp.shape == (14000,)
    sel = np.array([0, 1, 13106, 13107, 13108, 13999])  # around the batch boundary
    assert np.array_equal(lnp[sel], engine.log_prob(theta[sel]))
