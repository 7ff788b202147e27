"""GPU: the BASELINE.json configurations at FULL size against the CPU oracle.

C3 (24 eclipses x 1000 points), C4 (20 eclipses x 1000 points) and C5 (5000 points, 4x disc /
bright-spot element density) go through lfb_log_prob with their real light curves and surface grids;
eight walkers each are compared with the oracle (the oracle needs ~10-100 ms per light curve, so a
handful of walkers of the full-size tree is seconds).  The walkers are taken from a larger batch the
engine evaluated in one call, so the rows checked sit at the start, in the middle and at the end of a
batch.  Tolerances are the north-star ones (1e-7 relative chi-squared); PARITY UNPINNED against real
lfit.
"""
import numpy as np
import pytest

from oracle import oracle as O
from lfit_python_b200 import _cabi, workloads

pytestmark = pytest.mark.gpu


def oracle_layout(wl):
    return O.FlatLayout(wl.ndim, wl.npars, wl.gather, wl.consts, wl.prior_src, wl.prior_type, wl.prior_p1,
                        wl.prior_p2, wl.prior_norm, wl.prior_isvar, wl.lc_off, wl.lc_phase, wl.lc_width, wl.lc_y,
                        wl.lc_ye)


def check_against_oracle(eng, wl, n_batch, n_check=8, scatter=0.03):
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    theta = wl.walkers(n_batch, scatter=scatter, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR))
    lnp, chi = eng.log_prob(theta, return_chisq=True)
    assert chi.shape == (n_batch, wl.n_ecl)
    assert not np.isnan(lnp).any()
    sel = np.unique(np.linspace(0, n_batch - 1, n_check).astype(int))
    cfg = O.config(**wl.grid)
    ref, rchi = O.log_prob(oracle_layout(wl), theta[sel], cfg=cfg, return_chisq=True)
    fin = np.isfinite(ref)
    assert fin.all()                                     # walkers were drawn inside the priors
    assert np.array_equal(fin, np.isfinite(lnp[sel]))
    np.testing.assert_allclose(chi[sel], rchi, rtol=1e-7)    # every leaf's chi-squared
    np.testing.assert_allclose(lnp[sel], ref, rtol=1e-7)
    # the batch the rows came from does not matter: the same rows alone give the same bits
    assert np.array_equal(lnp[sel], eng.log_prob(theta[sel]))
    return theta, lnp


def test_config3_full_size_against_oracle(engine):
    wl = workloads.config(2)          # 3 bands x 8 eclipses, 1000 points each, ndim 300
    assert wl.n_ecl == 24 and wl.n_ph == 1000 and wl.ndim == 300
    check_against_oracle(engine, wl, n_batch=64)


def test_config4_full_size_against_oracle(engine):
    wl = workloads.config(3)          # 20 eclipses (7/7/6 over three bands), 1000 points each
    assert wl.n_ecl == 20 and wl.n_ph == 1000
    check_against_oracle(engine, wl, n_batch=64)


def test_config5_full_size_against_oracle():
    wl = workloads.config(4)          # 1 eclipse, 5000 points, 4x disc / bright-spot density
    assert wl.n_ph == 5000 and wl.grid == dict(n_disc_r=50, n_disc_th=80, n_bs=800)
    eng = _cabi.Engine(0, **wl.grid)
    try:
        theta, lnp = check_against_oracle(eng, wl, n_batch=96, scatter=0.05)
        # calc_flux on the 5000-point grid, one walker, against the oracle flux point by point
        pars = wl.cv_pars(theta[3], 0)
        st, ref = O.calc_flux(pars, wl.lc_phase, wl.lc_width, cfg=O.config(**wl.grid))
        assert st == 0
        f = eng.calc_flux(pars, wl.lc_phase, wl.lc_width)
        assert np.max(np.abs(f - ref) / np.abs(ref)) < 1e-9
    finally:
        eng.close()


def test_config4_many_batches_bitwise():
    """C4's shape with enough jobs for several batches on both lanes: a walker's result does not depend
    on where in which batch it sits (what makes N GPUs == 1 GPU bit for bit)."""
    wl = workloads.config(3, n_ph=200)
    eng = _cabi.Engine(0)
    try:
        wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
        wl.apply(eng)
        theta = wl.walkers(16384, scatter=0.02)       # 327 680 jobs: three batches
        lnp = eng.log_prob(theta)
        sel = np.array([0, 1, 5461, 5462, 8191, 8192, 10922, 10923, 16383])
        assert np.array_equal(lnp[sel], eng.log_prob(theta[sel]))
        halves = np.concatenate([eng.log_prob(theta[:8000]), eng.log_prob(theta[8000:])])
        assert np.array_equal(lnp, halves)
    finally:
        eng.close()
