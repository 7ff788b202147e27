"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances are the north-star ones: 1e-9 relative flux, 1e-7 relative chi-squared;
validity masks must be identical.  PARITY UNPINNED against real lfit.
"""
import numpy as np
import pytest

from oracle import oracle as O
from lfit_python_b200 import _cabi, workloads

pytestmark = pytest.mark.gpu

TESTCV = [0.333, 0.333, 0.333, 0.05, 0.1, 0.0607135, 0.6, 0.4, 0.0139, 0.039, 157.0, 0.2, 0.2, 0.0,
          2.0, 1.0, 120.0, 1.0]  # testCV.py:17-61 with dphi = findphi(0.1, 86.9)


def rel_flux_err(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_angle_of_a_unit_vector(engine):
    """The solver's replacement for atan2 (FP32 arctangent -> one of 129 tabulated angles -> arcsine series of
    the remainder, roche_device.cuh: angle_of): every part of the circle, the seams between table rows, and
    vectors that are unit only to rounding, as the Newton iteration leaves them."""
    rng = np.random.default_rng(3)
    seams = (np.arange(-64, 64) + 0.5) * np.pi / 64
    ang = np.concatenate([rng.uniform(-np.pi, np.pi, 200000), seams, seams + 1e-9, seams - 1e-9,
                          np.arange(-64, 65) * np.pi / 64 * (1 - 1e-16), [0.0, 1e-300, -1e-12, 3.0, -3.1]])
    ang = ang[np.abs(ang) < np.pi - 1e-6]           # (at +-pi either sign of the answer is right: not compared)
    norm = 1.0 + rng.uniform(-4e-16, 4e-16, ang.shape)
    out, ok = engine.roche(_cabi.ROCHE_ANGLE, np.cos(ang) * norm, np.sin(ang) * norm)
    assert ok.all()
    ref = np.arctan2(np.sin(ang), np.cos(ang))
    assert np.max(np.abs(out[:, 0] - ref)) < 1e-15


def test_roche_scalars(engine):
    q = np.array([0.05, 0.1, 0.1037, 0.2, 0.5, 1.0, 2.5])
    out, ok = engine.roche(_cabi.ROCHE_XL1, q)
    assert ok.all()
    np.testing.assert_allclose(out[:, 0], [O.xl1(v) for v in q], rtol=1e-14)
    out, ok = engine.roche(_cabi.ROCHE_FINDPHI, q, 90.0)
    np.testing.assert_allclose(out[:, 0], [O.findphi(v, 90.0) for v in q], rtol=1e-12)
    out, ok = engine.roche(_cabi.ROCHE_FINDPHI, q[:4], 86.9)
    np.testing.assert_allclose(out[:, 0], [O.findphi(v, 86.9) for v in q[:4]], rtol=1e-10)
    out, ok = engine.roche(_cabi.ROCHE_FINDI, q[:5], 0.0392)
    assert ok.all()
    np.testing.assert_allclose(out[:, 0], [O.findi(v, 0.0392) for v in q[:5]], rtol=1e-12)
    out, ok = engine.roche(_cabi.ROCHE_FINDI, [0.1], 0.2)  # wider than the edge-on eclipse
    assert not ok[0]
    rad = np.array([0.2953, 0.5214, 0.4]) * O.xl1(0.1037)
    out, ok = engine.roche(_cabi.ROCHE_BSPOT, [0.1037] * 3, rad)
    assert ok.all()
    for k in range(3):
        np.testing.assert_allclose(out[k], O.bspot(0.1037, rad[k]), rtol=0, atol=1e-10)
    out, ok = engine.roche(_cabi.ROCHE_BSPOT, [0.1037], [0.01])  # inside closest approach
    assert not ok[0]
    out, ok = engine.roche(_cabi.ROCHE_XL1, [-1.0])
    assert not ok[0]


@pytest.mark.parametrize("npars", [14, 18])
def test_calc_flux_testcv(engine, npars):
    phi = np.linspace(-0.5, 0.5, 1000)
    w = np.mean(np.diff(phi)) * np.ones_like(phi) / 2
    pars = TESTCV[:npars]
    st, ref, rcomp = O.calc_flux(pars, phi, w, components=True)
    assert st == 0
    tot, comp = engine.calc_flux(pars, phi, w, components=True)
    assert rel_flux_err(tot, ref) < 1e-9
    for k in range(4):
        np.testing.assert_allclose(comp[k], rcomp[k], rtol=0, atol=1e-9 * np.max(np.abs(rcomp[k])))


def test_calc_flux_example_eclipses(engine):
    """All six parameter sets of test_data/mcmc_input.dat, simple and complex BS."""
    wl = workloads.config(2)
    phi = np.linspace(-0.2, 0.3, 300)
    w = np.mean(np.diff(phi)) * np.ones_like(phi) / 2
    pars = np.array([wl.cv_pars(wl.p0, e) for e in range(0, wl.n_ecl, 4)])
    tot = engine.calc_flux(pars, phi, w)
    for k, p in enumerate(pars):
        st, ref = O.calc_flux(p, phi, w)
        assert st == 0
        assert rel_flux_err(tot[k], ref) < 1e-9
        st, ref = O.calc_flux(p[:14], phi, w)
        assert rel_flux_err(engine.calc_flux(p[:14], phi, w), ref) < 1e-9


def test_calc_flux_invalid_is_nan(engine):
    phi = np.linspace(-0.2, 0.2, 50)
    bad = []
    p = list(TESTCV); p[4] = -0.1; bad.append(p)          # q < 0
    p = list(TESTCV); p[5] = 0.2; bad.append(p)           # dphi wider than any inclination gives
    p = list(TESTCV); p[6] = 0.01; bad.append(p)          # disc inside the white dwarf / stream misses
    p = list(TESTCV); p[9] = -1.0; bad.append(p)          # negative spot scale
    p = list(TESTCV); p[0] = np.nan; bad.append(p)
    tot = engine.calc_flux(np.array(bad), phi, None)
    assert np.isnan(tot).all()
    for p in bad:
        st, ref = O.calc_flux(p, phi, None)
        assert st != 0 and np.isnan(ref).all()


def test_component_flags(engine):
    phi = np.linspace(-0.3, 0.3, 200)
    w = np.full_like(phi, 0.001)
    pars = list(TESTCV)
    pars[5] = 86.9
    for flags in (_cabi.FLAG_INCL, _cabi.FLAG_INCL | _cabi.FLAG_SKIP_BS | _cabi.FLAG_SKIP_DONOR,
                  _cabi.FLAG_INCL | _cabi.FLAG_SKIP_WD | _cabi.FLAG_SKIP_DISC):
        st, ref = O.calc_flux(pars, phi, w, flags=flags)
        assert st == 0
        assert rel_flux_err(engine.calc_flux(pars, phi, w, flags=flags) + 1.0, ref + 1.0) < 1e-9


def _setup(engine, wl):
    wl.make_data(lambda p, x, w: O.calc_flux(p, x, w)[1])
    wl.apply(engine)
    return O.FlatLayout(wl.ndim, wl.npars, wl.gather, wl.consts, wl.prior_src, wl.prior_type, wl.prior_p1,
                        wl.prior_p2, wl.prior_norm, wl.prior_isvar, wl.lc_off, wl.lc_phase, wl.lc_width,
                        wl.lc_y, wl.lc_ye)


@pytest.mark.parametrize("cfg", [0, 1])
def test_log_prob_single_eclipse(engine, cfg):
    wl = workloads.config(cfg, n_ph=300 if cfg == 0 else 400)
    lay = _setup(engine, wl)
    theta = wl.walkers(48, scatter=0.05)
    theta[0] = wl.p0
    for what in (_cabi.LN_PRIOR, _cabi.LN_LIKE, _cabi.LN_PROB):
        ref, rchi = O.log_prob(lay, theta, what=what, return_chisq=True)
        got, chi = engine.log_prob(theta, what=what, return_chisq=True)
        assert np.array_equal(np.isfinite(ref), np.isfinite(got)), "validity masks differ"
        fin = np.isfinite(ref)
        assert fin.sum() > 5
        if what == _cabi.LN_PRIOR:
            np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-12)
        else:
            np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-7)
            m = np.isfinite(rchi)
            assert np.array_equal(m, np.isfinite(chi))
            np.testing.assert_allclose(chi[m], rchi[m], rtol=1e-7)
        assert not np.isnan(got).any()


def test_log_prob_tree(engine):
    """3 bands x 2 eclipses sharing q/dphi/rwd (config 3 shrunk)."""
    wl = workloads.config(2, ecl_per_band=2, n_ph=250)
    lay = _setup(engine, wl)
    theta = wl.walkers(24, scatter=0.03)
    theta[0] = wl.p0
    ref, rchi = O.log_prob(lay, theta, what=2, return_chisq=True)
    got, chi = engine.log_prob(theta, what=2, return_chisq=True)
    assert np.array_equal(np.isfinite(ref), np.isfinite(got))
    fin = np.isfinite(ref)
    assert fin.sum() > 3
    np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-7)
    m = np.isfinite(rchi)
    np.testing.assert_allclose(chi[m], rchi[m], rtol=1e-7)


def test_element_solver_on_the_device_matches_the_oracle(engine):
    """Stage (1) alone through lfb_ingress_egress: disc-plane, sky-plane and far-away elements (behind the
    donor, eclipsed around phase 0.5) against the oracle's Newton and scan + bisection solvers."""
    rng = np.random.default_rng(21)
    n = 3000
    q = rng.uniform(0.03, 1.0, n)
    inc = rng.uniform(65.0, 90.0, n)
    pts = np.zeros((n, 5))
    r = 10.0 ** rng.uniform(-1.5, 0.9, n)
    az = rng.uniform(0, 2 * np.pi, n)
    pts[:, 0], pts[:, 1] = r * np.cos(az), r * np.sin(az)
    pts[::3, 2] = rng.uniform(-0.02, 0.02, len(pts[::3]))
    sky = np.arange(n) % 5 == 0                      # white-dwarf tiles: offsets fixed on the sky
    pts[sky, :3] = 0.0
    pts[sky, 3] = rng.uniform(-0.03, 0.03, sky.sum())
    pts[sky, 4] = rng.uniform(-0.03, 0.03, sky.sum())
    out, ok = engine.ingress_egress(q, inc, pts)
    worst, necl = 0.0, 0
    for k in range(n):
        solver = O.SOLVER_ROBUST if k % 8 == 0 else O.SOLVER_NEWTON
        ref = O.ingress_egress(q[k], inc[k], tuple(pts[k, :3]), xi=pts[k, 3], eta=pts[k, 4], solver=solver)
        assert (ref is not None) == bool(ok[k]), (k, q[k], inc[k], pts[k])
        if ref is not None:
            necl += 1
            worst = max(worst, abs(ref[0] - out[k, 0]), abs(ref[1] - out[k, 1]))
    assert necl > n // 4 and worst < 1e-11


@pytest.mark.parametrize("cfg,n,n_ph", [(1, 1500, 160), (0, 800, 200), (2, 300, 120)])
def test_whole_prior_box(engine, cfg, n, n_ph):
    """Walkers drawn uniformly between the prior limits (long strips reaching behind the donor, grazing
    inclinations, huge discs ...): validity masks identical, chi-squared to 1e-7 (tools/wide_sweep.py)."""
    wl = workloads.config(cfg, n_ph=n_ph)
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    lay = O.FlatLayout(wl.ndim, wl.npars, wl.gather, wl.consts, wl.prior_src, wl.prior_type, wl.prior_p1, wl.prior_p2,
                       wl.prior_norm, wl.prior_isvar, wl.lc_off, wl.lc_phase, wl.lc_width, wl.lc_y, wl.lc_ye)
    rng = np.random.default_rng(5 + cfg)
    lo, hi = wl.prior_p1.copy(), wl.prior_p2.copy()
    gauss = np.isin(wl.prior_type, (0, 1))
    lo[gauss], hi[gauss] = wl.prior_p1[gauss] - 3 * wl.prior_p2[gauss], wl.prior_p1[gauss] + 3 * wl.prior_p2[gauss]
    theta = lo + (hi - lo) * rng.random((n, wl.ndim))
    keep = rng.random((n // 2, wl.ndim)) < 0.7      # half of them: mostly the truth, a few wide-open parameters
    theta[:n // 2] = np.where(keep, wl.p0, theta[:n // 2])
    ref, rchi = O.log_prob(lay, theta, what=_cabi.LN_LIKE, return_chisq=True)
    got, chi = engine.log_prob(theta, what=_cabi.LN_LIKE, return_chisq=True)
    assert np.array_equal(np.isfinite(rchi), np.isfinite(chi))
    assert np.array_equal(np.isfinite(O.log_prob(lay, theta, what=_cabi.LN_PRIOR)),
                          np.isfinite(engine.log_prob(theta, what=_cabi.LN_PRIOR)))
    m = np.isfinite(rchi)
    assert m.sum() > n // 3
    np.testing.assert_allclose(chi[m], rchi[m], rtol=1e-7)
    assert not np.isnan(got).any()
