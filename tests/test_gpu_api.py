"""GPU: the reference-facing Python surface (lfit / roche shims, tree, vectorised wrappers)
against the oracle and against itself, plus edge cases of the C ABI."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from lfit_python_b200 import _cabi, lfit, mcmcfit, roche, workloads
from lfit_python_b200.CVModel import construct_model
from lfit_python_b200.flatten import FlatLayout, VectorModel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

from test_tree import write_input  # noqa: E402  (same synthetic input file as the CPU tests)


def olay(L):
    return O.FlatLayout(L.ndim, L.npars, L.gather, L.consts, L.prior_src, L.prior_type, L.prior_p1, L.prior_p2,
                        L.prior_norm, L.prior_isvar, L.lc_off, L.lc_phase, L.lc_width, L.lc_y, L.lc_ye)


def test_golden_calc_flux(engine):
    g = np.load(os.path.join(GOLD, "calc_flux.npz"))
    for k in range(g["pars"].shape[0]):
        p = g["pars"][k]
        p = p[~np.isnan(p)]
        tot, comp = engine.calc_flux(p, g["phase"], g["width"], components=True)
        assert np.max(np.abs(tot - g["total"][k]) / g["total"][k]) < 1e-9
        assert np.allclose(comp, g["comp"][k], rtol=0, atol=1e-9 * g["total"][k].max())


def test_golden_log_prob(engine):
    g = np.load(os.path.join(GOLD, "log_prob.npz"))
    wl = workloads.config(2, n_bands=2, ecl_per_band=2, n_ph=90, phase_range=(-0.15, 0.2))
    wl.lc_y = g["lc_y"]
    wl.apply(engine)
    for what, name in ((0, "ln_prior"), (1, "ln_like"), (2, "ln_prob")):
        v, chi = engine.log_prob(g["theta"], what=what, return_chisq=True)
        assert np.array_equal(np.isfinite(v), np.isfinite(g[name]))
        fin = np.isfinite(v)
        assert np.allclose(v[fin], g[name][fin], rtol=1e-7)
        assert np.array_equal(np.isnan(chi), np.isnan(g[name + "_chisq"]))
        m = np.isfinite(g[name + "_chisq"])
        assert np.allclose(chi[m], g[name + "_chisq"][m], rtol=1e-7)


def test_golden_roche(engine):
    r = np.load(os.path.join(GOLD, "roche.npz"))
    assert np.allclose(roche.xl1(r["q"]), r["xl1"], rtol=1e-14)
    assert np.allclose(roche.findphi(r["q"], 90.0), r["maxphi"], rtol=1e-12)
    assert np.allclose(roche.findi(r["q"], 0.6 * r["maxphi"]), r["incl"], rtol=1e-11)
    out, ok = engine.roche(_cabi.ROCHE_BSPOT, r["q"], 0.4 * r["xl1"])
    assert ok.all() and np.allclose(out, r["bspot"], rtol=0, atol=1e-10)


def test_roche_shim_scalars_and_errors():
    assert abs(roche.xl1(0.1037) - 0.714609543404) < 1e-11
    assert abs(roche.findi(0.1037, 0.0392) - 81.2221) < 2e-4
    x, y, vx, vy = roche.bspot(0.1037, 0.2953 * roche.xl1(0.1037))
    assert abs(np.hypot(x, y) - 0.2953 * roche.xl1(0.1037)) < 1e-12
    with pytest.raises(AssertionError):
        roche.xl1(-1.0)
    with pytest.raises(Exception):
        roche.bspot(0.1037, 0.01)
    with pytest.raises(Exception):
        roche.findi(0.1, 0.3)


def test_lfit_components_sum_to_cv():
    """testCV.py: weighted sum of the unit components equals CV.calcFlux."""
    q, inc = 0.1, 86.9
    phi = np.linspace(-0.5, 0.5, 400)
    width = np.mean(np.diff(phi)) * np.ones_like(phi) / 2.
    xl1 = roche.xl1(q)
    dphi = roche.findphi(q, inc)
    rwd = 0.01 / xl1
    w = lfit.PyWhiteDwarf(rwd, 0.4)
    d = lfit.PyDisc(q, rwd, 0.6, 0.2, 1000)
    s = lfit.PySpot(q, 0.6, 157.0, 0.2, 0.039, exp1=2.0, exp2=1.0, tilt=120.0, yaw=1.0, complex=True)
    rs = lfit.PyDonor(q, 400)
    ywd, yd, ys, yrs = (c.calcFlux(q, inc, phi, width) for c in (w, d, s, rs))
    pars = [0.333, 0.333, 0.333, 0.05, q, dphi, 0.6, 0.4, rwd, 0.039, 157.0, 0.2, 0.2, 0.0, 2.0, 1.0, 120.0, 1.0]
    cv = lfit.CV(pars)
    flux = cv.calcFlux(pars, phi, width)
    assert np.allclose(0.333 * (ywd + yd + ys) + 0.05 * yrs, flux, rtol=1e-9)
    assert np.allclose(cv.ywd, 0.333 * ywd, atol=1e-12) and np.allclose(cv.yrs, 0.05 * yrs, atol=1e-12)
    assert ywd.max() == pytest.approx(1.0, abs=1e-14) and abs(ywd.min()) < 1e-14  # 2^-56 fixed-point sums
    assert flux.shape == phi.shape and cv(pars, phi, width).shape == phi.shape
    st, ref = O.calc_flux(pars, phi, width)
    assert np.max(np.abs(flux - ref) / ref) < 1e-9
    with pytest.raises(ValueError):
        lfit.CV(pars[:15])
    bad = list(pars)
    bad[5] = 0.2
    with pytest.raises(ValueError):
        cv.calcFlux(bad, phi, width)
    # a denser disc grid than the default engine's
    d2 = lfit.PyDisc(q, rwd, 0.6, 0.2, 4000).calcFlux(q, inc, phi, width)
    assert 0 < np.max(np.abs(d2 - yd)) < 5e-3


def test_tree_scalar_path_equals_vector_path(tmp_path, engine):
    m = construct_model(write_input(tmp_path))
    vec = VectorModel(m)
    L = vec.layout
    lay = olay(L)
    p0 = np.asarray(m.dynasty_par_vals)
    rng = np.random.default_rng(5)
    theta = p0 * (1 + 0.02 * rng.standard_normal((16, L.ndim)))
    theta[0] = p0
    for name, what in (("ln_prior", 0), ("ln_like", 1), ("ln_prob", 2)):
        ref = O.log_prob(lay, theta, what=what)
        got = getattr(mcmcfit, name)(theta, m)               # (n, ndim): one CUDA call
        assert np.array_equal(np.isfinite(ref), np.isfinite(got))
        fin = np.isfinite(ref)
        assert np.allclose(got[fin], ref[fin], rtol=1e-7)
        for k in (0, 1, 5):                                   # 1-D: the reference's scalar tree walk
            one = getattr(mcmcfit, name)(theta[k], m)
            assert np.isfinite(one) == np.isfinite(ref[k])
            if np.isfinite(one):
                assert one == pytest.approx(ref[k], rel=1e-7)
    assert np.allclose(m.dynasty_par_vals, theta[5])          # the scalar path mutates the tree, as the reference does
    m.dynasty_par_vals = p0
    assert m.chisq() == pytest.approx(np.sum(vec.chisq(p0)), rel=1e-9)
    flx, ywd, ys, yrs, yd = m.children[0].children[0].calcComponents()
    assert np.allclose(flx, ywd + ys + yrs + yd, rtol=1e-12)
    with pytest.raises(ValueError):
        vec.ln_prob(theta[:, :-1])


def test_invalid_walkers_are_minus_inf_never_nan(engine):
    wl = workloads.config(1, n_ph=120)
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    lay = olay(wl)
    theta = np.tile(wl.p0, (9, 1))
    idx = {n.split("_")[0]: i for i, n in enumerate(wl.names)}
    theta[1, idx["q"]] = -0.1            # no Roche geometry
    theta[2, idx["dphi"]] = 0.09         # wider than the edge-on eclipse for this q
    theta[3, idx["rdisc"]] = 0.69        # beyond the 3:1 resonance
    theta[4, idx["scale"]] = 0.19        # spot larger than 3 white-dwarf radii
    theta[5, idx["az"]] = 51.0           # strip too far from the disc tangent
    theta[6, idx["ulimb"]] = 0.4         # 100 sigma from its Gaussian prior
    theta[7, idx["q"]] = np.nan
    theta[8, idx["fis"]] = 1.5           # outside its uniform prior
    got = engine.log_prob(theta)
    ref = O.log_prob(lay, theta)
    assert np.isfinite(got[0]) and np.all(got[1:] == -np.inf) and not np.isnan(got).any()
    assert np.array_equal(np.isfinite(ref), np.isfinite(got))
    like = engine.log_prob(theta, what=_cabi.LN_LIKE)
    rlike = O.log_prob(lay, theta, what=1)
    assert np.array_equal(np.isfinite(rlike), np.isfinite(like)) and not np.isnan(like).any()
    fin = np.isfinite(rlike)
    assert fin.sum() >= 5 and np.allclose(like[fin], rlike[fin], rtol=1e-7)


def test_results_are_bitwise_reproducible(engine):
    wl = workloads.config(1, n_ph=333)
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    theta = wl.walkers(64, scatter=0.03)
    a = engine.log_prob(theta)
    b = engine.log_prob(theta)
    c = engine.log_prob(theta[::-1].copy())[::-1]            # batch position must not matter
    d = np.concatenate([engine.log_prob(theta[:17]), engine.log_prob(theta[17:])])  # nor the split (= sharding)
    assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)


@pytest.mark.parametrize("phases", ["unsorted", "wrapped", "ragged", "single"])
def test_light_curve_shapes(engine, phases):
    rng = np.random.default_rng(8)
    if phases == "unsorted":
        x = rng.permutation(np.linspace(-0.3, 0.3, 257))
        w = np.full_like(x, 0.001)
    elif phases == "wrapped":
        x = np.linspace(0.7, 1.3, 200)           # cycle number not removed
        w = np.full_like(x, 0.0015)
    elif phases == "ragged":
        x = np.sort(rng.uniform(-0.5, 0.5, 301))  # uneven sampling, uneven exposures, whole orbit
        w = rng.uniform(0.0, 0.004, x.size)
    else:
        x, w = np.array([0.013]), np.array([0.002])
    pars = workloads.config(1).cv_pars(workloads.config(1).p0, 0)
    st, ref = O.calc_flux(pars, x, w)
    got = engine.calc_flux(pars, x, w)
    assert st == 0 and np.max(np.abs(got - ref) / ref) < 1e-9


def test_empty_and_mixed_length_light_curves(engine):
    wl = workloads.Workload("mixed", 1, 3, 50, phase_range=(-0.1, 0.15))
    # eclipse lengths 50, 0, 7
    n = [50, 0, 7]
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    x = np.concatenate([np.linspace(-0.1, 0.15, k) for k in n])
    w = np.full_like(x, 0.001)
    wl.lc_off, wl.lc_phase, wl.lc_width, wl.lc_ye = off, x, w, np.full_like(x, 0.004)
    wl.lc_y = np.full_like(x, 0.2)
    wl.apply(engine)
    lay = olay(wl)
    theta = wl.walkers(6, scatter=0.01)
    got, chi = engine.log_prob(theta, what=_cabi.LN_LIKE, return_chisq=True)
    ref, rchi = O.log_prob(lay, theta, what=1, return_chisq=True)
    assert np.all(chi[:, 1] == 0.0) and np.allclose(chi, rchi, rtol=1e-7) and np.allclose(got, ref, rtol=1e-7)
    assert engine.log_prob(np.empty((0, wl.ndim))).shape == (0,)


@pytest.mark.parametrize("grid", [dict(n_quad=1), dict(n_quad=5), dict(n_disc_r=50, n_disc_th=80, n_bs=800),
                                  dict(n_wd_rings=6, n_donor_th=12, donor_ulimb=0.6, donor_gdexp=0.25)])
def test_other_surface_grids(grid):
    eng = _cabi.Engine(0, **grid)
    cfg = O.config(**grid)
    wl = workloads.config(1)
    pars = wl.cv_pars(wl.p0, 0)
    x = np.linspace(-0.2, 0.3, 150)
    w = np.full_like(x, 0.0017)
    st, ref = O.calc_flux(pars, x, w, cfg=cfg)
    got = eng.calc_flux(pars, x, w)
    assert st == 0 and np.max(np.abs(got - ref) / ref) < 1e-9
    eng.close()


def test_call_order_errors():
    eng = _cabi.Engine(0)
    with pytest.raises(_cabi.EngineError, match="set_layout"):
        eng._check(eng._lib.lfb_log_prob(eng._h, 2, 1, np.zeros(18).ctypes.data, np.zeros(1).ctypes.data, None, None), "lfb_log_prob")
    wl = workloads.config(1, n_ph=20)
    eng.set_layout(wl.ndim, wl.npars, wl.gather, wl.consts)
    eng.ndim, eng.n_ecl = wl.ndim, wl.n_ecl
    with pytest.raises(_cabi.EngineError, match="set_lightcurves"):
        eng.log_prob(wl.p0[None, :])
    assert np.isfinite(eng.log_prob(wl.p0[None, :], what=_cabi.LN_PRIOR))[0] or True
    with pytest.raises(_cabi.EngineError):
        eng.set_layout(wl.ndim, 15, wl.gather, wl.consts)
    g = wl.gather.copy()
    g[0, 0] = 99
    with pytest.raises(_cabi.EngineError, match="out of range"):
        eng.set_layout(wl.ndim, wl.npars, g, wl.consts)
    with pytest.raises(_cabi.EngineError, match="grid"):
        _cabi.Engine(0, n_disc_th=41)
    eng.close()


def test_device_resident_buffers(engine):
    torch = pytest.importorskip("torch")
    wl = workloads.config(1, n_ph=100)
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    theta = wl.walkers(40, scatter=0.02)
    ref = engine.log_prob(theta)
    t = torch.from_numpy(theta).cuda()
    out = torch.empty(40, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        engine.log_prob_device(t.data_ptr(), 40, out.data_ptr(), stream=s.cuda_stream)
    s.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    ms = engine.last_stage_ms()
    assert ms["total"] > 0 and engine.launch_count > 0
    engine.set_trace(True)
    try:
        assert np.array_equal(engine.log_prob(theta), ref)
        tr = engine.last_trace_ms()
        assert tr["flux_kernel"] > 0 and tr["elements_kernel<1> disc"] > 0 and len(tr) == 13
    finally:
        engine.set_trace(False)


def test_device_resident_sampler_keeps_its_books(engine):
    from lfit_python_b200 import mcmc_utils

    wl = workloads.config(1, n_ph=150)
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    p0 = wl.walkers(64, scatter=0.01, ln_prior_fn=lambda t: engine.log_prob(t, what=_cabi.LN_PRIOR))
    s = mcmc_utils.DeviceSampler(engine, 64, seed=5)
    lnp0 = engine.log_prob(p0)
    pos, lnp, _ = s.run_mcmc(p0, 15)
    assert np.array_equal(lnp, engine.log_prob(pos))          # stored ln_prob belongs to the stored position
    assert np.isfinite(lnp).all() and 0.05 < s.acceptance_fraction.mean() < 0.95
    assert np.median(lnp) > np.median(lnp0) - 2 * wl.ndim      # relaxes from a tight ball to the posterior width, no further
    pos2, lnp2, _ = s.run_mcmc(None, 5)                        # continues from the resident state
    assert s.iterations == 20 and np.array_equal(lnp2, engine.log_prob(pos2))
    s.close()


def test_small_repeated_calls_replay_a_cuda_graph_safely():
    """From the second identical small call on, lfb_log_prob replays a captured CUDA graph: results must stay
    bit-identical, follow new inputs in the same buffers, and survive re-allocation and re-configuration."""
    import torch
    eng = _cabi.Engine(0)
    try:
        wl = workloads.config(0, n_ph=150)
        wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
        wl.apply(eng)
        theta = wl.walkers(64, scatter=0.02)
        t = torch.from_numpy(theta).cuda()
        out = torch.empty(64, dtype=torch.float64, device="cuda")
        s = torch.cuda.Stream()

        def call(n=64):
            eng.log_prob_device(t.data_ptr(), n, out.data_ptr(), stream=s.cuda_stream)
            s.synchronize()
            return out[:n].cpu().numpy().copy()

        first = call()
        l0 = eng.launch_count
        for _ in range(4):                                   # plain, captured, replayed, replayed
            assert np.array_equal(call(), first)
        per_call = (eng.launch_count - l0) // 4
        assert per_call >= 10                                # the replayed kernels are counted too
        assert eng.last_stage_ms()["total"] > 0 and eng.last_stage_ms()["flux"] == -1.0   # no stage events in a graph
        t.copy_(torch.from_numpy(theta[::-1].copy()))        # new walkers in the same buffer
        assert np.array_equal(call(), first[::-1])
        big = wl.walkers(3000, scatter=0.02)                 # a large call re-allocates the lane buffers
        ref_big = eng.log_prob(big)
        assert np.array_equal(call(), first[::-1])
        assert np.array_equal(call(), first[::-1])
        assert np.array_equal(eng.log_prob(big), ref_big)
        wl2 = workloads.config(0, n_ph=90)                   # new light curves: captured passes are dropped
        wl2.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
        wl2.apply(eng)
        a = call()
        assert not np.array_equal(a, first[::-1])
        for _ in range(3):
            assert np.array_equal(call(), a)
        assert np.array_equal(call(17), a[:17])              # another size is another graph
        assert np.array_equal(call(17), a[:17])
        assert np.array_equal(call(17), a[:17])
    finally:
        eng.close()


def test_page_locked_host_buffers_are_used_as_they_stand():
    """theta and the output in pinned host memory (copied from / to directly) give what pageable arrays give
    (staged through the engine's own pinned buffers), bit for bit; so do device outputs with pinned theta."""
    import torch
    eng = _cabi.Engine(0)
    wl = workloads.config(1, n_ph=150)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    theta = wl.walkers(1500, scatter=0.05, seed=8)          # two lanes; some walkers outside the priors
    ref, rchi = eng.log_prob(theta, return_chisq=True)
    tp = torch.from_numpy(theta).pin_memory()
    op = torch.empty(theta.shape[0], dtype=torch.float64).pin_memory()
    for _ in range(3):                                      # (the third identical call may replay a graph)
        op.fill_(7.0)
        got = eng.log_prob(tp.numpy(), out=op.numpy())
        assert got is not None and np.array_equal(op.numpy(), ref, equal_nan=True)
    small = theta[:300]                                     # one batch: the CUDA-graph path
    sp = torch.from_numpy(small).pin_memory()
    for _ in range(4):
        assert np.array_equal(eng.log_prob(sp.numpy()), ref[:300], equal_nan=True)
    d_out = torch.empty(theta.shape[0], dtype=torch.float64, device="cuda")
    eng.log_prob_device(tp.data_ptr(), theta.shape[0], d_out.data_ptr())
    tp.fill_(0.0)                                           # the call has finished reading the pinned buffer
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), ref, equal_nan=True)
    eng.close()
