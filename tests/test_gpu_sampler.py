"""GPU: the device-resident stretch-move sampler (csrc/sampler.cuh through lfb_sampler_*) against its
numpy restatement (oracle/stretch.py) driving the same engine, the chain recorder / writer, and the
statistics of what it samples.  Replaces emcee.EnsembleSampler + pool and the loops of
/root/reference/mcmc_utils.py:114-183."""
import numpy as np
import pytest

from oracle.stretch import StretchOracle
from lfit_python_b200 import _cabi, mcmc_utils, workloads

pytestmark = pytest.mark.gpu


def setup(engine, cfg=1, n_ph=150, **kw):
    wl = workloads.config(cfg, n_ph=n_ph, **kw)
    wl.make_data(lambda p, x, w: engine.calc_flux(p, x, w))
    wl.apply(engine)
    return wl


@pytest.mark.parametrize("nwalkers,steps", [(64, 12), (2500, 3)])   # CUDA-graph replay / plain launches, two lanes
def test_device_sampler_equals_numpy_restatement(engine, nwalkers, steps):
    wl = setup(engine)
    p0 = wl.walkers(nwalkers, scatter=0.01, seed=4, ln_prior_fn=lambda t: engine.log_prob(t, what=_cabi.LN_PRIOR))
    dev = mcmc_utils.DeviceSampler(engine, nwalkers, seed=77)
    ref = StretchOracle(lambda t: engine.log_prob(t), nwalkers, wl.ndim, seed=77)
    dev.set_state(p0)
    ref.set_state(p0)
    for k in (1, steps - 1):          # the first step runs plainly, later ones replay the graph (small ensembles)
        dev.run(k)
        ref.run(k)
        pos, lnp = dev.get_state()
        assert np.array_equal(pos, ref.pos) and np.array_equal(lnp, ref.lnp)
        assert np.array_equal(dev.naccepted, ref.naccepted)
    assert dev.iterations == steps and 0 < dev.naccepted.sum() < nwalkers * steps
    assert np.array_equal(lnp, engine.log_prob(pos))          # stored ln_prob belongs to the stored position
    dev.close()


def test_device_sampler_sharded_calls_equal_whole_ensemble_calls(engine):
    """half_begin / half_end over two slices (what two ranks would do, here on one GPU) == run()."""
    import torch
    wl = setup(engine)
    nw = 80
    p0 = wl.walkers(nw, scatter=0.01, seed=9, ln_prior_fn=lambda t: engine.log_prob(t, what=_cabi.LN_PRIOR))
    a = mcmc_utils.DeviceSampler(engine, nw, seed=3)
    b = mcmc_utils.DeviceSampler(engine, nw, seed=3)
    a.set_state(p0)
    b.set_state(p0)
    a.run(4)
    half, world = nw // 2, 3
    slot = -(-half // world)
    lib, chk = b._lib, engine._check
    gathered = torch.zeros((world, slot, wl.ndim + 2), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    from lfit_python_b200.parallel import shard_bounds
    for _ in range(4):
        for h in (0, 1):
            for r in range(world):
                lo, hi = shard_bounds(half, r, world)
                chk(lib.lfb_sampler_half_begin(b._s, h, lo, hi, gathered[r].data_ptr(), None), "half_begin")
            chk(lib.lfb_sampler_half_end(b._s, h, gathered.data_ptr(), world, slot, None), "half_end")
    pa, la = a.get_state()
    pb, lb = b.get_state()
    assert np.array_equal(pa, pb) and np.array_equal(la, lb) and np.array_equal(a.naccepted, b.naccepted)
    assert a.iterations == b.iterations == 4
    a.close()
    b.close()


def test_chain_recorder_and_writer(engine, tmp_path):
    wl = setup(engine, cfg=0)
    nw = 50
    p0 = wl.walkers(nw, scatter=0.01, seed=2, ln_prior_fn=lambda t: engine.log_prob(t, what=_cabi.LN_PRIOR))
    s = mcmc_utils.DeviceSampler(engine, nw, seed=5)
    path = tmp_path / "chain_prod.txt"
    names = "walker_no " + " ".join(wl.names) + " ln_prob"
    chain = mcmc_utils.run_mcmc_save_device(s, p0, 23, str(path), col_names=names, block=8)   # blocks of 8, 8, 7
    assert chain.shape == (nw, 23, wl.ndim + 1)
    pos, lnp = s.get_state()
    assert np.array_equal(chain[:, -1, :wl.ndim], pos) and np.array_equal(chain[:, -1, wl.ndim], lnp)
    # every recorded ln_prob belongs to its recorded position
    assert np.array_equal(chain[:, 11, wl.ndim], engine.log_prob(np.ascontiguousarray(chain[:, 11, :wl.ndim])))
    # the file holds the reference's bytes (mcmc_utils.py:157-164) and its reader round-trips (:252-272)
    text = path.read_text()
    ref = names + "\n" + "".join(mcmc_utils.format_step_python(chain[:, k, :wl.ndim], chain[:, k, wl.ndim]) for k in range(23))
    assert text == ref
    back = mcmc_utils.readchain(str(path))
    assert back.shape == (nw, 23, wl.ndim + 1) and np.array_equal(back[:, :, :wl.ndim], chain[:, :, :wl.ndim])
    # the generic loops take the device sampler too
    mcmc_utils.run_mcmc_save(s, None, 5, None, str(path), col_names=names)
    assert len(path.read_text().splitlines()) == 1 + 5 * nw
    pos2, lnp2, _ = mcmc_utils.run_burnin(s, pos, 3)
    assert s.iterations == 23 + 5 + 3 and np.array_equal(lnp2, engine.log_prob(pos2))
    s.close()


def test_device_sampler_samples_the_prior(engine):
    """what = LN_PRIOR: the target is the prior, whose independent margins are known --
    ulimb ~ N(0.284, 0.001) and wdFlux ~ U(0.001, 0.2) (test_data/mcmc_input.dat:53-55)."""
    wl = setup(engine, cfg=0)
    nw = 64
    p0 = wl.walkers(nw, scatter=0.05, seed=1, ln_prior_fn=lambda t: engine.log_prob(t, what=_cabi.LN_PRIOR))
    # (the ball is 1e-6 wide in ulimb, mcmcfit.py:220, and the stretch move widens a dimension only by a factor
    # of order one per step: start these two margins at their prior width)
    rng = np.random.default_rng(0)
    p0[:, wl.names.index("ulimb_b0")] = 0.284 + 0.001 * rng.standard_normal(nw)
    p0[:, wl.names.index("wdFlux_b0")] = rng.uniform(0.002, 0.199, nw)
    s = mcmc_utils.DeviceSampler(engine, nw, seed=12, what=_cabi.LN_PRIOR)
    s.set_state(p0)
    s.run(3000)
    rows = np.concatenate(list(s.run_block(6000, block=500)), axis=0)[::10]
    ul = rows[:, :, wl.names.index("ulimb_b0")].ravel()
    wd = rows[:, :, wl.names.index("wdFlux_b0")].ravel()
    assert abs(ul.mean() - 0.284) < 1e-4 and abs(ul.std() / 0.001 - 1.0) < 0.1
    assert abs(wd.mean() - 0.1005) < 0.006 and abs(wd.std() / (0.199 / np.sqrt(12.0)) - 1.0) < 0.1
    assert wd.min() > 0.001 and wd.max() < 0.2 and np.isfinite(rows[:, :, -1]).all()
    assert 0.05 < s.acceptance_fraction.mean() < 0.9
    s.close()
