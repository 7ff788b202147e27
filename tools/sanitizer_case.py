"""A small case for compute-sanitizer: every kernel of a log-probability pass (two lanes), calc_flux and the device sampler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lfit_python_b200 import _cabi, workloads, mcmc_utils

eng = _cabi.Engine(0)
wl = workloads.config(1, n_ph=300)
wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
wl.apply(eng)
theta = wl.walkers(1200, scatter=0.05, seed=3)
for _ in range(3):
    lnp = eng.log_prob(theta)
print("finite", int(np.isfinite(lnp).sum()), "of", len(lnp))
s = mcmc_utils.DeviceSampler(eng, 2400, seed=1)
s.set_state(wl.walkers(2400, scatter=0.01, seed=4))
s.run(3)
print("sampler ok", s.get_state()[1][:3])
s.close()
eng.close()
