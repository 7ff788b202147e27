#!/usr/bin/env python
"""Soak test: many log-probability calls of changing size, chi-squared and GP modes interleaved, on one
engine; results must not depend on the history and device memory must stay flat.

    PYTHONPATH=. python tools/soak.py [iterations]
"""
import sys

import numpy as np
import torch

from lfit_python_b200 import _cabi, workloads


def main(iters=400):
    wl = workloads.config(1, n_ph=700)
    eng = _cabi.Engine(0)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    theta = wl.walkers(3000, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR))
    ref = eng.log_prob(theta)
    wl.apply_gp(eng)
    ref_gp = eng.log_prob(theta)
    rng = np.random.default_rng(1)
    free0 = None
    for it in range(iters):
        gp = bool(it & 1)
        (wl.apply_gp if gp else wl.apply)(eng)
        n = int(rng.integers(1, 3000))
        lo = int(rng.integers(0, 3000 - n + 1))
        got = eng.log_prob(theta[lo:lo + n])
        want = (ref_gp if gp else ref)[lo:lo + n]
        if not np.array_equal(got, want):
            bad = np.where(got != want)[0]
            print("iteration %d (gp=%s, n=%d): %d values differ, first %r vs %r" % (it, gp, n, bad.size, got[bad[0]], want[bad[0]]))
            return 1
        if it == 20:
            free0 = torch.cuda.mem_get_info()[0]
    free1 = torch.cuda.mem_get_info()[0]
    print("%d calls bit-identical to the first evaluation; device memory drift %.1f MB" % (iters, (free0 - free1) / 1e6))
    return 0 if abs(free0 - free1) < 64e6 else 1


if __name__ == "__main__":
    sys.exit(main(*[int(a) for a in sys.argv[1:]]))
