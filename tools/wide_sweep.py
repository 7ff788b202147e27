#!/usr/bin/env python
"""Parity sweep over the whole prior box (not just around the truth): validity masks and chi-squared of
the CUDA path against the CPU oracle for walkers drawn uniformly between the prior limits.

    PYTHONPATH=. python tools/wide_sweep.py [n_walkers] [n_points] [config]
"""
import sys
import time

import numpy as np

from oracle import oracle as O
from lfit_python_b200 import _cabi, workloads


def main(n=2000, n_ph=200, cfg=1, seed=5, shape=0):
    wl = workloads.config(cfg, n_ph=n_ph)
    eng = _cabi.Engine(0, **wl.grid)
    srng = np.random.default_rng(100 + shape)
    if shape:  # other light-curve shapes: irregular / unsorted / several cycles / uneven exposures
        for e in range(wl.n_ecl):
            sl = slice(wl.lc_off[e], wl.lc_off[e + 1])
            if shape == 1:
                x = np.sort(srng.uniform(-0.5, 0.5, n_ph))
            elif shape == 2:
                x = srng.permutation(np.linspace(-0.3, 0.4, n_ph))
            elif shape == 3:
                x = np.linspace(-1.2, 1.3, n_ph)
            else:
                x = np.sort(srng.uniform(0.7, 1.4, n_ph))
            wl.lc_phase[sl] = x
            wl.lc_width[sl] = srng.uniform(0.0, 0.004, n_ph) if shape in (1, 4) else np.full(n_ph, 0.5 * (x.max() - x.min()) / n_ph)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    rng = np.random.default_rng(seed)
    lo, hi = wl.prior_p1.copy(), wl.prior_p2.copy()
    gauss = np.isin(wl.prior_type, (0, 1))
    lo[gauss], hi[gauss] = wl.prior_p1[gauss] - 3 * wl.prior_p2[gauss], wl.prior_p1[gauss] + 3 * wl.prior_p2[gauss]
    theta = lo + (hi - lo) * rng.random((n, wl.ndim))
    # half of them: the truth for everything but a few wide-open parameters, so that more models are valid
    k = n // 2
    keep = rng.random((k, wl.ndim)) < 0.7
    theta[:k] = np.where(keep, wl.p0, theta[:k])
    lay = O.FlatLayout(wl.ndim, wl.npars, wl.gather, wl.consts, wl.prior_src, wl.prior_type, wl.prior_p1, wl.prior_p2,
                       wl.prior_norm, wl.prior_isvar, wl.lc_off, wl.lc_phase, wl.lc_width, wl.lc_y, wl.lc_ye)
    t0 = time.time()
    ocfg = O.config(**wl.grid)
    ref, rchi = O.log_prob(lay, theta, what=_cabi.LN_LIKE, return_chisq=True, cfg=ocfg)
    t1 = time.time()
    got, chi = eng.log_prob(theta, what=_cabi.LN_LIKE, return_chisq=True)
    pri_ref = O.log_prob(lay, theta, what=_cabi.LN_PRIOR, cfg=ocfg)
    pri = eng.log_prob(theta, what=_cabi.LN_PRIOR)
    fin_r, fin_g = np.isfinite(rchi[:, 0]), np.isfinite(chi[:, 0])
    print("walkers %d, oracle %.1f s; valid models: oracle %d, cuda %d; mask mismatches %d; prior mask mismatches %d" % (
        n, t1 - t0, fin_r.sum(), fin_g.sum(), (fin_r != fin_g).sum(), (np.isfinite(pri_ref) != np.isfinite(pri)).sum()))
    both = fin_r & fin_g
    rel = np.abs(chi[both, 0] - rchi[both, 0]) / np.abs(rchi[both, 0])
    print("chi-squared relative difference: max %.3g, 99.9%% %.3g, median %.3g; > 1e-7: %d" % (
        rel.max(), np.quantile(rel, 0.999), np.median(rel), (rel > 1e-7).sum()))
    bad = np.where(both)[0][rel > 1e-7]
    for i in bad[:5]:
        print("  walker", i, "chi", chi[i, 0], rchi[i, 0], "theta", np.array2string(theta[i], precision=5, max_line_width=200))
    for i in np.where(fin_r != fin_g)[0][:5]:
        print("  mask", i, chi[i, 0], rchi[i, 0], np.array2string(theta[i], precision=5, max_line_width=200))
    return int((fin_r != fin_g).sum() + (rel > 1e-7).sum())


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    sys.exit(1 if main(*a) else 0)
