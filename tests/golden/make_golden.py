#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the CPU oracle (robust solver) in THIS container.

There are no reference outputs to pin against: `lfit` / `trm.roche` are not under
/root/reference and the reference holds no expected model output.
The vectors below therefore freeze the oracle itself (robust bisection solver, direct sums),
so that (a) the fast Newton oracle, (b) the CUDA path and (c) later refactors are all held
to the same numbers.  First-principles known answers for the Roche scalars are in
tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from lfit_python_b200 import workloads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    robust = O.config(solver=O.SOLVER_ROBUST)
    # 1. calcFlux curves: testCV.py parameters (simple + complex) and the six example eclipses
    phi = np.linspace(-0.25, 0.25, 120)
    width = np.full_like(phi, 0.0015)
    testcv = [0.333, 0.333, 0.333, 0.05, 0.1, 0.0607135, 0.6, 0.4, 0.0139, 0.039, 157.0, 0.2, 0.2, 0.0,
              2.0, 1.0, 120.0, 1.0]
    wl = workloads.config(2)
    sets = [testcv, testcv[:14]] + [list(wl.cv_pars(wl.p0, e)) for e in (0, 1, 2, 3, 4, 5)]
    sets.append(list(wl.cv_pars(wl.p0, 0))[:14])
    pars = np.full((len(sets), 18), np.nan)
    tot = np.empty((len(sets), phi.size))
    comp = np.empty((len(sets), 4, phi.size))
    for k, p in enumerate(sets):
        pars[k, : len(p)] = p
        st, t, c = O.calc_flux(p, phi, width, cfg=robust, components=True)
        assert st == 0
        tot[k], comp[k] = t, np.asarray(c)
    np.savez_compressed(os.path.join(HERE, "calc_flux.npz"), phase=phi, width=width, pars=pars, total=tot, comp=comp)

    # 2. ln_prior / ln_like / ln_prob of a small tree (2 bands x 2 eclipses, complex BS)
    wl = workloads.config(2, n_bands=2, ecl_per_band=2, n_ph=90, phase_range=(-0.15, 0.2))
    wl.make_data(lambda p, x, w: O.calc_flux(p, x, w, cfg=robust)[1], seed=7)
    lay = O.FlatLayout(wl.ndim, wl.npars, wl.gather, wl.consts, wl.prior_src, wl.prior_type, wl.prior_p1,
                       wl.prior_p2, wl.prior_norm, wl.prior_isvar, wl.lc_off, wl.lc_phase, wl.lc_width, wl.lc_y,
                       wl.lc_ye)
    theta = wl.walkers(24, scatter=0.04, seed=11)
    theta[0] = wl.p0
    out = {}
    for what, name in ((0, "ln_prior"), (1, "ln_like"), (2, "ln_prob")):
        v, chi = O.log_prob(lay, theta, what=what, cfg=robust, return_chisq=True)
        out[name] = v
        out[name + "_chisq"] = chi
    np.savez_compressed(os.path.join(HERE, "log_prob.npz"), theta=theta, lc_y=wl.lc_y, **out)

    # 3. Roche scalars over a grid of q
    q = np.array([0.03, 0.05, 0.1, 0.1037, 0.2, 0.35, 0.5, 0.8, 1.0, 1.7])
    xl1 = np.array([O.xl1(v) for v in q])
    maxphi = np.array([O.findphi(v, 90.0) for v in q])
    inc = np.array([O.findi(v, 0.6 * m) for v, m in zip(q, maxphi)])
    spot = np.array([O.bspot(v, 0.4 * x) for v, x in zip(q, xl1)])
    ie = []
    for v in q[:6]:
        for p0 in ((0.0, 0.0, 0.0), (0.2, 0.1, 0.0), (-0.15, 0.2, 0.0), (0.1, -0.25, 0.01)):
            r = O.ingress_egress(v, 84.0, p0, solver=O.SOLVER_ROBUST)
            ie.append([v, 84.0, *p0, *(r if r else (np.nan, np.nan))])
    np.savez_compressed(os.path.join(HERE, "roche.npz"), q=q, xl1=xl1, maxphi=maxphi, incl=inc, bspot=spot,
                        ingress_egress=np.asarray(ie))
    # 4. Gaussian-process likelihood (dense Cholesky statement of the reference's george kernel) and the
    #    white-dwarf contact phases behind its change points
    rng = np.random.default_rng(2024)
    cases = []
    for n, span in ((40, (-0.2, 0.3)), (150, (-0.5, 0.5)), (90, (0.05, 0.45)), (200, (-1.2, 1.3)), (1, (0.0, 0.0))):
        x = np.sort(rng.uniform(span[0], span[1], n)) if n > 1 else np.array([0.01])
        ye = rng.uniform(0.002, 0.006, n)
        hyper = np.exp([rng.uniform(-12, -7), rng.uniform(-12, -7), rng.uniform(-8, -3)])
        gaps = np.asarray(O.gp_changepoints(x, rng.uniform(0.02, 0.08), rng.normal(0, 0.002)), dtype=np.float64).reshape(-1, 2)
        r = rng.normal(0.0, 0.006, n)
        cases.append(dict(x=x, ye=ye, resid=r, hyper=hyper, gaps=gaps, lnl=O.gp_log_like(x, ye, r, *hyper, gaps.tolist())))
    gp = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            gp["%s_%d" % (k, i)] = v
    wq = np.array([[0.1037, 0.0392, 0.0187], [0.3, 0.06, 0.012], [0.05, 0.03, 0.02], [0.7, 0.09, 0.01]])
    wd = np.array([[O.findi(q_, d_), *O.wdphases(q_, O.findi(q_, d_), r_, 10), O.gp_dist_cp(q_, d_, r_, 10)] for q_, d_, r_ in wq])
    np.savez_compressed(os.path.join(HERE, "gp.npz"), n_cases=len(cases), wd_in=wq, wd_out=wd, **gp)
    print("wrote calc_flux.npz, log_prob.npz, roche.npz, gp.npz")


if __name__ == "__main__":
    main()
