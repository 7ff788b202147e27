#!/usr/bin/env python
"""Ingress / egress phases of the white-dwarf tiles of one parameter set: CUDA (as the pipeline solves them)
against the oracle's robust solver, and how close each boundary comes to an exposure sample."""
import sys

import numpy as np

from oracle import oracle as O
from lfit_python_b200 import _cabi

pars = np.load(sys.argv[1])
n_ph, shape = int(sys.argv[2]), int(sys.argv[3])
q, dphi, rwd, phi0 = pars[4], pars[5], pars[8], pars[13]
eng = _cabi.Engine(0)
inc = eng.roche(_cabi.ROCHE_FINDI, q, dphi)[0][0, 0]
xl1 = O.xl1(q)
n = 10
pts = []
for t in range(2 * n * n):
    k = int(np.sqrt(0.5 * t))
    while 2 * k * k > t: k -= 1
    while 2 * (k + 1) * (k + 1) <= t: k += 1
    r, q1 = t - 2 * k * k, 2 * k + 1
    nk = 4 * q1
    j = r if r < q1 else r + 2 * q1
    rho = np.sqrt(0.5 * ((k / n) ** 2 + ((k + 1) / n) ** 2))
    a = (j + 0.5) * 2 * np.pi / nk
    pts.append([0, 0, 0, rwd * xl1 * rho * np.cos(a), rwd * xl1 * rho * np.sin(a)])
pts = np.array(pts)
got, ok = eng.ingress_egress(q, inc, pts)
ref = np.array([O.ingress_egress(q, inc, (0, 0, 0), xi=p[3], eta=p[4], solver=O.SOLVER_ROBUST) or (np.nan, np.nan) for p in pts])
refn = np.array([O.ingress_egress(q, inc, (0, 0, 0), xi=p[3], eta=p[4], solver=O.SOLVER_NEWTON) or (np.nan, np.nan) for p in pts])
print("inclination", inc, "tiles", len(pts), "eclipsed cuda", int(ok.sum()), "oracle", int(np.isfinite(ref[:, 0]).sum()))
d = np.abs(got - ref)
print("max |cuda - oracle robust| ingress %.3e egress %.3e ; vs oracle newton %.3e" % (np.nanmax(d[:, 0]), np.nanmax(d[:, 1]), np.nanmax(np.abs(got - refn))))
worst = np.argsort(-np.nanmax(d, axis=1))[:5]
for i in worst:
    print("  tile", i, "xi,eta", pts[i, 3:], "cuda", got[i], "oracle", ref[i], "newton", refn[i])
eng.close()
