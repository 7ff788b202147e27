#!/usr/bin/env python
"""Where does one parameter set's CUDA light curve differ from the oracle's?  Components, phases, and the
ingress / egress phases of the surface elements behind the difference.

    PYTHONPATH=. python tools/diag_walker.py n_ph shape  p0 p1 ... p17
"""
import sys

import numpy as np

from oracle import oracle as O
from lfit_python_b200 import _cabi


def main(n_ph, shape, pars):
    eng = _cabi.Engine(0)
    if shape == 3:
        x = np.linspace(-1.2, 1.3, n_ph)
    else:
        x = np.linspace(-0.5, 0.5, n_ph)
    w = np.full(n_ph, 0.5 * (x.max() - x.min()) / n_ph) if shape == 3 else np.full(n_ph, np.mean(np.diff(x)) / 2)
    tot, comp = eng.calc_flux(pars, x, w, components=True)
    st, rtot, rcomp = O.calc_flux(pars, x, w, components=True)
    print("status", st, "max |dflux|", np.max(np.abs(tot - rtot)), "rel", np.max(np.abs(tot - rtot) / np.abs(rtot)))
    for k, nm in enumerate(("wd", "disc", "spot", "donor")):
        d = np.abs(comp[k] - np.asarray(rcomp[k]))
        j = int(np.argmax(d))
        print("  %-5s max diff %.3e at phase %.6f (value %.6e)" % (nm, d[j], x[j], rcomp[k][j]))
        bad = np.where(d > 1e-9 * max(np.max(np.abs(rcomp[k])), 1e-300))[0]
        if len(bad):
            print("        points off by > 1e-9 of the component's peak:", len(bad), "phases", x[bad][:6])
    eng.close()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), np.array([float(v) for v in sys.argv[3:]]))
