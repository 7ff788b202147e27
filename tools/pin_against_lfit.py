#!/usr/bin/env python
"""Pin the oracle (and with it the CUDA path) against REAL lfit / trm.roche, wherever those are installed.

The reference's arithmetic lives in two third-party packages that are not vendored, pinned or installable in the
build container (`import lfit`, `from trm import roche`: /root/reference/CVModel.py:13,15), so every parity claim of
this repo is "CUDA == oracle", never "== lfit" (DESIGN.md section 1).  This tool is the hook that closes the gap
the moment a machine has them:

    python tools/pin_against_lfit.py            # needs lfit + trm.roche; writes tests/golden/lfit_pin.npz

It dumps, from the real packages, what the reference's call sites consume:
  * lfit.CV(pars).calcFlux(pars, phase, width) and the ywd / yd / ys / yrs attributes for testCV.py's parameter set
    (testCV.py:17-63) and the six parameter sets of test_data/mcmc_input.dat (restated in lfit_python_b200/workloads.py),
    simple (14) and complex (18 parameters), on the reference's own dummy grid (CVModel.py:880-889) with
    width = mean(diff(phase)) / 2 (CVModel.py:64);
  * the unit-normalised component curves of lfit.PyWhiteDwarf / PyDisc / PySpot / PyDonor (testCV.py:27-49);
  * roche.xl1 / findi / findphi / bspot / wdphases at the values the tree asks for (CVModel.py:222,288,460,559,562).

tests/test_lfit_pin.py is skipped while the file is absent; once it exists it reports, per component, the largest
relative difference between the oracle and real lfit (this is where the discretisation choices of DESIGN.md section 2
-- tile counts, quadrature order, donor darkening constants, normalisations -- get fitted), and pins the Roche
scalars, which have no free choices, to 1e-6.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "lfit_pin.npz")

TESTCV = dict(q=0.1, inc=86.9, rwd_a=0.01, rdisc=0.6, dexp=0.2, az=157.0, fis=0.2, scale=0.039, exp1=2.0, exp2=1.0,
              tilt=120.0, yaw=1.0, ulimb=0.4)   # testCV.py:17-43


def main():
    try:
        import lfit
        from trm import roche
    except ImportError as exc:
        print("pin_against_lfit: real lfit / trm.roche are not importable here (%s); nothing written." % exc)
        return 2
    from lfit_python_b200 import workloads
    out = {}
    phase = np.linspace(-0.5, 0.5, 1000)
    width = np.full_like(phase, np.mean(np.diff(phase)) / 2.0)
    out["phase"], out["width"] = phase, width
    # --- lfit.CV on the example's parameter sets
    sets = []
    wl = workloads.Workload("pin", 3, 2, 10, complex_bs=True)
    for e in range(wl.n_ecl):
        sets.append(wl.cv_pars(wl.p0, e))
    t = TESTCV
    xl1 = roche.xl1(t["q"])
    dphi = roche.findphi(t["q"], t["inc"])
    sets.append(np.array([1 / 3, 1 / 3, 1 / 3, 0.05, t["q"], dphi, t["rdisc"], t["ulimb"], t["rwd_a"] / xl1, t["scale"],
                          t["az"], t["fis"], t["dexp"], 0.0, t["exp1"], t["exp2"], t["tilt"], t["yaw"]]))
    out["n_sets"] = len(sets)
    for k, pars in enumerate(sets):
        for npar in (14, 18):
            p = np.asarray(pars[:npar], dtype=float)
            cv = lfit.CV(p)
            flux = np.asarray(cv.calcFlux(p, phase, width))
            out["pars_%d_%d" % (npar, k)] = p
            out["flux_%d_%d" % (npar, k)] = flux
            out["comp_%d_%d" % (npar, k)] = np.asarray([cv.ywd, cv.yd, cv.ys, cv.yrs])
    # --- the component classes, as testCV.py drives them
    wd = lfit.PyWhiteDwarf(t["rwd_a"] / xl1, t["ulimb"])
    disc = lfit.PyDisc(t["q"], t["rwd_a"] / xl1, t["rdisc"], t["dexp"], 1000)
    spot = lfit.PySpot(t["q"], t["rdisc"], t["az"], t["fis"], t["scale"])
    donor = lfit.PyDonor(t["q"], 400)
    out["unit_wd"] = np.asarray(wd.calcFlux(t["q"], t["inc"], phase, width))
    out["unit_disc"] = np.asarray(disc.calcFlux(t["q"], t["inc"], phase, width))
    out["unit_spot"] = np.asarray(spot.calcFlux(t["q"], t["inc"], phase, width))
    out["unit_donor"] = np.asarray(donor.calcFlux(t["q"], t["inc"], phase, width))
    # --- trm.roche
    qs = np.array([0.05, 0.1, 0.1037, 0.2, 0.5, 1.0])
    out["roche_q"] = qs
    out["roche_xl1"] = np.array([roche.xl1(q) for q in qs])
    out["roche_findphi90"] = np.array([roche.findphi(q, 90.0) for q in qs])
    out["roche_findi_0392"] = np.array([roche.findi(q, 0.0392) for q in qs[:5]])
    rad = np.array([0.2953, 0.5214, 0.4]) * roche.xl1(0.1037)
    out["roche_bspot_rad"] = rad
    out["roche_bspot"] = np.array([roche.bspot(0.1037, r) for r in rad])
    inc = roche.findi(0.1037, 0.0392)
    out["roche_wdphases"] = np.array(roche.wdphases(0.1037, inc, 0.0187 * roche.xl1(0.1037), ntheta=10))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
