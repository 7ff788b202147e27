"""Oracle (and, with a GPU, the CUDA path) against REAL lfit / trm.roche -- runs only once someone with those
packages has produced tests/golden/lfit_pin.npz with tools/pin_against_lfit.py.  Until then parity stays
"unpinned"  and these tests are skipped."""
import os

import numpy as np
import pytest

from oracle import oracle as O

PIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lfit_pin.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PIN), reason="no tests/golden/lfit_pin.npz: real lfit was never available "
                                "(tools/pin_against_lfit.py writes it where lfit and trm.roche are installed)")


def _pin():
    return np.load(PIN)


def test_roche_scalars_against_trm_roche():
    g = _pin()
    np.testing.assert_allclose([O.xl1(q) for q in g["roche_q"]], g["roche_xl1"], rtol=1e-6)
    np.testing.assert_allclose([O.findphi(q, 90.0) for q in g["roche_q"]], g["roche_findphi90"], rtol=1e-5)
    np.testing.assert_allclose([O.findi(q, 0.0392) for q in g["roche_q"][:5]], g["roche_findi_0392"], rtol=1e-5)
    for rad, want in zip(g["roche_bspot_rad"], g["roche_bspot"]):
        np.testing.assert_allclose(O.bspot(0.1037, rad), want, rtol=0, atol=1e-4)


def test_calc_flux_against_lfit_reports_the_gap(capsys):
    """Reports the largest relative difference per parameter set and component; the discretisation constants of
    DESIGN.md section 2 are to be fitted until this reaches the north-star 1e-9."""
    g = _pin()
    worst = 0.0
    for k in range(int(g["n_sets"])):
        for npar in (14, 18):
            pars, want = g["pars_%d_%d" % (npar, k)], g["flux_%d_%d" % (npar, k)]
            st, tot, comp = O.calc_flux(pars, g["phase"], g["width"], components=True)
            assert st == 0
            rel = np.max(np.abs(tot - want) / np.maximum(np.abs(want), 1e-300))
            per = [np.max(np.abs(np.asarray(c) - w)) / max(np.max(np.abs(w)), 1e-300) for c, w in zip(comp, g["comp_%d_%d" % (npar, k)])]
            print("set %d (%d pars): flux %.3e; wd %.3e disc %.3e spot %.3e donor %.3e" % ((k, npar, rel) + tuple(per)))
            worst = max(worst, rel)
    print("largest relative flux difference oracle vs lfit: %.3e" % worst)
    assert worst < 0.05     # the same model; the north-star bar (1e-9) is what the constants are to be fitted to
