"""`trm.roche` replacement for the calls the tree makes (all evaluated on the GPU).

Reference call sites: roche.xl1(q) /root/reference/CVModel.py:222, roche.bspot(q, rad)
CVModel.py:288, roche.findphi(q, 90) CVModel.py:460, roche.findi(q, dphi) CVModel.py:559,
roche.wdphases(q, inc, rwd, ntheta=10) CVModel.py:562.
Failures raise RocheError, which is both an AssertionError (what CVModel.py:223 catches
for xl1) and therefore an Exception (what CVModel.py:309,475 catch).
"""
import numpy as np

from . import _cabi


class RocheError(AssertionError):
    pass


def _call(which, a, b=None):
    eng = _cabi.default_engine()
    scalar = np.ndim(a) == 0
    out, ok = eng.roche(which, a, b)
    if not ok.all():
        raise RocheError("roche: no solution for %s" % ("q=%r" % (a,) if b is None else "(%r, %r)" % (a, b)))
    return out, scalar


def xl1(q):
    """Distance of the inner Lagrangian point from the primary, units of the separation."""
    out, scalar = _call(_cabi.ROCHE_XL1, q)
    return float(out[0, 0]) if scalar else out[:, 0]


def findphi(q, iangle):
    """Full phase width of the eclipse of the white-dwarf centre at inclination iangle (deg)."""
    out, scalar = _call(_cabi.ROCHE_FINDPHI, q, iangle)
    return float(out[0, 0]) if scalar else out[:, 0]


def findi(q, deltaphi):
    """Inclination (deg) at which the white-dwarf centre is eclipsed for deltaphi of the orbit."""
    out, scalar = _call(_cabi.ROCHE_FINDI, q, deltaphi)
    return float(out[0, 0]) if scalar else out[:, 0]


def bspot(q, rad):
    """(x, y, vx, vy) where the ballistic stream from L1 first reaches radius rad (units of a);
    raises if it never does."""
    out, scalar = _call(_cabi.ROCHE_BSPOT, q, rad)
    return tuple(float(v) for v in out[0]) if scalar else tuple(out.T)


def wdphases(q, iangle, r1, r2=-1, ntheta=100):
    """(phi3, phi4): third and fourth contact phases of the white dwarf of radius r1 (units of a)
    at inclination iangle (deg): the earliest and latest egress over ntheta points on its limb.
    r2 (donor radius; -1 = Roche-lobe filling) is accepted for signature compatibility only."""
    if r2 != -1:
        raise NotImplementedError("wdphases: only a Roche-lobe filling donor (r2 = -1)")
    out, ok = _cabi.default_engine().wdphases(q, iangle, r1, ntheta)
    if not ok.all():
        raise RocheError("wdphases: the white dwarf is not fully eclipsed for (%r, %r, %r)" % (q, iangle, r1))
    return (float(out[0, 0]), float(out[0, 1])) if np.ndim(q) == 0 else (out[:, 0], out[:, 1])
