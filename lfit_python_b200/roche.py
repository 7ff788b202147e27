"""`trm.roche` replacement for the calls the tree makes (all evaluated on the GPU).

Reference call sites: roche.xl1(q) /root/reference/CVModel.py:222, roche.bspot(q, rad)
CVModel.py:288, roche.findphi(q, 90) CVModel.py:460, roche.findi(q, dphi) CVModel.py:561.
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
