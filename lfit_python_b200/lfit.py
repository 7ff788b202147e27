"""`lfit` replacement: the CV eclipse model and its four components, evaluated on the GPU.

Reference call sites: lfit.CV(pars) /root/reference/CVModel.py:128; cv.calcFlux(pars, phase,
width) CVModel.py:138,154, plot_lc_model.py:134; attributes ywd/yd/ys/yrs CVModel.py:155,
plot_lc_model.py:135-138; component classes testCV.py:27-49, fitEcl.py:21-24.
Parameter order and meaning: README.md:24-43.  There is no CPU fallback.
"""
import numpy as np

from . import _cabi

_DEFAULT_DISC = 1000   # testCV.py:31
_DEFAULT_DONOR = 400   # testCV.py:43
_engines = {}


def _engine(**grid):
    """Engines are cached per surface-grid setting (the default grid is the process-wide one)."""
    if not grid:
        return _cabi.default_engine()
    key = tuple(sorted(grid.items()))
    if key not in _engines:
        _engines[key] = _cabi.Engine(0, **grid)
    return _engines[key]


def _width(phi, width):
    phi = np.ascontiguousarray(phi, dtype=np.float64).ravel()
    if width is None:
        # "If it's not defined, the software will infer the bin width from the data" (README.md:63)
        width = np.mean(np.diff(phi)) / 2.0 if phi.size > 1 else 0.0
    return phi, np.broadcast_to(np.asarray(width, dtype=np.float64), phi.shape)


class CV:
    """lfit.CV: white dwarf + disc + bright spot + donor.  14 parameters select the simple
    bright-spot model, 18 the complex one (exp1, exp2, tilt, yaw)."""

    def __init__(self, pars, **grid):
        pars = np.asarray(pars, dtype=np.float64)
        if pars.shape[0] not in (14, 18):
            raise ValueError("CV takes 14 (simple BS) or 18 (complex BS) parameters, got %d" % pars.shape[0])
        self._grid = grid  # the engine (and with it the GPU) is only touched by calcFlux
        self.pars = pars.copy()
        self.ywd = self.yd = self.ys = self.yrs = None

    def calcFlux(self, pars, phi, width=None):
        phi, width = _width(phi, width)
        tot, comp = _engine(**self._grid).calc_flux(np.asarray(pars, dtype=np.float64), phi, width, components=True)
        self.pars = np.asarray(pars, dtype=np.float64).copy()
        self.ywd, self.yd, self.ys, self.yrs = comp[0], comp[1], comp[2], comp[3]
        if np.isnan(tot).any():
            raise ValueError("CV.calcFlux: these parameters admit no model")
        return tot

    __call__ = calcFlux


class _Component:
    _skip = 0
    _grid = {}

    def _flux(self, pars, phi, width):
        phi, width = _width(phi, width)
        out = _engine(**self._grid).calc_flux(np.asarray(pars, dtype=np.float64), phi, width,
                                  flags=_cabi.FLAG_INCL | self._skip)
        if np.isnan(out).any():
            raise ValueError("%s.calcFlux: these parameters admit no model" % type(self).__name__)
        return out


def _pars(**kw):
    p = dict(wdFlux=0.0, dFlux=0.0, sFlux=0.0, rsFlux=0.0, q=0.1, dphi=80.0, rdisc=0.5, ulimb=0.3, rwd=0.01,
             scale=0.02, az=120.0, fis=0.2, dexp=0.5, phi0=0.0, exp1=2.0, exp2=1.0, tilt=90.0, yaw=0.0)
    p.update(kw)
    order = ["wdFlux", "dFlux", "sFlux", "rsFlux", "q", "dphi", "rdisc", "ulimb", "rwd", "scale", "az", "fis",
             "dexp", "phi0", "exp1", "exp2", "tilt", "yaw"]
    return [p[k] for k in order]


class PyWhiteDwarf(_Component):
    """Limb-darkened white dwarf of radius r_wd (units of xl1), unit flux out of eclipse."""
    _skip = _cabi.FLAG_SKIP_DISC | _cabi.FLAG_SKIP_BS | _cabi.FLAG_SKIP_DONOR

    def __init__(self, r_wd, ulimb):
        self.r_wd, self.ulimb = float(r_wd), float(ulimb)

    def calcFlux(self, q, inc, phi, width=None):
        return self._flux(_pars(wdFlux=1.0, q=q, dphi=inc, rwd=self.r_wd, ulimb=self.ulimb), phi, width)


class PyDisc(_Component):
    """Flat disc between r_in and r_out (units of xl1), brightness r^-exp, unit flux out of eclipse."""
    _skip = _cabi.FLAG_SKIP_WD | _cabi.FLAG_SKIP_BS | _cabi.FLAG_SKIP_DONOR

    def __init__(self, q, r_in, r_out, exp, nelem=_DEFAULT_DISC):
        self.q, self.r_in, self.r_out, self.exp = float(q), float(r_in), float(r_out), float(exp)
        self._grid = {} if int(nelem) == _DEFAULT_DISC else dict(n_disc_th=40, n_disc_r=max(1, int(round(nelem / 40.0))))

    def calcFlux(self, q, inc, phi, width=None):
        return self._flux(_pars(dFlux=1.0, q=q, dphi=inc, rwd=self.r_in, rdisc=self.r_out, dexp=self.exp), phi, width)


class PySpot(_Component):
    """Bright-spot strip at the stream impact on a disc of radius rdisc; simple (exp1 = 2,
    exp2 = 1, beamed along the strip normal) or complex profile."""
    _skip = _cabi.FLAG_SKIP_WD | _cabi.FLAG_SKIP_DISC | _cabi.FLAG_SKIP_DONOR

    def __init__(self, q, rdisc, az, frac, scale, exp1=2.0, exp2=1.0, tilt=90.0, yaw=0.0, complex=False):
        self.args = dict(rdisc=float(rdisc), az=float(az), fis=float(frac), scale=float(scale))
        if complex:
            self.args.update(exp1=float(exp1), exp2=float(exp2), tilt=float(tilt), yaw=float(yaw))

    def calcFlux(self, q, inc, phi, width=None):
        return self._flux(_pars(sFlux=1.0, q=q, dphi=inc, **self.args), phi, width)


class PyDonor(_Component):
    """Roche-lobe-filling donor: ellipsoidal modulation, unit flux at quadrature."""
    _skip = _cabi.FLAG_SKIP_WD | _cabi.FLAG_SKIP_DISC | _cabi.FLAG_SKIP_BS

    def __init__(self, q, nelem=_DEFAULT_DONOR):
        self.q = float(q)
        self._grid = {} if int(nelem) == _DEFAULT_DONOR else dict(n_donor_th=max(4, int(round(np.sqrt(np.pi * nelem / 4.0)))))

    def calcFlux(self, q, inc, phi, width=None):
        return self._flux(_pars(rsFlux=1.0, q=q, dphi=inc), phi, width)
