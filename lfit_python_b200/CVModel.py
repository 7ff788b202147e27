"""CV-specific tree nodes: Lightcurve, Simple/ComplexEclipse, Band, LCModel, construct_model.

Host-side mirror of /root/reference/CVModel.py.  Each eclipse leaf evaluates
lfit.CV.calcFlux on the GPU (lfit_python_b200.lfit); the whole tree can be flattened with
`model.vectorised()` (flatten.py) into one batched CUDA log-probability for every walker.
The Gaussian-process variants (CVModel.py:494-711) evaluate george's likelihood exactly on the GPU
(csrc/gp_device.cuh) instead of through george's approximate HODLR solver.
"""
import os

import numpy as np

from . import lfit, roche
from .configobj import ConfigObj
from .model import Node, Param, extract_par_and_key


class Lightcurve:
    """Observed data of one eclipse: phase x, flux y, error ye, exposure half-width w."""

    def __init__(self, name, x, y, ye, w=None):
        self.name = name
        self.fname = None
        if w is None:
            w = np.mean(np.diff(x)) * np.ones_like(x) / 2.
        self.x, self.y, self.ye, self.w = x, y, ye, w

    @property
    def n_data(self):
        return self.x.shape[0]

    @classmethod
    def from_calib(cls, fname, name=None):
        """Read `phase flux error` columns ('#' comments; space, comma or pipe separated),
        drop NaN fluxes, width = half the mean phase step of the untrimmed curve
        (CVModel.py:41-72)."""
        data = None
        for delimiter in (None, ',', '|'):
            try:
                data = np.loadtxt(fname, delimiter=delimiter, comments='#')
                break
            except ValueError:
                continue
        if data is None:
            raise ValueError("could not parse calib file {}".format(fname))
        phase, flux, error = data.T[:3]
        keep = ~np.isnan(flux)
        phase, flux, error = phase[keep], flux[keep], error[keep]
        width = np.mean(np.diff(phase)) * np.ones_like(phase) / 2.
        lc = cls(os.path.split(fname)[1] if name is None else name, phase, flux, error, width)
        lc.fname = fname
        return lc

    def trim(self, lo, hi):
        """Keep lo < x < hi."""
        m = (self.x > lo) & (self.x < hi)
        self.x, self.y, self.ye, self.w = self.x[m], self.y[m], self.ye[m], self.w[m]


class SimpleEclipse(Node):
    """Leaf holding one eclipse, simple bright-spot model (CVModel.py:87-354)."""

    node_par_names = ('dFlux', 'sFlux', 'rdisc', 'scale', 'az', 'fis', 'dexp', 'phi0')

    def __init__(self, lightcurve, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if isinstance(lightcurve, Lightcurve):
            self.lc = lightcurve
        elif isinstance(lightcurve, str):
            self.lc = Lightcurve.from_calib(lightcurve)
        else:
            raise TypeError("Argument lightcurve is not a string or Lightcurve! Got {}".format(lightcurve))
        self.cv = lfit.CV(self.cv_parlist)

    @property
    def cv_parnames(self):
        return ['wdFlux', 'dFlux', 'sFlux', 'rsFlux', 'q', 'dphi', 'rdisc', 'ulimb', 'rwd', 'scale', 'az', 'fis',
                'dexp', 'phi0']

    @property
    def cv_parlist(self):
        """The CV parameter list gathered from this leaf, its band and the root."""
        pd = self.ancestor_param_dict
        return [pd[key].currVal for key in self.cv_parnames]

    def calcFlux(self):
        """Model flux at the data phases; NaN if the parameters admit no model (CVModel.py:132-147)."""
        try:
            return self.cv.calcFlux(self.cv_parlist, self.lc.x, self.lc.w)
        except Exception as e:
            self.log("SimpleEclipse.calcFlux", "Error: {}; parlist: {}".format(e, self.cv_parlist))
            return np.nan

    def calcComponents(self):
        """(total, wdFlux, sFlux, rsFlux, dFlux) curves (CVModel.py:149-155)."""
        flx = self.cv.calcFlux(self.cv_parlist, self.lc.x, self.lc.w)
        return flx, self.cv.ywd, self.cv.ys, self.cv.yrs, self.cv.yd

    def chisq(self):
        flx = self.calcFlux()
        if np.any(np.isnan(flx)):
            return np.inf
        return np.sum(((self.lc.y - flx) / self.lc.ye) ** 2)

    def ln_like(self):
        return -0.5 * self.chisq()

    def ln_prior(self, verbose=False, *args, **kwargs):
        """Physical validity of this eclipse, then the Param priors (CVModel.py:193-324):
        disc below the 3:1 resonance (rdisc * xl1 <= 0.46), spot scale within a factor 3 of the
        white-dwarf radius, stream reaching the disc, strip azimuth within 80 deg of the disc
        tangent at the impact point (clamped to [0, 178])."""
        pd = self.ancestor_param_dict
        q = pd['q'].currVal
        try:
            xl1 = roche.xl1(q)
        except AssertionError:
            if verbose:
                print("Failed to get the L1 point!")
            return -np.inf
        rdisc_a = pd['rdisc'].currVal * xl1
        if verbose:
            print("rDisc: {:.4f} || Max: {:.4f}".format(rdisc_a, 0.46))
        if rdisc_a > 0.46:
            return -np.inf
        rwd, scale = pd['rwd'].currVal, pd['scale'].currVal
        if verbose:
            print("Scale: {:.4f} || Limits: {:.4f} -> {:.4f}".format(scale, rwd / 3., rwd * 3.))
        if scale > rwd * 3. or scale < rwd / 3.:
            return -np.inf
        try:
            az = pd['az'].currVal
            x, y, _, _ = roche.bspot(q, rdisc_a)
            alpha = np.degrees(np.arctan2(y, x))
            if alpha < 0:
                alpha = 90 - alpha
            tangent = alpha + 90
            if az < max(0, tangent - 80.0) or az > min(178, tangent + 80.0):
                return -np.inf
        except Exception as err:
            if verbose:
                print(err)
                print("The mass stream of leaf {} does not intersect the disc!".format(self.name))
            return -np.inf
        return super().ln_prior(verbose=verbose, *args, **kwargs)


class ComplexEclipse(SimpleEclipse):
    """Leaf using the complex bright-spot model.  Node order ends (..., yaw, tilt) while the CV
    order ends (..., tilt, yaw) (CVModel.py:376-390)."""

    node_par_names = ('dFlux', 'sFlux', 'rdisc', 'scale', 'az', 'fis', 'dexp', 'phi0', 'exp1', 'exp2', 'yaw', 'tilt')

    @property
    def cv_parnames(self):
        return ['wdFlux', 'dFlux', 'sFlux', 'rsFlux', 'q', 'dphi', 'rdisc', 'ulimb', 'rwd', 'scale', 'az', 'fis',
                'dexp', 'phi0', 'exp1', 'exp2', 'tilt', 'yaw']


class Band(Node):
    """An observing band: white-dwarf and donor fluxes and limb darkening shared by its eclipses."""

    node_par_names = ('wdFlux', 'rsFlux', 'ulimb')

    @property
    def eclipses(self):
        return list(self.search_node_type("Eclipse"))


class LCModel(Node):
    """Root of the tree: q, dphi, rwd (CVModel.py:418-491)."""

    node_par_names = ('q', 'dphi', 'rwd')

    @property
    def eclipses(self):
        return list(self.search_node_type("Eclipse"))

    def ln_prior(self, verbose=False):
        """dphi must stay 1e-6 below the eclipse width at i = 90 deg for this q, else -inf."""
        dphi, q = self.dphi.currVal, self.q.currVal
        try:
            maxphi = roche.findphi(q, 90.0)
            if dphi > maxphi - 1e-6:
                if verbose:
                    print("{} has a dphi out of tolerance!\nq: {:.3f}\ndphi: {:.3f}, max: {:.3f} - {:.3g}".format(
                        self.name, q, dphi, maxphi, 1e-6))
                return -np.inf
        except Exception as error:
            if verbose:
                print("Failed to calculate a value of dphi at node {} || Exception: {}".format(self.name, repr(error)))
            return -np.inf
        return super().ln_prior(verbose=verbose)

    def vectorised(self, engine=None, device=0, **grid):
        """Flatten the tree once into the batched CUDA log-probability (flatten.VectorModel)."""
        from .flatten import VectorModel
        return VectorModel(self, engine=engine, device=device, **grid)


class GPLCModel(LCModel):
    """Root of a tree whose eclipses are judged by a Gaussian process over their residuals
    (CVModel.py:494-514): adds the ln of the in-eclipse amplitude, of the extra out-of-eclipse
    amplitude and of the time scale."""

    node_par_names = LCModel.node_par_names + ('ln_ampin_gp', 'ln_ampout_gp', 'ln_tau_gp')


class SimpleGPEclipse(SimpleEclipse):
    """SimpleEclipse whose ln_like is the GP likelihood of its residuals (CVModel.py:517-696)."""

    # unrealistic start values: the first calcChangepoints computes dist_cp (CVModel.py:518-526)
    _olddphi = 9e99
    _oldq = 9e99
    _oldrwd = 9e99
    _dist_cp = 9e99

    def dist_cp(self):
        """Distance of the change points from mid-eclipse, (dphi + dpwd) / 2, with the reference's
        cache: recomputed only when dphi, q or rwd moved by more than 120 % of their value since
        the last computation (CVModel.py:548-579) -- in practice once."""
        pd = self.ancestor_param_dict
        dphi, q, rwd = pd['dphi'], pd['q'], pd['rwd']
        dphi_change = np.fabs(self._olddphi - dphi.currVal) / dphi.currVal
        q_change = np.fabs(self._oldq - q.currVal) / q.currVal
        rwd_change = np.fabs(self._oldrwd - rwd.currVal) / rwd.currVal
        if (dphi_change > 1.2) or (q_change > 1.2) or (rwd_change > 1.2):
            inc = roche.findi(q.currVal, dphi.currVal)
            phi3, phi4 = roche.wdphases(q.currVal, inc, rwd.currVal, ntheta=10)
            self._dist_cp = (dphi.currVal + (phi4 - phi3)) / 2.
            self._oldq, self._olddphi, self._oldrwd = q.currVal, dphi.currVal, rwd.currVal
        return self._dist_cp

    def calcChangepoints(self):
        """[[egress of the previous eclipse, ingress of this one], ...] for every cycle the light
        curve touches (CVModel.py:527-601)."""
        dist_cp = self.dist_cp()
        phi0 = self.ancestor_param_dict['phi0'].currVal
        x = self.lc.x
        eclipses = [e for e in range(int(np.floor(x.min())), int(np.ceil(x.max())) + 1)
                    if e > x.min() and e < 1 + x.max()]
        return [[(e - 1) + dist_cp + phi0, e - dist_cp + phi0] for e in eclipses]

    def ln_like(self):
        """ln L of the residuals under ampin * Matern32(tau) + ampout * Matern32(tau) between the
        change points, with the data errors on the diagonal (CVModel.py:603-696)."""
        flx = self.calcFlux()
        residuals = self.lc.y - flx
        if np.any(np.isinf(residuals)) or np.any(np.isnan(residuals)):
            return -np.inf
        pd = self.ancestor_param_dict
        hyper = [np.exp(pd[k].currVal) for k in ('ln_ampin_gp', 'ln_ampout_gp', 'ln_tau_gp')]
        gaps = self.calcChangepoints()
        order = np.argsort(self.lc.x, kind='stable')
        from . import _cabi
        out = _cabi.default_engine().gp_loglike(self.lc.x[order], self.lc.ye[order], residuals[order], hyper, gaps)
        return float(out[0])


class ComplexGPEclipse(SimpleGPEclipse):
    """As SimpleGPEclipse with the complex bright spot (CVModel.py:699-711)."""

    node_par_names = ComplexEclipse.node_par_names
    cv_parnames = ComplexEclipse.cv_parnames


def construct_model(input_file, debug=False, nodata=False):
    """Parse an mcmc_input.dat into a model tree (CVModel.py:713-924): LCModel 'core' -> one Band
    per band label -> one Simple/Complex(GP)Eclipse per eclipse label (GPLCModel root and GP leaves for
    useGP = 1), in order of first appearance
    in the file; `neclipses` truncates; bands left without eclipses are pruned."""
    cfg = ConfigObj(input_file)
    is_complex = bool(int(cfg['complex']))
    use_gp = bool(int(cfg.get('useGP', 0)))
    neclipses = int(cfg['neclipses']) if 'neclipses' in cfg else 9999

    root_cls = GPLCModel if use_gp else LCModel
    model = root_cls('core', [Param.fromString(n, cfg[n]) for n in root_cls.node_par_names], DEBUG=debug)
    if use_gp:
        ecl_cls = ComplexGPEclipse if is_complex else SimpleGPEclipse
    else:
        ecl_cls = ComplexEclipse if is_complex else SimpleEclipse
    ecl_pars = ecl_cls.node_par_names

    bands, eclipses = [], []
    with open(input_file, 'r') as f:
        for line in f:
            words = line.strip().split()
            if not words:
                continue
            key = words[0]
            if any(key.startswith(p) for p in Band.node_par_names):
                label = extract_par_and_key(key)[1]
                if label not in bands:
                    bands.append(label)
            if any(key.startswith(p) for p in ecl_pars):
                label = extract_par_and_key(key)[1]
                if label not in eclipses:
                    eclipses.append(label)

    for label in bands:
        Band(label, [Param.fromString(p, cfg["{}_{}".format(p, label)]) for p in Band.node_par_names], parent=model)

    lo, hi = float(cfg['phi_start']), float(cfg['phi_end'])
    base = os.path.dirname(os.path.abspath(input_file))
    for label in eclipses[:neclipses]:
        params = [Param.fromString(p, cfg["{}_{}".format(p, label)]) for p in ecl_pars]
        if nodata:
            x = np.linspace(-0.5, 0.5, 1000)
            lc = Lightcurve("Dummy_Data_{}".format(label), x, np.zeros_like(x), np.ones_like(x))
        else:
            fname = cfg['file_{}'.format(label)]
            if not os.path.isabs(fname) and not os.path.exists(fname):
                fname = os.path.join(base, fname)
            lc = Lightcurve.from_calib(fname)
            lc.trim(lo, hi)
        band = model.search_Node('Band', cfg['band_{}'.format(label)])
        ecl_cls(lc, label, params, parent=band)

    model.children = [b for b in model.children if len(b.children)]
    return model
