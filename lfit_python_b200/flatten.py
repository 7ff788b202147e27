"""Flatten a model tree once into the arrays the CUDA engine evaluates for every walker.

Replaces, on the sampling path, the per-walker Python of the reference: the recursive
scatter of the parameter vector into Param.currVal (model.py:586-603), the gather of each
leaf's CV parameter list from its ancestors (CVModel.py:335-354, model.py:706-712) and the
recursive ln_prior / ln_like sums (model.py:382-498).  The tree is read, never mutated.
"""
import numpy as np

from . import _cabi

NPAR = _cabi.NPAR


def leaves(node):
    """Leaf nodes in depth-first order (the order `__call_recursive_func__` visits them)."""
    if node.is_leaf:
        return [node]
    out = []
    for child in node.children:
        out.extend(leaves(child))
    return out


class FlatLayout:
    """Index maps and tables of a tree: everything the C ABI's set_layout / set_priors /
    set_lightcurves take.  Pure host data; no GPU needed to build it."""

    def __init__(self, model):
        params, labels = model.__get_descendant_params__()
        self.names = [p.name + "_" + l for p, l in zip(params, labels) if p.isVar]
        self.ndim = len(self.names)
        src_of = {}
        consts = []
        col = 0
        for p in params:
            if p.isVar:
                src_of[id(p)] = col
                col += 1
            else:
                consts.append(float(p.currVal))
                src_of[id(p)] = -len(consts)
        self.consts = np.asarray(consts, dtype=np.float64)
        self.p0 = np.asarray([p.currVal for p in params if p.isVar], dtype=np.float64)
        self.prior_src = np.asarray([src_of[id(p)] for p in params], dtype=np.int32)
        self.prior_type = np.asarray([_cabi.PRIOR_CODES[p.prior.type] for p in params], dtype=np.int32)
        self.prior_p1 = np.asarray([p.prior.p1 for p in params], dtype=np.float64)
        self.prior_p2 = np.asarray([p.prior.p2 for p in params], dtype=np.float64)
        self.prior_norm = np.asarray([getattr(p.prior, "normalise", 1.0) for p in params], dtype=np.float64)
        self.prior_isvar = np.asarray([int(bool(p.isVar)) for p in params], dtype=np.int32)
        self.eclipses = leaves(model)
        if not self.eclipses or not hasattr(self.eclipses[0], "cv_parnames"):
            raise TypeError("the tree has no eclipse leaves")
        npars = {len(e.cv_parnames) for e in self.eclipses}
        if len(npars) != 1:
            raise ValueError("eclipses mix simple and complex bright-spot models")
        self.npars = npars.pop()
        gather = np.zeros((len(self.eclipses), NPAR), dtype=np.int32)
        for k, ecl in enumerate(self.eclipses):
            pd = ecl.ancestor_param_dict
            for j, nm in enumerate(ecl.cv_parnames):
                gather[k, j] = src_of[id(pd[nm])]
        self.gather = gather
        self.n_ecl = len(self.eclipses)
        # Gaussian-process likelihood (GPLCModel, CVModel.py:494-711): where the three hyper-parameters
        # come from, and every eclipse's change-point distance at the tree's current values (the
        # reference computes it on the first evaluation and keeps it, CVModel.py:548-579)
        self.gp = all(hasattr(e, "calcChangepoints") for e in self.eclipses)
        if self.gp:
            pd = self.eclipses[0].ancestor_param_dict
            self.gp_src = np.asarray([src_of[id(pd[k])] for k in ('ln_ampin_gp', 'ln_ampout_gp', 'ln_tau_gp')],
                                     dtype=np.int32)
            self.gp_dist = np.asarray([e.dist_cp() for e in self.eclipses], dtype=np.float64)
        elif any(hasattr(e, "calcChangepoints") for e in self.eclipses):
            raise ValueError("the tree mixes GP and chi-squared eclipses")
        n = [e.lc.n_data for e in self.eclipses]
        self.lc_off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        cat = lambda attr: np.concatenate([np.asarray(getattr(e.lc, attr), dtype=np.float64) for e in self.eclipses])
        self.lc_phase, self.lc_width, self.lc_y, self.lc_ye = cat("x"), cat("w"), cat("y"), cat("ye")

    def apply(self, engine):
        engine.set_layout(self.ndim, self.npars, self.gather, self.consts)
        engine.set_priors(self.prior_src, self.prior_type, self.prior_p1, self.prior_p2, self.prior_norm,
                          self.prior_isvar)
        engine.set_lightcurves(self.lc_off, self.lc_phase, self.lc_width, self.lc_y, self.lc_ye)
        if self.gp:
            engine.set_gp(self.gp_src, self.gp_dist)
        else:
            engine.set_gp()


class VectorModel:
    """Batched log-probability of a model tree on one GPU: `ln_prob(theta)` with theta of shape
    (n, ndim) returns n values, usable as emcee's `vectorize=True` log-probability."""

    def __init__(self, model, engine=None, device=0, **grid):
        self.model = model
        self.layout = FlatLayout(model)
        self.engine = engine if engine is not None else _cabi.Engine(device, **grid)
        self.layout.apply(self.engine)
        self.ndim = self.layout.ndim
        self.names = self.layout.names

    def _eval(self, theta, what, return_chisq=False):
        theta = np.asarray(theta, dtype=np.float64)
        single = theta.ndim == 1
        theta = np.atleast_2d(theta)
        if theta.shape[1] != self.ndim:
            raise ValueError('Wrong vector length on {} - Expected {}, got {}'.format(
                self.model.name, self.ndim, theta.shape[1]))
        out = self.engine.log_prob(theta, what=what, return_chisq=return_chisq)
        if return_chisq:
            return (out[0][0], out[1][0]) if single else out
        return out[0] if single else out

    def ln_prior(self, theta):
        return self._eval(theta, _cabi.LN_PRIOR)

    def ln_like(self, theta):
        return self._eval(theta, _cabi.LN_LIKE)

    def ln_prob(self, theta):
        return self._eval(theta, _cabi.LN_PROB)

    __call__ = ln_prob

    def ln_like_and_prob(self, theta):
        """(ln_like, ln_prob) of every row from ONE pass -- what the reference hands ptemcee as (logl, logp)
        (mcmcfit.py:264-270 passes ln_like and ln_prob).  Rows the priors reject get ln_like = 0 (they are not
        evaluated, as in ptemcee's evaluator)."""
        lnp, chi = self._eval(np.atleast_2d(theta), _cabi.LN_PROB, return_chisq=True)
        skipped = np.isnan(chi).any(axis=1)
        like = np.where(skipped, 0.0, -0.5 * np.where(np.isnan(chi), 0.0, chi).sum(axis=1))
        return like, lnp

    def chisq(self, theta):
        """Per-eclipse chi-squared, shape (n, n_ecl) (SimpleEclipse.chisq for every leaf); for a GP
        tree, -2 ln L of every leaf."""
        return self._eval(theta, _cabi.LN_LIKE, return_chisq=True)[1]
