"""lfit_python_b200 -- the LFIT CV eclipse model and its log-probability on B200 (sm_100a).

Hot path only: lfit.CV.calcFlux, chi-squared, priors and the tree's
ln_prob for every emcee walker, behind the reference's own interfaces:

    lfit_python_b200.lfit      CV, PyWhiteDwarf, PyDisc, PySpot, PyDonor   (was: import lfit)
    lfit_python_b200.roche     xl1, findi, findphi, bspot                  (was: from trm import roche)
    lfit_python_b200.model     Prior, Param, Node
    lfit_python_b200.CVModel   Lightcurve, Simple/ComplexEclipse, Band, LCModel, construct_model
    lfit_python_b200.mcmcfit   ln_prior / ln_like / ln_prob(param_vector, model), 1-D or (n, ndim)
    lfit_python_b200.flatten   FlatLayout, VectorModel (the tree flattened once)
    lfit_python_b200._cabi     ctypes binding of include/lfit_b200.h

Importing the package needs no GPU; any evaluation does (there is no CPU fallback).
"""
from . import _cabi  # noqa: F401

__all__ = ["_cabi", "lfit", "roche", "model", "CVModel", "mcmcfit", "mcmc_utils", "flatten", "parallel", "workloads"]
