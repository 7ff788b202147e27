"""The fit driver's log-probability wrappers, vectorised (mcmcfit.py:30-48 of the reference).

`ln_prob(param_vector, model)` keeps the reference signature.  A 1-D vector keeps the scalar
semantics (set the vector on the tree, call model.ln_prob()); an (n, ndim) matrix is evaluated
for all rows in one CUDA call, which is what `emcee.EnsembleSampler(..., vectorize=True)`
passes.  See `main()` for the driver (python -m lfit_python_b200.mcmcfit input.dat).
"""
import argparse
import sys

import numpy as np

from . import mcmc_utils as utils
from .CVModel import construct_model, extract_par_and_key
from .configobj import ConfigObj


def _vector(model):
    vec = getattr(model, "_vector_model", None)
    if vec is None:
        vec = model.vectorised()
        model._vector_model = vec
    return vec


def _dispatch(name, param_vector, model):
    param_vector = np.asarray(param_vector, dtype=np.float64)
    if param_vector.ndim == 2:
        return getattr(_vector(model), name)(param_vector)
    model.dynasty_par_vals = param_vector
    return getattr(model, name)()


def ln_prior(param_vector, model):
    return _dispatch("ln_prior", param_vector, model)


def ln_prob(param_vector, model):
    return _dispatch("ln_prob", param_vector, model)


def ln_like(param_vector, model):
    return _dispatch("ln_like", param_vector, model)


# per-parameter multipliers of the walker scatter (mcmcfit.py:208-246)
SCATTER_FRACTION = {'q': 1, 'rwd': 1, 'dphi': 0.2, 'dFlux': 1, 'sFlux': 1, 'wdFlux': 1, 'rsFlux': 1, 'rdisc': 1,
                    'ulimb': 1e-6, 'scale': 1, 'fis': 1, 'dexp': 1, 'phi0': 1, 'az': 1, 'exp1': 1, 'exp2': 1,
                    'yaw': 1, 'tilt': 1}


def scatter_vector(model, scatter, comp_scat=True):
    out = np.full(len(model.dynasty_par_names), float(scatter))
    if comp_scat:
        for i, name in enumerate(model.dynasty_par_names):
            key, _ = extract_par_and_key(name)
            if key.startswith('ln'):
                continue
            out[i] *= SCATTER_FRACTION[key]
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="Execute an MCMC fit to a dataset (CUDA log-probability).")
    ap.add_argument("input", help="The filename for the MCMC parameters' input file.")
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--quiet", action="store_true", help="accepted for compatibility; nothing is plotted")
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args(argv)

    model = construct_model(args.input, args.debug)
    cfg = ConfigObj(args.input)
    nburn, nprod, nwalkers = int(cfg['nburn']), int(cfg['nprod']), int(cfg['nwalkers'])
    scatter_1, scatter_2 = float(cfg['first_scatter']), float(cfg['second_scatter'])
    to_fit = int(cfg['fit'])
    double_burnin = bool(int(cfg['double_burnin']))
    comp_scat = bool(int(cfg['comp_scat']))
    use_pt = bool(int(cfg.get('usePT', 0)))
    ntemps = int(cfg.get('ntemps', 1))

    eclipses = model.search_node_type('Eclipse')
    dof = int(np.sum([e.lc.n_data for e in eclipses]) - len(model.dynasty_par_names) - 1)
    pars = np.asarray(model.dynasty_par_vals)
    vec = _vector(model)
    chisq0 = float(np.sum(vec.chisq(pars)))
    print("\n\nInitial guess has a chisq of {:.3f} ({:d} D.o.F.).".format(chisq0, dof))
    print("a ln_prior of {:.3f}".format(float(vec.ln_prior(pars))))
    print("a ln_like of {:.3f}".format(float(vec.ln_like(pars))))
    print("a ln_prob of {:.3f}".format(float(vec.ln_prob(pars))))
    if np.isinf(vec.ln_prior(pars)):
        print("ERROR: Starting position violates priors!")
        model.ln_prior(verbose=True)
        sys.exit(1)
    if not to_fit:
        return model

    npars = len(pars)
    print("\n\nThe MCMC has {:d} variables and {:d} walkers".format(npars, nwalkers))
    print("(It should have at least 2*npars, {:d} walkers)".format(2 * npars))
    if nwalkers < 2 * npars:
        sys.exit(1)
    rng = np.random.default_rng(args.seed)
    s1 = scatter_vector(model, scatter_1, comp_scat)
    col_names = "walker_no " + ' '.join(model.dynasty_par_names) + ' ln_prob'
    if use_pt:
        # parallel tempering (mcmcfit.py:251-270): ntemps x nwalkers rows per half-step are one CUDA pass.  As in
        # the reference, ptemcee's "likelihood" is ln_like and its "prior" is ln_prob (sic).
        print("MCMC using parallel tempering at {} levels, for {} total walkers.".format(ntemps, nwalkers * ntemps))
        p0 = utils.initialise_walkers_pt(pars, s1, nwalkers, ntemps, ln_prior, model, rng=rng)
        sampler = utils.PTSampler(nwalkers, npars, ln_like, ln_prob, loglargs=(model,), logpargs=(model,), ntemps=ntemps,
                                  vectorize=True, joint=vec.ln_like_and_prob, rng=rng)
        print("\n\nExecuting the burn-in phase...")
        pos, prob, state = utils.run_burnin(sampler, p0, nburn)
        if double_burnin:
            print("Executing the second burn-in phase")
            p0 = utils.initialise_walkers_pt(pos[0][np.argmax(prob[0])], s1 * (scatter_2 / scatter_1), nwalkers, ntemps,
                                             ln_prior, model, rng=rng)
            pos, prob, state = utils.run_burnin(sampler, p0, nburn)
        sampler.reset()
        print("Starting the main MCMC chain.")
        utils.run_ptmcmc_save(sampler, pos, nprod, "chain_prod.txt", col_names=col_names)
        return sampler
    p0 = utils.initialise_walkers(pars, s1, nwalkers, ln_prior, model, rng=rng)
    sampler = utils.EnsembleSampler(nwalkers, npars, ln_prob, args=(model,), vectorize=True, rng=rng)
    print("\n\nExecuting the burn-in phase...")
    pos, prob, state = utils.run_burnin(sampler, p0, nburn)
    if double_burnin:
        print("Executing the second burn-in phase")
        p0 = utils.initialise_walkers(pos[np.argmax(prob)], s1 * (scatter_2 / scatter_1), nwalkers, ln_prior, model,
                                      rng=rng)
        pos, prob, state = utils.run_burnin(sampler, p0, nburn)
    sampler.reset()
    print("Starting the main MCMC chain.")
    utils.run_mcmc_save(sampler, pos, nprod, state, "chain_prod.txt", col_names=col_names)
    return sampler


if __name__ == "__main__":
    main()
