"""End-to-end inference on the GPU with this package's own parts: a compact stand-in for the
reference's `simulate.simulate()` driver (simulate.py:17-198) for users without bilby/emcee.

    python -m sterne_b200.infer 57444 psr.inits psr.pmpar.in psr.par [more.pmpar.in more.par ...] \\
           --shares '[[0],[-1],...]' --nwalkers 4096 --iterations 2000 --outdir outdir

Reads the reference's input files, builds the priors of the `.inits` file (Uniform / Gaussian /
Sine, sin(incl) limits; the sin(incl) Gaussian terms go into the likelihood as in
simulate.py:379-383), runs the fused stretch-move sampler (three launches per half-step), and
writes `posterior_samples.dat` (parameter columns + log_likelihood + log_prior, the layout the
reference's post-processing reads, simulate.py:253-254) and `bayesian_estimates.dat`.
The sampler differs from the reference's (bilby's emcee wrapper with 30 walkers): same
algorithm family, far larger ensembles, different random numbers.
"""
import argparse
import ast
import os

import numpy as np

from . import readers, summary
from .likelihood import Gaussianlikelihood
from .sampler import BoxPrior, FusedEnsembleSampler
from .shares import get_parameters_from_shares


def split_args(pmparin, parfile, *args):
    """simulate.py:107-126: positional pmpar.in / .par (or '') pairs."""
    pmparins, parfiles = [], []
    for arg in (pmparin, parfile) + tuple(args):
        if ('.pmpar.in' in arg) and arg.endswith('.par'):
            raise ValueError('arg does not follow naming convention: %r' % arg)
        if '.pmpar.in' in arg:
            pmparins.append(arg)
        elif arg.endswith('.par') or arg == '':
            parfiles.append(arg)
    if len(pmparins) != len(parfiles):
        raise ValueError('Unequal number of parfiles provided for pmparins.')
    return pmparins, parfiles


def default_shares(n, efac):
    """simulate.py:134-135,169-170"""
    s = [list(range(n)), [0] * n, [0] * n, [0] * n, [0] * n, [0] * n, [0] * n, [0] * n, list(range(n))]
    if not efac:
        s[1], s[2] = [-1] * n, [-1] * n
    return s


def run(refepoch, initsfile, pmparin, parfile, *args, shares=None, pmparin_preliminaries=None,
        a1dot_constraints=False, nwalkers=1024, iterations=1000, burn=None, thin=1, outdir='outdir',
        earth_provider=None, earth_xyz=None, device=0, seed=0, init_scale=1e-3):
    """device: a CUDA ordinal, or a list of them -- the ensemble is then shared among those GPUs of this
    box from this one process (FusedEnsembleSampler / stn_multi_ensemble_run), same chain as on one GPU."""
    if not os.path.exists(initsfile):
        raise FileNotFoundError('%s does not exist' % initsfile)
    pmparins, parfiles = split_args(pmparin, parfile, *args)
    n = len(pmparins)
    if shares is None:
        shares = default_shares(n, pmparin_preliminaries is not None)
    elif pmparin_preliminaries is None:
        shares = [list(r) for r in shares]
        shares[1], shares[2] = [-1] * n, [-1] * n                     # simulate.py:169-170
    LoD_timing = readers.create_list_of_dict_timing(parfiles)
    LoD_VLBI = readers.create_list_of_dict_VLBI(pmparins, pmparin_preliminaries)
    limits, constraints = readers.read_inits(initsfile)
    names = list(get_parameters_from_shares(shares))
    missing = [k for k in names if k not in limits]
    if missing:
        raise KeyError('the .inits file has no prior for %r' % missing)
    keys = [k for k in limits if k in names]                           # prior (file) order, as bilby uses
    prior = BoxPrior({k: limits[k] for k in keys}, constraints, keys=keys)
    like = Gaussianlikelihood(refepoch, LoD_timing, LoD_VLBI, shares, None, constraints, a1dot_constraints,
                              earth_provider=earth_provider, earth_xyz=earth_xyz, device=device)
    # walkers start in a small ball around the prior centre (Uniform / Sine) or mean (Gaussian)
    rng = np.random.default_rng(seed)
    centre = np.where(np.array(prior.kinds) == 'Gaussian', prior.a, 0.5 * (prior.a + prior.b))
    width = np.where(np.array(prior.kinds) == 'Gaussian', prior.b, prior.b - prior.a)
    p0 = centre + init_scale * width * rng.standard_normal((nwalkers, prior.ndim))
    sampler = FusedEnsembleSampler(like, prior, nwalkers, seed=seed)
    burn = iterations // 2 if burn is None else burn
    sampler.run(p0, burn, store=False)
    sampler.run(sampler.last[0], iterations - burn, thin=thin)
    chain = sampler.chain.reshape(-1, prior.ndim).cpu().numpy()
    lnprob = sampler.lnprob.reshape(-1).cpu().numpy()
    lnprior = prior.log_prob(chain)
    os.makedirs(outdir, exist_ok=True)
    samplefile = os.path.join(outdir, 'posterior_samples.dat')
    with open(samplefile, 'w') as f:
        f.write(' '.join(keys + ['log_likelihood', 'log_prior']) + '\n')
        np.savetxt(f, np.column_stack([chain, lnprob - lnprior, lnprior]), fmt='%.17g')
    med, chi_sq, rchsq, estimates = summary.make_a_summary_of_bayesian_inference(
        samplefile, like, samples=sampler.chain.reshape(-1, prior.ndim), names=keys)        # summarised on the device
    result = dict(keys=keys, samplefile=samplefile, estimates=estimates, medians=med, chi_sq=chi_sq, rchsq=rchsq,
                  acceptance_fraction=sampler.acceptance_fraction, n_samples=len(chain))
    like.close()
    return result


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('refepoch', type=float)
    ap.add_argument('initsfile')
    ap.add_argument('files', nargs='+', help="pmpar.in / .par pairs ('' for no parfile)")
    ap.add_argument('--shares', type=ast.literal_eval, default=None)
    ap.add_argument('--preliminaries', nargs='*', default=None)
    ap.add_argument('--a1dot-constraints', type=ast.literal_eval, default=False)
    ap.add_argument('--nwalkers', type=int, default=1024)
    ap.add_argument('--iterations', type=int, default=1000)
    ap.add_argument('--outdir', default='outdir')
    ap.add_argument('--device', type=int, nargs='+', default=[0], help='CUDA ordinal(s); several share the ensemble')
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args(argv)
    res = run(a.refepoch, a.initsfile, *a.files, shares=a.shares, pmparin_preliminaries=a.preliminaries,
              a1dot_constraints=a.a1dot_constraints, nwalkers=a.nwalkers, iterations=a.iterations, outdir=a.outdir,
              device=a.device if len(a.device) > 1 else a.device[0], seed=a.seed)
    print(open(res['estimates']).read())
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
