"""Posterior summaries: medians, 1-sigma ranges, the periodic estimate
for om_asc, correlation coefficients and the reduced chi-square at the median -- the content
of the reference's `bayesian_estimates.dat` (simulate.py:244-301, others.py:227-379), with
the chi-square evaluated by the device kernel (stn_chisq) instead of the Python model loop.

The text written is format-compatible with the reference's, so downstream scripts that
parse `bayesian_estimates.dat` keep working.
"""
import math
import numpy as np


def deg_float2dms(degree):
    """dd:mm:ss.sssssss (others.py:238-248)"""
    sign = np.sign(degree)
    degree = float(abs(degree))
    d = math.floor(degree)
    m = math.floor((degree - d) * 60)
    s = ((degree - d) * 60 - m) * 60
    dms = '%02d:%02d:%010.7f' % (d, m, s)
    return '-' + dms if sign == -1 else dms


def _significant_volume(n, confidencelevel):
    CL = confidencelevel
    if CL < 1:
        return int(CL * n)
    if CL < 10:
        return int(math.erf(CL / 2 ** 0.5) * n)
    return int(CL)


def sample2median(samples):
    """others.py:355-362"""
    a = np.sort(np.asarray(samples, dtype=np.float64))
    n = len(a)
    return 0.5 * (a[n // 2 - 1] + a[n // 2]) if n % 2 == 0 else a[(n - 1) // 2]


def sample2median_range(samples, confidencelevel):
    """Central interval holding erf(CL/sqrt 2) of the samples (others.py:363-374)."""
    a = np.sort(np.asarray(samples, dtype=np.float64))
    SV = _significant_volume(len(a), confidencelevel)
    start = int((len(a) - SV) / 2 - 1)
    return a[start], a[start + SV]


def _narrowest_interval(sorted_like, SV):
    """(value, error, median) of the narrowest interval spanning SV samples of an already
    ordered array (others.py:283-306): first minimum of a[i+SV]-a[i], i < len-SV-1."""
    a = sorted_like
    m = len(a) - SV - 1
    widths = a[SV:SV + m] - a[:m]
    j = int(np.argmin(widths))
    lo, hi = a[j], a[j + SV]
    return 0.5 * (lo + hi), 0.5 * (hi - lo), a[int(j + SV / 2.)]


def periodic_sample2estimate(samples, period=360, confidencelevel=1):
    """Estimate for a periodic parameter (om_asc): over 360 cut positions, re-order the sorted
    samples so the cut is at the head, take the narrowest 1-sigma interval, keep the cut with
    the smallest error (others.py:308-345).  Returns (median, upper error, lower error)."""
    a = np.sort(np.asarray(samples, dtype=np.float64)) % period
    SV = _significant_volume(len(a), confidencelevel)
    best = (None, float('inf'), None)
    for threshold in np.arange(0, period, period / 360.):
        k = int(np.abs(a - threshold).argmin())
        # the reference calls sorted() on the re-ordered list before scanning it
        reordered = np.sort(np.concatenate((a[k:] - period, a[:k])))
        value, error, median = _narrowest_interval(reordered, SV)
        if error < best[1]:
            best = (value, error, median)
    value, error, median = best
    return median % period, value + error - median, median - (value - error)


def read_posterior_samples(samplefile):
    """bilby's posterior_samples.dat: header line of names, whitespace-separated rows; the last
    two columns are log_likelihood and log_prior (simulate.py:253-254)."""
    with open(samplefile) as f:
        lines = [l for l in f.read().splitlines() if l.strip() and not l.startswith('#')]
    names = lines[0].split()
    data = np.array([[float(x) for x in l.split()] for l in lines[1:]], dtype=np.float64)
    return names, data


def _is_cuda_tensor(x):
    return type(x).__module__.startswith('torch') and getattr(x, 'is_cuda', False)


class _DeviceStats(object):
    """Order statistics and correlations of an (n, P) CUDA chain computed on the device: one sort of all
    columns, the same index arithmetic as the host functions above, only the few numbers needed come back.
    Large ensembles produce millions of samples; the reference sorts every column (and, for om_asc, 360
    re-ordered copies of it) in numpy (others.py:308-374)."""

    def __init__(self, samples):
        import torch
        self.torch = torch
        self.x = samples.to(torch.float64)
        self.sorted = torch.sort(self.x, dim=0).values
        self.n = int(self.x.shape[0])
        self.corr = torch.corrcoef(self.x.T).cpu().numpy() if self.x.shape[1] > 1 else np.ones((1, 1))

    def median(self, i):
        a, n = self.sorted[:, i], self.n
        return float(0.5 * (a[n // 2 - 1] + a[n // 2])) if n % 2 == 0 else float(a[(n - 1) // 2])

    def median_range(self, i, confidencelevel):
        a = self.sorted[:, i]
        SV = _significant_volume(self.n, confidencelevel)
        start = int((self.n - SV) / 2 - 1)
        return float(a[start]), float(a[start + SV])

    def periodic(self, i, period=360, confidencelevel=1):
        torch = self.torch
        a = self.sorted[:, i] % period
        SV = _significant_volume(self.n, confidencelevel)
        m = self.n - SV - 1
        best = (None, float('inf'), None)
        for threshold in np.arange(0, period, period / 360.):
            k = int(torch.argmin(torch.abs(a - threshold)))
            r = torch.sort(torch.cat((a[k:] - period, a[:k]))).values
            widths = r[SV:SV + m] - r[:m]
            j = int(torch.argmin(widths))                         # first minimum, as numpy's argmin
            lo, hi = float(r[j]), float(r[j + SV])
            value, error, median = 0.5 * (lo + hi), 0.5 * (hi - lo), float(r[int(j + SV / 2.)])
            if error < best[1]:
                best = (value, error, median)
        value, error, median = best
        return median % period, value + error - median, median - (value - error)


def brief_summary(names, samples):
    """(dict_median, text) for parameter columns `names` of the (n, len(names)) sample array -- a numpy array,
    or a CUDA tensor (the chain of sampler.FusedEnsembleSampler), in which case sorting, order statistics
    and correlations run on the device."""
    dict_median = {}
    out = ['#Medians of the simulated samples:\n',
           '#(Units: px in mas; mu_a and mu_d in mas/yr; incl in rad.)\n']
    dev = _DeviceStats(samples) if _is_cuda_tensor(samples) else None
    for i, p in enumerate(names):
        col = None if dev is not None else samples[:, i]
        if 'om_asc' in p:
            med, up, low = dev.periodic(i) if dev is not None else periodic_sample2estimate(col)
            dict_median[p] = med
            out.append('%s = %f + %f - %f (deg)\n' % (p, med, up, low))
            continue
        med = dev.median(i) if dev is not None else sample2median(col)
        lo, hi = dev.median_range(i, 1) if dev is not None else sample2median_range(col, 1)
        dict_median[p] = med
        if 'ra' in p:
            k = 180 / np.pi / 15
            out.append('%s = %s + %f - %f (ms)\n' % (p, deg_float2dms(med * k), (hi - med) * k * 3600 * 1000,
                                                    (med - lo) * k * 3600 * 1000))
        elif 'dec' in p:
            k = 180 / np.pi
            out.append('%s = %s + %f - %f (mas)\n' % (p, deg_float2dms(med * k), (hi - med) * k * 3600 * 1000,
                                                     (med - lo) * k * 3600 * 1000))
        else:
            out.append('%s = %f + %f - %f\n' % (p, med, hi - med, med - lo))
    out.append('\n#Correlation coefficients:\n')
    for i in range(1, len(names)):
        for j in range(i):
            r = dev.corr[j, i] if dev is not None else np.corrcoef(samples[:, j], samples[:, i])[0, 1]
            out.append('r__%s__%s = %f\n' % (names[j], names[i], r))
    return dict_median, ''.join(out)


def make_a_summary_of_bayesian_inference(samplefile, likelihood, outputfile=None, samples=None, names=None):
    """Writes bayesian_estimates.dat next to `samplefile` (simulate.py:244-250): the brief
    summary plus chi-square / reduced chi-square at the medians (device kernel) and the
    reference epoch.  `samples` / `names`: the chain itself (numpy or CUDA tensor, parameter columns only)
    when it is still at hand -- a CUDA chain is summarised on the device; otherwise the file is read.
    Returns (dict_median, chi_sq, reduced chi_sq, path)."""
    if samples is None:
        names, data = read_posterior_samples(samplefile)
        names, data = names[:-2], data[:, :-2]
    else:
        data = samples
    dict_median, text = brief_summary(names, data)
    chi_sq, rchsq = likelihood.calculate_reduced_chi_square(dict_median)
    text += '\nchi-square = %f\nreduced chi-square = %f\n' % (chi_sq, rchsq)
    text += 'The reference epoch is MJD %d \n' % likelihood.refepoch
    if outputfile is None:
        outputfile = samplefile.replace('posterior_samples', 'bayesian_estimates')
    with open(outputfile, 'w') as f:
        f.write(text)
    return dict_median, chi_sq, rchsq, outputfile
