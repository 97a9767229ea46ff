"""Parameter naming and parameter -> source routing.

Mirrors priors.get_parameters_from_shares / group_elements_by_same_values
(priors.py:268-302), priors.parameter_name_to_pmparin_indice (priors.py:159-173) and
positions.filter_dictionary_of_parameter_with_index (positions.py:32-60), and adds the
one thing the device needs: the routing resolved ONCE into an integer column table.
"""
import numpy as np

from .constants import ROOTS, SENTINEL


def group_elements_by_same_values(alist):
    """Suffix strings ('_0_2', '_1', ...) for one row of `shares` (priors.py:279-302):
    sources with the same non-negative value share a parameter; negative = absent."""
    values = [int(v) for v in alist]
    suffixes, seen = [], set()
    for i, v in enumerate(values):
        if i in seen or v < 0:
            continue
        members = [j for j, w in enumerate(values) if w == v]
        seen.update(members)
        suffixes.append('_' + '_'.join(str(j) for j in members))
    return suffixes


def get_parameters_from_shares(shares):
    """Ordered {name: None} dict; its key order is the default column order of the
    batched (N, P) API (priors.py:268-277)."""
    if len(shares) != len(ROOTS):
        raise ValueError('shares must have %d rows (one per root %s), got %d'
                         % (len(ROOTS), list(ROOTS), len(shares)))
    parameters = {}
    for root, row in zip(ROOTS, shares):
        for suffix in group_elements_by_same_values(row):
            parameters[root + suffix] = None
    return parameters


def parameter_name_to_pmparin_indice(name):
    """'mu_a_0_2_3' -> ([0, 2, 3], 'mu_a')   (priors.py:159-173)"""
    indice, root = [], []
    for token in name.split('_'):
        try:
            indice.append(int(token))
        except ValueError:
            root.append(token)
    return indice, '_'.join(root)


def filter_dictionary_of_parameter_with_index(dict_of_parameters, filter_index):
    """Parameters of source `filter_index`, missing roots filled with -999
    (positions.py:32-60).  Kept for API compatibility; the hot path uses column_table()."""
    key = str(filter_index)
    kept = {k: v for k, v in dict_of_parameters.items() if key in k.split('_')}
    for root in ROOTS:
        if not any(root in k for k in kept):
            kept[root] = SENTINEL
    return kept


def column_table(names, n_sources):
    """idx[s, r] = column (position in `names`) of root r for source s, or -1.

    This is filter_dictionary_of_parameter_with_index + the positional pick from the
    sorted key list (positions.py:20-27, simulate.py:304-308) evaluated once."""
    names = list(names)
    idx = -np.ones((n_sources, len(ROOTS)), dtype=np.int32)
    for c, name in enumerate(names):
        srcs, root = parameter_name_to_pmparin_indice(name)
        if root not in ROOTS:
            raise ValueError('unknown parameter root in %r' % name)
        r = ROOTS.index(root)
        for s in srcs:
            if s < 0 or s >= n_sources:
                raise ValueError('parameter %r names source %d but there are %d sources' % (name, s, n_sources))
            if idx[s, r] >= 0:
                raise ValueError('source %d has two %r parameters (%r and %r)' % (s, root, names[idx[s, r]], name))
            idx[s, r] = c
    return idx
