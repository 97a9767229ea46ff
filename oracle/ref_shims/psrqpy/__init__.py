"""Stand-in for the absent psrqpy package (see ../README.md)."""


class QueryATNF(object):
    def __init__(self, *a, **k):
        raise NotImplementedError('psrqpy stand-in: no network query')
