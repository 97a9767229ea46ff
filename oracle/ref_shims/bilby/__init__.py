"""Stand-in for the absent bilby package: just enough for the reference to import (see ../README.md)."""
from . import core
from .core import prior


class Likelihood(object):
    def __init__(self, parameters=None):
        self.parameters = parameters if parameters is not None else {}

    def log_likelihood(self):
        raise NotImplementedError


def run_sampler(*args, **kwargs):
    raise NotImplementedError('bilby stand-in: no sampler')


class result(object):
    @staticmethod
    def read_in_result(*args, **kwargs):
        raise NotImplementedError('bilby stand-in')
