from .base import Prior


class PriorDict(dict):
    def __init__(self, conversion_function=None):
        dict.__init__(self)
        self.conversion_function = conversion_function


class _Named(Prior):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class Uniform(_Named):
    pass


class Gaussian(_Named):
    pass


class Sine(_Named):
    pass


class Constraint(_Named):
    pass
