"""Minimal unit algebra standing in for astropy.units (see ../README.md)."""
import math
import numpy as np

_BASES = ('s', 'm', 'rad')


class Unit(object):
    __array_ufunc__ = None          # numpy scalars defer to our reflected operators

    def __init__(self, scale, dims):
        self.scale = float(scale)
        self.dims = tuple(dims)

    # unit algebra
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale * other.scale, [a + b for a, b in zip(self.dims, other.dims)])
        if isinstance(other, Quantity):
            return Quantity(other.value, self * other.unit)
        return Quantity(float(other), self)

    def __rmul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit * self)
        return Quantity(float(other), self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale / other.scale, [a - b for a, b in zip(self.dims, other.dims)])
        return Quantity(1.0 / float(other), self)

    def __rtruediv__(self, other):
        inv = Unit(1.0 / self.scale, [-a for a in self.dims])
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit * inv)
        return Quantity(float(other), inv)

    def __pow__(self, p):
        return Unit(self.scale ** p, [a * p for a in self.dims])

    def same_dims(self, other):
        return all(abs(a - b) < 1e-12 for a, b in zip(self.dims, other.dims))

    def to(self, other, value=1.0):
        if not self.same_dims(other):
            raise ValueError('unit conversion between different dimensions')
        return value * (self.scale / other.scale)

    def is_dimensionless(self):
        return all(abs(a) < 1e-12 for a in self.dims)


dimensionless = Unit(1.0, (0, 0, 0))
s = Unit(1.0, (1, 0, 0))
m = Unit(1.0, (0, 1, 0))
rad = Unit(1.0, (0, 0, 1))
d = Unit(86400.0, (1, 0, 0))
yr = Unit(31557600.0, (1, 0, 0))
AU = Unit(149597870700.0, (0, 1, 0))
deg = Unit(math.pi / 180.0, (0, 0, 1))
arcsec = Unit((math.pi / 180.0) / 3600.0, (0, 0, 1))
mas = Unit((math.pi / 180.0) / 3600.0 * 0.001, (0, 0, 1))


def _q(x):
    return x if isinstance(x, Quantity) else Quantity(float(x), dimensionless)


class Quantity(float):
    """A float that carries a unit.  Being a float subclass, numpy array constructors
    (np.mat / np.asmatrix of nested lists) take its value, as they do for astropy's
    ndarray-based Quantity."""

    def __new__(cls, value, unit=dimensionless):
        obj = float.__new__(cls, value)
        obj.unit = unit
        return obj

    @property
    def value(self):
        return float(self)

    def to(self, unit):
        return Quantity(self.unit.to(unit, float(self)), unit)

    # multiplicative
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(float(self), self.unit * other)
        o = _q(other)
        return Quantity(float(self) * float(o), self.unit * o.unit)

    def __rmul__(self, other):
        o = _q(other)
        return Quantity(float(o) * float(self), o.unit * self.unit)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(float(self), self.unit / other)
        o = _q(other)
        return Quantity(float(self) / float(o), self.unit / o.unit)

    def __rtruediv__(self, other):
        o = _q(other)
        return Quantity(float(o) / float(self), o.unit / self.unit)

    def __pow__(self, p):
        return Quantity(float(self) ** p, self.unit ** p)

    # additive: the right operand is converted to the left operand's unit
    def _conv(self, other):
        o = _q(other)
        if not self.unit.same_dims(o.unit):
            raise ValueError('incompatible units in addition')
        if o.unit.scale == self.unit.scale:
            return float(o)
        return float(o) * (o.unit.scale / self.unit.scale)

    def __add__(self, other):
        return Quantity(float(self) + self._conv(other), self.unit)

    def __sub__(self, other):
        return Quantity(float(self) - self._conv(other), self.unit)

    def __radd__(self, other):
        return _q(other).__add__(self)

    def __rsub__(self, other):
        return _q(other).__sub__(self)

    def __neg__(self):
        return Quantity(-float(self), self.unit)

    def __abs__(self):
        return Quantity(abs(float(self)), self.unit)

    def _angle_rad(self):
        if self.unit.is_dimensionless():
            return float(self)
        if not self.unit.same_dims(rad):
            raise ValueError('trigonometric function of a non-angle')
        if self.unit.scale == 1.0:
            return float(self)
        return float(self) * (self.unit.scale / rad.scale)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs:
            return NotImplemented
        name = ufunc.__name__
        if name in ('sin', 'cos', 'tan'):
            return Quantity(float(getattr(np, name)(inputs[0]._angle_rad())), dimensionless)
        if name == 'arctan2':
            a, b = _q(inputs[0]), _q(inputs[1])
            return Quantity(float(np.arctan2(float(a), a._conv(b))), rad)
        if name == 'arctan':
            return Quantity(float(np.arctan(float(inputs[0]))), rad)
        a, b = (inputs + (None,))[:2]
        if name == 'multiply':
            return _q(a).__mul__(b) if not isinstance(a, Quantity) else a.__mul__(b)
        if name in ('true_divide', 'divide'):
            return _q(a).__truediv__(b)
        if name == 'add':
            return _q(a).__add__(b)
        if name == 'subtract':
            return _q(a).__sub__(b)
        if name == 'power':
            return _q(a).__pow__(float(b))
        if name == 'negative':
            return _q(a).__neg__()
        return NotImplemented

    def __repr__(self):
        return 'Quantity(%r, scale=%r dims=%r)' % (float(self), self.unit.scale, self.unit.dims)
