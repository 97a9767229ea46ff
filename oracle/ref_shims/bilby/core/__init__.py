from . import prior
