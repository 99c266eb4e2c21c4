from .boptim import boptimizer  # noqa: F401
from . import acqfunc  # noqa: F401
