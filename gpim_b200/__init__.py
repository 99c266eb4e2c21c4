"""
gpim_b200 -- B200-native exact-GP-on-grids engine behind the GPim API.

Public surface = the reference's gpim/__init__.py:1-5 restricted to the accelerated path:
``utils`` (grid / data-layout helpers), ``reconstructor`` (exact and inducing-point GP), ``boptimizer``,
``skreconstructor`` with ``ski=False`` (GPyTorch's exact-GP semantics) and ``vreconstructor`` with
``independent=True`` (independent multi-output GP).
"""
from . import gprutils as utils  # noqa: F401
from .gpreg.gpr import reconstructor  # noqa: F401
from .gpreg.skgpr import skreconstructor  # noqa: F401
from .gpreg.vgpr import vreconstructor  # noqa: F401
from .gpbayes.boptim import boptimizer  # noqa: F401

__version__ = "0.2.0"
