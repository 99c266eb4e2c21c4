"""Alias so that code written against the reference (`import gpim`) runs on the B200 engine."""
from gpim_b200 import utils, reconstructor, boptimizer, skreconstructor, vreconstructor  # noqa: F401
from gpim_b200 import gprutils  # noqa: F401
