from .gpr import reconstructor, ExactGPModel  # noqa: F401
