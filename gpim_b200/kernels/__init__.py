from .gp_kernels import get_kernel, KernelState  # noqa: F401
