"""
Kernel factory: replaces gpim/kernels/pyro_kernels.py:14-96.  The covariance math itself lives in
CUDA (csrc/common.cuh cov_from_r2); what stays on the host is the bookkeeping Pyro did: Uniform
priors -> interval constraints, the prior draw that initialises the MAP value, and the
unconstrained parametrisation Adam steps on.
"""
import warnings

import torch
from torch.distributions import constraints, transform_to

from .._lib import KERNEL_IDS


class KernelState:
    """Host mirror of a pyro.contrib.gp kernel with Delta-MAP guides for variance / lengthscale.

    Attributes the reference touches from outside (gpr.py:195-197,214-216,318; boptim.py:319):
    ``lengthscale``, ``variance``, ``lengthscale_map``, ``variance_map`` (CPU tensors).
    """

    def __init__(self, name, input_dim, lengthscale, amp, dtype, variance0, lengthscale0):
        self.name = name
        self.kernel_id = KERNEL_IDS[name]
        self.input_dim = input_dim
        self.dtype = dtype
        self.amp_lo, self.amp_hi = amp
        self.ls_lo, self.ls_hi = lengthscale
        # one lengthscale for all dimensions: scalar bounds [lo, hi], or one-element lists [[lo], [hi]] on d > 1
        # inputs (pyro broadcasts a shape-[1] lengthscale over the input dimensions)
        self.isotropic = self.ls_lo.dim() == 0 or (self.ls_lo.numel() == 1 and input_dim > 1)
        self.n_ls = 1 if self.isotropic else int(self.ls_lo.numel())
        if not self.isotropic and self.n_ls != input_dim:
            raise ValueError(f"lengthscale bounds have {self.n_ls} entries per side; expected a scalar, one entry "
                             f"or one per input dimension ({input_dim})")
        if self.ls_hi.numel() != self.ls_lo.numel():
            raise ValueError("lower and upper lengthscale bounds differ in length")
        self._tf_v = transform_to(constraints.interval(self.amp_lo, self.amp_hi))
        self._tf_l = transform_to(constraints.interval(self.ls_lo, self.ls_hi))
        # unconstrained storage = transform_to(constraint).inv(value), as PyroParam does
        self.u_variance = self._tf_v.inv(variance0)
        self.u_lengthscale = self._tf_l.inv(lengthscale0)
        self.u_noise = torch.zeros((), dtype=dtype)            # noise = 1.0 (GPRegression default)
        self.u_scale_mixture = torch.zeros((), dtype=dtype)    # RationalQuadratic scale_mixture = 1.0
        self._theta = None                                      # constrained values after the last sync

    # constrained views -----------------------------------------------------------------
    @property
    def variance(self):
        return self._tf_v(self.u_variance)

    @property
    def lengthscale(self):
        return self._tf_l(self.u_lengthscale)

    variance_map = variance
    lengthscale_map = lengthscale

    @property
    def scale_mixture(self):
        return self.u_scale_mixture.exp()

    # packing for the C ABI -------------------------------------------------------------
    def pack_u(self):
        """{variance, noise, scale_mixture, lengthscale[n_ls]} unconstrained."""
        return torch.cat([self.u_variance.reshape(1), self.u_noise.reshape(1), self.u_scale_mixture.reshape(1),
                          self.u_lengthscale.reshape(-1)]).to(self.dtype)

    def unpack_u(self, u):
        u = u.detach().cpu().to(self.dtype)
        self.u_variance = u[0].clone()
        self.u_noise = u[1].clone()
        self.u_scale_mixture = u[2].clone()
        self.u_lengthscale = (u[3].clone().reshape(self.ls_lo.shape) if self.isotropic else u[3:3 + self.n_ls].clone())

    def pack_theta(self):
        """Constrained {variance, noise, scale_mixture, lengthscale[d]} (isotropic value repeated)."""
        ls = self.lengthscale.reshape(-1)
        if self.isotropic:
            ls = ls.expand(self.input_dim)
        return torch.cat([self.variance.reshape(1), self.u_noise.exp().reshape(1), self.scale_mixture.reshape(1),
                          ls]).to(self.dtype)

    def bounds(self):
        return [float(self.amp_lo), float(self.amp_hi)] + [float(v) for v in self.ls_lo.reshape(-1)] + \
            [float(v) for v in self.ls_hi.reshape(-1)]


def get_kernel(kernel_type, input_dim, lengthscale, use_gpu=False, **kwargs):
    """Same call as pyro_kernels.get_kernel (pyro_kernels.py:14).  Draw order follows the
    reference: variance prior first, lengthscale prior second (:81-94), on the CPU generator when
    use_gpu is False and on the CUDA generator otherwise (the reference's default-tensor-type
    switch, :46-49, moves the Uniform parameters and hence torch.rand to that device)."""
    precision = kwargs.get("precision", "double")
    dtype = torch.float32 if precision == "single" else torch.float64
    if kernel_type not in KERNEL_IDS:
        print('Select one of the currently available kernels:', '"RBF", "RationalQuadratic", "Matern52"')
        raise KeyError(kernel_type)
    amp = kwargs.get("amplitude")
    amp = [1e-4, 10.] if amp is None else amp
    dev = "cuda" if (use_gpu and torch.cuda.is_available()) else "cpu"
    as_t = lambda v: torch.as_tensor(v, dtype=dtype)
    amp_t = (as_t(amp[0]), as_t(amp[1]))
    ls_t = (as_t(lengthscale[0]), as_t(lengthscale[1]))
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", category=UserWarning)
        r_v = torch.rand(amp_t[0].shape, dtype=dtype, device=dev).cpu()
        r_l = torch.rand(ls_t[0].shape, dtype=dtype, device=dev).cpu()
    variance0 = amp_t[0] + r_v * (amp_t[1] - amp_t[0])
    lengthscale0 = ls_t[0] + r_l * (ls_t[1] - ls_t[0])
    return KernelState(kernel_type, input_dim, ls_t, amp_t, dtype, variance0, lengthscale0)
