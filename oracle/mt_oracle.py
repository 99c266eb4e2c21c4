"""
oracle/mt_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain torch + autograd) of ``gpim.vreconstructor(..., independent=True)``: the independent-output
branch of gpim/gpreg/vgpr.py.  Nothing under ``gpim_b200/`` imports this module; it checks
``gpg_fit_adam_mt`` and ``gpim_b200/gpreg/vgpr.py`` in tests/test_gpu_mt.py.

PARITY UNPINNED.  The arithmetic lives in ``gpytorch`` (``>= 0.3.6``; not installed in this image, no lock file) and
the reference has no test for vreconstructor, so there is no golden vector.  What is restated is the published
algorithm of the classes the reference composes, and -- because nothing reference-held pins it -- the loss and the
prediction are additionally verified by an algebraically independent route in tests/test_mt_oracle.py (the joint
N T-dimensional block-diagonal MultivariateNormal of torch.distributions and joint Gaussian conditioning):

  vgpr.py:119          MultitaskGaussianLikelihood(num_tasks)  (rank 0): noise_t = task_noises[t] + noise,
                       both softplus(raw) + 1e-4 (GreaterThan(1e-4)), raw 0
  vgpr.py:340-354      ivgprmodel: ConstantMean(batch_shape=[T])  -> one constant per output, raw 0
                       kernel.batch_shape = [T] is assigned AFTER the base kernel was built
                       (gpytorch_kernels.py:60-69), so its raw_lengthscale keeps shape (1, n_ls): ONE lengthscale
                       shared by all outputs;  ScaleKernel(kernel, batch_shape=[T]) -> one outputscale per output
                       MultitaskMultivariateNormal.from_batch_mvn: outputs are independent given the parameters
  gpytorch_kernels.py:55-57  lengthscale_constraint = Interval(lo, hi) when bounds are given, else GPyTorch's default
                       Positive(): lengthscale = softplus(raw)
  vgpr.py:169-179      Adam over model.parameters(); loss = -ExactMarginalLogLikelihood
                         = -sum_t log N(y_t; c_t, s_t K_l + noise_t I) / (N T)      (num_data = N T)
  vgpr.py:183-187      after every step base_kernel.lengthscale.tolist()[0] is recorded
  vgpr.py:218-225      predict: 100 rsample() draws of likelihood(model(Xtest)), their mean and sqrt(var): Monte-Carlo
                       estimates of  mean_t = c_t + k*^T A_t^-1 (y_t - c_t),
                       var_t = s_t k** - k*^T A_t^-1 k* + noise_t,  A_t = s_t K_l + noise_t I  -- computed here in closed
                       form (``predict``) and, for the statistical check, as the reference does (``predict_mc``).

Where GPyTorch itself approximates (``fast_pred_var``, conjugate gradients beyond 800 training points) this
restatement computes the exact quantity.  No random numbers are drawn in training (all raw parameters start at 0).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .gp_oracle import to_rows
from .sk_oracle import sk_kernel_matrix

KERNEL_NAMES = ("RBF", "Matern52")


def vector_training_rows(X, y):
    """gprutils.prepare_training_data(vector_valued=True), gprutils.py:49-55: rows with any NaN dropped."""
    Xr = to_rows(np.asarray(X, dtype=np.float64))
    Xr = Xr[~np.isnan(Xr).any(axis=1)]
    yr = np.asarray(y, dtype=np.float64).reshape(-1, np.shape(y)[-1])
    yr = yr[~np.isnan(yr).any(axis=1)]
    return Xr, yr


class MTOracleGP:
    """vreconstructor(independent=True) restated: same constructor arguments that matter on this branch."""

    def __init__(self, X, y, Xtest=None, kernel="RBF", lengthscale=None, learning_rate=0.1, iterations=50,
                 precision="double", isotropic=False):
        if kernel not in KERNEL_NAMES:
            raise KeyError(kernel)
        self.dtype = torch.float32 if precision == "single" else torch.float64
        self.kernel_name = kernel
        dim = np.ndim(y) - 1
        Xr, yr = vector_training_rows(X, y)
        self.X = torch.from_numpy(Xr).to(self.dtype)
        self.Y = torch.from_numpy(yr).to(self.dtype)                       # [N, T]
        self.T = self.Y.shape[1]
        n_ls = 1 if isotropic else dim
        if lengthscale is not None:
            t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=self.dtype).reshape(-1)
            self.ls_lo, self.ls_hi = t(lengthscale[0]), t(lengthscale[1])
            assert self.ls_lo.numel() == n_ls
        else:
            self.ls_lo = self.ls_hi = None
        z = lambda n: torch.zeros(n, dtype=self.dtype, requires_grad=True)
        self.raw_outputscale, self.raw_task_noises, self.raw_noise = z(self.T), z(self.T), z(1)
        self.constant, self.raw_lengthscale = z(self.T), z(n_ls)
        self.params = [self.raw_task_noises, self.raw_noise, self.constant, self.raw_lengthscale, self.raw_outputscale]
        self.learning_rate, self.iterations = learning_rate, iterations
        self.fulldims = (np.asarray(Xtest).shape[1:] if Xtest is not None else np.asarray(X).shape[1:]) + (self.T,)
        self.Xtest = torch.from_numpy(to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype) if Xtest is not None else None
        self.lscales, self.losses = [], []

    def theta(self):
        """-> outputscale [T], total noise [T], constant [T], lengthscale [n_ls]"""
        s = F.softplus(self.raw_outputscale)
        noise = (F.softplus(self.raw_task_noises) + 1e-4) + (F.softplus(self.raw_noise) + 1e-4)
        if self.ls_lo is None:
            ls = F.softplus(self.raw_lengthscale)
        else:
            ls = self.ls_lo + (self.ls_hi - self.ls_lo) * torch.sigmoid(self.raw_lengthscale)
        return s, noise, self.constant, ls

    def loss(self):
        s, noise, c, ls = self.theta()
        N = self.X.shape[0]
        K = sk_kernel_matrix(self.kernel_name, self.X, self.X, ls)
        eye = torch.eye(N, dtype=self.dtype)
        total = 0.0
        for t in range(self.T):
            L = torch.linalg.cholesky(s[t] * K + noise[t] * eye)
            a = torch.linalg.solve_triangular(L, (self.Y[:, t] - c[t]).unsqueeze(-1), upper=False).squeeze(-1)
            total = total + 0.5 * a @ a + torch.log(torch.diagonal(L)).sum() + 0.5 * N * math.log(2.0 * math.pi)
        return total / (N * self.T)

    def train(self, learning_rate=None, iterations=None):
        lr = self.learning_rate if learning_rate is None else learning_rate
        iters = self.iterations if iterations is None else iterations
        opt = torch.optim.Adam([{"params": self.params}], lr=lr)          # vgpr.py:169-170
        for _ in range(iters):
            opt.zero_grad()
            loss = self.loss()
            loss.backward()
            opt.step()
            self.lscales.append(self.theta()[3].detach().tolist())
            self.losses.append(float(loss.detach()))
        return self

    def predict_rows(self, Xs):
        """closed-form (mean, sd) of the noisy predictive distribution, [M, T] each"""
        with torch.no_grad():
            s, noise, c, ls = self.theta()
            N = self.X.shape[0]
            K = sk_kernel_matrix(self.kernel_name, self.X, self.X, ls)
            Ks = sk_kernel_matrix(self.kernel_name, self.X, Xs, ls)
            eye = torch.eye(N, dtype=self.dtype)
            means, sds = [], []
            for t in range(self.T):
                L = torch.linalg.cholesky(s[t] * K + noise[t] * eye)
                pack = torch.linalg.solve_triangular(L, torch.cat(((self.Y[:, t] - c[t]).unsqueeze(-1), s[t] * Ks), dim=1),
                                                     upper=False)
                means.append(c[t] + pack[:, 0] @ pack[:, 1:])
                sds.append((s[t] - pack[:, 1:].pow(2).sum(0) + noise[t]).sqrt())
        return torch.stack(means, dim=1), torch.stack(sds, dim=1)

    def predict(self, Xtest=None):
        Xs = self.Xtest if Xtest is None else torch.from_numpy(to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype)
        mean, sd = self.predict_rows(Xs)
        shape = self.fulldims if Xtest is None else np.asarray(Xtest).shape[1:] + (self.T,)
        return mean.numpy().reshape(shape), sd.numpy().reshape(shape)

    def predict_mc(self, n_samples=100, seed=0, Xtest=None):
        """vgpr.py:218-225: mean and sqrt(var) of n_samples draws of the noisy predictive distribution.  The outputs are
        independent and only per-point moments are taken, so draws from the per-point marginals have the same law."""
        Xs = self.Xtest if Xtest is None else torch.from_numpy(to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype)
        mean, sd = self.predict_rows(Xs)
        g = torch.Generator().manual_seed(seed)
        draws = mean[None] + sd[None] * torch.randn((n_samples,) + tuple(mean.shape), dtype=self.dtype, generator=g)
        return draws.mean(0).numpy(), draws.var(0).sqrt().numpy()

    def run(self):
        self.train()
        mean, sd = self.predict()
        return mean, sd, {"lengthscale": self.lscales}
