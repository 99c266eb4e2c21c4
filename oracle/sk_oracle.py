"""
oracle/sk_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain torch + autograd) of ``gpim.skreconstructor(..., ski=False)`` for the 'RBF' and 'Matern52'
kernels: the exact-GP branch of gpim/gpreg/skgpr.py.  Nothing under ``gpim_b200/`` imports this
module; it checks ``gpg_fit_adam_sk`` and ``gpim_b200/gpreg/skgpr.py`` in tests/test_gpu_sk.py.

PARITY UNPINNED.  The arithmetic lives in ``gpytorch`` (``>= 0.3.6``, requirements.txt / setup.py; not installed in
this image, no lock file) and the reference has no test for skreconstructor, so there is no golden vector.  What is
restated is the published algorithm of the classes the reference composes:

  skgpr.py:143        GaussianLikelihood()            noise = softplus(raw_noise) + 1e-4   (GreaterThan(1e-4)), raw 0
  skgpr.py:399-436    ExactGP with ConstantMean()     constant, raw 0, unconstrained
                      ScaleKernel(kernel)             outputscale = softplus(raw_outputscale)  (Positive), raw 0
  gpytorch_kernels.py:55-69  RBFKernel / MaternKernel(nu=2.5 default), ard_num_dims = input_dim (1 if isotropic),
                      lengthscale_constraint = Interval(lo, hi): lengthscale = lo + (hi - lo) sigmoid(raw), raw 0
                      RBF: exp(-r^2 / 2);  Matern: (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r),  r = |x - z| / lengthscale
  skgpr.py:186-196    Adam over model.parameters(); loss = -ExactMarginalLogLikelihood = -log N(y; c, v K + noise I) / N
  skgpr.py:197-222    after every step: lengthscale.tolist()[0] and noise.item() are recorded
  skgpr.py:283-326    predict: likelihood(model(Xtest)): mean = c + K*^T A^-1 (y - c),
                      var = v - diag(K*^T A^-1 K*) + noise, A = v K + noise I;  sd = sqrt(var)

Where GPyTorch itself approximates -- ``fast_pred_var`` (LOVE, a rank-``maxroot`` Lanczos approximation of the
predictive variance, skgpr.py:285) and, beyond ``max_cholesky_size`` = 800 training points, conjugate gradients with a
stochastic Lanczos log-determinant -- this restatement (and the CUDA path) computes the exact quantity those
approximate.  No random numbers are drawn on this branch (every raw parameter starts at 0), so ``seed`` has no effect.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .gp_oracle import to_rows, training_rows

KERNEL_NAMES = ("RBF", "Matern52")


def sk_kernel_matrix(kernel_name, X, Z, lengthscale):
    """Base kernel (no outputscale): GPyTorch RBFKernel / MaternKernel(nu=2.5) on x / lengthscale."""
    Xs, Zs = X / lengthscale, Z / lengthscale
    diff = Xs.unsqueeze(1) - Zs.unsqueeze(0)
    r2 = diff.pow(2).sum(-1)
    if kernel_name == "RBF":
        return torch.exp(-0.5 * r2)
    if kernel_name == "Matern52":
        r = torch.sqrt(r2.clamp_min(1e-30))
        return (1.0 + math.sqrt(5.0) * r + 5.0 / 3.0 * r2) * torch.exp(-math.sqrt(5.0) * r)
    raise KeyError(kernel_name)


class SKOracleGP:
    """skreconstructor(ski=False) restated: same constructor arguments that matter on this branch."""

    def __init__(self, X, y, Xtest=None, kernel="RBF", lengthscale=None, learning_rate=0.1, iterations=50,
                 precision="double", isotropic=False):
        if kernel not in KERNEL_NAMES:
            raise KeyError(kernel)
        self.dtype = torch.float32 if precision == "single" else torch.float64
        npf = np.float32 if precision == "single" else np.float64
        self.kernel_name = kernel
        dim = np.ndim(y)
        Xr, yr = training_rows(np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64))
        self.X = torch.from_numpy(Xr).to(self.dtype)
        self.y = torch.from_numpy(yr).to(self.dtype)
        if lengthscale is None:                                   # skgpr.py:138-143
            lmean = npf(np.mean(np.shape(y)) / 2)
            lengthscale = [0.0, lmean] if isotropic else [[0.0] * dim, [lmean] * dim]
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=npf), dtype=self.dtype).reshape(-1)
        self.ls_lo, self.ls_hi = t(lengthscale[0]), t(lengthscale[1])
        n_ls = 1 if isotropic else dim
        assert self.ls_lo.numel() == n_ls
        z = lambda n: torch.zeros(n, dtype=self.dtype, requires_grad=True)
        self.raw_outputscale, self.raw_noise, self.constant, self.raw_lengthscale = z(1), z(1), z(1), z(n_ls)
        self.params = [self.raw_noise, self.constant, self.raw_outputscale, self.raw_lengthscale]
        self.learning_rate, self.iterations = learning_rate, iterations
        self.fulldims = np.asarray(Xtest).shape[1:] if Xtest is not None else np.asarray(X).shape[1:]
        self.Xtest = torch.from_numpy(to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype) if Xtest is not None else None
        self.lscales, self.noise_all, self.losses = [], [], []

    def theta(self):
        v = F.softplus(self.raw_outputscale)[0]
        noise = F.softplus(self.raw_noise)[0] + 1e-4
        ls = self.ls_lo + (self.ls_hi - self.ls_lo) * torch.sigmoid(self.raw_lengthscale)
        return v, noise, self.constant[0], ls

    def loss(self):
        v, noise, c, ls = self.theta()
        N = self.X.shape[0]
        A = v * sk_kernel_matrix(self.kernel_name, self.X, self.X, ls) + noise * torch.eye(N, dtype=self.dtype)
        L = torch.linalg.cholesky(A)
        a = torch.linalg.solve_triangular(L, (self.y - c).unsqueeze(-1), upper=False).squeeze(-1)
        nll = 0.5 * a @ a + torch.log(torch.diagonal(L)).sum() + 0.5 * N * math.log(2.0 * math.pi)
        return nll / N

    def train(self, learning_rate=None, iterations=None):
        lr = self.learning_rate if learning_rate is None else learning_rate
        iters = self.iterations if iterations is None else iterations
        opt = torch.optim.Adam([{"params": self.params}], lr=lr)          # skgpr.py:186-187
        for _ in range(iters):
            opt.zero_grad()
            loss = self.loss()
            loss.backward()
            opt.step()
            v, noise, c, ls = self.theta()
            self.lscales.append(ls.detach().tolist())
            self.noise_all.append(float(noise.detach()))
            self.losses.append(float(loss.detach()))
        return self

    def predict(self, Xtest=None):
        Xs = self.Xtest if Xtest is None else torch.from_numpy(to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype)
        with torch.no_grad():
            v, noise, c, ls = self.theta()
            N = self.X.shape[0]
            A = v * sk_kernel_matrix(self.kernel_name, self.X, self.X, ls) + noise * torch.eye(N, dtype=self.dtype)
            L = torch.linalg.cholesky(A)
            Ks = v * sk_kernel_matrix(self.kernel_name, self.X, Xs, ls)
            pack = torch.linalg.solve_triangular(L, torch.cat(((self.y - c).unsqueeze(-1), Ks), dim=1), upper=False)
            mean = c + pack[:, 0] @ pack[:, 1:]
            var = v - pack[:, 1:].pow(2).sum(0) + noise
        shape = self.fulldims if Xtest is None else np.asarray(Xtest).shape[1:]
        return mean.numpy().reshape(shape), var.sqrt().numpy().reshape(shape)

    def run(self):
        self.train()
        mean, sd = self.predict()
        return mean, sd, {"lengthscale": self.lscales, "noise": self.noise_all}
