"""
oracle/sparse_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain torch + autograd) of the reference's inducing-point path, the first item of the
"next" list: ``gpim.reconstructor(..., sparse=True)`` (gpim/gpreg/gpr.py:145-155, 198-199)
over Pyro's ``SparseGPRegression`` with its default approximation (VFE, Titsias 2009).  Nothing under
``gpim_b200/`` imports this module; it is the checker of the CUDA inducing-point path (csrc/sparse.cuh,
gpg_sparse_* in include/gpgrid.h), used by tests/test_gpu_sparse.py.

PARITY UNPINNED: the reference's tests do not exercise ``sparse=True`` and the stored notebook runs used the
CUDA generator, so there is no golden vector.  The arithmetic restates the published algorithm of
``pyro.contrib.gp.models.sgpr.SparseGPRegression`` (pyro-ppl >= 0.4.1, requirements.txt:6) at the reference's
call sites:

  model()    Kuu = k(Xu) + jitter I;  Luu = chol(Kuu);  W = (Luu^-1 k(Xu, X))^T  (N x M);  D = noise 1
             trace_term = clamp(sum(diag Kff - rowsum W^2) / noise, 0)
             loss = -log N(y; 0, W W^T + D) + trace_term / 2      (LowRankMultivariateNormal + pyro.factor)
  forward()  W_Dinv = W^T / D;  K = W_Dinv W + I;  L = chol(K);  Ws = Luu^-1 k(Xu, X*)
             [a | Q] = L^-1 [W_Dinv y | Ws];  loc = a^T Q
             var = diag k(X*, X*) (+ noise) - colsum Ws^2 + colsum Q^2
  trained    kernel variance / lengthscale (Uniform priors -> interval constraints, MAP), noise (positive),
             and the inducing inputs Xu themselves (plain Parameter), Adam, one fresh optimiser per train().
  gpr.py     indpoints default len(X) // 10 (at least 1, at most len(X));  Xu = X[::len(X) // indpoints];
             hyperparams["inducing_points"] gets Xu after every step;  predict uses noiseless=False.
"""
import math

import numpy as np
import torch

from .gp_oracle import OracleGP, kernel_matrix, to_rows


def vfe_loss(kernel_name, X, y, Xu, v, ls, noise, a, jitter):
    """SparseGPRegression.model's objective (minus the constant Uniform log-priors) as a function of the CONSTRAINED
    hyper-parameters and the inducing inputs -- differentiable, so autograd gives the gradients the CUDA path's
    closed form must reproduce."""
    N, M = X.shape[0], Xu.shape[0]
    dtype = X.dtype
    Kuu = kernel_matrix(kernel_name, Xu, Xu, v, ls, a) + jitter * torch.eye(M, dtype=dtype)
    Luu = torch.linalg.cholesky(Kuu)
    Kuf = kernel_matrix(kernel_name, Xu, X, v, ls, a)
    W = torch.linalg.solve_triangular(Luu, Kuf, upper=False).t()              # N x M
    Kffdiag = v.expand(N)                                      # stationary kernels: k(x, x) = variance
    trace_term = ((Kffdiag - W.pow(2).sum(-1)).sum() / noise).clamp(min=0)
    # log N(y; 0, W W^T + noise I) by the matrix-determinant / Woodbury lemmas (LowRankMultivariateNormal)
    A = W.t() @ W / noise + torch.eye(M, dtype=dtype)
    LA = torch.linalg.cholesky(A)
    Wty = W.t() @ y / noise
    c = torch.linalg.solve_triangular(LA, Wty.unsqueeze(-1), upper=False).squeeze(-1)
    quad = y @ y / noise - c @ c
    logdet = N * torch.log(noise) + 2.0 * torch.log(torch.diagonal(LA)).sum()
    log_prob = -0.5 * (quad + logdet + N * math.log(2.0 * math.pi))
    return -log_prob + 0.5 * trace_term


def vfe_predict(kernel_name, X, y, Xu, Xs, v, ls, noise, a, jitter):
    """SparseGPRegression.forward(Xs, full_cov=False, noiseless=False) -> (loc, var) tensors."""
    dtype = X.dtype
    M = Xu.shape[0]
    Kuu = kernel_matrix(kernel_name, Xu, Xu, v, ls, a) + jitter * torch.eye(M, dtype=dtype)
    Luu = torch.linalg.cholesky(Kuu)
    Kuf = kernel_matrix(kernel_name, Xu, X, v, ls, a)
    W = torch.linalg.solve_triangular(Luu, Kuf, upper=False).t()
    W_Dinv = W.t() / noise                                     # M x N
    K = W_Dinv @ W + torch.eye(M, dtype=dtype)
    L = torch.linalg.cholesky(K)
    W_Dinv_y = W_Dinv @ y.unsqueeze(-1)
    Kus = kernel_matrix(kernel_name, Xu, Xs, v, ls, a)
    Ws = torch.linalg.solve_triangular(Luu, Kus, upper=False)
    pack = torch.linalg.solve_triangular(L, torch.cat((W_Dinv_y, Ws), dim=1), upper=False)
    loc = (pack[:, :1].t() @ pack[:, 1:]).squeeze(0)
    var = v + noise - Ws.pow(2).sum(0) + pack[:, 1:].pow(2).sum(0)
    return loc, var


class SparseOracleGP(OracleGP):
    """reconstructor(sparse=True) restated.  Reuses OracleGP for the data layout, the prior draws and the
    constrained parametrisation; adds the inducing inputs and the VFE objective."""

    def __init__(self, X, y, Xtest=None, indpoints=None, **kwargs):
        super().__init__(X, y, Xtest, **kwargs)
        n = len(self.X)
        if indpoints is None:                                  # gpr.py:146-148
            indpoints = n // 10
            indpoints = indpoints + 1 if indpoints == 0 else indpoints
        else:                                                  # gpr.py:149-150
            indpoints = n if indpoints > n else indpoints
        self.Xu = self.X[::n // indpoints].clone().requires_grad_(True)      # gpr.py:151
        self.params.append(self.Xu)
        self.indpoints_all = []

    # ---------------------------------------------------------------------------------------
    def _theta(self):
        v = self.tf_v(self.u_v)
        ls = self.tf_l(self.u_l)
        noise = self.tf_n(self.u_n)
        a = self.tf_n(self.u_a) if self.kernel_name == "RationalQuadratic" else torch.ones((), dtype=self.dtype)
        return v, ls, noise, a

    def loss(self):
        """-ELBO of the MAP guide up to the constant Uniform log-priors: VFE bound."""
        v, ls, noise, a = self._theta()
        return vfe_loss(self.kernel_name, self.X, self.y, self.Xu, v, ls, noise, a, self.jitter)

    def train(self, learning_rate=None, iterations=None):
        lr = self.learning_rate if learning_rate is None else learning_rate
        iters = self.iterations if iterations is None else iterations
        opt = torch.optim.Adam(self.params, lr=lr)             # fresh optimiser, warm parameters (gpr.py:184-185)
        self.losses = []
        for _ in range(iters):
            opt.zero_grad()
            loss = self.loss()
            loss.backward()
            opt.step()
            v, ls, noise, _ = self._theta()
            self.lscales.append(ls.detach().tolist())          # recorded AFTER the step (gpr.py:194-199)
            self.amp_all.append(float(v.detach()))
            self.noise_all.append(float(noise.detach()))
            self.indpoints_all.append(self.Xu.detach().numpy().copy())
            self.losses.append(float(loss.detach()))
        return self

    def predict(self, Xtest=None):
        """(mean, sd) shaped like the test grid; SparseGPRegression.forward(full_cov=False, noiseless=False)."""
        Xs = self.Xtest if Xtest is None else torch.from_numpy(to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype)
        with torch.no_grad():
            v, ls, noise, a = self._theta()
            loc, var = vfe_predict(self.kernel_name, self.X, self.y, self.Xu, Xs, v, ls, noise, a, self.jitter)
        shape = self.fulldims if Xtest is None else np.asarray(Xtest).shape[1:]
        return loc.numpy().reshape(shape), var.sqrt().numpy().reshape(shape)

    def run(self):
        self.train()
        mean, sd = self.predict()
        return mean, sd, {"lengthscale": self.lscales, "noise": self.noise_all, "variance": self.amp_all,
                          "inducing_points": self.indpoints_all}
