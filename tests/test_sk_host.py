"""
CPU test of the HOST side of ``gpim.skreconstructor(ski=False)``: the engine is replaced by a stand-in that answers
``fit_adam(gpytorch_params=True)`` / ``factorize`` / ``predict`` through oracle/sk_oracle.py's arithmetic.  Under test:
the raw-parameter packing, the lengthscale-bound layout, the trajectory -> hyperparams bookkeeping, the constant
mean being removed before the factorisation and added back after the prediction -- not the CUDA arithmetic
(tests/test_gpu_sk.py).
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import workloads as W
from oracle import gp_oracle as O
from oracle.sk_oracle import SKOracleGP, sk_kernel_matrix

NAMES = {0: "RBF", 1: "Matern52"}


class StandInEngine:
    device = torch.device("cpu")

    @staticmethod
    def _theta(u, bounds, n_ls, d):
        lo = torch.tensor(bounds[2:2 + n_ls], dtype=u.dtype)
        hi = torch.tensor(bounds[2 + n_ls:2 + 2 * n_ls], dtype=u.dtype)
        ls = lo + (hi - lo) * torch.sigmoid(u[3:3 + n_ls])
        return F.softplus(u[0]), F.softplus(u[1]) + 1e-4, u[2], (ls.expand(d) if n_ls == 1 else ls)

    def fit_adam(self, kernel_id, X, y, jitter, u, bounds, n_ls, iters, lr, gpytorch_params=False):
        assert gpytorch_params and jitter == 0.0
        N, d = X.shape
        up = u.clone().requires_grad_(True)
        opt = torch.optim.Adam([up], lr=lr)
        traj = torch.zeros(max(iters, 1), 4 + d, dtype=X.dtype)
        for it in range(iters):
            opt.zero_grad()
            v, n, c, ls = self._theta(up, bounds, n_ls, d)
            A = v * sk_kernel_matrix(NAMES[kernel_id], X, X, ls) + n * torch.eye(N, dtype=X.dtype)
            L = torch.linalg.cholesky(A)
            a = torch.linalg.solve_triangular(L, (y - c).unsqueeze(-1), upper=False).squeeze(-1)
            loss = (0.5 * a @ a + torch.log(torch.diagonal(L)).sum() + 0.5 * N * math.log(2 * math.pi)) / N
            loss.backward()
            opt.step()
            with torch.no_grad():
                v, n, c, ls = self._theta(up, bounds, n_ls, d)
                traj[it] = torch.cat([v.reshape(1), n.reshape(1), c.reshape(1), ls, loss.detach().reshape(1)])
        with torch.no_grad():
            u.copy_(up)
            v, n, c, ls = self._theta(u, bounds, n_ls, d)
        return traj[:iters], torch.cat([v.reshape(1), n.reshape(1), c.reshape(1), ls]), torch.zeros(1, dtype=torch.int32)

    def factorize(self, kernel_id, theta, X, y, jitter):
        return {"args": (kernel_id, theta.clone(), X, y.clone()), "info": torch.zeros(1, dtype=torch.int32)}

    def predict(self, kernel_id, theta, X, fac, Xs):
        kid, th, X0, yc = fac["args"]
        N = X0.shape[0]
        A = th[0] * sk_kernel_matrix(NAMES[kid], X0, X0, th[3:]) + th[1] * torch.eye(N, dtype=X0.dtype)
        L = torch.linalg.cholesky(A)
        Ks = th[0] * sk_kernel_matrix(NAMES[kid], X0, Xs, th[3:])
        pack = torch.linalg.solve_triangular(L, torch.cat((yc.unsqueeze(-1), Ks), dim=1), upper=False)
        return pack[:, 0] @ pack[:, 1:], (th[0] - pack[:, 1:].pow(2).sum(0) + th[1]).sqrt()


@pytest.mark.parametrize("kernel,iso", [("RBF", False), ("Matern52", True)])
def test_sk_reconstructor_glue_follows_the_oracle(monkeypatch, kernel, iso):
    from gpim_b200.gpreg import skgpr
    monkeypatch.setattr(skgpr, "get_engine", lambda device=None: StandInEngine())
    R = W.dummy_blob(14, 60) + 0.3
    Xs, Xf = O.sparse_grid(R), O.full_grid(R)
    ls = [1.0, 10.0] if iso else [[1.0, 1.0], [10.0, 10.0]]
    kw = dict(kernel=kernel, lengthscale=ls, learning_rate=0.1, iterations=8, isotropic=iso)
    m0, s0, hp0 = SKOracleGP(Xs, R, Xf, **kw).run()
    rec = skgpr.skreconstructor(Xs, R, Xf, ski=False, verbose=0, **kw)
    assert float(rec.model.likelihood.noise_covar.noise) == pytest.approx(math.log(2.0) + 1e-4)
    assert rec.model.covar_module.base_kernel.lengthscale.shape == (1, 1 if iso else 2)
    assert float(rec.model.covar_module.base_kernel.lengthscale[0, 0]) == pytest.approx(5.5)      # Interval midpoint
    m1, s1, hp1 = rec.run()
    assert m1.shape == s1.shape == R.shape
    np.testing.assert_allclose(m1, m0, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(s1, s0, rtol=1e-8)
    np.testing.assert_allclose(np.array(hp1["noise"]), np.array(hp0["noise"]), rtol=1e-10)
    np.testing.assert_allclose(np.array(hp1["lengthscale"]), np.array(hp0["lengthscale"]), rtol=1e-10)
    assert abs(float(rec.model.mean_module.constant)) > 1e-3                                        # it was trained


@pytest.mark.parametrize("kernel", ["RBF", "Matern52"])
def test_sk_oracle_reproduces_its_committed_training_vectors(kernel, golden_dir):
    """tests/golden/oracle_sk_train_*.npz (make_oracle_vectors.py): the restatement must keep reproducing them."""
    import os
    g = np.load(os.path.join(golden_dir, f"oracle_sk_train_{kernel}.npz"))
    R = g["R"]
    ora = SKOracleGP(O.sparse_grid(R), R, O.full_grid(R), kernel=kernel, lengthscale=[[1.0, 1.0], [10.0, 10.0]],
                     learning_rate=0.1, iterations=15)
    mean, sd, hp = ora.run()
    np.testing.assert_allclose(np.array(hp["noise"]), g["noise"], rtol=1e-9)
    np.testing.assert_allclose(np.array(hp["lengthscale"]), g["lengthscale"], rtol=1e-9)
    np.testing.assert_allclose(np.array(ora.losses), g["loss"], rtol=1e-9)
    np.testing.assert_allclose(mean, g["mean"], rtol=0, atol=1e-9 * np.abs(g["mean"]).max())
    np.testing.assert_allclose(sd, g["sd"], rtol=1e-9)


@pytest.mark.parametrize("kernel", ["RBF", "Matern52"])
def test_sk_oracle_loss_and_prediction_by_the_dense_route(kernel):
    """Nothing reference-held pins the GPyTorch-semantics oracle, so it is verified by an ALGEBRAICALLY INDEPENDENT
    route: the marginal likelihood from torch.distributions.MultivariateNormal on the dense covariance
    (no hand-rolled Cholesky / log-determinant), the prediction from torch.linalg.solve on the full system."""
    R = W.dummy_blob(14, 60) + 0.3
    Xs, Xf = O.sparse_grid(R), O.full_grid(R)
    g = SKOracleGP(Xs, R, Xf, kernel=kernel, lengthscale=[[1.0, 1.0], [10.0, 10.0]], learning_rate=0.1, iterations=5).train()
    with torch.no_grad():
        v, noise, c, ls = g.theta()
        N = g.X.shape[0]
        A = v * sk_kernel_matrix(kernel, g.X, g.X, ls) + noise * torch.eye(N, dtype=torch.float64)
        mvn = torch.distributions.MultivariateNormal(c.expand(N), covariance_matrix=A)
        assert float(g.loss()) == pytest.approx(float(-mvn.log_prob(g.y) / N), rel=1e-10)
        Xt = g.Xtest[::5]
        Ks = v * sk_kernel_matrix(kernel, g.X, Xt, ls)
        mean = c + Ks.t() @ torch.linalg.solve(A, g.y - c)
        var = v - (Ks * torch.linalg.solve(A, Ks)).sum(0) + noise
    m0, s0 = g.predict()
    np.testing.assert_allclose(m0.reshape(-1)[::5], mean.numpy(), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(s0.reshape(-1)[::5], var.sqrt().numpy(), rtol=1e-8)
