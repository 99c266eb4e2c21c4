"""
CPU checks of oracle/mt_oracle.py (vreconstructor(independent=True) restated) and of the host side of
``gpim.vreconstructor``.  Nothing reference-held pins this oracle (gpytorch is absent, the reference has no test for
the class), so its loss and its prediction are verified here by an ALGEBRAICALLY INDEPENDENT route: the joint
N T-dimensional Gaussian of torch.distributions (block-diagonal covariance + the multitask noise) for the marginal
likelihood, joint Gaussian conditioning for the prediction, and finite differences for the gradient.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import workloads as W
from oracle import gp_oracle as O
from oracle.mt_oracle import MTOracleGP
from oracle.sk_oracle import sk_kernel_matrix, SKOracleGP

NAMES = {0: "RBF", 1: "Matern52"}


def vector_field(n=12, knock=40, seed=1):
    """(n, n, 2) observations with NaN knock-outs: two different smooth components."""
    R = W.dummy_blob(n, knock, seed) + 0.3
    y = np.stack([R, 0.5 - 0.7 * R + 0.05 * np.cos(np.arange(n))[:, None]], axis=-1)
    return y


def grids(y):
    X_full = O.full_grid(y[..., 0])
    X_sparse = X_full.copy().astype(np.float64)
    X_sparse[:, np.isnan(y).any(axis=-1)] = np.nan
    return X_sparse, X_full


@pytest.mark.parametrize("kernel", ["RBF", "Matern52"])
def test_loss_equals_the_joint_multitask_gaussian(kernel):
    y = vector_field()
    Xs, Xf = grids(y)
    ora = MTOracleGP(Xs, y, Xf, kernel=kernel, lengthscale=[[1.0, 1.0], [8.0, 8.0]])
    with torch.no_grad():
        ora.raw_outputscale += torch.tensor([0.3, -0.4], dtype=torch.float64)
        ora.raw_task_noises += torch.tensor([-2.0, -1.0], dtype=torch.float64)
        ora.raw_noise -= 3.0
        ora.constant += torch.tensor([0.2, -0.1], dtype=torch.float64)
        ora.raw_lengthscale += torch.tensor([0.5, -0.3], dtype=torch.float64)
    s, noise, c, ls = [t.detach() for t in ora.theta()]
    N, T = ora.Y.shape
    K = sk_kernel_matrix(kernel, ora.X, ora.X, ls)
    # joint covariance in GPyTorch's interleaved (point-major) layout: entry ((i, t), (j, t')) = delta_tt' s_t K_ij
    cov = torch.zeros(N * T, N * T, dtype=torch.float64)
    for t in range(T):
        cov[t::T, t::T] = s[t] * K + noise[t] * torch.eye(N, dtype=torch.float64)
    mvn = torch.distributions.MultivariateNormal(c.repeat(N), covariance_matrix=cov)
    want = -mvn.log_prob(ora.Y.reshape(-1)) / (N * T)
    assert float(ora.loss()) == pytest.approx(float(want), rel=1e-10)
    # one shared lengthscale, T outputscales / constants / task noises + one global noise
    assert [p.numel() for p in ora.params] == [T, 1, T, 2, T]


def test_gradient_matches_finite_differences():
    y = vector_field(10, 25)
    Xs, Xf = grids(y)
    ora = MTOracleGP(Xs, y, Xf, kernel="RBF", lengthscale=None)              # softplus lengthscale
    loss = ora.loss()
    loss.backward()
    for p in ora.params:
        for k in range(p.numel()):
            g = float(p.grad.reshape(-1)[k])
            with torch.no_grad():
                p.reshape(-1)[k] += 1e-6
                up = float(ora.loss())
                p.reshape(-1)[k] -= 2e-6
                dn = float(ora.loss())
                p.reshape(-1)[k] += 1e-6
            assert g == pytest.approx((up - dn) / 2e-6, rel=2e-5, abs=1e-9)


def test_prediction_equals_joint_conditioning_and_mc_estimator_converges():
    y = vector_field()
    Xs, Xf = grids(y)
    ora = MTOracleGP(Xs, y, Xf, kernel="RBF", lengthscale=[[1.0, 1.0], [8.0, 8.0]], iterations=10).train()
    mean, sd = ora.predict()
    s, noise, c, ls = [t.detach() for t in ora.theta()]
    Xt = ora.Xtest[:7]
    for t in range(ora.T):
        A = s[t] * sk_kernel_matrix("RBF", ora.X, ora.X, ls) + noise[t] * torch.eye(ora.X.shape[0], dtype=torch.float64)
        Ks = s[t] * sk_kernel_matrix("RBF", ora.X, Xt, ls)
        m = c[t] + Ks.t() @ torch.linalg.solve(A, ora.Y[:, t] - c[t])
        v = s[t] - (Ks * torch.linalg.solve(A, Ks)).sum(0) + noise[t]
        np.testing.assert_allclose(mean.reshape(-1, ora.T)[:7, t], m.numpy(), rtol=1e-9)
        np.testing.assert_allclose(sd.reshape(-1, ora.T)[:7, t], v.sqrt().numpy(), rtol=1e-9)
    # the reference's estimator (100 draws per point) scatters around the closed form with the expected spread
    mm, ss = ora.predict_mc(100, seed=3)
    z = (mm - mean.reshape(mm.shape)) / (sd.reshape(mm.shape) / 10.0)
    assert abs(z.mean()) < 0.2 and 0.85 < z.std() < 1.15
    rel = ss / sd.reshape(ss.shape) - 1.0
    assert abs(rel.mean()) < 0.02 and 0.05 < rel.std() < 0.09            # sd of a sample sd: 1 / sqrt(2 * 99) = 0.071


def test_single_output_is_the_scalar_model_with_a_split_noise():
    """T = 1: the same likelihood as skreconstructor's model when noise_sk = task_noise + noise."""
    R = W.dummy_blob(12, 40, 1) + 0.3
    y = R[..., None]
    Xs, Xf = O.sparse_grid(R), O.full_grid(R)
    mt = MTOracleGP(Xs, y, Xf, kernel="Matern52", lengthscale=[[1.0, 1.0], [8.0, 8.0]])
    sk = SKOracleGP(Xs, R, Xf, kernel="Matern52", lengthscale=[[1.0, 1.0], [8.0, 8.0]])
    with torch.no_grad():
        # softplus(raw_sk) + 1e-4 == 2 (softplus(0) + 1e-4)
        target = 2 * (math.log(2.0) + 1e-4) - 1e-4
        sk.raw_noise.fill_(math.log(math.expm1(target)))
    assert float(mt.loss()) == pytest.approx(float(sk.loss()), rel=1e-12)


# ---------------------------------------------------------------------------------------------
# host side of gpim.vreconstructor with a stand-in engine that answers through the oracle's arithmetic
# ---------------------------------------------------------------------------------------------
class StandInEngine:
    device = torch.device("cpu")

    def fit_adam_mt(self, kernel_id, X, Y, jitter, u, ls_bounds, n_ls, iters, lr):
        N, d = X.shape
        T = Y.shape[0]
        assert jitter == 0.0 and u.numel() == 3 * T + 1 + n_ls and torch.all(u == 0)
        up = u.clone().requires_grad_(True)
        opt = torch.optim.Adam([up], lr=lr)
        traj = torch.zeros(max(iters, 1), d + 1, dtype=X.dtype)

        def theta(v):
            s = F.softplus(v[:T])
            noise = (F.softplus(v[T:2 * T]) + 1e-4) + (F.softplus(v[2 * T]) + 1e-4)
            c = v[2 * T + 1:3 * T + 1]
            raw = v[3 * T + 1:]
            if ls_bounds is None:
                ls = F.softplus(raw)
            else:
                lo, hi = torch.tensor(ls_bounds[:n_ls], dtype=X.dtype), torch.tensor(ls_bounds[n_ls:], dtype=X.dtype)
                ls = lo + (hi - lo) * torch.sigmoid(raw)
            return s, noise, c, (ls.expand(d) if n_ls == 1 else ls)

        for it in range(iters):
            opt.zero_grad()
            s, noise, c, ls = theta(up)
            K = sk_kernel_matrix(NAMES[kernel_id], X, X, ls)
            tot = 0.0
            for t in range(T):
                L = torch.linalg.cholesky(s[t] * K + noise[t] * torch.eye(N, dtype=X.dtype))
                a = torch.linalg.solve_triangular(L, (Y[t] - c[t]).unsqueeze(-1), upper=False).squeeze(-1)
                tot = tot + 0.5 * a @ a + torch.log(torch.diagonal(L)).sum() + 0.5 * N * math.log(2 * math.pi)
            loss = tot / (N * T)
            loss.backward()
            opt.step()
            with torch.no_grad():
                traj[it] = torch.cat([theta(up)[3], loss.detach().reshape(1)])
        with torch.no_grad():
            u.copy_(up)
            s, noise, c, ls = theta(u)
            th = torch.cat([s[:, None], noise[:, None], c[:, None], ls[None, :].expand(T, -1)], dim=1)
        return traj[:iters], th, torch.zeros(1, dtype=torch.int32)

    def factorize(self, kernel_id, theta, X, y, jitter):
        return {"args": (kernel_id, theta.clone(), X, y.clone()), "info": torch.zeros(1, dtype=torch.int32)}

    def predict(self, kernel_id, theta, X, fac, Xs, mean=None, sd=None):
        kid, th, X0, yc = fac["args"]
        N = X0.shape[0]
        A = th[0] * sk_kernel_matrix(NAMES[kid], X0, X0, th[3:]) + th[1] * torch.eye(N, dtype=X0.dtype)
        L = torch.linalg.cholesky(A)
        Ks = th[0] * sk_kernel_matrix(NAMES[kid], X0, Xs, th[3:])
        pack = torch.linalg.solve_triangular(L, torch.cat((yc.unsqueeze(-1), Ks), dim=1), upper=False)
        mean.copy_(pack[:, 0] @ pack[:, 1:])
        sd.copy_((th[0] - pack[:, 1:].pow(2).sum(0) + th[1]).sqrt())
        return mean, sd


@pytest.mark.parametrize("kernel,iso,bounds", [("RBF", False, True), ("Matern52", True, True), ("RBF", False, False)])
def test_vreconstructor_glue_follows_the_oracle(monkeypatch, kernel, iso, bounds):
    from gpim_b200.gpreg import vgpr
    monkeypatch.setattr(vgpr, "get_engine", lambda device=None: StandInEngine())
    y = vector_field()
    Xs, Xf = grids(y)
    ls = None if not bounds else ([1.0, 8.0] if iso else [[1.0, 1.0], [8.0, 8.0]])
    kw = dict(kernel=kernel, lengthscale=ls, learning_rate=0.1, iterations=6, isotropic=iso)
    m0, s0, hp0 = MTOracleGP(Xs, y, Xf, **kw).run()
    rec = vgpr.vreconstructor(Xs, y, Xf, independent=True, verbose=0, **kw)
    assert rec.model.covar_module.base_kernel.lengthscale.shape == (1, 1 if iso else 2)
    assert rec.model.likelihood.task_noises.shape == (2,) and rec.model.likelihood.noise.shape == (1,)
    m1, s1, hp1 = rec.run()
    assert m1.shape == s1.shape == y.shape
    np.testing.assert_allclose(m1, m0, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(s1, s0, rtol=1e-8)
    np.testing.assert_allclose(np.array(hp1["lengthscale"]), np.array(hp0["lengthscale"]), rtol=1e-10)
    assert np.abs(rec.model.mean_module.constant.numpy()).min() > 1e-4            # both constants were trained
    # the reference's Monte-Carlo estimator on request
    mm, ss = rec.predict(mc_samples=100)
    assert mm.shape == m1.shape and 0.01 < np.abs(mm - m1).max() / np.abs(s1).max() < 1.0


def test_vreconstructor_correlated_outputs_are_not_on_the_path():
    import gpim
    with pytest.raises(NotImplementedError):
        gpim.vreconstructor(None, None, independent=False)
