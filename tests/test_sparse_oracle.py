"""
CPU tests of oracle/sparse_oracle.py, the checker waiting for the inducing-point path (not built
yet).  No reference golden exists for sparse=True, so these pin the restatement through properties of the
VFE approximation itself: it collapses onto the exact GP when every training point is an inducing point, its
objective upper-bounds the exact negative log marginal likelihood, and training lowers it.
"""
import numpy as np
import torch

import workloads as W
from oracle import gp_oracle as O
from oracle.sparse_oracle import SparseOracleGP


def _problem():
    R = W.dummy_blob(20, 200)
    return R, O.sparse_grid(R), O.full_grid(R)


def test_inducing_point_selection_follows_gpr_py():
    R, Xs, Xf = _problem()
    n = int((~np.isnan(R)).sum())
    g = SparseOracleGP(Xs, R, Xf, kernel="RBF")
    step = n // (n // 10)
    assert len(g.Xu) == len(range(0, n, step))                 # X[::len(X) // indpoints], gpr.py:151
    np.testing.assert_array_equal(g.Xu.detach().numpy(), g.X.numpy()[::step])
    assert len(SparseOracleGP(Xs, R, Xf, indpoints=10 ** 6).Xu) == n      # capped at len(X), gpr.py:149-150
    assert len(SparseOracleGP(Xs, R, Xf, indpoints=7).Xu) == len(range(0, n, n // 7))


def test_all_points_inducing_recovers_the_exact_gp():
    R, Xs, Xf = _problem()
    kw = dict(kernel="RBF", seed=3, jitter=1e-8)
    exact = O.OracleGP(Xs, R, Xf, **kw)
    sparse = SparseOracleGP(Xs, R, Xf, indpoints=10 ** 6, **kw)
    # same seed -> same prior draw; give both a sensible noise so that the comparison is well conditioned
    for g in (exact, sparse):
        with torch.no_grad():
            g.u_n.fill_(np.log(0.05))
    m0, s0 = exact.predict()
    m1, s1 = sparse.predict()
    np.testing.assert_allclose(m1, m0, rtol=0, atol=2e-5 * np.abs(m0).max())
    np.testing.assert_allclose(s1, s0, rtol=2e-5)
    # and the bound is tight: VFE objective == exact negative log marginal likelihood
    np.testing.assert_allclose(float(sparse.loss().detach()), float(exact.nll().detach()), rtol=1e-5)


def test_vfe_objective_upper_bounds_the_exact_nll_and_training_lowers_it():
    R, Xs, Xf = _problem()
    kw = dict(kernel="Matern52", seed=1, learning_rate=0.1, iterations=15)
    exact = O.OracleGP(Xs, R, Xf, **kw)
    sparse = SparseOracleGP(Xs, R, Xf, indpoints=12, **kw)
    assert float(sparse.loss().detach()) >= float(exact.nll().detach()) - 1e-9               # variational bound
    xu0 = sparse.Xu.detach().numpy().copy()
    mean, sd, hp = sparse.run()
    assert sparse.losses[-1] < sparse.losses[0]
    assert mean.shape == sd.shape == R.shape and np.isfinite(mean).all() and (sd > 0).all()
    assert len(hp["inducing_points"]) == len(hp["noise"]) == len(hp["variance"]) == len(hp["lengthscale"]) == 15
    assert np.abs(hp["inducing_points"][-1] - xu0).max() > 1e-3            # the inducing inputs are trained too
    # recorded AFTER the step: entry 0 is one Adam step away from noise = 1 (exp(-/+ lr))
    assert abs(abs(np.log(hp["noise"][0])) - 0.1) < 1e-6


# oracle-generated fixtures (tests/golden/make_oracle_vectors.py): the restatement must keep reproducing them
import os  # noqa: E402

import pytest  # noqa: E402


@pytest.mark.parametrize("kernel", ["RBF", "Matern52", "RationalQuadratic"])
def test_sparse_oracle_reproduces_its_committed_training_vectors(kernel, golden_dir):
    g = np.load(os.path.join(golden_dir, f"oracle_sparse_train_{kernel}.npz"))
    R = g["R"]
    ora = SparseOracleGP(O.sparse_grid(R), R, O.full_grid(R), indpoints=14, kernel=kernel, learning_rate=0.1,
                         iterations=15, seed=2)
    mean, sd, hp = ora.run()
    for key in ("variance", "noise", "lengthscale"):
        np.testing.assert_allclose(np.array(hp[key]), g[key], rtol=1e-8)
    np.testing.assert_allclose(np.array(hp["inducing_points"]), g["inducing_points"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(np.array(ora.losses), g["loss"], rtol=1e-9)
    np.testing.assert_allclose(mean, g["mean"], rtol=0, atol=1e-8 * np.abs(g["mean"]).max())
    np.testing.assert_allclose(sd, g["sd"], rtol=1e-8)


# ---------------------------------------------------------------------------------------------------------------
# The closed form the CUDA path evaluates (csrc/sparse.cuh header), written out with the same
# intermediates and in the same order as csrc/sparse_driver.cuh, against autograd of the restated objective.
# ---------------------------------------------------------------------------------------------------------------
def _closed_form(kname, X, y, Xu, v, ls, s2, al, jitter):
    from oracle.gp_oracle import kernel_matrix
    import math
    N, m = X.shape[0], Xu.shape[0]
    I = torch.eye(m, dtype=X.dtype)
    Luu = torch.linalg.cholesky(kernel_matrix(kname, Xu, Xu, v, ls, al) + jitter * I)
    Ui = torch.linalg.inv(Luu)
    B = Ui @ kernel_matrix(kname, Xu, X, v, ls, al)               # m x N
    S = B @ B.T
    Ap = S / s2 + I
    LA = torch.linalg.cholesky(Ap)
    LAi = torch.linalg.inv(LA)
    beta = B @ y
    c0 = LAi @ beta
    a0 = LAi.T @ c0
    a = a0 / s2
    rho = y - B.T @ a
    w = Ui.T @ a
    Ainv = LAi.T @ LAi
    Phi = 2 * I - Ainv - Ap - torch.outer(a, a)
    H = Ainv - I
    Guu = -0.5 * Ui.T @ (Phi @ Ui)                                # dF/dKuu
    Guf = ((Ui.T @ H) @ B - torch.outer(w, rho)) / s2             # dF/dKuf
    yy, T_ = y @ y, N * v - torch.diagonal(S).sum()
    loss = 0.5 * (yy / s2 - (c0 @ c0) / s2 ** 2 + N * torch.log(s2) + 2 * torch.log(torch.diagonal(LA)).sum()
                  + N * math.log(2 * math.pi)) + 0.5 * torch.clamp(T_ / s2, min=0)
    ab, aa = (a0 @ beta) / s2, (a0 @ a0) / s2 ** 2
    ds2 = (-yy / (2 * s2 ** 2) + ab / s2 ** 2 - (ab / s2 - aa) / (2 * s2) + N / (2 * s2)
           - (m - torch.diagonal(Ainv).sum()) / (2 * s2) - T_ / (2 * s2 ** 2))
    return loss, ds2, Guu, Guf


@pytest.mark.parametrize("kernel", ["RBF", "Matern52", "RationalQuadratic"])
def test_closed_form_gradient_of_the_vfe_bound_matches_autograd(kernel):
    from oracle.gp_oracle import kernel_matrix
    from oracle.sparse_oracle import vfe_loss
    rng = np.random.RandomState(1)
    n1 = 14
    R = rng.rand(n1, n1)
    R[rng.rand(n1, n1) < 0.3] = np.nan
    Xg = np.mgrid[:n1, :n1].astype(float)
    Xg[:, np.isnan(R)] = np.nan
    o = SparseOracleGP(Xg, R, kernel=kernel, lengthscale=[[1., 1.], [5., 5.]], indpoints=12, jitter=1e-5)
    N = o.X.shape[0]
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    # autograd of the restated objective w.r.t. the constrained hyper-parameters and the inducing inputs
    v, ls, s2, al = (leaf(t) for t in o._theta())
    Xu = leaf(o.Xu)
    ref = vfe_loss(kernel, o.X, o.y, Xu, v, ls, s2, al, o.jitter)
    ref.backward()
    # closed form: sensitivities of the two kernel matrices + explicit noise / variance terms; the chain onto the
    # kernel parameters and Xu (sgp_kgrad_kernel on the device) is taken by autograd of sum(G * K) here
    loss, ds2, Guu, Guf = _closed_form(kernel, o.X, o.y, Xu.detach(), v.detach(), ls.detach(), s2.detach(), al.detach(), o.jitter)
    v2, ls2, al2, Xu2 = leaf(v), leaf(ls), leaf(al), leaf(Xu)
    chain = ((Guu * kernel_matrix(kernel, Xu2, Xu2, v2, ls2, al2)).sum()
             + (Guf * kernel_matrix(kernel, Xu2, o.X, v2, ls2, al2)).sum() + N * v2 / (2 * s2.detach()))
    chain.backward()
    assert abs(float(loss) - float(ref.detach())) < 1e-10 * abs(float(ref.detach()))
    assert abs(float(ds2) - float(s2.grad)) < 1e-9 * abs(float(s2.grad))
    assert abs(float(v2.grad) - float(v.grad)) < 1e-9 * abs(float(v.grad))
    np.testing.assert_allclose(ls2.grad.numpy(), ls.grad.numpy(), rtol=1e-9)
    np.testing.assert_allclose(Xu2.grad.numpy(), Xu.grad.numpy(), rtol=0, atol=1e-9 * float(Xu.grad.abs().max()))
    if kernel == "RationalQuadratic":
        assert abs(float(al2.grad) - float(al.grad)) < 1e-8 * abs(float(al.grad))
    assert float((Guu - Guu.T).abs().max()) < 1e-9 * float(Guu.abs().max())       # the kernel's factor 2 on dF/dXu relies on it


@pytest.mark.parametrize("kernel", ["RBF", "Matern52", "RationalQuadratic"])
def test_vfe_objective_and_prediction_by_the_dense_route(kernel):
    """No reference test or golden pins sparse=True, so the restatement is verified by an ALGEBRAICALLY INDEPENDENT
    route: the N x N covariance Qff + noise I of the Nystrom model built explicitly (torch.linalg.solve, no Woodbury /
    matrix-determinant lemma), its log-density from torch.distributions, the trace term from the dense Kff - Qff, and
    Titsias' predictive equations with the m x m matrix Sigma = (Kuu + Kuf Kfu / noise)^-1 -- none of the
    low-rank factorisations the oracle (and the CUDA path) go through."""
    from oracle.gp_oracle import kernel_matrix
    from oracle.sparse_oracle import vfe_loss, vfe_predict
    R, Xs, Xf = _problem()
    g = SparseOracleGP(Xs, R, Xf, kernel=kernel, indpoints=17, seed=2, learning_rate=0.1, iterations=4).train()
    with torch.no_grad():
        v, ls, noise, a = g._theta()
        X, y, Xu = g.X, g.y, g.Xu.detach()
        N, m = X.shape[0], Xu.shape[0]
        Kuu = kernel_matrix(kernel, Xu, Xu, v, ls, a) + g.jitter * torch.eye(m, dtype=torch.float64)
        Kuf = kernel_matrix(kernel, Xu, X, v, ls, a)
        Qff = Kuf.t() @ torch.linalg.solve(Kuu, Kuf)
        mvn = torch.distributions.MultivariateNormal(torch.zeros(N, dtype=torch.float64),
                                                     covariance_matrix=Qff + noise * torch.eye(N, dtype=torch.float64))
        Kff_diag = kernel_matrix(kernel, X, X, v, ls, a).diagonal()
        dense = -mvn.log_prob(y) + 0.5 * torch.clamp((Kff_diag - Qff.diagonal()).sum() / noise, min=0)
        got = vfe_loss(kernel, X, y, Xu, v, ls, noise, a, g.jitter)
        np.testing.assert_allclose(float(got), float(dense), rtol=1e-9)
        Xt = g.Xtest[::7]
        Kus = kernel_matrix(kernel, Xu, Xt, v, ls, a)
        Sigma = torch.linalg.inv(Kuu + Kuf @ Kuf.t() / noise)
        mean = Kus.t() @ Sigma @ (Kuf @ y) / noise
        var = v + noise - (Kus * torch.linalg.solve(Kuu, Kus)).sum(0) + (Kus * (Sigma @ Kus)).sum(0)
        loc, var0 = vfe_predict(kernel, X, y, Xu, Xt, v, ls, noise, a, g.jitter)
        np.testing.assert_allclose(loc.numpy(), mean.numpy(), rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(var0.numpy(), var.numpy(), rtol=1e-7)
