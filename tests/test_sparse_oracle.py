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
