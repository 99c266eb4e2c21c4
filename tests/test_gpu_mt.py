"""
GPU parity tests (-m gpu) of ``gpim.vreconstructor(independent=True)`` -- GPyTorch's batch-independent multi-output
exact GP on the engine: ``gpg_fit_adam_mt`` + per-output ``gpg_factorize`` / ``gpg_predict`` against
oracle/mt_oracle.py on the same inputs.  fp64: 1e-6 on the trajectory and the reconstruction; fp32: mean 1e-4 /
sd 1e-3 at equal hyper-parameters (BASELINE.json).  The oracle is pinned by the restated library formulas and by the
independent-route checks of tests/test_mt_oracle.py (no GPyTorch in the image, no reference test for this class).
"""
import numpy as np
import pytest
import torch

import workloads as W
from oracle import gp_oracle as O
from oracle.mt_oracle import MTOracleGP

pytestmark = pytest.mark.gpu


def relinf(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def vector_field(R):
    """(..., 2) observations from a scalar field with NaN gaps: two different smooth components."""
    ramp = 0.05 * np.cos(np.arange(R.shape[0])).reshape((-1,) + (1,) * (R.ndim - 1))
    return np.stack([R + 0.3, 0.5 - 0.7 * R + ramp], axis=-1)


def grids(y):
    X_full = O.full_grid(y[..., 0])
    X_sparse = X_full.copy().astype(np.float64)
    X_sparse[:, np.isnan(y).any(axis=-1)] = np.nan
    return X_sparse, X_full


@pytest.mark.parametrize("kernel,iso,bounds", [("RBF", False, True), ("Matern52", False, True), ("RBF", True, True),
                                               ("Matern52", False, False)])
def test_mt_run_matches_oracle(kernel, iso, bounds):
    import gpim
    y = vector_field(W.dummy_blob(20, 200))
    Xs, Xf = grids(y)
    ls = None if not bounds else ([1.0, 10.0] if iso else [[1.0, 1.0], [10.0, 10.0]])
    kw = dict(kernel=kernel, lengthscale=ls, learning_rate=0.1, iterations=25, isotropic=iso)
    ref = MTOracleGP(Xs, y, Xf, **kw)
    m0, s0, hp0 = ref.run()
    rec = gpim.vreconstructor(Xs, y, Xf, independent=True, verbose=0, **kw)
    m1, s1, hp1 = rec.run()
    assert m1.shape == s1.shape == y.shape and sorted(hp1) == ["lengthscale"]
    assert np.array(hp1["lengthscale"]).shape == (25, 1 if iso else 2)
    np.testing.assert_allclose(np.array(hp1["lengthscale"]), np.array(hp0["lengthscale"]), rtol=1e-6)
    np.testing.assert_allclose(np.array(rec.loss_all), np.array(ref.losses), rtol=1e-7, atol=1e-9)
    s, noise, c, l = [t.detach().numpy() for t in ref.theta()]
    np.testing.assert_allclose(rec.model.mean_module.constant.numpy(), c, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(rec.model.covar_module.outputscale.numpy(), s, rtol=1e-6)
    tot = rec.model.likelihood.task_noises.numpy() + rec.model.likelihood.noise.numpy()
    np.testing.assert_allclose(tot, noise, rtol=1e-6)
    assert relinf(m1, m0) < 1e-6 and relinf(s1, s0) < 1e-6


def test_mt_single_precision_3d_three_outputs():
    import gpim
    R = W.hyperspectral((10, 10, 8))
    y = np.stack([R, 1.0 - R, 0.3 * R + 0.1], axis=-1)
    Xs, Xf = grids(y)
    kw = dict(kernel="RBF", lengthscale=[[1., 1., 1.], [10., 10., 10.]], learning_rate=0.1, iterations=15)
    ref = MTOracleGP(Xs, y, Xf, **kw)
    m0, s0, hp0 = ref.run()
    rec = gpim.vreconstructor(Xs, y, Xf, independent=True, verbose=0, precision="single", **kw)
    m1, s1, hp1 = rec.run()
    assert m1.dtype == np.float32 and m1.shape == y.shape
    np.testing.assert_allclose(np.array(hp1["lengthscale"]), np.array(hp0["lengthscale"]), rtol=2e-3)
    # prediction parity at the SAME hyper-parameters: feed the fp64 checker the trained fp32 raw values
    u = rec.model._u.detach().cpu().double()
    T = 3
    with torch.no_grad():
        ref.raw_outputscale.copy_(u[:T]); ref.raw_task_noises.copy_(u[T:2 * T]); ref.raw_noise.copy_(u[2 * T:2 * T + 1])
        ref.constant.copy_(u[2 * T + 1:3 * T + 1]); ref.raw_lengthscale.copy_(u[3 * T + 1:])
    m2, s2 = ref.predict()
    assert relinf(m1, m2) < 1e-4 and relinf(s1, s2) < 1e-3


def test_mt_monte_carlo_estimator_scatters_around_the_closed_form():
    """The reference returns the mean / sd of 100 draws (vgpr.py:218-225); predict(mc_samples=100) reproduces that
    estimator and it scatters around the closed form with the spread the sampling theory predicts."""
    import gpim
    y = vector_field(W.dummy_blob(24, 250))
    Xs, Xf = grids(y)
    rec = gpim.vreconstructor(Xs, y, Xf, independent=True, verbose=0, lengthscale=[[1., 1.], [10., 10.]], iterations=10)
    rec.train()
    mean, sd = rec.predict()
    torch.manual_seed(0)
    mm, ss = rec.predict(mc_samples=100)
    z = (mm - mean) / (sd / 10.0)
    assert abs(z.mean()) < 0.15 and 0.9 < z.std() < 1.1
    rel = ss / sd - 1.0
    assert abs(rel.mean()) < 0.02 and 0.055 < rel.std() < 0.085          # 1 / sqrt(2 * 99) = 0.071


def test_mt_tensor_core_sizes_track_fp64():
    """N >= 1024 in fp32: every per-output pass of the Adam loop runs the tcgen05 factorisation."""
    import gpim
    y = vector_field(W.spiral_scan(128) - 0.3)
    Xs, Xf = grids(y)
    kw = dict(kernel="RBF", lengthscale=[[1., 1.], [6., 6.]], learning_rate=0.05, iterations=6, independent=True, verbose=0)
    a = gpim.vreconstructor(Xs, y, Xf, precision="double", **kw)
    b = gpim.vreconstructor(Xs, y, Xf, precision="single", **kw)
    assert a.model._X.shape[0] >= 1024
    ma, sa, ha = a.run()
    mb, sb, hb = b.run()
    np.testing.assert_allclose(np.array(hb["lengthscale"]), np.array(ha["lengthscale"]), rtol=2e-3)
    assert relinf(mb, ma) < 5e-3 and relinf(sb, sa) < 5e-3


def test_mt_out_of_path_branches_raise():
    import gpim
    y = vector_field(W.dummy_blob(12, 30))
    Xs, Xf = grids(y)
    with pytest.raises(NotImplementedError):
        gpim.vreconstructor(Xs, y, Xf)                               # independent=False is the reference's default
    with pytest.raises(NotImplementedError):
        gpim.vreconstructor(Xs, y, Xf, independent=True, kernel="Spectral")
    with pytest.raises(KeyError):
        gpim.vreconstructor(Xs, y, Xf, independent=True, kernel="nope")
