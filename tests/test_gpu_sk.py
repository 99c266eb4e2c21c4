"""
GPU parity tests (-m gpu) of ``gpim.skreconstructor(ski=False)`` -- GPyTorch's exact-GP semantics on
the engine: ``gpg_fit_adam_sk`` + ``gpg_factorize`` / ``gpg_predict`` against oracle/sk_oracle.py (ExactGP with
ConstantMean, ScaleKernel, Interval-constrained lengthscales, GaussianLikelihood and the per-datum marginal
likelihood, restated) on the same inputs.  fp64: 1e-6 on the trajectory; fp32: mean 1e-4 / sd 1e-3 (BASELINE.json).
Parity is pinned by the restated library formulas only (no GPyTorch in the image, no reference test for this class).
"""
import numpy as np
import pytest
import torch

import workloads as W
from oracle import gp_oracle as O
from oracle.sk_oracle import SKOracleGP

pytestmark = pytest.mark.gpu


def relinf(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("kernel,iso", [("RBF", False), ("Matern52", False), ("RBF", True), ("Matern52", True)])
def test_sk_run_matches_oracle(kernel, iso):
    import gpim
    R = W.dummy_blob(20, 200) + 0.3                   # a non-zero level, so that the constant mean has work to do
    Xs, Xf = gpim.utils.get_sparse_grid(R), gpim.utils.get_full_grid(R)
    ls = [1.0, 10.0] if iso else [[1.0, 1.0], [10.0, 10.0]]
    kw = dict(kernel=kernel, lengthscale=ls, learning_rate=0.1, iterations=30, isotropic=iso)
    ref = SKOracleGP(O.sparse_grid(R), R, O.full_grid(R), **kw)
    m0, s0, hp0 = ref.run()
    rec = gpim.skreconstructor(Xs, R, Xf, ski=False, verbose=0, **kw)
    m1, s1, hp1 = rec.run()
    assert m1.shape == s1.shape == R.shape
    assert sorted(hp1) == ["lengthscale", "noise"] and len(hp1["noise"]) == 30
    assert np.array(hp1["lengthscale"]).shape == (30, 1 if iso else 2)
    np.testing.assert_allclose(np.array(hp1["noise"]), np.array(hp0["noise"]), rtol=1e-6)
    np.testing.assert_allclose(np.array(hp1["lengthscale"]), np.array(hp0["lengthscale"]), rtol=1e-6)
    np.testing.assert_allclose(np.array(rec.loss_all), np.array(ref.losses), rtol=1e-7, atol=1e-9)
    v, noise, c, l = ref.theta()
    assert abs(float(rec.model.mean_module.constant) - float(c.detach())) < 1e-6
    assert abs(float(rec.model.covar_module.outputscale) - float(v.detach())) < 1e-6 * float(v.detach())
    assert relinf(m1, m0) < 1e-6 and relinf(s1, s0) < 1e-6
    # first recorded noise is one Adam step away from softplus(0) + 1e-4
    assert abs(hp1["noise"][0] - (np.log1p(np.exp(-0.1)) + 1e-4)) < 1e-6 or abs(hp1["noise"][0] - (np.log1p(np.exp(0.1)) + 1e-4)) < 1e-6


def test_sk_single_precision_3d_and_warm_restart():
    import gpim
    R = W.hyperspectral((10, 10, 8))
    Xs, Xf = gpim.utils.get_sparse_grid(R), gpim.utils.get_full_grid(R)
    kw = dict(kernel="Matern52", lengthscale=[[1., 1., 1.], [10., 10., 10.]], learning_rate=0.1, iterations=25)
    ref = SKOracleGP(O.sparse_grid(R), R, O.full_grid(R), **kw)
    m0, s0, hp0 = ref.run()
    rec = gpim.skreconstructor(Xs, R, Xf, ski=False, verbose=0, precision="single", **kw)
    m1, s1, hp1 = rec.run()
    assert m1.dtype == np.float32 and m1.shape == R.shape
    np.testing.assert_allclose(np.array(hp1["noise"]), np.array(hp0["noise"]), rtol=2e-3)
    np.testing.assert_allclose(np.array(hp1["lengthscale"]), np.array(hp0["lengthscale"]), rtol=2e-3)
    # prediction parity at the SAME hyper-parameters: feed the fp64 checker the trained fp32 raw values
    u = rec.model.kernel.u.double()
    with torch.no_grad():
        ref.raw_outputscale[0], ref.raw_noise[0], ref.constant[0] = u[0], u[1], u[2]
        ref.raw_lengthscale.copy_(u[3:])
    m2, s2 = ref.predict()
    assert relinf(m1, m2) < 1e-4 and relinf(s1, s2) < 1e-3
    # a second train() continues from the trained values with a fresh optimiser (skgpr.py:186-187)
    ref.train(iterations=4)
    rec.train(iterations=4)
    assert len(hp1["noise"]) == 29
    np.testing.assert_allclose(np.array(hp1["noise"][-4:]), np.array(ref.noise_all[-4:]), rtol=5e-3)


def test_sk_tensor_core_sizes_track_fp64():
    """N >= 1024 in fp32 runs the tcgen05 factorisation inside the Adam loop: the trajectory follows the engine's
    own fp64 run."""
    import gpim
    R = W.spiral_scan(128)
    Xs, Xf = gpim.utils.get_sparse_grid(R), gpim.utils.get_full_grid(R)
    kw = dict(kernel="RBF", lengthscale=[[1., 1.], [6., 6.]], learning_rate=0.05, iterations=8, ski=False, verbose=0)
    a = gpim.skreconstructor(Xs, R, Xf, precision="double", **kw)
    b = gpim.skreconstructor(Xs, R, Xf, precision="single", **kw)
    assert a.model.X.shape[0] >= 1024
    ma, sa, ha = a.run()
    mb, sb, hb = b.run()
    np.testing.assert_allclose(np.array(hb["noise"]), np.array(ha["noise"]), rtol=2e-3)
    np.testing.assert_allclose(np.array(hb["lengthscale"]), np.array(ha["lengthscale"]), rtol=2e-3)
    assert relinf(mb, ma) < 5e-3 and relinf(sb, sa) < 5e-3


def test_sk_out_of_path_branches_raise():
    import gpim
    R = W.dummy_blob(12, 30)
    Xs, Xf = gpim.utils.get_sparse_grid(R), gpim.utils.get_full_grid(R)
    with pytest.raises(NotImplementedError):
        gpim.skreconstructor(Xs, R, Xf)                              # ski=True is the reference's default
    with pytest.raises(NotImplementedError):
        gpim.skreconstructor(Xs, R, Xf, kernel="Spectral")
    with pytest.raises(KeyError):
        gpim.skreconstructor(Xs, R, Xf, kernel="nope", ski=False)


@pytest.mark.parametrize("kernel", ["RBF", "Matern52"])
def test_sk_run_matches_committed_training_vectors(kernel, golden_dir):
    """tests/golden/oracle_sk_train_*.npz: gpim.skreconstructor(ski=False).run() in fp64 against the frozen
    trajectory and reconstruction -- no oracle in the loop."""
    import os
    import gpim
    g = np.load(os.path.join(golden_dir, f"oracle_sk_train_{kernel}.npz"))
    R = g["R"]
    rec = gpim.skreconstructor(gpim.utils.get_sparse_grid(R), R, gpim.utils.get_full_grid(R), kernel=kernel, ski=False,
                               lengthscale=[[1.0, 1.0], [10.0, 10.0]], learning_rate=0.1, iterations=15, verbose=0)
    mean, sd, hp = rec.run()
    np.testing.assert_allclose(np.array(hp["noise"]), g["noise"], rtol=1e-6)
    np.testing.assert_allclose(np.array(hp["lengthscale"]), g["lengthscale"], rtol=1e-6)
    np.testing.assert_allclose(np.array(rec.loss_all), g["loss"], rtol=1e-7, atol=1e-9)
    assert relinf(mean, g["mean"]) < 1e-6 and relinf(sd, g["sd"]) < 1e-6
