"""
GPU parity tests (-m gpu) of the inducing-point path: the gpg_sparse_* entry points of libgpgrid.so
and ``gpim.reconstructor(sparse=True)`` against oracle/sparse_oracle.py (pyro's SparseGPRegression, VFE, restated)
on the same seeded inputs.  fp64: agreement to summation order; fp32: mean 1e-4 / sd 1e-3 (BASELINE.json) in the
inf-norm relative sense against the fp64 oracle.  Parity is pinned by the restated library formulas only -- the
reference has no test or golden vector for sparse=True (oracle/sparse_oracle.py header).
"""
import numpy as np
import pytest
import torch

import workloads as W
from oracle import gp_oracle as O
from oracle.sparse_oracle import SparseOracleGP, vfe_loss, vfe_predict

pytestmark = pytest.mark.gpu

KERNELS = ["RBF", "Matern52", "RationalQuadratic"]


@pytest.fixture(scope="module")
def eng():
    from gpim_b200._lib import get_engine
    return get_engine()


def relinf(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def problem(n, d, m, seed):
    """n scattered points in [0, 20]^d with a smooth response, m inducing inputs picked like gpr.py:151."""
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d) * 20.0
    y = np.sin(X[:, 0] / 3.0) * np.cos(X[:, -1] / 4.0) + 0.05 * rng.randn(n)
    Xu = X[::n // m].copy()
    return X, y, Xu


def theta_tensor(d, dtype, variance=0.7, noise=0.05, alpha=1.3, ls=(2.5, 3.0, 4.0, 5.0)):
    return torch.tensor([variance, noise, alpha, *ls[:d]], dtype=dtype)


def oracle_loss_grad(kernel, X, y, Xu, th, jitter):
    d = X.shape[1]
    v, s2, a = (th[i].double().clone().requires_grad_(True) for i in range(3))
    ls = th[3:3 + d].double().clone().requires_grad_(True)
    Xu_t = torch.tensor(Xu, dtype=torch.float64, requires_grad=True)
    loss = vfe_loss(kernel, torch.tensor(X), torch.tensor(y), Xu_t, v, ls, s2, a, jitter)
    loss.backward()
    ga = a.grad if a.grad is not None else torch.zeros(())
    return float(loss.detach()), np.array([float(v.grad), float(s2.grad), float(ga), *ls.grad.tolist()]), Xu_t.grad.numpy()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("d,n,m", [(2, 400, 40), (3, 500, 23), (2, 1500, 300)])
def test_sparse_loss_grad_matches_autograd(eng, kernel, d, n, m):
    from gpim_b200._lib import KERNEL_IDS
    X, y, Xu = problem(n, d, m, 0)
    th = theta_tensor(d, torch.float64)
    ref_loss, ref_g, ref_gxu = oracle_loss_grad(kernel, X, y, Xu, th, 1e-5)
    loss, grad, gxu, info = eng.sparse_loss_grad(KERNEL_IDS[kernel], th.cuda(), torch.tensor(X).cuda(), torch.tensor(y).cuda(),
                                                 torch.tensor(Xu).cuda(), 1e-5)
    assert int(info.item()) == 0
    assert abs(float(loss.item()) - ref_loss) <= 1e-9 * abs(ref_loss)
    g = grad.cpu().numpy()
    if kernel != "RationalQuadratic":
        g[2] = 0.0                                  # the scale mixture is not a parameter of the other kernels
    np.testing.assert_allclose(g, ref_g, rtol=0, atol=1e-8 * np.abs(ref_g).max())
    np.testing.assert_allclose(gxu.cpu().numpy(), ref_gxu, rtol=0, atol=1e-8 * np.abs(ref_gxu).max())


@pytest.mark.parametrize("kernel", KERNELS)
def test_sparse_loss_grad_fp32(eng, kernel):
    from gpim_b200._lib import KERNEL_IDS
    X, y, Xu = problem(800, 2, 80, 1)
    th = theta_tensor(2, torch.float32)
    Xf, yf, Xuf = X.astype(np.float32), y.astype(np.float32), Xu.astype(np.float32)
    ref_loss, ref_g, ref_gxu = oracle_loss_grad(kernel, Xf.astype(np.float64), yf.astype(np.float64), Xuf.astype(np.float64),
                                                th, 1e-4)
    loss, grad, gxu, info = eng.sparse_loss_grad(KERNEL_IDS[kernel], th.cuda(), torch.tensor(Xf).cuda(), torch.tensor(yf).cuda(),
                                                 torch.tensor(Xuf).cuda(), 1e-4)
    assert int(info.item()) == 0
    # the objective is a sum of O(N) terms of both signs (it is -29 here, with N log(noise) / 2 = -1200):
    # the fp32 error is measured against that scale, not against the cancelled total
    assert abs(float(loss.item()) - ref_loss) <= 1e-5 * len(yf) * abs(np.log(float(th[1])))
    g = grad.cpu().numpy().astype(np.float64)
    if kernel != "RationalQuadratic":
        g[2] = 0.0
    np.testing.assert_allclose(g, ref_g, rtol=0, atol=5e-3 * np.abs(ref_g).max())
    np.testing.assert_allclose(gxu.cpu().numpy(), ref_gxu, rtol=0, atol=5e-3 * np.abs(ref_gxu).max())


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("precision,tol_m,tol_s", [("double", 1e-9, 1e-9), ("single", 1e-4, 1e-3)])
@pytest.mark.parametrize("d,n,m,M", [(2, 600, 60, 1000), (3, 700, 150, 333)])
def test_sparse_predict_matches_oracle(eng, kernel, precision, tol_m, tol_s, d, n, m, M):
    from gpim_b200._lib import KERNEL_IDS, torch_dtype
    dtype = torch_dtype(precision)
    X, y, Xu = problem(n, d, m, 2)
    Xs = np.random.RandomState(3).rand(M, d) * 20.0
    Xs[5] = np.nan                                  # NaN rows -> NaN outputs (predict(X_sparse) in EI / POI)
    X, y, Xu, Xs = (np.asarray(torch.tensor(a, dtype=dtype).double()) for a in (X, y, Xu, Xs))
    th = theta_tensor(d, dtype)
    jitter = 1e-5 if precision == "double" else 1e-4
    loc, var = vfe_predict(kernel, torch.tensor(X), torch.tensor(y), torch.tensor(Xu), torch.tensor(Xs), th[0].double(),
                           th[3:].double(), th[1].double(), th[2].double(), jitter)
    ref_m, ref_s = loc.numpy(), var.sqrt().numpy()
    kid = KERNEL_IDS[kernel]
    dev = lambda a: torch.tensor(a, dtype=dtype).cuda()
    fac = eng.sparse_factorize(kid, th.cuda(), dev(X), dev(y), dev(Xu), jitter)
    assert int(fac["info"].item()) == 0
    mean, sd = eng.sparse_predict(kid, th.cuda(), dev(Xu), fac, dev(Xs))
    mean, sd = mean.cpu().numpy(), sd.cpu().numpy()
    assert np.isnan(mean[5]) and np.isnan(sd[5]) and np.isnan(ref_m[5])
    ok = ~np.isnan(ref_m)
    assert np.isfinite(mean[ok]).all() and np.isfinite(sd[ok]).all()
    assert relinf(mean[ok], ref_m[ok]) < tol_m
    assert relinf(sd[ok], ref_s[ok]) < tol_s


def test_sparse_with_every_point_inducing_recovers_the_exact_engine(eng):
    """Property, no oracle in the loop: with Xu = X the VFE posterior IS the exact posterior (up to the jitter)."""
    from gpim_b200._lib import KERNEL_IDS
    X, y, _ = problem(300, 2, 30, 4)
    Xs = np.random.RandomState(5).rand(500, 2) * 20.0
    th = theta_tensor(2, torch.float64).cuda()
    Xd, yd, Xsd = torch.tensor(X).cuda(), torch.tensor(y).cuda(), torch.tensor(Xs).cuda()
    kid = KERNEL_IDS["RBF"]
    m0, s0 = eng.predict(kid, th, Xd, eng.factorize(kid, th, Xd, yd, 1e-9), Xsd)
    m1, s1 = eng.sparse_predict(kid, th, Xd, eng.sparse_factorize(kid, th, Xd, yd, Xd, 1e-9), Xsd)
    assert relinf(m1.cpu(), m0.cpu()) < 1e-5
    assert relinf(s1.cpu(), s0.cpu()) < 1e-5


@pytest.mark.parametrize("kernel,iso", [("RBF", False), ("Matern52", True), ("RationalQuadratic", False)])
def test_sparse_run_matches_oracle(kernel, iso):
    """gpim.reconstructor(sparse=True).run() against the oracle: same inducing-point selection, same trajectory of
    hyper-parameters AND inducing inputs over 25 Adam steps (fp64), same prediction."""
    import gpim
    R = W.dummy_blob(20, 200)
    Xs, Xf = gpim.utils.get_sparse_grid(R), gpim.utils.get_full_grid(R)
    ls = [1.0, 8.0] if iso else [[1.0, 1.0], [8.0, 8.0]]
    kw = dict(kernel=kernel, lengthscale=ls, learning_rate=0.1, iterations=25, seed=1, isotropic=iso)
    ref = SparseOracleGP(O.sparse_grid(R), R, O.full_grid(R), indpoints=15, **kw)
    m0, s0, hp0 = ref.run()
    rec = gpim.reconstructor(Xs, R, Xf, sparse=True, indpoints=15, verbose=0, **kw)
    assert rec.model.Xu.shape == ref.Xu.shape
    m1, s1, hp1 = rec.run()
    assert m1.shape == s1.shape == R.shape
    assert len(hp1["inducing_points"]) == len(hp1["noise"]) == 25
    np.testing.assert_allclose(np.array(hp1["variance"]), np.array(hp0["variance"]), rtol=1e-6)
    np.testing.assert_allclose(np.array(hp1["noise"]), np.array(hp0["noise"]), rtol=1e-6)
    np.testing.assert_allclose(np.array(hp1["lengthscale"]), np.array(hp0["lengthscale"]), rtol=1e-6)
    np.testing.assert_allclose(np.array(hp1["inducing_points"]), np.array(hp0["inducing_points"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(np.array(rec.loss_all), np.array(ref.losses), rtol=1e-8)
    assert relinf(m1, m0) < 1e-6 and relinf(s1, s0) < 1e-6
    # a second train() continues from the trained values with a fresh optimiser (gpr.py:184-185)
    ref.train(iterations=5)
    rec.train(iterations=5)
    np.testing.assert_allclose(np.array(hp1["noise"][-5:]), np.array(ref.noise_all[-5:]), rtol=1e-6)
    assert len(hp1["inducing_points"]) == 30


def test_sparse_default_indpoints_3d_single_precision():
    """sparse=True with the default number of inducing points (len(X) // 10) on a 3-D grid in fp32: runs, finite,
    and agrees with the fp64 oracle fed the trained fp32 values."""
    import gpim
    R = W.hyperspectral((10, 10, 8))
    Xs, Xf = gpim.utils.get_sparse_grid(R), gpim.utils.get_full_grid(R)
    rec = gpim.reconstructor(Xs, R, Xf, kernel="Matern52", lengthscale=[[1., 1., 1.], [10., 10., 10.]], sparse=True,
                             learning_rate=0.1, iterations=30, verbose=0, precision="single", jitter=1e-4)
    n = rec.model.X.shape[0]
    assert rec.model.Xu.shape[0] == len(range(0, n, n // (n // 10)))
    mean, sd, hp = rec.run()
    assert mean.shape == R.shape and np.isfinite(mean).all() and np.isfinite(sd).all() and (sd > 0).all()
    assert rec.loss_all[-1] < rec.loss_all[0]
    th = rec.model.theta.double().cpu()
    loc, var = vfe_predict("Matern52", rec.model.X.double().cpu(), rec.model.y.double().cpu(), rec.model.Xu.double().cpu(),
                           torch.tensor(O.to_rows(Xf)), th[0], th[3:], th[1], th[2], 1e-4)
    assert relinf(mean.ravel(), loc.numpy()) < 1e-4
    assert relinf(sd.ravel(), var.sqrt().numpy()) < 1e-3


def test_boptimizer_with_sparse_surrogate(tmp_path):
    """boptimizer(sparse=True) (boptim.py:167-237 passes sparse / indpoints through): the surrogate swaps X / y in
    place, keeps its inducing inputs and retrains."""
    import gpim
    np.random.seed(0)
    n = 24
    xx, yy = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    truth = np.exp(-((xx - 8) ** 2 + (yy - 15) ** 2) / 30.0)
    y_seed = np.full((n, n), np.nan)
    idx = np.random.randint(0, n, (60, 2))
    y_seed[idx[:, 0], idx[:, 1]] = truth[idx[:, 0], idx[:, 1]]
    X_seed, X_full = gpim.utils.get_sparse_grid(y_seed), gpim.utils.get_full_grid(y_seed)
    bo = gpim.boptimizer(X_seed, y_seed, X_full, lambda ind: truth[ind[0], ind[1]], acquisition_function="ei",
                         exploration_steps=3, gp_iterations=20, sparse=True, indpoints=12, verbose=0,
                         filename=str(tmp_path / "bo"))
    m0 = bo.surrogate_model.model.Xu.shape[0]
    bo.run()
    assert len(bo.indices_all) == 3 and bo.surrogate_model.model.Xu.shape[0] == m0
    assert int((~np.isnan(bo.target_func_vals[-1])).sum()) == int((~np.isnan(y_seed)).sum()) + 3
    assert np.isfinite(bo.gp_predictions[-1][0]).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_sparse_run_matches_committed_training_vectors(kernel, golden_dir):
    """tests/golden/oracle_sparse_train_*.npz: gpim.reconstructor(sparse=True).run() in fp64 against the frozen
    trajectory (hyper-parameters, inducing inputs, loss) and reconstruction -- no oracle in the loop."""
    import os
    import gpim
    g = np.load(os.path.join(golden_dir, f"oracle_sparse_train_{kernel}.npz"))
    R = g["R"]
    rec = gpim.reconstructor(gpim.utils.get_sparse_grid(R), R, gpim.utils.get_full_grid(R), kernel=kernel, sparse=True,
                             indpoints=14, learning_rate=0.1, iterations=15, verbose=0, seed=2)
    mean, sd, hp = rec.run()
    for key in ("variance", "noise", "lengthscale"):
        np.testing.assert_allclose(np.array(hp[key]), g[key], rtol=1e-6)
    np.testing.assert_allclose(np.array(hp["inducing_points"]), g["inducing_points"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(np.array(rec.loss_all), g["loss"], rtol=1e-8)
    assert relinf(mean, g["mean"]) < 1e-6 and relinf(sd, g["sd"]) < 1e-6


def test_sparse_full_size_properties_c2(eng):
    """BASELINE.json configs[1] at FULL size with the notebook setting (256 x 256 spiral, N = 7688, 769 inducing
    points; the oracle's autograd would need minutes here): size-independent properties of the VFE path."""
    from gpim_b200._lib import KERNEL_IDS
    R = W.spiral_scan(256)
    X, y = O.training_rows(O.sparse_grid(R), R)
    N = len(y)
    ft = W.FIXED_THETA
    kid = KERNEL_IDS["RBF"]
    th = [ft["variance"], ft["noise"], 1.0, ft["lengthscale"], ft["lengthscale"]]
    th64 = torch.tensor(th, dtype=torch.float64).cuda()
    X64, y64 = torch.tensor(X).cuda(), torch.tensor(y).cuda()
    Xu64 = X64[::N // (N // 10)].clone()
    assert Xu64.shape[0] == 769
    # (1) the collapsed bound is an upper bound of the exact negative log marginal likelihood (engine's own fp64 path)
    loss, grad, gxu, info = eng.sparse_loss_grad(kid, th64, X64, y64, Xu64, 1e-5)
    nll, _, info2 = eng.nll_grad(kid, th64, X64, y64, 1e-5)
    assert int(info.item()) == 0 and int(info2.item()) == 0
    assert float(loss.item()) >= float(nll.item())
    # (2) the closed-form gradient is the derivative of the loss: central differences along one lengthscale,
    #     the noise and one inducing coordinate
    def loss_at(theta, Xu):
        return float(eng.sparse_loss_grad(kid, theta, X64, y64, Xu, 1e-5)[0].item())
    for p, h in ((3, 1e-4), (1, 1e-6)):
        tp, tm = th64.clone(), th64.clone()
        tp[p] += h
        tm[p] -= h
        fd = (loss_at(tp, Xu64) - loss_at(tm, Xu64)) / (2 * h)
        assert abs(fd - float(grad[p].item())) <= 1e-4 * abs(fd) + 1e-6 * abs(float(loss.item()))
    up, um = Xu64.clone(), Xu64.clone()
    up[100, 1] += 1e-4
    um[100, 1] -= 1e-4
    fd = (loss_at(th64, up) - loss_at(th64, um)) / 2e-4
    assert abs(fd - float(gxu[100, 1].item())) <= 1e-4 * abs(fd) + 1e-6 * abs(float(loss.item()))
    # (3) fp32 prediction against the engine's fp64 one on a sample of the dense grid, north-star tolerances;
    #     variance bounds: noise <= var <= variance + noise
    Xf = torch.tensor(O.to_rows(O.full_grid(R))).cuda()[::4].contiguous()
    m64, s64 = eng.sparse_predict(kid, th64, Xu64, eng.sparse_factorize(kid, th64, X64, y64, Xu64, 1e-5), Xf)
    th32, X32, y32, Xu32 = th64.float(), X64.float(), y64.float(), Xu64.float()
    fac32 = eng.sparse_factorize(kid, th32, X32, y32, Xu32, 1e-5)
    assert int(fac32["info"].item()) == 0
    m32, s32 = eng.sparse_predict(kid, th32, Xu32, fac32, Xf.float())
    assert relinf(m32.cpu(), m64.cpu()) < 1e-4 and relinf(s32.cpu(), s64.cpu()) < 1e-3
    var = (s64 ** 2).cpu().numpy()
    assert var.min() >= ft["noise"] * (1 - 1e-6) and var.max() <= (ft["variance"] + ft["noise"]) * (1 + 1e-6)
    # (4) linearity in y: the factors do not depend on y, so sd is bit-identical and the mean scales
    m2, s2 = eng.sparse_predict(kid, th64, Xu64, eng.sparse_factorize(kid, th64, X64, -3.0 * y64, Xu64, 1e-5), Xf)
    assert torch.equal(s2, s64) and relinf(m2.cpu(), -3.0 * m64.cpu()) < 1e-9


def test_sparse_tensor_core_products_match_the_simt_route(eng):
    """fp32 at C2 size (m = 769, N = 7688): B = Luu^-1 k(Xu, X) and dF/dKuf = T2 B run on the tcgen05 split-fp16 GEMM.
    Loss and gradients (theta and inducing inputs) against the engine's fp64 path; the tensor-core route must be as
    close to it as the all-SIMT fp32 route (GPG_OPT_GEMM_PATH = 1) is."""
    from gpim_b200._lib import KERNEL_IDS, OPT_GEMM_PATH
    R = W.spiral_scan(256)
    X, y = O.training_rows(O.sparse_grid(R), R)
    N = len(y)
    ft = W.FIXED_THETA
    th = [ft["variance"], ft["noise"], 1.0, ft["lengthscale"], ft["lengthscale"]]
    for kernel in ("RBF", "Matern52"):
        kid = KERNEL_IDS[kernel]
        th64 = torch.tensor(th, dtype=torch.float64).cuda()
        X64, y64 = torch.tensor(X).cuda(), torch.tensor(y).cuda()
        Xu64 = X64[::N // (N // 10)].clone()
        l64, g64, x64, _ = eng.sparse_loss_grad(kid, th64, X64, y64, Xu64, 1e-4)
        g64, x64 = g64.cpu().numpy(), x64.cpu().numpy()
        err = {}
        try:
            for path in (1, 0):
                eng.set_option(OPT_GEMM_PATH, path)
                l32, g32, x32, info = eng.sparse_loss_grad(kid, th64.float(), X64.float(), y64.float(), Xu64.float(), 1e-4)
                assert int(info.item()) == 0
                g = g32.double().cpu().numpy()
                sel = [0, 1, 3, 4]                    # variance, noise, lengthscales (no scale mixture here)
                err[path] = (abs(float(l32.item()) - float(l64.item())),
                             float(np.abs(g[sel] - g64[sel]).max() / np.abs(g64[sel]).max()),
                             float(np.abs(x32.double().cpu().numpy() - x64).max() / np.abs(x64).max()))
        finally:
            eng.set_option(OPT_GEMM_PATH, 0)
        print(kernel, "fp32 vs fp64 (loss abs, grad theta rel, grad Xu rel): SIMT", err[1], "tcgen05", err[0])
        scale = 1e-5 * N * abs(np.log(ft["noise"]))
        assert err[0][0] <= max(2 * err[1][0], scale)
        assert err[0][1] <= max(2 * err[1][1], 2e-3)
        assert err[0][2] <= max(2 * err[1][2], 2e-3)


def test_predict_model_sharded_single_rank_equals_predict_sd():
    """gpim_b200.sharded.predict_model_sharded on one rank (no process group) is model.predict_sd for the exact, the
    inducing-point and the GPyTorch-semantics model (the world-size-2 plumbing is tests/test_sharded.py on gloo)."""
    import gpim
    from gpim_b200 import sharded
    R = W.dummy_blob(16, 100)
    Xs, Xf = gpim.utils.get_sparse_grid(R), gpim.utils.get_full_grid(R)
    rows = torch.tensor(O.to_rows(Xf))
    kw = dict(kernel="RBF", learning_rate=0.1, iterations=3, verbose=0)
    for rec in (gpim.reconstructor(Xs, R, Xf, **kw), gpim.reconstructor(Xs, R, Xf, sparse=True, indpoints=12, **kw),
                gpim.skreconstructor(Xs, R, Xf, ski=False, **kw)):
        rec.train()
        m0, s0 = rec.model.predict_sd(rows)
        m1, s1 = sharded.predict_model_sharded(rec.model, rows)
        assert torch.equal(m0, m1) and torch.equal(s0, s1)
