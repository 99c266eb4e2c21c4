"""Pins oracle/gp_oracle.py against the reference's own golden vectors (CPU only).

* test/test_boptim.py:42-58 + test/test_data/test_{ei,poi,cb}.npy  (copied byte-for-byte to
  tests/golden/ref_test_*.npy by tests/golden/make_golden.py)
* stored outputs of examples/notebooks/GP_based_exploration_exploitation.ipynb cell 13
* shape / no-NaN checks of test/test_gpreg.py:24-36
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O


def _boptim_setup():
    # test/test_boptim.py:17-39
    def trial_func(idx, x0=5, y0=10, fwhm=4.5):
        return np.exp(-4 * np.log(2) * ((idx[0] - x0) ** 2 + (idx[1] - y0) ** 2) / fwhm ** 2)
    np.random.seed(0)
    x = np.arange(0, 25, 1.0)
    y = x[:, np.newaxis]
    Z = trial_func([y, x])
    idx = np.random.randint(0, Z.shape[0], size=(2, 5))
    Zs = np.ones_like(Z) * np.nan
    Zs[idx[0], idx[1]] = Z[idx[0], idx[1]]
    return trial_func, Zs


@pytest.mark.parametrize("acqf", ["ei", "poi", "cb"])
def test_oracle_reproduces_reference_bo_golden(acqf, golden_dir):
    torch.set_num_threads(1)          # N <= 25: threading only adds overhead
    trial_func, Zs = _boptim_setup()
    out = O.bo_run(O.sparse_grid(Zs), Zs, O.full_grid(Zs), trial_func, acqf, exploration_steps=20)
    expected = np.load(os.path.join(golden_dir, f"ref_test_{acqf}.npy"))
    got = out["target_func_vals"][-1]
    assert np.array_equal(np.isnan(got), np.isnan(expected))
    np.testing.assert_allclose(got, expected)


def test_oracle_reproduces_notebook_known_answers(golden_dir):
    torch.set_num_threads(1)
    kat = json.load(open(os.path.join(golden_dir, "notebook_kat_ei.json")))["trainings"]

    def trial_func(idx):
        def f(x0, y0, a, b, fwhm):
            return np.exp(-4 * np.log(2) * (a * (idx[0] - x0) ** 2 + b * (idx[1] - y0) ** 2) / fwhm ** 2)
        return f(5, 10, 1, 1, 4.5) + f(10, 8, 0.75, 1.5, 7) + f(18, 18, 1, 1.5, 10)
    x, y = np.meshgrid(np.linspace(0, 24, 25), np.linspace(0, 24, 25), indexing="ij")
    Z = trial_func([x, y])
    np.random.seed(42)
    Zs = np.ones_like(Z) * np.nan
    for i in np.random.randint(0, Z.shape[0], size=(5, 2)):
        Zs[tuple(i)] = trial_func(i)
    seen = []
    O.bo_run(O.sparse_grid(Zs), Zs, O.full_grid(Zs), trial_func, "ei", exploration_steps=3,
             on_train=lambda gp: seen.append((gp.amp_all[-1], gp.lscales[-1], gp.noise_all[-1])))
    assert len(seen) == 4
    for (amp, ls, noise), ref in zip(seen, kat):
        assert np.around(amp, 4) == ref["amp"]
        np.testing.assert_array_equal(np.around(ls, 4), ref["lengthscale"])
        assert float(np.around(noise, 7)) == ref["noise"]


def _dummy_data():
    # test/test_gpreg.py:9-21
    np.random.seed(0)
    xx, yy = np.meshgrid(np.arange(0, 100, 5), np.arange(0, 100, 5))
    Z = np.exp(-((xx - 25) ** 2 + (yy - 50) ** 2) / 300)
    for _ in range(200):
        Z[np.random.randint(Z.shape[0]), np.random.randint(Z.shape[1])] = np.nan
    return Z


@pytest.mark.parametrize("kernel", ["RBF", "Matern52", "RationalQuadratic"])
def test_oracle_gpr_2d_sanity(kernel):
    R = _dummy_data()
    mean, sd, hp = O.OracleGP(O.sparse_grid(R), R, O.full_grid(R), kernel=kernel,
                              learning_rate=0.1, iterations=2).run()
    assert mean.shape == sd.shape == R.shape
    assert not np.isnan(mean).any() and not np.isnan(sd).any()
    assert len(hp["lengthscale"]) == 2
    # values are recorded AFTER the Adam step: noise_0 = exp(0 +- lr)  (GP_BEPFM.ipynb:590)
    assert hp["noise"][0] == pytest.approx(np.exp(0.1)) or hp["noise"][0] == pytest.approx(np.exp(-0.1))


def test_oracle_prior_draw_order():
    """variance first, then lengthscale (pyro_kernels.py:81-94); SURVEY appendix A values."""
    _, Zs = _boptim_setup()
    gp = O.OracleGP(O.sparse_grid(Zs), Zs, O.full_grid(Zs))
    v, l, n, _ = gp.theta()
    assert v.item() == pytest.approx(9.70053, abs=1e-5)
    np.testing.assert_allclose(l.detach().numpy(), [8.84775, 5.74229], atol=1e-5)
    assert n.item() == pytest.approx(1.0)


def test_oracle_unknown_kernel_raises():
    _, Zs = _boptim_setup()
    with pytest.raises(KeyError):
        O.OracleGP(O.sparse_grid(Zs), Zs, kernel="Periodic")


# ---------------------------------------------------------------------------------------------
# oracle-generated fixtures (tests/golden/make_oracle_vectors.py): the oracle must keep reproducing them
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["RBF", "Matern52", "RationalQuadratic"])
@pytest.mark.parametrize("case", ["predict2d", "predict3d"])
def test_oracle_reproduces_its_committed_predict_vectors(kernel, case, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, f"oracle_{case}_{kernel}.npz"))
    v, noise, mix, jitter = g["theta"]
    mean, sd, _ = O.predict_fixed_theta(kernel, g["X"], g["y"], g["Xs"], float(v), g["lengthscale"], float(noise),
                                        jitter=float(jitter), scale_mixture=float(mix))
    np.testing.assert_allclose(mean, g["mean"], rtol=0, atol=1e-10 * np.abs(g["mean"]).max())
    np.testing.assert_allclose(sd, g["sd"], rtol=1e-10)


@pytest.mark.parametrize("kernel", ["RBF", "Matern52", "RationalQuadratic"])
def test_oracle_reproduces_its_committed_training_vectors(kernel, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, f"oracle_train_{kernel}.npz"))
    R = g["R"]
    ora = O.OracleGP(O.sparse_grid(R), R, O.full_grid(R), kernel=kernel, learning_rate=0.1, iterations=20, seed=2)
    mean, sd, hp = ora.run()
    np.testing.assert_allclose(np.array(hp["variance"]), g["variance"], rtol=1e-8)
    np.testing.assert_allclose(np.array(hp["noise"]), g["noise"], rtol=1e-8)
    np.testing.assert_allclose(np.array(hp["lengthscale"]), g["lengthscale"], rtol=1e-8)
    np.testing.assert_allclose(mean, g["mean"], rtol=0, atol=1e-8 * np.abs(g["mean"]).max())
    np.testing.assert_allclose(sd, g["sd"], rtol=1e-8)
