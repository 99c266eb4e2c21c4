"""
``skreconstructor`` with the exact-GP semantics of GPyTorch on the B200 engine.

Drop-in for the ``ski=False`` branch of the reference's gpim/gpreg/skgpr.py (:22-326) with the 'RBF' and 'Matern52'
kernels: gpytorch.models.ExactGP with ConstantMean, ScaleKernel(RBFKernel | MaternKernel) under an
Interval lengthscale constraint (gpytorch_kernels.py:55-69) and GaussianLikelihood, trained with Adam on
-ExactMarginalLogLikelihood (skgpr.py:186-196).  Same constructor arguments, ``train`` / ``predict`` / ``run`` and
``hyperparams`` = {"lengthscale", "noise"} as there.

* ``train``   -> one ``gpg_fit_adam_sk`` call (the Adam loop in GPyTorch's raw parametrisation on the device);
* ``predict`` -> ``gpg_factorize`` on y - constant + ``gpg_predict``; the constant mean is added back.  The predictive
  variance is EXACT: the reference asks GPyTorch for ``fast_pred_var`` (a rank-``maxroot`` Lanczos approximation,
  skgpr.py:285) and GPyTorch switches to conjugate gradients beyond 800 training points; both approximate what is
  computed here.

Not on this path (NotImplementedError): ``ski=True`` (KISS-GP grid interpolation, the reference's default) and the
'Spectral' mixture kernel.
"""
import time
import warnings

import numpy as np
import torch
import torch.nn.functional as F

from .. import gprutils
from .._lib import KERNEL_IDS, get_engine
from .gpr import ExactGPModel


class _Namespace:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class SKKernelState:
    """Host mirror of the GPyTorch parameters: raw {outputscale, noise, mean constant, lengthscale[n_ls]}, all
    initialised to 0 as GPyTorch does (no prior draws on this branch)."""

    def __init__(self, name, input_dim, lengthscale, dtype, isotropic):
        if name not in ("RBF", "Matern52"):
            if name == "Spectral":
                raise NotImplementedError("the spectral mixture kernel is not on the accelerated exact-GP path")
            print('Select one of the currently available kernels:', '"RBF", "Matern52", "Spectral"')
            raise KeyError(name)
        self.name, self.kernel_id, self.input_dim, self.dtype = name, KERNEL_IDS[name], input_dim, dtype
        self.isotropic = bool(isotropic)
        self.n_ls = 1 if self.isotropic else input_dim
        self.ls_lo = torch.as_tensor(lengthscale[0], dtype=dtype).reshape(-1)
        self.ls_hi = torch.as_tensor(lengthscale[1], dtype=dtype).reshape(-1)
        if self.ls_lo.numel() != self.n_ls or self.ls_hi.numel() != self.n_ls:
            raise ValueError("lengthscale bounds do not match the number of kernel lengthscales")
        self.u = torch.zeros(3 + self.n_ls, dtype=dtype)

    # constrained views (CPU tensors) ---------------------------------------------------
    @property
    def outputscale(self):
        return F.softplus(self.u[0])

    @property
    def noise(self):
        return F.softplus(self.u[1]) + 1e-4

    @property
    def constant(self):
        return self.u[2]

    @property
    def lengthscale(self):
        """Shape (1, n_ls), like gpytorch's kernel.lengthscale."""
        return (self.ls_lo + (self.ls_hi - self.ls_lo) * torch.sigmoid(self.u[3:])).reshape(1, -1)

    # packing for the C ABI -------------------------------------------------------------
    def pack_u(self):
        return self.u.clone()

    def unpack_u(self, u):
        self.u = u.detach().cpu().to(self.dtype).clone()

    def pack_theta(self):
        ls = self.lengthscale.reshape(-1)
        if self.isotropic:
            ls = ls.expand(self.input_dim)
        return torch.cat([self.outputscale.reshape(1), self.noise.reshape(1), self.constant.reshape(1), ls]).to(self.dtype)

    def bounds(self):
        return [0.0, 1.0] + [float(v) for v in self.ls_lo] + [float(v) for v in self.ls_hi]


class SKExactGPModel(ExactGPModel):
    """What ``skreconstructor.model`` exposes: the attribute paths of the GPyTorch model the reference reads
    (skgpr.py:197-222,371): ``covar_module.base_kernel.lengthscale``, ``covar_module.outputscale``,
    ``likelihood.noise_covar.noise``, ``mean_module.constant``, ``parameters()``, ``train()`` / ``eval()``."""

    def __init__(self, X, y, kernel, engine=None):
        super().__init__(X, y, kernel, jitter=0.0, engine=engine)

    @property
    def noise(self):
        return self.kernel.noise

    @property
    def covar_module(self):
        k = self.kernel
        return _Namespace(base_kernel=_Namespace(lengthscale=k.lengthscale), outputscale=k.outputscale)

    @property
    def likelihood(self):
        n = self.kernel.noise.reshape(1)
        return _Namespace(noise_covar=_Namespace(noise=n), noise=n)

    @property
    def mean_module(self):
        return _Namespace(constant=self.kernel.constant)

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def fit(self, iterations, learning_rate):
        k = self.kernel
        if self._X.shape[0] != self._y.shape[0]:
            raise ValueError("X and y have different numbers of rows")
        traj, theta, info = self.engine.fit_adam(k.kernel_id, self._X, self._y, 0.0, self._u, k.bounds(), k.n_ls,
                                                 iterations, learning_rate, gpytorch_params=True)
        traj_host = traj.cpu()
        self.last_info = int(info.item())
        if iterations > 0:
            self._theta = theta
            k.unpack_u(self._u)
        self._factor = None
        if self.last_info != 0:
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not "
                f"positive-definite (the leading minor of order {self.last_info} is not positive-definite).")
        return traj_host

    def factor(self, check=True):
        fresh = self._factor is None
        if fresh:                                    # ConstantMean: the GP is on y - constant
            yc = self._y - self._theta[2]
            self._factor = self.engine.factorize(self.kernel.kernel_id, self._theta, self._X, yc, 0.0)
        if fresh and check:
            self._raise_if_not_pd(self._factor)
        return self._factor, fresh

    def mean_shift(self):
        """The constant mean (device scalar) that goes back onto a prediction made on y - constant."""
        return self._theta[2]

    def predict_sd(self, Xnew):
        mean, sd = super().predict_sd(Xnew)
        return mean + self.mean_shift(), sd


class skreconstructor:
    """
    GP regression with GPyTorch's exact-GP semantics (``ski=False``) for 2D-4D grids.

    Args:
        X (ndarray): grid indices (c, N, M[, L[, K]]); missing points are NaN
        y (ndarray): observations (N, M[, L[, K]]); missing points are NaN
        Xtest (ndarray): "test" grid indices
        kernel (str): 'RBF' or 'Matern52'
        lengthscale (list): [lo, hi] bounds (isotropic) or [[lo]*c, [hi]*c]; default [0, mean(y.shape) / 2]
        ski (bool): grid-interpolation kernel; only ``False`` is implemented here (the reference defaults to True)
        learning_rate (float), iterations (int): Adam settings (defaults 0.1 / 50 as skgpr.py:87-88)
        use_gpu, verbose, seed: as the reference (the engine always computes on the GPU; no random numbers are drawn)
        **isotropic, **precision ('single' | 'double'), **num_batches (accepted; the engine tiles the test grid itself)
    """

    def __init__(self, X, y, Xtest=None, kernel='RBF', lengthscale=None, ski=True, learning_rate=.1, iterations=50,
                 use_gpu=1, verbose=1, seed=0, **kwargs):
        self.precision = kwargs.get("precision", "double")
        npfloat_ = np.float32 if self.precision == "single" else np.float64
        dtype = torch.float32 if self.precision == "single" else torch.float64
        if ski and kernel != "Spectral":
            raise NotImplementedError(
                "ski=True (KISS-GP grid interpolation, skgpr.py:437-441) is not on the accelerated exact-GP path; "
                "pass ski=False")
        engine = get_engine()
        torch.manual_seed(seed)
        input_dim = np.ndim(y)
        self.fulldims = Xtest.shape[1:] if Xtest is not None else X.shape[1:]
        self.X, self.y = gprutils.prepare_training_data(X, y, precision=self.precision)
        self.Xtest = gprutils.prepare_test_data(Xtest, precision=self.precision) if Xtest is not None else None
        self.do_ski = False
        isotropic = kwargs.get("isotropic")
        if lengthscale is None and not isotropic:
            lmean = npfloat_(np.mean(y.shape) / 2)
            lengthscale = [[0. for _ in range(input_dim)], [lmean for _ in range(input_dim)]]
        elif lengthscale is None and isotropic:
            lengthscale = [0., npfloat_(np.mean(y.shape) / 2)]
        kern = SKKernelState(kernel, input_dim, lengthscale, dtype, isotropic)
        self.model = SKExactGPModel(self.X, self.y, kern, engine=engine)
        self.likelihood = self.model.likelihood
        self.iterations = iterations
        self.num_batches = kwargs.get("num_batches", 1)
        self.learning_rate = learning_rate
        self.noise_all, self.lscales, self.loss_all = [], [], []
        self.hyperparams = {"lengthscale": self.lscales, "noise": self.noise_all}
        self.verbose = verbose

    def train(self, **kwargs):
        """Trains the model: **learning_rate, **iterations, **verbose as in skgpr.py:170-272."""
        if kwargs.get("learning_rate") is not None:
            self.learning_rate = kwargs.get("learning_rate")
        if kwargs.get("iterations") is not None:
            self.iterations = kwargs.get("iterations")
        if kwargs.get("verbose") is not None:
            self.verbose = kwargs.get("verbose")
        if self.verbose:
            print('Model training...')
        start_time = time.time()
        traj = self.model.fit(self.iterations, self.learning_rate).double().numpy()
        d = self.model.X.shape[1]
        n_ls = self.model.kernel.n_ls
        for i, row in enumerate(traj):
            self.lscales.append(row[3:3 + n_ls].tolist())          # lengthscale.tolist()[0], skgpr.py:204-206
            self.noise_all.append(float(row[1]))
            self.loss_all.append(float(row[3 + d]))
            if self.verbose == 2 and (i % 10 == 0 or i == self.iterations - 1):
                print('iter: {} ... loss: {} ... length: {} ... noise: {} ...'.format(
                    i, np.around(self.loss_all[-1], 4), np.around(self.lscales[-1], 4), np.around(self.noise_all[-1], 7)))
        if self.verbose:
            print('training completed in {} s'.format(np.round(time.time() - start_time, 2)))
            if self.lscales:
                print('Final parameter values:\n', 'lengthscale: {}, noise: {}'.format(
                    np.around(self.lscales[-1], 4), np.around(self.noise_all[-1], 7)))
        return

    def predict(self, Xtest=None, **kwargs):
        """Predictive mean and standard deviation (numpy, shaped like the test grid); skgpr.py:274-326."""
        if Xtest is None and self.Xtest is None:
            warnings.warn("No test data provided. Using training data for prediction", UserWarning)
            self.Xtest = self.X
        elif Xtest is not None:
            self.Xtest = gprutils.prepare_test_data(Xtest, precision=self.precision)
            self.fulldims = Xtest.shape[1:]
        if kwargs.get("verbose") is not None:
            self.verbose = kwargs.get("verbose")
        if kwargs.get("num_batches") is not None:
            self.num_batches = kwargs.get("num_batches")
        if kwargs.get("max_root") is not None:
            self.max_root = kwargs.get("max_root")
        if self.verbose:
            print('Calculating predictive mean and uncertainty...')
        mean_d, sd_d = self.model.predict_sd(self.Xtest)
        self._last_pred_device = (mean_d, sd_d)
        both = torch.stack((mean_d, sd_d)).cpu().numpy()
        mean, sd = both[0], both[1]
        if mean.size == int(np.prod(self.fulldims)):
            mean, sd = mean.reshape(self.fulldims), sd.reshape(self.fulldims)
        if self.verbose:
            print("\nDone")
        return mean, sd

    def run(self):
        """train() then predict(); returns (mean, sd, hyperparams) as skgpr.py:328-340."""
        self.train()
        mean, sd = self.predict()
        return mean, sd, self.hyperparams

    def step(self, acquisition_function=None, batch_size=100, batch_update=False, lscale=None, **kwargs):
        """The reference's step() (skgpr.py:342-397) ends in a call to the non-existent gprutils.acquisition."""
        raise NotImplementedError("skreconstructor.step is dead code in the reference (skgpr.py:394); use gpim.boptimizer")
