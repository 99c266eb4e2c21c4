"""
``vreconstructor(independent=True)``: GP regression for vector-valued functions with independent output
dimensions, on the B200 engine.

Drop-in for the ``independent=True`` branch of the reference's gpim/gpreg/vgpr.py (:19-354): same constructor
arguments, ``train`` / ``predict`` / ``run`` and ``hyperparams = {"lengthscale"}``.  The reference's model
(``ivgprmodel``, vgpr.py:320-354) is GPyTorch's batch-independent exact GP: T outputs on ONE shared set of inputs,
ONE shared lengthscale (the base kernel's parameter is created before ``kernel.batch_shape`` is overwritten,
vgpr.py:346), per-output ``ScaleKernel`` outputscales and ``ConstantMean`` constants, and a
``MultitaskGaussianLikelihood`` whose noise for output t is ``task_noise_t + noise`` -- i.e. T exact GPs that differ
in (outputscale, noise, constant) and are coupled only through the lengthscale and the global noise.

* ``train``   -> one ``gpg_fit_adam_mt`` call: per Adam iteration T passes of the exact-GP hot path (K assembly,
  Cholesky, inverse, solves, K^-1, fused gradient reduction) and one combined step on the raw parameters, all on the
  device; the lengthscale trajectory comes back in a single copy.
* ``predict`` -> per output ``gpg_factorize`` on y_t - c_t + ``gpg_predict``.  The reference estimates mean and sd
  from ``n_samples`` = 100 draws of the noisy predictive distribution (vgpr.py:218-225); what those estimates converge
  to is the closed form returned here,  mean_t = c_t + k*^T K_t^-1 (y_t - c_t),
  sd_t = sqrt(s_t k** - k*^T K_t^-1 k* + noise_t).  ``predict(..., mc_samples=n)`` reproduces the reference's estimator
  (sample mean and unbiased sample sd of n draws per point) for callers who depend on its Monte-Carlo character.

``independent=False`` (``MultitaskKernel``: a Kronecker-structured joint GP over all outputs) is not on the
accelerated exact-GP path and raises NotImplementedError.
"""
import time
import warnings

import numpy as np
import torch
import torch.nn.functional as F

from .. import gprutils
from .._lib import KERNEL_IDS, get_engine


class _Namespace:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class IndependentMultitaskGPModel:
    """What ``vreconstructor.model`` exposes: the attribute paths the reference reads on its GPyTorch model
    (vgpr.py:183-187): ``covar_module.base_kernel.lengthscale`` (shape (1, n_ls)), ``covar_module.outputscale`` (T),
    ``mean_module.constant`` (T), ``likelihood.task_noises`` / ``.noise``, ``parameters()``, ``train()`` / ``eval()``."""

    def __init__(self, X, Y, kernel_name, input_dim, lengthscale, dtype, isotropic, engine=None):
        if kernel_name not in ("RBF", "Matern52"):
            if kernel_name == "Spectral":
                raise NotImplementedError("the spectral mixture kernel is not on the accelerated exact-GP path")
            print('Select one of the currently available kernels:', '"RBF", "Matern52", "Spectral"')
            raise KeyError(kernel_name)
        self.engine = engine or get_engine()
        self.kernel_id = KERNEL_IDS[kernel_name]
        self.dtype = dtype
        self.input_dim = input_dim
        self.n_ls = 1 if isotropic else input_dim
        self.T = Y.shape[1]
        dev = self.engine.device
        self._X = X.to(dev, dtype).contiguous()
        self._Y = Y.to(dev, dtype).t().contiguous()                   # task-major [T, N]
        if lengthscale is not None:                                   # gpytorch.constraints.Interval
            lo = torch.as_tensor(lengthscale[0], dtype=torch.float64).reshape(-1)
            hi = torch.as_tensor(lengthscale[1], dtype=torch.float64).reshape(-1)
            if lo.numel() != self.n_ls or hi.numel() != self.n_ls:
                raise ValueError("lengthscale bounds do not match the number of kernel lengthscales")
            self.ls_bounds = [float(v) for v in lo] + [float(v) for v in hi]
        else:                                                         # GPyTorch's default: Positive (softplus)
            self.ls_bounds = None
        self._u = torch.zeros(3 * self.T + 1 + self.n_ls, dtype=dtype, device=dev)      # GPyTorch initialises raw = 0
        self._theta = None
        self._factors = None
        self.last_info = 0
        self._refresh_theta()

    # constrained views ---------------------------------------------------------------------
    def _raw(self):
        u, T = self._u.detach().cpu().double(), self.T
        return u[:T], u[T:2 * T], u[2 * T], u[2 * T + 1:3 * T + 1], u[3 * T + 1:]

    def _lengthscale(self, u_ls):
        if self.ls_bounds is None:
            return F.softplus(u_ls)
        lo = torch.tensor(self.ls_bounds[:self.n_ls], dtype=torch.float64)
        hi = torch.tensor(self.ls_bounds[self.n_ls:], dtype=torch.float64)
        return lo + (hi - lo) * torch.sigmoid(u_ls)

    def _refresh_theta(self):
        """theta [T, 3 + d] = {outputscale_t, task_noise_t + noise, constant_t, lengthscale[d]} from the raw parameters."""
        us, utn, un, c, uls = self._raw()
        ls = self._lengthscale(uls)
        if self.n_ls == 1:
            ls = ls.expand(self.input_dim)
        noise = (F.softplus(utn) + 1e-4) + (F.softplus(un) + 1e-4)
        th = torch.cat([F.softplus(us)[:, None], noise[:, None], c[:, None], ls[None, :].expand(self.T, -1)], dim=1)
        self._theta = th.to(self.dtype).to(self.engine.device).contiguous()
        self._factors = None

    @property
    def covar_module(self):
        us, _, _, _, uls = self._raw()
        return _Namespace(base_kernel=_Namespace(lengthscale=self._lengthscale(uls).reshape(1, -1).to(self.dtype)),
                          outputscale=F.softplus(us).to(self.dtype))

    @property
    def mean_module(self):
        return _Namespace(constant=self._raw()[3].to(self.dtype))

    @property
    def likelihood(self):
        _, utn, un, _, _ = self._raw()
        return _Namespace(task_noises=(F.softplus(utn) + 1e-4).to(self.dtype), noise=(F.softplus(un) + 1e-4).reshape(1).to(self.dtype))

    def parameters(self):
        yield self._u

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def cpu(self):
        return self

    # arithmetic ---------------------------------------------------------------------------
    def fit(self, iterations, learning_rate):
        """``iterations`` Adam steps on -mll (fresh optimiser state, warm parameters: vgpr.py:169-170).  Returns the
        trajectory as a CPU tensor [iterations, d + 1] = {lengthscale[d], loss}."""
        traj, theta, info = self.engine.fit_adam_mt(self.kernel_id, self._X, self._Y, 0.0, self._u, self.ls_bounds,
                                                    self.n_ls, iterations, learning_rate)
        traj_host = traj.cpu()
        self.last_info = int(info.item())
        if iterations > 0:
            self._theta = theta
        self._factors = None
        if self.last_info != 0:
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not "
                f"positive-definite (the leading minor of order {self.last_info} is not positive-definite).")
        return traj_host

    def predict_sd(self, Xnew):
        """Closed-form (mean, sd) of the noisy predictive distribution at the rows of Xnew: [M, T] each."""
        eng = self.engine
        Xnew = Xnew.to(eng.device, self.dtype).contiguous()
        M = Xnew.shape[0]
        mean = torch.empty(self.T, M, dtype=self.dtype, device=eng.device)
        sd = torch.empty(self.T, M, dtype=self.dtype, device=eng.device)
        infos = []
        for t in range(self.T):
            th = self._theta[t]
            yc = self._Y[t] - th[2]                                   # ConstantMean: the GP is on y_t - c_t
            fac = eng.factorize(self.kernel_id, th, self._X, yc, 0.0)
            eng.predict(self.kernel_id, th, self._X, fac, Xnew, mean=mean[t], sd=sd[t])
            mean[t] += th[2]
            infos.append(fac["info"])
        bad = [int(i.item()) for i in infos]                          # one sync, after everything has been enqueued
        if any(bad):
            self.last_info = next(b for b in bad if b)
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not "
                f"positive-definite (the leading minor of order {self.last_info} is not positive-definite).")
        return mean.t().contiguous(), sd.t().contiguous()


class vreconstructor:
    """
    Multi-output GP regression for vector-valued 2D/3D/4D functions with INDEPENDENT output dimensions.

    Args (as the reference, vgpr.py:19-84):
        X (ndarray): grid indices (c, N, M[, L[, K]])
        y (ndarray): observations (N, M[, L[, K]], d) with d output dimensions (NaN rows are dropped)
        Xtest (ndarray): "test" grid indices
        kernel (str): 'RBF' or 'Matern52'
        lengthscale (list of two lists): lower / upper bounds of the kernel lengthscale(s), or None
        independent (bool): must be True here (the correlated-output model is not on the accelerated path)
        learning_rate (float), iterations (int), use_gpu, verbose (int), seed (int)
        **isotropic, **precision, **num_batches (accepted; the engine tiles the test set itself), **maxroot (ignored:
          the predictive variance is exact, no Lanczos decomposition)
    """

    def __init__(self, X, y, Xtest=None, kernel='RBF', lengthscale=None, independent=False, learning_rate=.1,
                 iterations=50, use_gpu=1, verbose=1, seed=0, **kwargs):
        if not independent:
            raise NotImplementedError(
                "gpim.vreconstructor(independent=False) (GPyTorch MultitaskKernel: correlated outputs) is outside the "
                "accelerated exact-GP path of this engine; independent=True is supported")
        self.precision = kwargs.get("precision", "double")
        dtype = torch.float32 if self.precision == "single" else torch.float64
        engine = get_engine()
        torch.manual_seed(seed)
        input_dim = np.ndim(y) - 1
        Xr, yr = gprutils.prepare_training_data(X, y, vector_valued=True, precision=self.precision)
        num_tasks = yr.shape[-1]
        self.fulldims = (Xtest.shape[1:] if Xtest is not None else np.shape(X)[1:]) + (num_tasks,)
        self.X, self.y = Xr, yr
        self.Xtest = gprutils.prepare_test_data(Xtest, precision=self.precision) if Xtest is not None else None
        self.model = IndependentMultitaskGPModel(Xr, yr, kernel, input_dim, lengthscale, dtype,
                                                 bool(kwargs.get("isotropic")), engine=engine)
        self.likelihood = self.model.likelihood
        self.iterations = iterations
        self.num_batches = kwargs.get("num_batches", 1)
        self.learning_rate = learning_rate
        self.independent = independent
        self.lscales, self.loss_all = [], []
        self.hyperparams = {"lengthscale": self.lscales}
        self.verbose = verbose

    def train(self, **kwargs):
        """Trains the model: **learning_rate, **iterations, **verbose as in vgpr.py:142-196."""
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
        d, n_ls = self.model.input_dim, self.model.n_ls
        for i, row in enumerate(traj):
            self.lscales.append(row[:d].tolist() if n_ls > 1 else [float(row[0])])
            self.loss_all.append(float(row[d]))
            if self.verbose == 2 and (i % 10 == 0 or i == self.iterations - 1):
                print('iter: {} ...'.format(i), 'loss: {} ...'.format(np.around(self.loss_all[-1], 4)),
                      'length: {} ...'.format(np.around(self.lscales[-1], 4)))
        if self.verbose:
            print('training completed in {} s'.format(np.round(time.time() - start_time, 2)))
            if self.lscales:
                print('Final parameter values:\n', 'lengthscale: {}'.format(np.around(self.lscales[-1], 4)))
        return

    def predict(self, Xtest=None, **kwargs):
        """(mean, sd) shaped like the test grid + (d,); vgpr.py:198-264.  **mc_samples=n: the reference's estimator
        (mean and unbiased sd of n draws of the noisy predictive distribution per point) instead of its limit."""
        if Xtest is None and self.Xtest is None:
            warnings.warn("No test data provided. Using training data for prediction", UserWarning)
            self.Xtest = self.X
        elif Xtest is not None:
            self.Xtest = gprutils.prepare_test_data(Xtest, precision=self.precision)
            self.fulldims = Xtest.shape[1:] + (self.y.shape[-1],)
        if kwargs.get("verbose") is not None:
            self.verbose = kwargs.get("verbose")
        if kwargs.get("num_batches") is not None:
            self.num_batches = kwargs.get("num_batches")
        if self.verbose:
            print('Calculating predictive mean and uncertainty...')
        mean_d, sd_d = self.model.predict_sd(self.Xtest)
        n = kwargs.get("mc_samples")
        if n:
            draws = mean_d[None] + sd_d[None] * torch.randn((int(n),) + tuple(mean_d.shape), dtype=mean_d.dtype, device=mean_d.device)
            mean_d, sd_d = draws.mean(dim=0), draws.var(dim=0).sqrt()
        both = torch.stack((mean_d, sd_d)).cpu().numpy()
        mean, sd = both[0], both[1]
        if mean.size == int(np.prod(self.fulldims)):
            mean, sd = mean.reshape(self.fulldims), sd.reshape(self.fulldims)
        if self.verbose:
            print("\nDone")
        return mean, sd

    def run(self):
        """train() then predict(); returns (mean, sd, hyperparams) as vgpr.py:266-278."""
        self.train()
        mean, sd = self.predict()
        return mean, sd, self.hyperparams
