"""
Gaussian-process reconstruction of sparse 2D images and 3D/4D hyperspectral grids on one B200.

Drop-in for the reference's gpim/gpreg/gpr.py ``reconstructor`` (:22-283): same constructor
arguments, same ``train`` / ``predict`` / ``run`` methods and return values.  Underneath, the
Pyro ``GPRegression`` object is replaced by :class:`ExactGPModel`, whose arithmetic is the CUDA
engine behind libgpgrid.so:

* ``train``  -> one ``gpg_fit_adam`` call: the whole Adam loop runs on the device, the
  hyper-parameter trajectory comes back in a single copy (the reference syncs three times per
  iteration, gpr.py:195-197);
* ``predict`` -> ``gpg_factorize`` (cached per (X, y, theta)) + ``gpg_predict`` tiled over the test
  grid (the reference refactorises on every call and materialises the N x M cross-kernel three
  times, gpr.py:248).

``sparse=True`` (gpr.py:145-155: pyro's SparseGPRegression, VFE) runs on :class:`SparseGPModel`:
``gpg_sparse_fit_adam`` trains the hyper-parameters AND the inducing inputs on the device,
``gpg_sparse_factorize`` + ``gpg_sparse_predict`` replace SparseGPRegression.forward.
"""
import time
import warnings

import numpy as np
import torch

from .. import gprutils
from .._lib import get_engine
from ..kernels import gp_kernels


class ExactGPModel:
    """What ``reconstructor.model`` exposes: the members of pyro's GPRegression the reference's
    host code touches: assignable ``X`` / ``y``, ``kernel``, ``noise``,
    ``parameters()``, ``cuda()`` / ``cpu()`` and ``__call__(Xnew, full_cov, noiseless)``."""

    def __init__(self, X, y, kernel, jitter=1e-6, engine=None):
        self.engine = engine or get_engine()
        self.kernel = kernel
        self.jitter = float(jitter)
        self._X = X.to(self.engine.device).contiguous()
        self._y = y.to(self.engine.device).contiguous()
        self._u = kernel.pack_u().to(self.engine.device)       # unconstrained, Adam steps on this
        self._theta = kernel.pack_theta().to(self.engine.device)
        self._factor = None
        self.last_info = 0

    # data are swapped in place by boptimizer.update_posterior (boptim.py:248-249) ------------
    @property
    def X(self):
        return self._X

    @X.setter
    def X(self, value):
        self._X = value.to(self.engine.device, self.kernel.dtype).contiguous()
        self._factor = None

    @property
    def y(self):
        return self._y

    @y.setter
    def y(self, value):
        self._y = value.to(self.engine.device, self.kernel.dtype).contiguous()
        self._factor = None

    @property
    def noise(self):
        return self.kernel.u_noise.exp()

    @property
    def theta(self):
        """Constrained {variance, noise, scale_mixture, lengthscale[d]} on the device."""
        return self._theta

    def set_theta(self, variance, lengthscale, noise, scale_mixture=1.0):
        """Fix the hyper-parameters (constrained values) without training."""
        d = self._X.shape[1]
        ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (d,))
        th = torch.tensor([variance, noise, scale_mixture, *ls], dtype=self.kernel.dtype)
        self._theta = th.to(self.engine.device)
        self._factor = None

    def load_unconstrained(self, u):
        """Restore the unconstrained hyper-parameters {variance, noise, scale_mixture, lengthscale[n_ls]} (a
        checkpoint of boptimizer.save_results) and the constrained values derived from them."""
        k = self.kernel
        k.unpack_u(torch.as_tensor(np.asarray(u), dtype=k.dtype))
        self._u = k.pack_u().to(self.engine.device)
        self._theta = k.pack_theta().to(self.engine.device)
        self._factor = None

    def parameters(self):
        yield self._u

    def cuda(self):
        return self

    def cpu(self):
        return self

    # training ----------------------------------------------------------------------------
    def fit(self, iterations, learning_rate):
        """``iterations`` Adam steps on -log p(y | theta) (fresh optimizer state, warm theta:
        gpr.py:184-185).  Returns the trajectory as a CPU tensor [iterations, 4 + d]:
        {variance, noise, scale_mixture, lengthscale[d], loss} recorded after each step."""
        k = self.kernel
        if self._X.shape[0] != self._y.shape[0]:
            raise ValueError("X and y have different numbers of rows")
        traj, theta, info = self.engine.fit_adam(k.kernel_id, self._X, self._y, self.jitter, self._u, k.bounds(),
                                                 k.n_ls, iterations, learning_rate)
        traj_host = traj.cpu()                          # the one device->host copy of train()
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

    # prediction --------------------------------------------------------------------------
    def _raise_if_not_pd(self, fac):
        self.last_info = int(fac["info"].item())
        if self.last_info != 0:
            self._factor = None
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not "
                f"positive-definite (the leading minor of order {self.last_info} is not positive-definite).")

    def factor(self, check=True):
        """Factor cache for the current (X, y, theta).  check=False defers the (synchronising) pivot check to the
        caller, so that the prediction kernels are enqueued right behind the factorisation."""
        fresh = self._factor is None
        if fresh:
            self._factor = self.engine.factorize(self.kernel.kernel_id, self._theta, self._X, self._y, self.jitter)
        if fresh and check:
            self._raise_if_not_pd(self._factor)
        return self._factor, fresh

    def __call__(self, Xnew, full_cov=False, noiseless=False):
        """(loc, var) at the rows of Xnew, as GPRegression.forward(full_cov=False)."""
        if full_cov:
            raise NotImplementedError("only the diagonal predictive variance is on the hot path")
        if Xnew.dim() != 2 or Xnew.shape[1] != self._X.shape[1]:
            raise ValueError("Train data and test data should have the same shape of features")
        Xnew = Xnew.to(self.engine.device, self.kernel.dtype).contiguous()
        mean, sd = self.predict_sd(Xnew)
        var = sd * sd
        if noiseless:
            var = var - self._theta[1]
        return mean, var

    def predict_sd(self, Xnew):
        Xnew = Xnew.to(self.engine.device, self.kernel.dtype).contiguous()
        fac, fresh = self.factor(check=False)
        out = self.engine.predict(self.kernel.kernel_id, self._theta, self._X, fac, Xnew)
        if fresh:
            self._raise_if_not_pd(fac)               # one sync, after everything has been enqueued
        return out


class SparseGPModel(ExactGPModel):
    """What ``reconstructor.model`` is for ``sparse=True``: the members of pyro's SparseGPRegression (default
    approximation "VFE") the reference touches -- everything ExactGPModel has plus ``Xu`` (gpr.py:199)."""

    def __init__(self, X, y, kernel, Xu, jitter=1e-6, engine=None):
        super().__init__(X, y, kernel, jitter=jitter, engine=engine)
        self._Xu = Xu.to(self.engine.device, kernel.dtype).contiguous().clone()
        self.approx = "VFE"

    @property
    def Xu(self):
        return self._Xu

    @Xu.setter
    def Xu(self, value):
        self._Xu = value.to(self.engine.device, self.kernel.dtype).contiguous().clone()
        self._factor = None

    def parameters(self):
        yield self._u
        yield self._Xu

    def fit(self, iterations, learning_rate, record_xu=True):
        """As ExactGPModel.fit, on the VFE objective, with the inducing inputs trained alongside.  Returns
        (trajectory [iterations, 4 + d], Xu after every step [iterations, m, d]) as CPU tensors."""
        k = self.kernel
        if self._X.shape[0] != self._y.shape[0]:
            raise ValueError("X and y have different numbers of rows")
        traj, xu_traj, theta, info = self.engine.sparse_fit_adam(
            k.kernel_id, self._X, self._y, self._Xu, self.jitter, self._u, k.bounds(), k.n_ls, iterations,
            learning_rate, record_xu=record_xu)
        traj_host = traj.cpu()
        xu_host = xu_traj.cpu() if xu_traj is not None else None
        self.last_info = int(info.item())
        if iterations > 0:
            self._theta = theta
            k.unpack_u(self._u)
        self._factor = None
        if self.last_info != 0:
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not "
                f"positive-definite (the leading minor of order {self.last_info} is not positive-definite).")
        return traj_host, xu_host

    def factor(self, check=True):
        fresh = self._factor is None
        if fresh:
            self._factor = self.engine.sparse_factorize(self.kernel.kernel_id, self._theta, self._X, self._y, self._Xu,
                                                        self.jitter)
        if fresh and check:
            self._raise_if_not_pd(self._factor)
        return self._factor, fresh

    def predict_sd(self, Xnew):
        Xnew = Xnew.to(self.engine.device, self.kernel.dtype).contiguous()
        fac, fresh = self.factor(check=False)
        out = self.engine.sparse_predict(self.kernel.kernel_id, self._theta, self._Xu, fac, Xnew)
        if fresh:
            self._raise_if_not_pd(fac)
        return out


class reconstructor:
    """
    GP-based reconstruction of sparse 2D images and 3D spectroscopic datasets.

    Args:
        X (ndarray): grid indices, shape (c, N, M) or (c, N, M, L); missing points are NaN
        y (ndarray): observations, shape (N, M) or (N, M, L); missing points are NaN
        Xtest (ndarray): "test" grid indices for prediction, shape (c, N', M'[, L'])
        kernel (str): 'RBF', 'Matern52' or 'RationalQuadratic'
        lengthscale (list): [lo, hi] (one shared lengthscale) or [[lo]*c, [hi]*c] (one per dimension)
        sparse (bool): inducing-point GP (pyro SparseGPRegression, VFE) instead of the exact one
        indpoints (int): number of inducing points (default len(X) // 10)
        learning_rate (float), iterations (int): Adam settings
        use_gpu (bool): the engine ALWAYS computes on the GPU; this flag only selects which
            generator the prior draws come from, so that use_gpu=False reproduces the reference's
            CPU runs (and its golden tests) and use_gpu=True its CUDA-generator runs
        verbose (int): 0, 1 or 2
        seed (int)
        **amplitude, **precision ('single' | 'double'), **jitter, **isotropic
        **shard: multi-GPU prediction under torch.distributed (see _shard_group); not in the reference
    """

    def __init__(self, X, y, Xtest=None, kernel='RBF', lengthscale=None, sparse=False, indpoints=None,
                 learning_rate=5e-2, iterations=1000, use_gpu=False, verbose=1, seed=0, **kwargs):
        self.precision = kwargs.get("precision", "double")
        npfloat_ = np.float32 if self.precision == "single" else np.float64
        self.verbose = verbose
        engine = get_engine()                            # raises when no CUDA device / library
        torch.manual_seed(seed)
        if use_gpu:
            torch.cuda.manual_seed_all(seed)
        input_dim = np.ndim(y)
        self.X, self.y = gprutils.prepare_training_data(X, y, precision=self.precision)
        self.do_sparse = sparse
        if lengthscale is None and not kwargs.get("isotropic"):
            lmean = npfloat_(np.mean(np.shape(y)) / 2)
            lengthscale = [[0. for _ in range(input_dim)], [lmean for _ in range(input_dim)]]
        elif lengthscale is None and kwargs.get("isotropic"):
            lengthscale = [0., npfloat_(np.mean(np.shape(y)) / 2)]
        kern = gp_kernels.get_kernel(kernel, input_dim, lengthscale, use_gpu,
                                     amplitude=kwargs.get('amplitude'), precision=self.precision)
        self.fulldims = Xtest.shape[1:] if Xtest is not None else np.shape(X)[1:]
        self.Xtest = gprutils.prepare_test_data(Xtest, precision=self.precision) if Xtest is not None else None
        jitter = kwargs.get("jitter", 1.0e-5)
        if not self.do_sparse:
            self.model = ExactGPModel(self.X, self.y, kern, jitter=jitter, engine=engine)
        else:                                             # gpr.py:145-155
            if indpoints is None:
                indpoints = len(self.X) // 10
                indpoints = indpoints + 1 if indpoints == 0 else indpoints
            else:
                indpoints = len(self.X) if indpoints > len(self.X) else indpoints
            Xu = self.X[::len(self.X) // indpoints]
            if self.verbose == 2:
                print("# of inducing points for sparse GP regression: {}".format(len(Xu)))
            self.model = SparseGPModel(self.X, self.y, kern, Xu, jitter=jitter, engine=engine)
        self.learning_rate = learning_rate
        self.iterations = iterations
        self.shard = kwargs.get("shard", "auto")
        self.indpoints_all = []
        self.lscales, self.noise_all, self.amp_all, self.loss_all = [], [], [], []
        self.hyperparams = {
            "lengthscale": self.lscales,
            "noise": self.noise_all,
            "variance": self.amp_all,
            "inducing_points": self.indpoints_all
        }

    def train(self, **kwargs):
        """Trains the model: **learning_rate, **iterations, **verbose as in gpr.py:170-217."""
        if kwargs.get("learning_rate") is not None:
            self.learning_rate = kwargs.get("learning_rate")
        if kwargs.get("iterations") is not None:
            self.iterations = kwargs.get("iterations")
        if kwargs.get("verbose") is not None:
            self.verbose = kwargs.get("verbose")
        start_time = time.time()
        if self.verbose:
            print('Model training...')
        if self.do_sparse:
            traj, xu_traj = self.model.fit(self.iterations, self.learning_rate)
            self.indpoints_all.extend(xu_traj.numpy())       # Xu after every step (gpr.py:198-199)
        else:
            traj = self.model.fit(self.iterations, self.learning_rate)
        traj = traj.double().numpy()
        d = self.model.X.shape[1]
        iso = self.model.kernel.isotropic
        for i, row in enumerate(traj):
            self.lscales.append(float(row[3]) if iso else row[3:3 + d].tolist())
            self.amp_all.append(float(row[0]))
            self.noise_all.append(float(row[1]))
            self.loss_all.append(float(row[3 + d]))
            if self.verbose == 2 and (i % 100 == 0 or i == self.iterations - 1):
                print('iter: {} ...'.format(i),
                      'loss: {} ...'.format(np.around(self.loss_all[-1], 4)),
                      'amp: {} ...'.format(np.around(self.amp_all[-1], 4)),
                      'length: {} ...'.format(np.around(self.lscales[-1], 4)),
                      'noise: {} ...'.format(np.around(self.noise_all[-1], 7)))
        if self.verbose:
            elapsed = time.time() - start_time
            if self.iterations > 100:
                print('average time per iteration: {} s'.format(np.round(elapsed / self.iterations, 5)))
            print('training completed in {} s'.format(np.round(elapsed, 2)))
            print('Final parameter values:\n',
                  'amp: {}, lengthscale: {}, noise: {}'.format(
                      np.around(self.model.kernel.variance_map.item(), 4),
                      np.around(self.model.kernel.lengthscale_map.tolist(), 4),
                      np.around(self.model.noise.item(), 7)))
        return

    def predict(self, Xtest=None, **kwargs):
        """Predictive mean and standard deviation (numpy, shaped like the test grid); gpr.py:219-255."""
        if Xtest is None and self.Xtest is None:
            warnings.warn("No test data provided. Using training data for prediction", UserWarning)
            self.Xtest = self.X
        elif Xtest is not None:
            self.Xtest = gprutils.prepare_test_data(Xtest, precision=self.precision)
            self.fulldims = Xtest.shape[1:]
        if kwargs.get("verbose") is not None:
            self.verbose = kwargs.get("verbose")
        if self.verbose:
            print("Calculating predictive mean and variance...", end=" ")
        if self._shard_group() is not None:
            # one process per GPU under torch.distributed (NCCL): every rank calls predict() (a collective), rank 0
            # factorises, the rows of X_full are tiled over the ranks and every rank gets the whole (mean, sd)
            from .. import sharded
            mean_d, sd_d = sharded.predict_model_sharded(self.model, self.Xtest, src=0, group=self._shard_group())
        else:
            mean_d, sd_d = self.model.predict_sd(self.Xtest)
        self._last_pred_device = (mean_d, sd_d)
        both = torch.stack((mean_d, sd_d)).cpu().numpy()          # one device->host copy, one synchronisation
        mean, sd = both[0], both[1]
        if mean.size == int(np.prod(self.fulldims)):
            mean, sd = mean.reshape(self.fulldims), sd.reshape(self.fulldims)
        if self.verbose:
            print("Done")
        return mean, sd

    def _shard_group(self):
        """The process group predict() shards over, or None.  ``shard=True`` / ``"auto"`` (default: auto = whenever
        torch.distributed runs on NCCL with more than one rank), ``shard=False`` keeps every rank independent,
        ``shard=<ProcessGroup>`` shards over that group.  Sharding makes predict() a collective call."""
        import torch.distributed as dist
        mode = self.shard
        if mode is False or mode is None or not (dist.is_available() and dist.is_initialized()):
            return None
        if mode is True or mode == "auto":
            if dist.get_world_size() < 2 or (mode == "auto" and dist.get_backend() != "nccl"):
                return None
            return dist.group.WORLD
        return mode if dist.get_world_size(mode) > 1 else None

    def run(self, **kwargs):
        """train() then predict(); returns (mean, sd, hyperparams) as gpr.py:257-283."""
        if kwargs.get("learning_rate") is not None:
            self.learning_rate = kwargs.get("learning_rate")
        if kwargs.get("iterations") is not None:
            self.iterations = kwargs.get("iterations")
        self.train(learning_rate=self.learning_rate, iterations=self.iterations)
        mean, sd = self.predict()
        return mean, sd, self.hyperparams

    def step(self, acquisition_function=None, batch_size=100, batch_update=False, lscale=None, **kwargs):
        """The reference's step() (gpr.py:285-329) ends in a call to the non-existent
        gprutils.acquisition (:326); it has been dead code since boptimizer replaced it."""
        raise NotImplementedError("reconstructor.step is dead code in the reference (gpr.py:326); use gpim.boptimizer")
