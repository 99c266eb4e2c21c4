"""
oracle/gp_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain torch, no CUDA, no ctypes) of the reference's exact-GP hot path so
that the CUDA engine in ``gpim_b200`` has something independent to be checked against.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module; nothing under ``gpim_b200/`` does.

Why a restatement: the arithmetic of ``gpim.reconstructor`` / ``gpim.boptimizer`` lives in
the un-vendored dependency ``pyro-ppl`` (pinned only as ``>=0.4.1`` in
/root/reference/requirements.txt:6 and setup.py:30), which is not installed in this image
and cannot be fetched.  The functions below restate the published algorithm of
``pyro.contrib.gp`` (kernels/isotropic.py, models/gpr.py, util.conditional, the Delta
"MAP" autoguide of parameterized.py, infer/trace_elbo.py) at the reference's own call
sites, each cited as reference file:line.

PARITY PINNING: this oracle reproduces (tests/test_oracle.py)
  * all three golden fixtures of the reference's own test-suite
    (test/test_boptim.py:42-58 -> test/test_data/test_{ei,poi,cb}.npy), and
  * the stored-output known answers of
    examples/notebooks/GP_based_exploration_exploitation.ipynb cell 13
    (hyper-parameters after each of the first trainings, 4 printed digits).
Unpinned by any reference test (formula-only): Matern52 / RationalQuadratic values, fp32
mode, 3-D/4-D inputs, mask / batch_update / dscale, isotropic lengthscale.
"""
import math
import time
import types

import numpy as np
import torch
from scipy.stats import norm
from torch.distributions import constraints, transform_to

KERNEL_NAMES = ("RBF", "RationalQuadratic", "Matern52")


# --------------------------------------------------------------------------------------
# data layout helpers (gprutils.py:23-85, 108-210)
# --------------------------------------------------------------------------------------
def full_grid(R, dense_x=1.0):
    """np.mgrid pixel coordinates, shape (c, *dims) -- gprutils.py:108-172 (no-extent branch)."""
    step = np.float64(dense_x)
    sl = tuple(slice(0, e, step) for e in R.shape)
    if not 2 <= len(sl) <= 4:
        raise NotImplementedError("Currently works only for 2D-4D sets")
    return np.array(np.mgrid[sl])


def sparse_grid(R):
    """Same coordinates with NaN wherever R is NaN -- gprutils.py:175-210."""
    if not np.isnan(R).any():
        raise NotImplementedError("Missing values in sparse data must be represented as NaNs")
    X = full_grid(R).astype(np.float64)
    c = X.shape[0]
    flat = X.reshape(c, -1)
    if R.ndim == 3 and not np.isnan(R[..., -1]).any():
        # whole-spectrum sparsity (gprutils.py:195-200): a spectrum is dropped if ANY z is NaN
        e1, e2, e3 = R.shape
        bad = np.isnan(R.reshape(e1 * e2, e3)).any(axis=1)
        X3 = X.reshape(c, e1 * e2, e3)
        X3[:, bad] = np.nan
        return X3.reshape(c, e1, e2, e3)
    if R.ndim not in (2, 3):
        raise NotImplementedError("Currently supports only 2D and 3D sets")
    flat[:, np.isnan(R.reshape(-1))] = np.nan
    return flat.reshape(X.shape)


def to_rows(X):
    """(c, *dims) -> (prod(dims), c) -- gprutils.py:49, 82."""
    return X.reshape(X.shape[0], -1).T


def training_rows(X, y):
    """Drop rows with any NaN coordinate / NaN observation -- gprutils.py:49-57."""
    Xr = to_rows(X)
    Xr = Xr[~np.isnan(Xr).any(axis=1)]
    yr = y.reshape(-1)
    yr = yr[~np.isnan(yr)]
    return Xr, yr


# --------------------------------------------------------------------------------------
# kernels (pyro.contrib.gp.kernels.isotropic; call sites pyro_kernels.py:58-68)
# --------------------------------------------------------------------------------------
def sq_scaled_dist(X, Z, lengthscale):
    """r2 = X2 - 2 X Z^T + Z2^T on lengthscale-scaled inputs, clamp(min=0)
    (Isotropy._square_scaled_dist)."""
    sX = X / lengthscale
    sZ = Z / lengthscale
    X2 = (sX ** 2).sum(1, keepdim=True)
    Z2 = (sZ ** 2).sum(1, keepdim=True)
    r2 = X2 - 2 * sX.matmul(sZ.t()) + Z2.t()
    return r2.clamp(min=0)


def kernel_matrix(name, X, Z, variance, lengthscale, scale_mixture=None):
    r2 = sq_scaled_dist(X, Z, lengthscale)
    if name == "RBF":
        return variance * torch.exp(-0.5 * r2)
    if name == "Matern52":
        r = (r2 + 1e-12).sqrt()                       # _torch_sqrt(x, eps=1e-12)
        s5r = 5 ** 0.5 * r
        return variance * (1 + s5r + (5.0 / 3) * r ** 2) * torch.exp(-s5r)
    if name == "RationalQuadratic":
        return variance * (1 + (0.5 / scale_mixture) * r2).pow(-scale_mixture)
    raise KeyError(name)


class OracleGP:
    """reconstructor restated (gpr.py:22-283) for the exact (sparse=False) model."""

    def __init__(self, X, y, Xtest=None, kernel="RBF", lengthscale=None,
                 learning_rate=5e-2, iterations=1000, seed=0, precision="double",
                 jitter=1e-5, amplitude=None, isotropic=False, num_threads=None):
        if kernel not in KERNEL_NAMES:
            raise KeyError(kernel)                     # pyro_kernels.py:70-75
        if num_threads:
            torch.set_num_threads(num_threads)
        self.dtype = torch.float32 if precision == "single" else torch.float64
        npf = np.float32 if precision == "single" else np.float64
        self.kernel_name = kernel
        torch.manual_seed(seed)                        # gpr.py:101
        dim = np.ndim(y)
        Xr, yr = training_rows(np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64))
        self.X = torch.from_numpy(Xr).to(self.dtype)
        self.y = torch.from_numpy(yr).to(self.dtype)
        if lengthscale is None:                        # gpr.py:118-123
            lmean = npf(np.mean(np.shape(y)) / 2)
            lengthscale = [0.0, lmean] if isotropic else [[0.0] * dim, [lmean] * dim]
        amp = [1e-4, 10.0] if amplitude is None else amplitude
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=npf), dtype=self.dtype)
        self.amp_lo, self.amp_hi = t(amp[0]), t(amp[1])
        self.ls_lo, self.ls_hi = t(lengthscale[0]), t(lengthscale[1])
        # prior draws, variance first then lengthscale (pyro_kernels.py:81-94; Uniform.sample)
        v0 = self.amp_lo + torch.rand(self.amp_lo.shape, dtype=self.dtype) * (self.amp_hi - self.amp_lo)
        l0 = self.ls_lo + torch.rand(self.ls_lo.shape, dtype=self.dtype) * (self.ls_hi - self.ls_lo)
        self.tf_v = transform_to(constraints.interval(self.amp_lo, self.amp_hi))
        self.tf_l = transform_to(constraints.interval(self.ls_lo, self.ls_hi))
        self.tf_n = transform_to(constraints.positive)
        self.u_v = self.tf_v.inv(v0).clone().requires_grad_(True)
        self.u_l = self.tf_l.inv(l0).clone().requires_grad_(True)
        self.u_n = self.tf_n.inv(torch.tensor(1.0, dtype=self.dtype)).clone().requires_grad_(True)
        self.params = [self.u_v, self.u_l, self.u_n]
        if kernel == "RationalQuadratic":              # scale_mixture: PyroParam(1.0, positive)
            self.u_a = torch.zeros((), dtype=self.dtype, requires_grad=True)
            self.params.append(self.u_a)
        self.jitter = jitter
        self.learning_rate, self.iterations = learning_rate, iterations
        self.fulldims = (Xtest.shape[1:] if Xtest is not None else np.shape(X)[1:])
        self.Xtest = None if Xtest is None else torch.from_numpy(
            to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype)
        self.lscales, self.amp_all, self.noise_all = [], [], []
        self.hyperparams = {"lengthscale": self.lscales, "noise": self.noise_all,
                            "variance": self.amp_all, "inducing_points": []}

    # -- constrained views -------------------------------------------------------------
    def theta(self):
        a = self.tf_n(self.u_a) if self.kernel_name == "RationalQuadratic" else None
        return self.tf_v(self.u_v), self.tf_l(self.u_l), self.tf_n(self.u_n), a

    def set_data(self, X, y):
        """boptim.py:243-249 (model.X / model.y swapped in place)."""
        Xr, yr = training_rows(X, y)
        self.X = torch.from_numpy(Xr).to(self.dtype)
        self.y = torch.from_numpy(yr).to(self.dtype)

    def _factor(self, v, l, n, a):
        N = self.X.shape[0]
        K = kernel_matrix(self.kernel_name, self.X, self.X, v, l, a)
        K = K + (self.jitter + n) * torch.eye(N, dtype=self.dtype)   # GPRegression.model
        return torch.linalg.cholesky(K)

    def nll(self):
        """-log N(y; 0, K + (noise+jitter) I); Uniform log-priors are constants (Trace_ELBO)."""
        v, l, n, a = self.theta()
        L = self._factor(v, l, n, a)
        alpha = torch.linalg.solve_triangular(L, self.y.unsqueeze(1), upper=False)
        N = self.X.shape[0]
        return 0.5 * (alpha ** 2).sum() + L.diagonal().log().sum() + 0.5 * N * math.log(2 * math.pi)

    def train(self, learning_rate=None, iterations=None):
        """gpr.py:170-217: fresh Adam each call, theta warm-started, record AFTER the step."""
        if learning_rate is not None:
            self.learning_rate = learning_rate
        if iterations is not None:
            self.iterations = iterations
        opt = torch.optim.Adam(self.params, lr=self.learning_rate)
        for _ in range(self.iterations):
            opt.zero_grad()
            loss = self.nll()
            loss.backward()
            opt.step()
            with torch.no_grad():
                v, l, n, _a = self.theta()
                self.lscales.append(l.tolist())
                self.amp_all.append(v.item())
                self.noise_all.append(n.item())

    def predict_rows(self, Xs):
        """GPRegression.forward + util.conditional, full_cov=False, noiseless=False (gpr.py:248)."""
        with torch.no_grad():
            v, l, n, a = self.theta()
            L = self._factor(v, l, n, a)
            Kfs = kernel_matrix(self.kernel_name, self.X, Xs, v, l, a)
            pack = torch.cat((self.y.unsqueeze(1), Kfs), dim=1)
            sol = torch.linalg.solve_triangular(L, pack, upper=False)
            vhat, W = sol[:, :1], sol[:, 1:].t()
            mean = W.matmul(vhat).squeeze(-1)
            var = (v - W.pow(2).sum(-1)).clamp(min=0) + n
            # rows with NaN coordinates propagate NaN (needed by EI's nanmax, acqfunc.py:57-59)
            bad = torch.isnan(Xs).any(dim=1)
            mean = torch.where(bad, torch.full_like(mean, float("nan")), mean)
            var = torch.where(bad, torch.full_like(var, float("nan")), var)
        return mean.numpy(), var.sqrt().numpy()

    def predict(self, Xtest=None):
        if Xtest is not None:
            self.Xtest = torch.from_numpy(to_rows(np.asarray(Xtest, dtype=np.float64))).to(self.dtype)
            self.fulldims = Xtest.shape[1:]
        elif self.Xtest is None:
            self.Xtest = self.X
        mean, sd = self.predict_rows(self.Xtest)
        return mean.reshape(self.fulldims), sd.reshape(self.fulldims)

    def run(self):
        self.train()
        mean, sd = self.predict()
        return mean, sd, self.hyperparams


# --------------------------------------------------------------------------------------
# acquisition functions (acqfunc.py:11-92) and the BO loop (boptim.py:239-470)
# --------------------------------------------------------------------------------------
def acq_cb(gp, X_full, X_sparse, alpha=0, beta=1, **_):
    mean, sd = gp.predict(X_full)
    return alpha * mean + beta * sd, (mean, sd)


def acq_ei(gp, X_full, X_sparse, xi=0.01, **_):
    mean, sd = gp.predict(X_full)
    mean_s, _sd = gp.predict(X_sparse)
    imp = mean - np.nanmax(mean_s) - xi
    z = imp / sd
    return imp * norm.cdf(z) + sd * norm.pdf(z), (mean, sd)


def acq_poi(gp, X_full, X_sparse, xi=0.01, **_):
    mean, sd = gp.predict(X_full)
    both = gp.predict(X_sparse)          # the reference keeps the (mean, sd) TUPLE, acqfunc.py:86
    z = (mean - np.nanmax(both) - xi) / sd
    return norm.cdf(z), (mean, sd)


ACQ = {"cb": acq_cb, "ei": acq_ei, "poi": acq_poi}


def rank_points(acq, batch_size, mask=None):
    """boptim.py:303-315: full argsort, reversed, top batch_size."""
    if mask is not None:
        acq = mask * acq
    order = np.argsort(acq.ravel())
    idx = np.stack(np.unravel_index(order, acq.shape), axis=1)
    vals = acq.ravel()[order]
    if mask is not None:
        keep = ~np.isnan(vals)
        vals, idx = vals[keep], idx[:keep.sum()]
    return vals[::-1][:batch_size].tolist(), idx[::-1][:batch_size].tolist()


def pick_unvisited(idx_list, val_list, visited, dscale=None, gamma=0.8, memory=10, exit_strategy=1):
    """boptim.py:378-429 restated: first candidate that is neither visited nor too close to
    the k-th most recent pick (dscale * gamma**k)."""
    ds = 0 if dscale is None else dscale

    def too_close(idx):
        prev = visited[-memory:]
        d = [np.linalg.norm(np.array(idx) - np.array(p)) for p in prev][::-1]
        lim = [ds * gamma ** k for k in range(len(prev))]
        return any(not (di > li) for di, li in zip(d, lim))

    j = 0
    if not visited:
        return idx_list[0], val_list[0]
    while (idx_list[j] in visited) or too_close(idx_list[j]):
        j += 1
        if j == len(idx_list):
            j = np.random.randint(0, len(idx_list)) if exit_strategy else -1
            break
    return idx_list[j], val_list[j]


def suppress_batch(val_list, idx_list, visited, batch_dscale, batch_out_max=10, dscale=None, gamma=0.8, memory=10,
                   exit_strategy=1):
    """boptim.py:326-376 restated (batch_update=True).  The list is cut at the first admissible candidate
    (pick_unvisited), then: take the largest remaining value, drop every candidate within batch_dscale of it
    (closed ball, Euclidean distance between index vectors), repeat until nothing is left; keep the first
    batch_out_max picks; if fewer were found, pad with uniformly random candidates of the cut list (np.random)."""
    _, val0 = pick_unvisited(idx_list, val_list, visited, dscale, gamma, memory, exit_strategy)
    start = int(np.where(np.array(val_list) == val0)[0][0])
    vals = np.array(val_list, dtype=float)[start:]
    pts = np.vstack(idx_list)[start:].astype(float)
    orig = vals.copy()
    dead = vals.min() - 1
    picks = []
    while True:
        b = int(np.argmax(vals))
        if not vals[b] > dead:
            break
        picks.append(b)
        vals[np.linalg.norm(pts - pts[b], axis=1) <= batch_dscale] = dead
    picks = picks[:batch_out_max]
    out_vals = [float(orig[b]) for b in picks]
    out_idx = [[int(c) for c in pts[b]] for b in picks]
    short = batch_out_max - len(picks)
    if short > 0:
        extra = np.random.randint(0, len(vals), short)
        out_idx.extend([[int(c) for c in pts[b]] for b in extra])
        out_vals.extend(orig[extra].tolist())
    return out_vals, out_idx


def bo_run(X_seed, y_seed, X_full, target, acquisition="cb", exploration_steps=10, batch_size=100,
           kernel="RBF", lengthscale=None, gp_iterations=1000, seed=0, learning_rate=5e-2,
           jitter=1e-6, precision="double", isotropic=False, mask=None, dscale=None,
           on_train=None, batch_update=False, batch_dscale=None, batch_out_max=10, gamma=0.8, memory=10,
           **acq_kw):
    """boptimizer.run restated (boptim.py:431-470), single-point steps and batch_update=True (every step measures the
    batch_out_max points suppress_batch returns; batch_dscale defaults to the mean kernel lengthscale, boptim.py:318-322).
    Returns dict(target_func_vals, gp_predictions, indices_all, vals_all, gp)."""
    gp = OracleGP(X_seed, y_seed, X_full, kernel, lengthscale, learning_rate, gp_iterations, seed,
                  precision=precision, jitter=jitter, isotropic=isotropic)
    X_sparse, y_sparse = X_seed.copy(), y_seed.copy()
    acq_fn = ACQ[acquisition] if isinstance(acquisition, str) else acquisition
    if not isinstance(acquisition, (str, types.FunctionType)):
        raise NotImplementedError
    vals_hist, preds, picked, picked_vals = [y_seed.copy()], [], [], []
    for e in range(exploration_steps):
        if e == 0:
            gp.train()
            if on_train:
                on_train(gp)
        acq, pred = acq_fn(gp, X_full, X_sparse, **acq_kw)
        preds.append(pred)
        vals, idxs = rank_points(acq, batch_size, mask)
        if batch_update:
            bd = float(torch.as_tensor(gp.theta()[1]).detach().double().mean()) if batch_dscale is None else batch_dscale
            bvals, binds = suppress_batch(vals, idxs, picked, bd, batch_out_max, dscale, gamma, memory)
            for ind in binds:
                y_sparse[tuple(ind)] = target(tuple(ind))
        else:
            ind, val = pick_unvisited(idxs, vals, picked, dscale, gamma, memory)
            y_sparse[tuple(ind)] = target(tuple(ind))
        X_sparse = sparse_grid(y_sparse)
        vals_hist.append(y_sparse.copy())
        gp.set_data(X_sparse, y_sparse)
        gp.train()
        if on_train:
            on_train(gp)
        if batch_update:
            picked.extend(binds)
            picked_vals.extend(bvals)
        else:
            picked.append(ind)
            picked_vals.append(val)
    return {"target_func_vals": vals_hist, "gp_predictions": preds, "indices_all": picked,
            "vals_all": picked_vals, "gp": gp}


# --------------------------------------------------------------------------------------
# fixed-theta predict used as the CPU baseline timer (bench.py cpu_baseline / --impl reference)
# --------------------------------------------------------------------------------------
def predict_fixed_theta(kernel, X, y, Xs, variance, lengthscale, noise, jitter=1e-5,
                        dtype=torch.float64, scale_mixture=1.0, device=None, to_numpy=True):
    """One reference-style predict() with given constrained theta: K, cholesky, K*, trsm, reduce.
    X (N,d), y (N,), Xs (M,d) numpy (or tensors).  Returns (mean, sd, seconds_by_stage).
    device="cuda" runs the same torch ops on the GPU (what the reference's use_gpu=True dispatches to:
    cuSOLVER potrf, cuBLAS trsm) -- bench.py's torch_cuda_baseline leg; stages are then synchronised."""
    dev = torch.device(device) if device is not None else torch.device("cpu")
    sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
    X = torch.as_tensor(X, dtype=dtype, device=dev)
    y = torch.as_tensor(y, dtype=dtype, device=dev)
    Xs = torch.as_tensor(Xs, dtype=dtype, device=dev)
    ls = torch.as_tensor(lengthscale, dtype=dtype, device=dev)
    t = {}
    sync(); t0 = time.perf_counter()
    K = kernel_matrix(kernel, X, X, variance, ls, scale_mixture)
    K.view(-1)[:: X.shape[0] + 1] += jitter + noise
    sync(); t["kmat"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    L = torch.linalg.cholesky(K)
    sync(); t["cholesky"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Kfs = kernel_matrix(kernel, X, Xs, variance, ls, scale_mixture)
    sync(); t["kcross"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    sol = torch.linalg.solve_triangular(L, torch.cat((y.unsqueeze(1), Kfs), dim=1), upper=False)
    sync(); t["trsm"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    vhat, W = sol[:, :1], sol[:, 1:].t()
    mean = W.matmul(vhat).squeeze(-1)
    var = (variance - W.pow(2).sum(-1)).clamp(min=0) + noise
    sd = var.sqrt()
    sync(); t["reduce"] = time.perf_counter() - t0
    if to_numpy:
        return mean.cpu().numpy(), sd.cpu().numpy(), t
    return mean, sd, t
