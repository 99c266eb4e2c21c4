"""
ctypes binding of libgpgrid.so (include/gpgrid.h) and the thin Engine wrapper the host code uses.

There is deliberately NO fallback: if the shared library is missing, or no CUDA device is visible,
every compute entry raises.  PyTorch is used only to own device memory and streams.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _build

GPG_OK, GPG_EINVAL, GPG_ENOTPD, GPG_ECUDA = 0, 1, 2, 3
GPG_F32, GPG_F64 = 0, 1
KERNEL_IDS = {"RBF": 0, "Matern52": 1, "RationalQuadratic": 2}
ACQ_IDS = {"cb": 0, "ei": 1, "poi": 2}
OPT_GEMM_PATH, OPT_PREDICT_CHUNK, OPT_STAGE_TIMING = 1, 2, 3
OPT_PANEL_REFINE, OPT_SYRK_CHUNK, OPT_FACTOR_ALGO, OPT_FIT_GRAPH, OPT_PANEL_MODE, OPT_OUTER_PANEL, OPT_COMPACT_SUPPORT = 4, 5, 6, 7, 8, 9, 10
OPT_INNER_LEFT, OPT_LOOKAHEAD, OPT_PANEL_WORKERS = 11, 12, 13
STAGES = ("kmat", "cholesky", "trtri", "solve", "kcross", "pgemm", "pfinal", "grad", "acq")

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double

# name -> argtypes; every symbol include/gpgrid.h declares (tests check the export list against this)
SIGNATURES = {
    "gpg_version": ([], C.c_int),
    "gpg_last_error": ([], C.c_char_p),
    "gpg_create": ([_i32, C.POINTER(_vp)], C.c_int),
    "gpg_destroy": ([_vp], C.c_int),
    "gpg_set_option": ([_vp, _i32, C.c_longlong], C.c_int),
    "gpg_launch_count": ([_vp], C.c_longlong),
    "gpg_workspace_bytes": ([_vp], C.c_size_t),
    "gpg_variance_gemm_macs": ([_vp, C.POINTER(_f64)], C.c_int),
    "gpg_stage_times": ([_vp, C.POINTER(_f64), C.POINTER(C.c_longlong)], C.c_int),
    "gpg_gemm_nt_f32": ([_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _f64, _f64, _f64, _f64, _vp], C.c_int),
    "gpg_kmat": ([_vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _f64, _i32, _vp, _i64, _vp], C.c_int),
    "gpg_cholesky": ([_vp, _i32, _vp, _i64, _i64, _vp, _vp], C.c_int),
    "gpg_trtri": ([_vp, _i32, _vp, _i64, _i64, _vp, _i64, _vp], C.c_int),
    "gpg_solve_vec": ([_vp, _i32, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp], C.c_int),
    "gpg_factorize": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _f64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                       _vp], C.c_int),
    "gpg_predict": ([_vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp], C.c_int),
    "gpg_predict_grid": ([_vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, C.POINTER(_i64),
                          C.POINTER(_f64), _i64, _i64, _vp, _vp, _vp], C.c_int),
    "gpg_nll_grad": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _f64, _vp, _vp, _vp, _vp], C.c_int),
    "gpg_fit_adam": ([_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _f64, _vp, C.POINTER(_f64), _i32, _f64,
                      _vp, _vp, _vp, _vp], C.c_int),
    "gpg_fit_adam_sk": ([_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _f64, _vp, C.POINTER(_f64), _i32, _f64,
                         _vp, _vp, _vp, _vp], C.c_int),
    "gpg_fit_adam_mt": ([_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _f64, _vp, C.POINTER(_f64), _i32, _f64,
                         _vp, _vp, _vp, _vp], C.c_int),
    "gpg_acq_sweep": ([_vp, _i32, _i32, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _i32, _vp, _vp, _vp, _vp, _vp],
                      C.c_int),
    "gpg_acq_select": ([_vp, _i32, _vp, _vp, _vp, _i32, _i32, C.POINTER(_i64), _vp, _i32, _i32, _f64, _f64, _i32, _f64,
                        _i32, _vp, _vp], C.c_int),
    "gpg_sparse_loss_grad": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _f64, _vp, _vp, _vp, _vp, _vp],
                             C.c_int),
    "gpg_sparse_fit_adam": ([_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _f64, _vp, C.POINTER(_f64), _i32,
                             _f64, _vp, _vp, _vp, _vp, _vp], C.c_int),
    "gpg_sparse_factorize": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _f64, _vp, _vp, _i64, _vp, _vp,
                              _vp, _vp, _vp], C.c_int),
    "gpg_sparse_predict": ([_vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp,
                            _vp], C.c_int),
    # multi-GPU (NCCL bound at run time inside the library)
    "gpg_comm_unique_id": ([_vp], C.c_int),
    "gpg_comm_init": ([_vp, _i32, _i32, _vp], C.c_int),
    "gpg_comm_destroy": ([_vp], C.c_int),
    "gpg_comm_info": ([_vp, C.POINTER(_i32), C.POINTER(_i32)], C.c_int),
    "gpg_predict_uses_planes": ([_vp, _i32, _i64, _i32], C.c_int),
    "gpg_bcast_factor": ([_vp, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp], C.c_int),
    "gpg_allgather_pred": ([_vp, _i32, _vp, _i64, _vp, _vp], C.c_int),
    "gpg_predict_sharded": ([_vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp,
                             _i64, _vp, _vp], C.c_int),
    "gpg_acq_sweep_sharded": ([_vp, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _f64, _f64, _f64, _f64, _i32, _vp, _vp, _vp,
                               _vp, _vp], C.c_int),
}

_LIB = None


def lib_path():
    return _build.LIB


def load_library(build_if_missing=True):
    """Loads (building if stale and nvcc is present) libgpgrid.so; raises if that is impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if build_if_missing and _build.is_stale():
        try:
            _build.build()
        except Exception as e:                       # noqa: BLE001
            if not os.path.exists(path):
                raise RuntimeError(f"libgpgrid.so is missing and could not be built: {e}") from e
    if not os.path.exists(path):
        raise RuntimeError("libgpgrid.so not found; run `python -m gpim_b200._build`")
    lib = C.CDLL(path)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)                       # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.gpg_version() < 120:
        raise RuntimeError("libgpgrid.so is older than this package")
    _LIB = lib
    return lib


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise ValueError("libgpgrid takes device pointers; got a CPU tensor")
    return C.c_void_p(t.data_ptr())


def _c(t):
    """The C ABI assumes dense row-major arrays."""
    return None if t is None else t.contiguous()


def np_dtype(precision):
    return np.float32 if precision == "single" else np.float64


def torch_dtype(precision):
    return torch.float32 if precision == "single" else torch.float64


class Engine:
    """One handle on one CUDA device.  All tensor arguments must live on that device."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError(
                "gpim_b200 needs a CUDA device (B200, sm_100a): this engine has no CPU path. "
                "The reference's use_gpu=False mode is reproduced numerically (same RNG stream), "
                "not by running on the host.")
        self.lib = load_library()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = C.c_void_p()
        self._check(self.lib.gpg_create(self.device.index, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.gpg_destroy(self.h)
                self.h = None
        except Exception:                             # noqa: BLE001
            pass

    # -- helpers ------------------------------------------------------------------------
    def _check(self, rc):
        if rc != GPG_OK:
            msg = self.lib.gpg_last_error().decode()
            if rc == GPG_EINVAL:
                raise ValueError(msg)
            raise RuntimeError(f"libgpgrid error {rc}: {msg}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _dt(t):
        if t.dtype == torch.float32:
            return GPG_F32
        if t.dtype == torch.float64:
            return GPG_F64
        raise TypeError(f"unsupported dtype {t.dtype}")

    def set_option(self, key, value):
        self._check(self.lib.gpg_set_option(self.h, key, int(value)))

    def launch_count(self):
        return int(self.lib.gpg_launch_count(self.h))

    def stage_times(self):
        """{stage: (device_ms, n_brackets)} since the last call; needs set_option(OPT_STAGE_TIMING, 1)."""
        ms = (_f64 * len(STAGES))()
        n = (C.c_longlong * len(STAGES))()
        self._check(self.lib.gpg_stage_times(self.h, ms, n))
        return {name: (ms[i], int(n[i])) for i, name in enumerate(STAGES)}

    def variance_gemm_macs(self):
        """MACs the variance GEMM executed under stage timing since the last call (tile granularity)."""
        v = _f64(0.0)
        self._check(self.lib.gpg_variance_gemm_macs(self.h, C.byref(v)))
        return float(v.value)

    def empty(self, *shape, dtype):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # -- entry points -------------------------------------------------------------------
    def gemm_nt(self, A, B, C=None, alpha=1.0, beta=0.0, scale_a=None, scale_b=None):
        """C = alpha * A @ B.T + beta * C on the split-fp16 tcgen05 kernel (f32 row-major)."""
        assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.stride(1) == 1 and B.stride(1) == 1
        M, K = A.shape
        N = B.shape[0]
        if C is None:
            C = self.empty(M, N, dtype=torch.float32)

        def pow2_scale(t):
            m = float(t.abs().max().item()) or 1.0
            return 2.0 ** np.floor(np.log2(16384.0 / m))
        sa = pow2_scale(A) if scale_a is None else scale_a
        sb = pow2_scale(B) if scale_b is None else scale_b
        self._check(self.lib.gpg_gemm_nt_f32(self.h, _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(C), C.stride(0),
                                             M, N, K, float(alpha), float(beta), float(sa), float(sb), self._stream()))
        return C

    def kmat(self, kernel_id, theta, X, Z=None, jitter=0.0, lower_only=False, out=None):
        theta, X, Z = _c(theta), _c(X), _c(Z)
        N, d = X.shape
        P = N if Z is None else Z.shape[0]
        if out is None:
            out = self.empty(N, P, dtype=X.dtype)
        self._check(self.lib.gpg_kmat(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), N, _ptr(Z), P,
                                      float(jitter), int(lower_only), _ptr(out), out.stride(0), self._stream()))
        return out

    def cholesky_(self, A, info=None):
        N = A.shape[0]
        if info is None:
            info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.gpg_cholesky(self.h, self._dt(A), _ptr(A), N, A.stride(0), _ptr(info), self._stream()))
        return A, info

    def trtri(self, L, out=None):
        N = L.shape[0]
        if out is None:
            out = self.empty(N, N, dtype=L.dtype)
        self._check(self.lib.gpg_trtri(self.h, self._dt(L), _ptr(L), N, L.stride(0), _ptr(out), out.stride(0),
                                       self._stream()))
        return out

    def solve_vec(self, L, Linv, y):
        y = _c(y)
        N = L.shape[0]
        assert L.stride(0) == Linv.stride(0)
        vhat, alpha = torch.empty_like(y), torch.empty_like(y)
        scalars = self.empty(2, dtype=y.dtype)
        self._check(self.lib.gpg_solve_vec(self.h, self._dt(L), _ptr(L), _ptr(Linv), N, L.stride(0), _ptr(y),
                                           _ptr(vhat), _ptr(alpha), _ptr(scalars), self._stream()))
        return vhat, alpha, scalars

    def alloc_factor(self, N, dtype, with_L=True):
        """Empty factor cache (what a non-factorising rank receives by broadcast, sharded.py)."""
        ld = (N + 63) // 64 * 64
        f32 = dtype == torch.float32
        return {"L": self.empty(N, ld, dtype=dtype) if with_L else None,
                "Linv": self.empty(N, ld, dtype=dtype),
                # tensor-core form of Linv (fp16 hi plane, lo plane) + operand scales; f32 only
                "wsplit": torch.empty(2, N, ld, dtype=torch.float16, device=self.device) if f32 else None,
                "scales": torch.zeros(16, dtype=torch.float32, device=self.device) if f32 else None,
                "vhat": self.empty(N, dtype=dtype), "alpha": self.empty(N, dtype=dtype),
                "scalars": self.empty(2, dtype=dtype),
                "info": torch.zeros(1, dtype=torch.int32, device=self.device), "ld": ld}

    def factorize(self, kernel_id, theta, X, y, jitter, out=None):
        """-> dict(L, Linv, vhat, alpha, scalars, info); N x N buffers padded to ld % 64 == 0."""
        theta, X, y = _c(theta), _c(X), _c(y)
        N, d = X.shape
        fac = out if out is not None else self.alloc_factor(N, X.dtype)
        L, Linv, vhat, alpha, scalars, info, ld = (fac[k] for k in ("L", "Linv", "vhat", "alpha", "scalars", "info", "ld"))
        self._check(self.lib.gpg_factorize(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), _ptr(y), N,
                                           float(jitter), _ptr(L), _ptr(Linv), ld, _ptr(vhat), _ptr(alpha),
                                           _ptr(scalars), _ptr(info), _ptr(fac.get("wsplit")), _ptr(fac.get("scales")),
                                           self._stream()))
        return fac

    def predict(self, kernel_id, theta, X, fac, Xs, mean=None, sd=None):
        theta, X, Xs = _c(theta), _c(X), _c(Xs)
        N, d = X.shape
        M = Xs.shape[0]
        if mean is None:
            mean = self.empty(M, dtype=X.dtype)
        if sd is None:
            sd = self.empty(M, dtype=X.dtype)
        if M == 0:
            return mean, sd
        self._check(self.lib.gpg_predict(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), N, _ptr(fac["Linv"]),
                                         fac["ld"], _ptr(fac["alpha"]), _ptr(fac.get("wsplit")), _ptr(fac.get("scales")),
                                         _ptr(Xs), M, _ptr(mean), _ptr(sd), self._stream()))
        return mean, sd

    def predict_grid(self, kernel_id, theta, X, fac, dims, step, j0, M, mean=None, sd=None):
        theta, X = _c(theta), _c(X)
        N, d = X.shape
        if mean is None:
            mean = self.empty(M, dtype=X.dtype)
        if sd is None:
            sd = self.empty(M, dtype=X.dtype)
        dims_c = (_i64 * d)(*[int(v) for v in dims])
        step_c = (_f64 * d)(*[float(v) for v in step])
        self._check(self.lib.gpg_predict_grid(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), N,
                                              _ptr(fac["Linv"]), fac["ld"], _ptr(fac["alpha"]),
                                              _ptr(fac.get("wsplit")), _ptr(fac.get("scales")), dims_c, step_c,
                                              int(j0), int(M), _ptr(mean), _ptr(sd), self._stream()))
        return mean, sd

    def nll_grad(self, kernel_id, theta, X, y, jitter):
        theta, X, y = _c(theta), _c(X), _c(y)
        N, d = X.shape
        nll = self.empty(1, dtype=X.dtype)
        grad = self.empty(3 + d, dtype=X.dtype)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.gpg_nll_grad(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), _ptr(y), N,
                                          float(jitter), _ptr(nll), _ptr(grad), _ptr(info), self._stream()))
        return nll, grad, info

    def fit_adam(self, kernel_id, X, y, jitter, u, bounds, n_ls, iters, lr, gpytorch_params=False):
        """u (device, in/out) layout {variance, noise, scale_mixture, lengthscale[n_ls]}; with gpytorch_params
        (skreconstructor) {outputscale, noise, mean constant, lengthscale[n_ls]} in GPyTorch's raw parametrisation.
        Returns (traj [iters, 4+d], theta [3+d], info)."""
        X, y = _c(X), _c(y)
        assert u.is_contiguous()
        N, d = X.shape
        traj = self.empty(max(iters, 1), 4 + d, dtype=X.dtype)
        theta = self.empty(3 + d, dtype=X.dtype)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        b = (_f64 * len(bounds))(*[float(v) for v in bounds])
        fn = self.lib.gpg_fit_adam_sk if gpytorch_params else self.lib.gpg_fit_adam
        self._check(fn(self.h, self._dt(X), kernel_id, d, n_ls, _ptr(X), _ptr(y), N, float(jitter),
                       _ptr(u), b, int(iters), float(lr), _ptr(traj), _ptr(theta), _ptr(info), self._stream()))
        return traj[:iters], theta, info

    def fit_adam_mt(self, kernel_id, X, Y, jitter, u, ls_bounds, n_ls, iters, lr):
        """vreconstructor(independent=True): Y [T, N] task-major, u raw {outputscale[T] | task noise[T] | noise |
        constant[T] | lengthscale[n_ls]} (device, in/out), ls_bounds [lo..., hi...] or None (softplus lengthscale).
        Returns (traj [iters, d + 1] = {lengthscale[d], loss}, theta [T, 3 + d], info)."""
        X, Y = _c(X), _c(Y)
        assert u.is_contiguous()
        N, d = X.shape
        T = Y.shape[0]
        assert Y.shape[1] == N and u.numel() == 3 * T + 1 + n_ls
        traj = self.empty(max(iters, 1), d + 1, dtype=X.dtype)
        theta = self.empty(T, 3 + d, dtype=X.dtype)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        b = (_f64 * len(ls_bounds))(*[float(v) for v in ls_bounds]) if ls_bounds is not None else None
        self._check(self.lib.gpg_fit_adam_mt(self.h, self._dt(X), kernel_id, d, n_ls, T, _ptr(X), _ptr(Y), N, float(jitter),
                                             _ptr(u), b, int(iters), float(lr), _ptr(traj), _ptr(theta), _ptr(info),
                                             self._stream()))
        return traj[:iters], theta, info

    def acq_sweep(self, acq_id, mean, sd, k, mu_best=0.0, xi=0.01, alpha=0.0, beta=1.0, mask=None, want_acq=False):
        mean, sd, mask = _c(mean), _c(sd), _c(mask)
        M = mean.numel()
        k = int(min(k, M))
        vals = self.empty(k, dtype=mean.dtype)
        idx = torch.empty(k, dtype=torch.int64, device=self.device)
        count = torch.zeros(1, dtype=torch.int32, device=self.device)
        acq = self.empty(M, dtype=mean.dtype) if want_acq else None
        self._check(self.lib.gpg_acq_sweep(self.h, self._dt(mean), acq_id, _ptr(mean), _ptr(sd), _ptr(mask), M,
                                           float(mu_best), float(xi), float(alpha), float(beta), k, _ptr(vals),
                                           _ptr(idx), _ptr(count), _ptr(acq), self._stream()))
        return vals, idx, count, acq

    def acq_select(self, vals, idx, count, dims, visited_flat, memory=10, dscale=0.0, gamma=0.8, batch=False,
                   batch_dscale=0.0, batch_out_max=10):
        """gpg_acq_select on the device-resident output of acq_sweep -> (first, start, picks list, nan_seen) on the host."""
        k = int(vals.numel())
        dims_c = (_i64 * len(dims))(*[int(v) for v in dims])
        nv = len(visited_flat)
        vis = torch.as_tensor(np.asarray(visited_flat, dtype=np.int64), device=self.device) if nv else None
        sel = torch.zeros(4 + max(int(batch_out_max), 0), dtype=torch.int32, device=self.device)
        self._check(self.lib.gpg_acq_select(self.h, self._dt(vals), _ptr(vals), _ptr(idx), _ptr(count), k, len(dims), dims_c,
                                            _ptr(vis), nv, int(memory), float(dscale), float(gamma), int(bool(batch)),
                                            float(batch_dscale), int(batch_out_max), _ptr(sel), self._stream()))
        out = sel.cpu().numpy()
        return int(out[0]), int(out[1]), [int(v) for v in out[4:4 + int(out[2])]], bool(out[3])

    # -- inducing-point GP (sparse=True) ----------------------------------------------------
    def sparse_loss_grad(self, kernel_id, theta, X, y, Xu, jitter):
        """-> (loss [1], grad_theta [3 + d], grad_Xu [m, d], info)."""
        theta, X, y, Xu = _c(theta), _c(X), _c(y), _c(Xu)
        N, d = X.shape
        m = Xu.shape[0]
        loss = self.empty(1, dtype=X.dtype)
        grad = self.empty(3 + d, dtype=X.dtype)
        gxu = self.empty(m, d, dtype=X.dtype)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.gpg_sparse_loss_grad(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), _ptr(y), N,
                                                  _ptr(Xu), m, float(jitter), _ptr(loss), _ptr(grad), _ptr(gxu),
                                                  _ptr(info), self._stream()))
        return loss, grad, gxu, info

    def sparse_fit_adam(self, kernel_id, X, y, Xu, jitter, u, bounds, n_ls, iters, lr, record_xu=True):
        """u and Xu (device, in/out).  Returns (traj [iters, 4+d], xu_traj [iters, m, d] | None, theta [3+d], info)."""
        X, y = _c(X), _c(y)
        assert u.is_contiguous() and Xu.is_contiguous()
        N, d = X.shape
        m = Xu.shape[0]
        traj = self.empty(max(iters, 1), 4 + d, dtype=X.dtype)
        xu_traj = self.empty(max(iters, 1), m, d, dtype=X.dtype) if record_xu else None
        theta = self.empty(3 + d, dtype=X.dtype)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        b = (_f64 * len(bounds))(*[float(v) for v in bounds])
        self._check(self.lib.gpg_sparse_fit_adam(self.h, self._dt(X), kernel_id, d, n_ls, _ptr(X), _ptr(y), N, _ptr(Xu),
                                                 m, float(jitter), _ptr(u), b, int(iters), float(lr), _ptr(traj),
                                                 _ptr(xu_traj), _ptr(theta), _ptr(info), self._stream()))
        return traj[:iters], (xu_traj[:iters] if record_xu else None), theta, info

    def alloc_sparse_factor(self, m, dtype):
        """Empty inducing-point factor cache (what a non-factorising rank receives by broadcast, sharded.py)."""
        ld = (m + 63) // 64 * 64
        f32 = dtype == torch.float32
        return {"Ui": self.empty(m, ld, dtype=dtype), "Pm": self.empty(m, ld, dtype=dtype),
                "w": self.empty(m, dtype=dtype), "info": torch.zeros(1, dtype=torch.int32, device=self.device), "ld": ld,
                # tensor-core form of the two factors (fp16 hi / lo planes of Ui and Pm) + operand scales; f32 only
                "split": torch.empty(4, m, ld, dtype=torch.float16, device=self.device) if f32 else None,
                "scales": torch.zeros(24, dtype=torch.float32, device=self.device) if f32 else None}

    def sparse_factorize(self, kernel_id, theta, X, y, Xu, jitter, out=None):
        """-> dict(Ui, Pm, w, info, ld[, split, scales]): the cache gpg_sparse_predict consumes."""
        theta, X, y, Xu = _c(theta), _c(X), _c(y), _c(Xu)
        N, d = X.shape
        m = Xu.shape[0]
        fac = out if out is not None else self.alloc_sparse_factor(m, X.dtype)
        self._check(self.lib.gpg_sparse_factorize(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), _ptr(y), N,
                                                  _ptr(Xu), m, float(jitter), _ptr(fac["Ui"]), _ptr(fac["Pm"]), fac["ld"],
                                                  _ptr(fac["w"]), _ptr(fac["info"]), _ptr(fac["split"]),
                                                  _ptr(fac["scales"]), self._stream()))
        return fac

    def sparse_predict(self, kernel_id, theta, Xu, fac, Xs):
        theta, Xu, Xs = _c(theta), _c(Xu), _c(Xs)
        m, d = Xu.shape
        M = Xs.shape[0]
        mean = self.empty(M, dtype=Xu.dtype)
        sd = self.empty(M, dtype=Xu.dtype)
        if M == 0:
            return mean, sd
        self._check(self.lib.gpg_sparse_predict(self.h, self._dt(Xu), kernel_id, d, _ptr(theta), _ptr(Xu), m,
                                                _ptr(fac["Ui"]), _ptr(fac["Pm"]), fac["ld"], _ptr(fac["w"]),
                                                _ptr(fac.get("split")), _ptr(fac.get("scales")), _ptr(Xs), M,
                                                _ptr(mean), _ptr(sd), self._stream()))
        return mean, sd


    # -- multi-GPU: one process per GPU, the library's own NCCL communicator -------------
    def comm_unique_id(self):
        """128-byte NCCL unique id (bytes); rank 0 draws it, every rank passes it to comm_init."""
        buf = C.create_string_buffer(128)
        self._check(self.lib.gpg_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.gpg_comm_init(self.h, int(nranks), int(rank), buf))

    def comm_destroy(self):
        self._check(self.lib.gpg_comm_destroy(self.h))

    def comm_info(self):
        """(nranks, rank) of the handle's communicator; (0, -1) without one."""
        n, r = _i32(0), _i32(-1)
        self._check(self.lib.gpg_comm_info(self.h, C.byref(n), C.byref(r)))
        return n.value, r.value

    def predict_uses_planes(self, dtype, N, have_planes=True):
        """gpg_predict's routing rule: True when it reads the fp16 planes of the cache, False when it reads Linv."""
        code = GPG_F32 if dtype == torch.float32 else GPG_F64
        return bool(self.lib.gpg_predict_uses_planes(self.h, code, int(N), int(bool(have_planes))))

    def bcast_factor(self, theta, X, fac, root=0):
        """Broadcast rank `root`'s factor cache (and theta, X) into the same-shaped buffers of every rank."""
        assert theta.is_contiguous() and X.is_contiguous()
        N, d = X.shape
        self._check(self.lib.gpg_bcast_factor(self.h, self._dt(X), d, N, fac["ld"], _ptr(theta), _ptr(X), _ptr(fac.get("Linv")),
                                              _ptr(fac["alpha"]), _ptr(fac.get("wsplit")), _ptr(fac.get("scales")),
                                              _ptr(fac["info"]), int(root), self._stream()))
        return fac

    def allgather_pred(self, local, out=None):
        """out[r] = rank r's `local` (same element count on every rank)."""
        local = local.contiguous()
        nranks, _ = self.comm_info()
        if out is None:
            out = torch.empty((nranks,) + tuple(local.shape), dtype=local.dtype, device=self.device)
        self._check(self.lib.gpg_allgather_pred(self.h, self._dt(local), _ptr(local), local.numel(), _ptr(out), self._stream()))
        return out

    def predict_sharded(self, kernel_id, theta, X, fac, Xs_local, M_pad, root=0, gather=True, pred_local=None, pred_all=None):
        """gpg_predict_sharded: broadcast of root's cache (pipelined) + this rank's tile + one all-gather.
        Returns (pred_local [2, M_pad], pred_all [nranks, 2, M_pad] or None)."""
        assert theta.is_contiguous() and X.is_contiguous()
        Xs_local = _c(Xs_local)
        N, d = X.shape
        M_local = Xs_local.shape[0]
        nranks, _ = self.comm_info()
        if pred_local is None:
            pred_local = torch.zeros(2, M_pad, dtype=X.dtype, device=self.device)
        if gather and pred_all is None:
            pred_all = torch.empty(nranks, 2, M_pad, dtype=X.dtype, device=self.device)
        self._check(self.lib.gpg_predict_sharded(self.h, self._dt(X), kernel_id, d, _ptr(theta), _ptr(X), N,
                                                 _ptr(fac.get("Linv")), fac["ld"], _ptr(fac["alpha"]), _ptr(fac.get("wsplit")),
                                                 _ptr(fac.get("scales")), _ptr(fac["info"]), int(root),
                                                 _ptr(Xs_local) if M_local else C.c_void_p(0), M_local, _ptr(pred_local),
                                                 int(M_pad), _ptr(pred_all) if gather else C.c_void_p(0), self._stream()))
        return pred_local, (pred_all if gather else None)

    def acq_sweep_sharded(self, acq_id, mean_local, sd_local, idx_offset, k, mu_best=0.0, xi=0.01, alpha=0.0, beta=1.0,
                          mask_local=None, want_acq=False):
        """Global top-k of the acquisition function over a grid whose tiles live on the ranks (global flat indices)."""
        mean_local, sd_local, mask_local = _c(mean_local), _c(sd_local), _c(mask_local)
        M_local = mean_local.numel()
        vals = self.empty(k, dtype=mean_local.dtype)
        idx = torch.empty(k, dtype=torch.int64, device=self.device)
        count = torch.zeros(1, dtype=torch.int32, device=self.device)
        acq = self.empty(M_local, dtype=mean_local.dtype) if want_acq else None
        null = C.c_void_p(0)
        self._check(self.lib.gpg_acq_sweep_sharded(self.h, self._dt(mean_local), acq_id, _ptr(mean_local) if M_local else null,
                                                   _ptr(sd_local) if M_local else null, _ptr(mask_local), M_local,
                                                   int(idx_offset), float(mu_best), float(xi), float(alpha), float(beta),
                                                   int(k), _ptr(vals), _ptr(idx), _ptr(count), _ptr(acq), self._stream()))
        return vals, idx, count, acq


_ENGINES = {}


def get_engine(device=None):
    """Process-wide engine per device (the reference is single-threaded with global state too, gpr.py:103-113)."""
    if not torch.cuda.is_available():
        return Engine(device)                         # raises with the explanatory message
    idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    if idx not in _ENGINES:
        _ENGINES[idx] = Engine(idx)
    return _ENGINES[idx]
