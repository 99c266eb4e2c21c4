"""
CPU test of the HOST side of the inducing-point path: ``gpim.reconstructor(sparse=True)`` with the engine replaced
by a stand-in that answers the three sparse entry points through oracle/sparse_oracle.py's functional forms.  What
is under test is the Python glue around the C-ABI -- inducing-point selection (gpr.py:145-151), the packing of the
unconstrained parameters and bounds, the trajectory layout, hyperparams["inducing_points"], warm restarts -- not
the arithmetic (that is tests/test_gpu_sparse.py, on the GPU, through libgpgrid.so).
"""
import numpy as np
import torch
from torch.distributions import constraints, transform_to

import workloads as W
from oracle import gp_oracle as O
from oracle.sparse_oracle import SparseOracleGP, vfe_loss, vfe_predict

NAMES = {0: "RBF", 1: "Matern52", 2: "RationalQuadratic"}


class StandInEngine:
    """Same call signatures as gpim_b200._lib.Engine's sparse_* methods, CPU tensors."""
    device = torch.device("cpu")

    def _theta(self, u, bounds, n_ls, d):
        tf_v = transform_to(constraints.interval(torch.tensor(bounds[0], dtype=u.dtype), torch.tensor(bounds[1], dtype=u.dtype)))
        lo = torch.tensor(bounds[2:2 + n_ls], dtype=u.dtype)
        hi = torch.tensor(bounds[2 + n_ls:2 + 2 * n_ls], dtype=u.dtype)
        tf_l = transform_to(constraints.interval(lo, hi))
        ls = tf_l(u[3:3 + n_ls])
        return tf_v(u[0]), u[1].exp(), u[2].exp(), (ls.expand(d) if n_ls == 1 else ls)

    def sparse_fit_adam(self, kernel_id, X, y, Xu, jitter, u, bounds, n_ls, iters, lr, record_xu=True):
        d = X.shape[1]
        is_rq = kernel_id == 2
        up = u.clone().requires_grad_(True)
        xp = Xu.clone().requires_grad_(True)
        opt = torch.optim.Adam([up, xp], lr=lr)
        traj = torch.zeros(max(iters, 1), 4 + d, dtype=X.dtype)
        xu_traj = torch.zeros(max(iters, 1), *Xu.shape, dtype=X.dtype)
        for it in range(iters):
            opt.zero_grad()
            v, n, a, ls = self._theta(up, bounds, n_ls, d)
            loss = vfe_loss(NAMES[kernel_id], X, y, xp, v, ls, n, a if is_rq else torch.ones((), dtype=X.dtype), jitter)
            loss.backward()
            if not is_rq:
                up.grad[2] = 0.0
            opt.step()
            with torch.no_grad():
                v, n, a, ls = self._theta(up, bounds, n_ls, d)
                traj[it] = torch.cat([v.reshape(1), n.reshape(1), a.reshape(1), ls, loss.detach().reshape(1)])
                xu_traj[it] = xp.detach()
        with torch.no_grad():
            u.copy_(up)
            Xu.copy_(xp)
            v, n, a, ls = self._theta(u, bounds, n_ls, d)
            theta = torch.cat([v.reshape(1), n.reshape(1), a.reshape(1), ls])
        return traj[:iters], (xu_traj[:iters] if record_xu else None), theta, torch.zeros(1, dtype=torch.int32)

    def sparse_factorize(self, kernel_id, theta, X, y, Xu, jitter):
        return {"args": (kernel_id, theta.clone(), X, y, Xu.clone(), jitter), "info": torch.zeros(1, dtype=torch.int32)}

    def sparse_predict(self, kernel_id, theta, Xu, fac, Xs):
        kid, th, X, y, Xu0, jitter = fac["args"]
        loc, var = vfe_predict(NAMES[kid], X, y, Xu0, Xs, th[0], th[3:], th[1], th[2], jitter)
        return loc, var.sqrt()


def _reconstructor(monkeypatch, *args, **kwargs):
    from gpim_b200.gpreg import gpr
    monkeypatch.setattr(gpr, "get_engine", lambda device=None: StandInEngine())
    return gpr.reconstructor(*args, **kwargs)


def test_sparse_reconstructor_glue_follows_the_oracle(monkeypatch):
    R = W.dummy_blob(16, 100)
    Xs, Xf = O.sparse_grid(R), O.full_grid(R)
    for kernel, iso in (("RBF", False), ("RationalQuadratic", True)):
        ls = [1.0, 8.0] if iso else [[1.0, 1.0], [8.0, 8.0]]
        kw = dict(kernel=kernel, lengthscale=ls, learning_rate=0.1, iterations=6, seed=1, isotropic=iso)
        ref = SparseOracleGP(Xs, R, Xf, indpoints=9, **kw)
        m0, s0, hp0 = ref.run()
        rec = _reconstructor(monkeypatch, Xs, R, Xf, sparse=True, indpoints=9, verbose=0, **kw)
        n = rec.model.X.shape[0]
        np.testing.assert_array_equal(rec.model.Xu.numpy(), rec.model.X.numpy()[::n // 9])     # gpr.py:151
        m1, s1, hp1 = rec.run()
        assert m1.shape == s1.shape == R.shape
        np.testing.assert_allclose(m1, m0, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(s1, s0, rtol=1e-9)
        for key in ("variance", "noise", "lengthscale"):
            np.testing.assert_allclose(np.array(hp1[key]), np.array(hp0[key]), rtol=1e-10)
        assert len(hp1["inducing_points"]) == 6 and hp1["inducing_points"][0].shape == (len(ref.Xu), 2)
        np.testing.assert_allclose(np.array(hp1["inducing_points"]), np.array(hp0["inducing_points"]), rtol=0, atol=1e-10)
        # warm restart: trained hyper-parameters and inducing inputs, fresh optimiser (gpr.py:184-185)
        ref.train(iterations=3)
        rec.train(iterations=3)
        np.testing.assert_allclose(np.array(hp1["noise"][-3:]), np.array(ref.noise_all[-3:]), rtol=1e-10)
        assert len(hp1["inducing_points"]) == 9


def test_sparse_default_and_capped_inducing_point_counts(monkeypatch):
    R = W.dummy_blob(16, 100)
    Xs, Xf = O.sparse_grid(R), O.full_grid(R)
    n = int((~np.isnan(R)).sum())
    rec = _reconstructor(monkeypatch, Xs, R, Xf, sparse=True, verbose=0)
    assert rec.model.Xu.shape[0] == len(range(0, n, n // (n // 10)))                    # default len(X) // 10
    rec = _reconstructor(monkeypatch, Xs, R, Xf, sparse=True, indpoints=10 ** 6, verbose=0)
    assert rec.model.Xu.shape[0] == n                                                    # capped, gpr.py:149-150
    assert [p.shape for p in rec.model.parameters()][1] == (n, 2)                        # Xu is a trained parameter
