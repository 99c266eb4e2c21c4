"""
CPU tests of the host-side mirror of the reference interface (no GPU, no library calls): grid / data-layout
helpers against the oracle's independent restatement, the kernel factory's prior-draw stream and constraint
bookkeeping, and the candidate filters of boptimizer (boptim.py:303-429) on hand-made inputs.
"""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as O
from gpim_b200 import gprutils
from gpim_b200.kernels import gp_kernels
from gpim_b200.gpbayes.boptim import boptimizer


def _sparse_image(shape, frac, seed):
    rng = np.random.RandomState(seed)
    R = rng.rand(*shape)
    R[rng.rand(*shape) < frac] = np.nan
    return R


# ---------------------------------------------------------------------------------------------
# gprutils (gprutils.py:23-210)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(7, 9), (5, 6, 4), (3, 4, 2, 5)])
def test_full_grid_matches_mgrid(shape):
    R = np.zeros(shape)
    X = gprutils.get_full_grid(R)
    assert X.shape == (len(shape),) + shape
    np.testing.assert_array_equal(X, np.array(np.mgrid[tuple(slice(0, e, 1.0) for e in shape)]))
    np.testing.assert_array_equal(X, O.full_grid(R))
    Xd = gprutils.get_full_grid(R, dense_x=0.5)
    assert Xd.shape[1] == 2 * shape[0] and Xd[0].max() == shape[0] - 0.5


def test_full_grid_extent_and_errors():
    R = np.zeros((10, 20))
    X = gprutils.get_full_grid(R, extent=[[2, 7], [0, 10]])
    assert X[0].min() == 2 and X[0].max() < 7 and X[1].min() == 0 and X[1].max() < 10
    assert X.shape == (2, 10, 20)           # step = 1 / (size // (hi - lo)) -> size points per axis
    with pytest.raises(NotImplementedError):
        gprutils.get_full_grid(np.zeros(5))
    with pytest.raises(NotImplementedError):
        gprutils.get_full_grid(np.zeros((2, 2, 2, 2, 2)))
    with pytest.raises(NotImplementedError):
        gprutils.get_sparse_grid(np.zeros((4, 4)))          # no NaNs: not a sparse image


@pytest.mark.parametrize("shape", [(12, 11), (6, 7, 5)])
def test_sparse_grid_and_training_rows_match_oracle(shape):
    R = _sparse_image(shape, 0.4, 1)
    Xs = gprutils.get_sparse_grid(R)
    np.testing.assert_array_equal(np.isnan(Xs).any(axis=0), np.isnan(R))
    np.testing.assert_array_equal(Xs, O.sparse_grid(R))
    X, y = gprutils.prepare_training_data(Xs, R)
    Xo, yo = O.training_rows(O.sparse_grid(R), R)
    assert X.dtype == torch.float64 and X.shape == (int((~np.isnan(R)).sum()), len(shape))
    np.testing.assert_array_equal(X.numpy(), Xo)
    np.testing.assert_array_equal(y.numpy(), yo)
    X32, y32 = gprutils.prepare_training_data(Xs, R, precision="single")
    assert X32.dtype == torch.float32 and y32.dtype == torch.float32
    Xt = gprutils.prepare_test_data(Xs)
    assert Xt.shape == (R.size, len(shape)) and torch.isnan(Xt).any()      # NaN rows are kept
    np.testing.assert_array_equal(Xt.numpy(), O.to_rows(O.sparse_grid(R)))


def test_whole_spectrum_sparsity_3d():
    """3-D data whose last slice is fully observed: a spectrum with any NaN loses all of its coordinates
    (gprutils.py:195-200)."""
    R = np.random.RandomState(0).rand(4, 5, 6)
    R[1, 2, 3] = np.nan
    R[3, 0, 0] = np.nan
    assert not np.isnan(R[..., -1]).any()
    Xs = gprutils.get_sparse_grid(R)
    gone = np.isnan(Xs).any(axis=0)
    assert gone[1, 2].all() and gone[3, 0].all() and gone.sum() == 12
    np.testing.assert_array_equal(Xs, O.sparse_grid(R))


def test_corrupt_and_edge_helpers():
    R = np.random.RandomState(2).rand(10, 10)
    X = gprutils.get_full_grid(R)
    Xc, Rc = gprutils.corrupt_data_xy(X, R, prob=0.3)
    Xc2, Rc2 = gprutils.corrupt_data_xy(X, R, prob=0.3)
    np.testing.assert_array_equal(np.isnan(Rc), np.isnan(Rc2))               # seeded like the reference (seed 0)
    np.testing.assert_array_equal(np.isnan(Xc).any(axis=0), np.isnan(Rc))
    assert 10 < np.isnan(Rc).sum() < 55
    R3 = np.random.RandomState(3).rand(6, 6, 4)
    X3c, R3c = gprutils.corrupt_data_xy(gprutils.get_full_grid(R3), R3, prob=0.5)
    whole = np.isnan(R3c).reshape(36, 4)
    assert (whole.all(axis=1) | (~whole).all(axis=1)).all()                  # whole spectra go, never parts
    edge = gprutils.open_edge_points(np.full((12, 12), np.nan), np.ones((12, 12)), s=3)
    assert np.nansum(edge) > 0 and np.isnan(edge[5, 5])
    with pytest.raises(NotImplementedError):
        gprutils.corrupt_data_xy(np.zeros((4, 2, 2, 2, 2)), np.zeros((2, 2, 2, 2)))


# ---------------------------------------------------------------------------------------------
# kernel factory (pyro_kernels.py:14-96)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "single"])
def test_kernel_factory_prior_draws_follow_the_reference_stream(precision):
    """variance first, then lengthscale, from torch's CPU generator at the chosen default dtype
    (seed 0 -> double 0.97005, [0.70782, 0.45938]; float 0.49626, [0.76822, 0.08848])."""
    torch.manual_seed(0)
    k = gp_kernels.get_kernel("RBF", 2, [[0., 0.], [12.5, 12.5]], precision=precision)
    if precision == "double":
        np.testing.assert_allclose(float(k.variance), 1e-4 + 0.970053 * (10 - 1e-4), rtol=1e-5)
        np.testing.assert_allclose(k.lengthscale.numpy(), [8.84775, 5.74229], rtol=1e-5)
    else:
        np.testing.assert_allclose(float(k.variance), 1e-4 + 0.49626 * (10 - 1e-4), rtol=1e-4)
        np.testing.assert_allclose(k.lengthscale.numpy(), [12.5 * 0.76822, 12.5 * 0.08848], rtol=1e-3)
    # the oracle draws the same numbers
    R = np.full((25, 25), np.nan)
    R[3, 4], R[10, 20] = 1.0, 2.0
    ora = O.OracleGP(O.sparse_grid(R), R, O.full_grid(R), kernel="RBF", precision=precision, seed=0)
    np.testing.assert_allclose(float(ora.tf_v(ora.u_v).detach()), float(k.variance), rtol=1e-6)
    np.testing.assert_allclose(ora.tf_l(ora.u_l).detach().numpy(), k.lengthscale.numpy(), rtol=1e-6)


def test_kernel_factory_bookkeeping():
    torch.manual_seed(1)
    k = gp_kernels.get_kernel("Matern52", 3, [[1., 1., 1.], [20., 20., 20.]], amplitude=[0.5, 2.0])
    assert not k.isotropic and k.n_ls == 3 and k.bounds() == [0.5, 2.0, 1., 1., 1., 20., 20., 20.]
    u = k.pack_u()
    assert u.shape == (6,) and float(u[1]) == 0.0 and float(u[2]) == 0.0          # noise = 1, scale_mixture = 1
    th = k.pack_theta()
    assert th.shape == (6,) and 0.5 <= float(th[0]) <= 2.0 and float(th[1]) == 1.0 and (th[3:] >= 1).all() and (th[3:] <= 20).all()
    k.unpack_u(u + 0.1)
    np.testing.assert_allclose(float(k.u_noise.exp()), np.exp(0.1))               # positive constraint = exp
    ki = gp_kernels.get_kernel("RBF", 2, [0., 6.0])
    assert ki.isotropic and ki.n_ls == 1 and ki.pack_theta().shape == (5,)
    assert float(ki.pack_theta()[3]) == float(ki.pack_theta()[4])                  # one lengthscale, repeated
    with pytest.raises(KeyError):
        gp_kernels.get_kernel("Periodic", 2, [[0., 0.], [1., 1.]])


def test_to_constrained_interval_accepts_both_spellings():
    u_l, u_v = torch.tensor([0.0, 2.0]), torch.tensor(-1.0)
    for key in ("lenghtscale_map_unconstrained", "lengthscale_map_unconstrained"):
        l, a = gprutils.to_constrained_interval({key: u_l, "variance_map_unconstrained": u_v}, [1., 4.], [1e-4, 10.])
        np.testing.assert_allclose(l.numpy(), 1 + 3 / (1 + np.exp(-u_l.numpy())), rtol=1e-6)
        np.testing.assert_allclose(float(a), 1e-4 + (10 - 1e-4) / (1 + np.e), rtol=1e-6)


# ---------------------------------------------------------------------------------------------
# boptimizer candidate logic (boptim.py:303-429) without a surrogate model
# ---------------------------------------------------------------------------------------------
def _bare_boptimizer(**kw):
    bo = object.__new__(boptimizer)
    defaults = dict(verbose=0, mask=None, batch_size=100, batch_out_max=3, dscale=None, gamma=0.8, points_mem=10,
                    exit_strategy=1, indices_all=[], vals_all=[])
    defaults.update(kw)
    for k, v in defaults.items():
        setattr(bo, k, v)
    return bo


def test_host_candidates_order_matches_reversed_argsort():
    rng = np.random.RandomState(0)
    acq = rng.rand(6, 7)
    acq[2, 3] = acq[4, 1]                                     # a tie: the larger flat index ranks first
    bo = _bare_boptimizer(batch_size=10)
    vals, idx = bo._host_candidates(acq, acq)
    order = np.argsort(acq.ravel())[::-1][:10]
    np.testing.assert_array_equal(vals, acq.ravel()[order])
    np.testing.assert_array_equal(idx, np.stack(np.unravel_index(order, acq.shape), axis=1))
    mask = np.ones_like(acq)
    mask[:3] = np.nan                                        # rows 0..2 are off limits
    bo = _bare_boptimizer(batch_size=5, mask=mask)
    vals, idx = bo._host_candidates(acq, acq)
    assert len(vals) == 5 and all(i[0] >= 3 for i in idx) and not np.isnan(vals).any()
    assert vals == sorted(vals, reverse=True)


def test_checkvalues_skips_visited_and_too_close_points():
    cand = [[5, 5], [5, 6], [9, 9], [0, 0]]
    vals = [4.0, 3.0, 2.0, 1.0]
    assert _bare_boptimizer().checkvalues(cand, vals) == ([5, 5], 4.0)                       # nothing visited yet
    assert _bare_boptimizer(indices_all=[[5, 5]]).checkvalues(cand, vals) == ([5, 6], 3.0)   # already measured
    bo = _bare_boptimizer(indices_all=[[5, 4]], dscale=3.0)
    assert bo.checkvalues(cand, vals) == ([9, 9], 2.0)                                       # within dscale of the last pick
    bo = _bare_boptimizer(indices_all=[[5, 4], [20, 20]], dscale=3.0, gamma=0.1)
    assert bo.checkvalues(cand, vals) == ([5, 5], 4.0)         # the older pick only counts with weight gamma**1
    bo = _bare_boptimizer(indices_all=[[5, 5], [5, 6], [9, 9], [0, 0]], exit_strategy=0)
    assert bo.checkvalues(cand, vals) == ([0, 0], 1.0)         # list exhausted: exit_strategy 0 -> last element


def test_update_points_ball_suppression():
    np.random.seed(0)
    cand = [[10, 10], [10, 11], [11, 10], [20, 20], [20, 21], [0, 0]]
    vals = [6.0, 5.0, 4.0, 3.0, 2.0, 1.0]
    bo = _bare_boptimizer(batch_out_max=3)
    v, idx = bo.update_points(vals, cand, dscale=2.5)
    assert idx == [[10, 10], [20, 20], [0, 0]] and v == [6.0, 3.0, 1.0]
    bo = _bare_boptimizer(batch_out_max=5)                    # fewer survivors than requested: random padding
    v, idx = bo.update_points(vals, cand, dscale=2.5)
    assert len(idx) == 5 and idx[:3] == [[10, 10], [20, 20], [0, 0]] and all(i in cand for i in idx[3:])


def test_checkpoint_carries_the_inducing_inputs_of_a_sparse_surrogate(tmp_path):
    """save_results() / resume() with a sparse surrogate: the trained inducing inputs are part of the engine state
    (host logic only: the surrogate is a stand-in object with the members the two methods touch)."""
    import types
    import torch

    def surrogate(xu_value):
        model = types.SimpleNamespace(_u=torch.tensor([0.1, -0.2, 0.0, 0.3, 0.4], dtype=torch.float64),
                                      Xu=torch.full((3, 2), xu_value, dtype=torch.float64), X=None, y=None)
        model.load_unconstrained = lambda u: setattr(model, "_u", torch.as_tensor(np.asarray(u)))
        return types.SimpleNamespace(model=model, train=lambda **kw: None)

    y = np.full((4, 4), np.nan)
    y[1, 2] = 0.5
    y[3, 0] = -1.0
    common = dict(filename=str(tmp_path / "bo"), extent=None, precision="double", exploration_steps=4,
                  gp_predictions=[(np.zeros((4, 4)), np.ones((4, 4)))], target_func_vals=[y.copy()],
                  indices_all=[[1, 2]], vals_all=[0.7])
    first = _bare_boptimizer(surrogate_model=surrogate(2.5), **common)
    first.save_results()
    saved = np.load(str(tmp_path / "bo.npy"), allow_pickle=True).item()
    assert {"gp_pred", "func_val", "inds_all", "vals_all"} <= set(saved)             # the reference's four keys
    np.testing.assert_array_equal(saved["engine_state"]["Xu"], np.full((3, 2), 2.5))
    second = _bare_boptimizer(surrogate_model=surrogate(-9.0), **{**common, "gp_predictions": [], "target_func_vals": [],
                                                                  "indices_all": [], "vals_all": []})
    second.resume()
    np.testing.assert_array_equal(second.surrogate_model.model.Xu.numpy(), np.full((3, 2), 2.5))
    np.testing.assert_allclose(second.surrogate_model.model._u.numpy(), [0.1, -0.2, 0.0, 0.3, 0.4])
    assert second.indices_all == [[1, 2]] and second._first_step == 1
    assert tuple(second.surrogate_model.model.X.shape) == (2, 2)                     # the two measured pixels


def test_resume_from_a_checkpoint_written_before_any_step_starts_at_step_zero(tmp_path):
    """a checkpoint with no completed step must not skip step 0 (and its initial training); one-element
    lengthscale bounds on d > 1 inputs pack as ONE shared lengthscale."""
    import types
    import torch
    from gpim_b200.kernels import gp_kernels
    trained = []
    model = types.SimpleNamespace(_u=torch.zeros(5, dtype=torch.float64), X=None, y=None)
    model.load_unconstrained = lambda u: setattr(model, "_u", torch.as_tensor(np.asarray(u)))
    sur = types.SimpleNamespace(model=model, train=lambda **kw: trained.append(1), hyperparams={"noise": []})
    y = np.full((4, 4), np.nan)
    y[1, 2] = 0.5
    common = dict(filename=str(tmp_path / "bo0"), extent=None, precision="double", exploration_steps=3,
                  gp_predictions=[], target_func_vals=[y.copy()], indices_all=[], vals_all=[], surrogate_model=sur)
    first = _bare_boptimizer(**common)
    first.save_results()
    second = _bare_boptimizer(**common)
    second.resume()
    assert second._first_step == 0 and not second._skip_initial_training
    k = gp_kernels.get_kernel("RBF", 3, [[1.0], [5.0]])
    assert k.isotropic and k.n_ls == 1 and k.pack_theta().numel() == 3 + 3 and k.pack_u().numel() == 3 + 1
    assert len(k.bounds()) == 2 + 2
    with pytest.raises(ValueError):
        gp_kernels.get_kernel("RBF", 3, [[1.0, 1.0], [5.0, 5.0]])


def _random_ranked_list(rng, n, shape):
    """n distinct grid points with descending values (a few ties), as next_point returns them"""
    flat = rng.choice(int(np.prod(shape)), size=n, replace=False)
    idx = np.stack(np.unravel_index(flat, shape), axis=1).tolist()
    vals = np.sort(rng.integers(0, max(3, n // 2), size=n).astype(float) / 7.0)[::-1].tolist()
    return vals, idx


def test_oracle_batch_suppression_against_an_independent_kdtree_route():
    """oracle.suppress_batch (the restatement of boptim.py:326-376) against a second implementation built on
    scipy's cKDTree ball queries, and the product's host path against both, on random ranked lists."""
    from scipy import spatial
    from oracle import gp_oracle as O
    rng = np.random.default_rng(5)
    for trial in range(40):
        shape = (17, 23) if trial % 2 == 0 else (7, 9, 5)
        n = int(rng.integers(5, 60))
        vals, idx = _random_ranked_list(rng, n, shape)
        visited = [idx[int(q)] for q in rng.choice(n, size=int(rng.integers(0, 4)), replace=False)]
        radius = float(rng.choice([1.0, 2.0, 2.5, 4.0, 6.3]))
        ds = None if trial % 3 == 0 else float(rng.choice([1.0, 3.0]))
        bmax = int(rng.choice([3, 200]))                 # 200: every greedy pick is kept, the rest is random padding
        np.random.seed(trial)
        ov, oi = O.suppress_batch(list(vals), [list(p) for p in idx], visited, radius, bmax, dscale=ds)
        # independent route: cut at the first admissible value, greedy over a KD-tree
        _, v0 = O.pick_unvisited(idx, vals, visited, ds)
        start = vals.index(v0)
        pts = np.array(idx[start:], dtype=float)
        tree = spatial.cKDTree(pts)
        alive = np.ones(len(pts), dtype=bool)
        want = []
        for q in range(len(pts)):
            if alive[q]:
                want.append(idx[start + q])
                alive[tree.query_ball_point(pts[q], radius)] = False
        want = want[:bmax]
        assert len(oi) == bmax and oi[:len(want)] == want
        assert ov[:len(want)] == [vals[start + idx[start:].index(p)] for p in want]
        bo = _bare_boptimizer(indices_all=[list(p) for p in visited], dscale=ds, batch_out_max=bmax)
        np.random.seed(trial)                            # same draws for the random padding
        hv, hi = bo.update_points(list(vals), [list(p) for p in idx], radius)
        assert hi == oi and hv == ov
