"""Host-side logic of the multi-GPU shard (gpim_b200/sharded.py) on the gloo backend, world_size 2,
CPU tensors, with a stand-in tile predictor (the CUDA engine is not needed for the plumbing)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpim_b200 import sharded


def test_tile_bounds_partition_exactly():
    for M in (0, 1, 7, 64, 300, 65537, 128 * 128 - 37):
        for world in (1, 2, 3, 8):
            edges = [sharded.tile_bounds(M, world, r) for r in range(world)]
            width = sharded.tile_width(M, world)
            assert edges[0][0] == 0 and edges[-1][1] == M
            for r, (a, b) in enumerate(edges):
                assert 0 <= b - a <= width and a == min(M, r * width)      # back to back, common padded width
                assert a % 128 == 0 or a == M                              # edges on the engine's 128-row groups
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c


def test_cyclic_rows_partition_and_merge():
    for M in (1, 127, 128, 129, 1000, 128 * 128 - 37, 65536):
        for world in (1, 2, 3, 8):
            w = sharded.tile_width(M, world)
            allp = torch.zeros(world, 2, max(w, 1))
            seen = torch.zeros(M)
            for r in range(world):
                idx = sharded.cyclic_rows(M, world, r)
                assert len(idx) <= w
                assert all(int(i) // 128 % world == r for i in idx[::128])      # whole 128-row groups, dealt round-robin
                allp[r, 0, :len(idx)] = idx.double().float()
                allp[r, 1, :len(idx)] = -idx.double().float()
                seen[idx] += 1
            assert bool((seen == 1).all())
            m = sharded.cyclic_merge(allp, M, world)
            assert torch.equal(m[0], torch.arange(M).float()) and torch.equal(m[1], -torch.arange(M).float())


def _worker(rank, world, port, M, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = 5
    Xs = torch.arange(M * 2, dtype=torch.float64).reshape(M, 2)

    def factorize():
        return {"Linv": torch.full((N, N), 3.0, dtype=torch.float64), "alpha": torch.arange(N, dtype=torch.float64)}

    def alloc():
        return {"Linv": torch.zeros(N, N, dtype=torch.float64), "alpha": torch.zeros(N, dtype=torch.float64)}

    def tile(fac, X):
        # depends on the broadcast factor AND on the tile rows
        return X[:, 0] * fac["Linv"][0, 0] + fac["alpha"].sum(), X[:, 1] + fac["Linv"].sum()

    mean, sd = sharded.predict_sharded(factorize, alloc, tile, Xs)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.stack([mean.numpy(), sd.numpy()]))

    # the inducing-point cache {Ui, Pm, w} (reconstructor(sparse=True)) shards the same way
    def sfactorize():
        return {"Ui": torch.full((N, N), 2.0, dtype=torch.float64), "Pm": torch.eye(N, dtype=torch.float64),
                "w": torch.arange(N, dtype=torch.float64), "ld": N}

    def salloc():
        return {"Ui": torch.zeros(N, N, dtype=torch.float64), "Pm": torch.zeros(N, N, dtype=torch.float64),
                "w": torch.zeros(N, dtype=torch.float64), "ld": N}

    def stile(fac, X):
        return X[:, 0] * fac["Ui"][1, 1] + fac["w"].sum(), X[:, 1] + fac["Pm"].sum()

    mean, sd = sharded.predict_sharded(sfactorize, salloc, stile, Xs)
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.stack([mean.numpy(), sd.numpy()]))
    dist.destroy_process_group()


class _StandInEngine:
    """CPU stand-in with the Engine methods predict_model_sharded touches."""
    device = torch.device("cpu")

    def alloc_factor(self, N, dtype, with_L=True):
        return {"Linv": torch.zeros(N, N, dtype=dtype), "alpha": torch.zeros(N, dtype=dtype)}

    def alloc_sparse_factor(self, m, dtype):
        return {"Ui": torch.zeros(m, m, dtype=dtype), "Pm": torch.zeros(m, m, dtype=dtype), "w": torch.zeros(m, dtype=dtype)}

    def predict(self, kid, theta, X, fac, Xs):
        return Xs[:, 0] * theta[0] + fac["alpha"].sum(), Xs[:, 1] + fac["Linv"].sum()

    def sparse_predict(self, kid, theta, Xu, fac, Xs):
        return Xs[:, 0] * theta[0] + fac["w"].sum() + Xu.sum(), Xs[:, 1] + fac["Ui"].sum()


class _StandInModel:
    def __init__(self, rank, sparse):
        import types
        self.engine = _StandInEngine()
        self.kernel = types.SimpleNamespace(dtype=torch.float64, kernel_id=0)
        self._X = torch.zeros(4, 2, dtype=torch.float64)
        self._theta = torch.full((5,), 2.0 if rank == 0 else -7.0, dtype=torch.float64)     # ranks disagree before the call
        self._factor = None
        if sparse:
            self._Xu = torch.full((3, 2), 1.0 if rank == 0 else 9.0, dtype=torch.float64)
            self.Xu = self._Xu

    def factor(self, check=True):
        if hasattr(self, "_Xu"):
            return {"Ui": torch.full((3, 3), 2.0, dtype=torch.float64), "Pm": torch.eye(3, dtype=torch.float64),
                    "w": torch.arange(3, dtype=torch.float64)}, True
        return {"Linv": torch.full((4, 4), 3.0, dtype=torch.float64), "alpha": torch.arange(4, dtype=torch.float64)}, True


def _model_worker(rank, world, port, M, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Xs = torch.arange(M * 2, dtype=torch.float64).reshape(M, 2)
    for tag, sparse in (("e", False), ("s", True)):
        mean, sd = sharded.predict_model_sharded(_StandInModel(rank, sparse), Xs)
        np.save(os.path.join(out_dir, f"{tag}{rank}.npy"), np.stack([mean.numpy(), sd.numpy()]))
    dist.destroy_process_group()


def test_predict_model_sharded_world2_gloo(tmp_path):
    """Model-level helper: rank 0's hyper-parameters and inducing inputs win, cache broadcast, tiles gathered."""
    M = 21
    port = 29900 + os.getpid() % 90
    mp.spawn(_model_worker, args=(2, port, M, str(tmp_path)), nprocs=2, join=True)
    Xs = np.arange(M * 2, dtype=np.float64).reshape(M, 2)
    want_e = np.stack([Xs[:, 0] * 2.0 + 6.0, Xs[:, 1] + 48.0])
    want_s = np.stack([Xs[:, 0] * 2.0 + 3.0 + 6.0, Xs[:, 1] + 18.0])
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"e{r}.npy"), want_e)
        np.testing.assert_array_equal(np.load(tmp_path / f"s{r}.npy"), want_s)


@pytest.mark.parametrize("M", [64, 37])
def test_sharded_predict_world2_gloo(tmp_path, M):
    port = 29600 + (os.getpid() + M) % 300
    mp.spawn(_worker, args=(2, port, M, str(tmp_path)), nprocs=2, join=True)
    Xs = np.arange(M * 2, dtype=np.float64).reshape(M, 2)
    want = np.stack([Xs[:, 0] * 3.0 + 10.0, Xs[:, 1] + 75.0])
    swant = np.stack([Xs[:, 0] * 2.0 + 10.0, Xs[:, 1] + 5.0])
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"r{r}.npy"), want)
        np.testing.assert_array_equal(np.load(tmp_path / f"s{r}.npy"), swant)


# ---------------------------------------------------------------------------------------------
# failure and non-zero source rank on the generic transport (gloo, CPU)
# ---------------------------------------------------------------------------------------------
def _edge_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, M = 4, 9
    Xs = torch.arange(M * 2, dtype=torch.float64).reshape(M, 2)
    calls = []

    def factorize_bad():                      # a failed factorisation must not raise on the source rank alone
        calls.append(rank)
        return {"Linv": torch.ones(N, N, dtype=torch.float64), "alpha": torch.ones(N, dtype=torch.float64),
                "info": torch.tensor([3], dtype=torch.int32)}

    def alloc():
        return {"Linv": torch.zeros(N, N, dtype=torch.float64), "alpha": torch.zeros(N, dtype=torch.float64),
                "info": torch.zeros(1, dtype=torch.int32)}

    def tile(fac, X):
        return X[:, 0] + fac["alpha"].sum(), X[:, 1] * fac["Linv"][0, 0]

    raised = False
    try:
        sharded.predict_sharded(factorize_bad, alloc, tile, Xs, src=1)
    except torch.linalg.LinAlgError as e:
        raised = "order 3" in str(e)
    assert calls == ([1] if rank == 1 else []), "only the source rank factorises"

    def factorize_ok():
        return {"Linv": torch.full((N, N), 2.0, dtype=torch.float64), "alpha": torch.arange(N, dtype=torch.float64),
                "info": torch.zeros(1, dtype=torch.int32)}

    mean, sd = sharded.predict_sharded(factorize_ok, alloc, tile, Xs, src=1)      # rank 1 is the source
    # a sub-group that excludes global rank 0 would need get_global_rank; here: the world group by explicit handle
    mean2, _ = sharded.predict_sharded(factorize_ok, alloc, tile, Xs, group=dist.group.WORLD, src=1)
    # candidate merge in the reference's order (descending value, NaN first, ties -> larger flat index)
    vals = [np.array([5.0, 3.0, 3.0]), np.array([np.nan, 3.0, 1.0])][rank]
    idx = [np.array([10, 4, 2]), np.array([7, 9, -1])][rank]
    mv, mi = sharded.topk_merge_generic(vals, idx, 4)
    np.save(os.path.join(out_dir, f"edge{rank}.npy"),
            np.concatenate([[float(raised)], mean.numpy(), sd.numpy(), mean2.numpy(), mi.astype(np.float64), np.nan_to_num(mv, nan=-99.0)]))
    dist.destroy_process_group()


def test_sharded_failure_is_raised_on_every_rank_and_src_is_honoured(tmp_path):
    M = 9
    port = 29300 + os.getpid() % 200
    mp.spawn(_edge_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    Xs = np.arange(M * 2, dtype=np.float64).reshape(M, 2)
    for r in range(2):
        got = np.load(tmp_path / f"edge{r}.npy")
        assert got[0] == 1.0, "LinAlgError naming the pivot must be raised on every rank"
        np.testing.assert_array_equal(got[1:1 + M], Xs[:, 0] + 6.0)
        np.testing.assert_array_equal(got[1 + M:1 + 2 * M], Xs[:, 1] * 2.0)
        np.testing.assert_array_equal(got[1 + 2 * M:1 + 3 * M], Xs[:, 0] + 6.0)
        np.testing.assert_array_equal(got[1 + 3 * M:1 + 3 * M + 4], [7, 10, 9, 4])
        np.testing.assert_array_equal(got[1 + 3 * M + 4:], [-99.0, 5.0, 3.0, 3.0])


# ---------------------------------------------------------------------------------------------
# native transport on real GPUs (-m gpu; needs 2 devices): sharded == unsharded, bit for bit
# ---------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import workloads as W
    from oracle import gp_oracle as O
    from gpim_b200._lib import get_engine, KERNEL_IDS, ACQ_IDS
    eng = get_engine(rank)
    res = {}
    for name, dtype in (("c1k", torch.float32), ("c1k", torch.float64)):
        wl = W.make_workload(name)
        X, y = O.training_rows(O.sparse_grid(wl["R"]), wl["R"])
        Xs = W.rows_of(wl["Xfull"])[: 128 * 128 - 37]                      # ragged tiles
        kid = KERNEL_IDS[wl["kernel"]]
        th = torch.tensor(wl["theta"], dtype=dtype).cuda()
        if rank != 0:
            th = th * 0 + 1.0                                              # the source's theta must win
        Xd, yd, Xsd = (torch.tensor(a, dtype=dtype).cuda() for a in (X, y, Xs))
        th0 = torch.tensor(wl["theta"], dtype=dtype).cuda()
        fac = eng.factorize(kid, th0, Xd, yd, wl["jitter"])
        m1, s1 = eng.predict(kid, th0, Xd, fac, Xsd)
        tag = "f32" if dtype == torch.float32 else "f64"
        ok = []
        for mode in ("replicate", "broadcast"):                                # every rank factorises / rank 0's cache travels
            thm = th.clone()
            mean, sd, info = sharded.predict_exact_sharded(eng, kid, thm, Xd, yd, wl["jitter"], Xsd, src=0, factor=mode)
            assert int(info.item()) == 0
            ok += [bool(torch.equal(mean, m1)), bool(torch.equal(sd, s1))]
        res[tag] = (all(ok[0::2]), all(ok[1::2]))
        # sharded acquisition sweep == single-device sweep over the whole grid
        lo, hi = sharded.tile_bounds(Xs.shape[0], world, rank)
        v0, i0, c0, _ = eng.acq_sweep(ACQ_IDS["ei"], m1, s1, 100, mu_best=float(m1.max()), xi=0.01)
        v1, i1, c1, _ = sharded.acq_topk_sharded(eng, ACQ_IDS["ei"], m1[lo:hi], s1[lo:hi], lo, 100,
                                                 mu_best=float(m1.max()), xi=0.01)
        res[tag + "_acq"] = (bool(torch.equal(i0, i1)), bool(torch.equal(v0, v1)), int(c0.item()) == int(c1.item()))
    # behind the kept API: reconstructor.predict shards by itself under NCCL
    import gpim
    wl = W.make_workload("c1k")
    rec = gpim.reconstructor(gpim.utils.get_sparse_grid(wl["R"]), wl["R"], wl["Xfull"], kernel="RBF", iterations=0,
                             verbose=0, precision="single", jitter=wl["jitter"], lengthscale=[[1., 1.], [20., 20.]])
    t = wl["theta"]
    rec.model.set_theta(t[0], t[3:], t[1], t[2])
    m_sh, s_sh = rec.predict(verbose=0)
    rec.shard = False
    rec.model._factor = None
    m_un, s_un = rec.predict(verbose=0)
    res["api"] = (bool(np.array_equal(m_sh, m_un)), bool(np.array_equal(s_sh, s_un)))
    np.save(os.path.join(out_dir, f"nccl{rank}.npy"), res, allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_equals_unsharded_on_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29100 + os.getpid() % 200
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = np.load(tmp_path / f"nccl{r}.npy", allow_pickle=True).item()
        for k, v in res.items():
            assert all(v), (r, k, v)
