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
    for M in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            edges = [sharded.tile_bounds(M, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == M
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and b - a >= d - c >= 0 and (b - a) - (d - c) <= 1


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
