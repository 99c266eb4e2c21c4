"""
Tile-sharded dense-grid prediction over the GPUs of one box.

Test points are independent given the factor cache {Linv, alpha, theta, X}: rank 0 alone
assembles K and factorises it, ONE broadcast ships the cache to the other ranks (NCCL over
NVLink when the tensors live on GPUs), every rank predicts its contiguous tile of X_full rows and
ONE all-gather returns (mean, sd).  The reference has no multi-device path at all; the
single-device semantics being sharded are those of reconstructor.predict (gpr.py:219-255).

The communication plumbing is written against torch.distributed only, so the same code runs on
the gloo backend with CPU tensors (tests/test_sharded.py drives it with a stand-in tile predictor).
"""
import torch
import torch.distributed as dist


def tile_bounds(M, world, rank):
    """Contiguous row tile [lo, hi) of rank `rank`; the first M % world ranks get one extra row."""
    base, extra = divmod(int(M), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_factor(fac, src=0, group=None):
    """Broadcast the tensors of a factor cache in place (every rank passes same-shaped buffers)."""
    # the tcgen05 predict path (f32, N >= 1024: gpg_predict's routing rule) reads only the fp16 planes of Linv;
    # the fp32 matrix travels only when the SIMT kernels will consume it
    planes = fac.get("wsplit")
    if "Ui" in fac:                  # inducing-point cache (gpg_sparse_factorize): two m x m factors and a vector
        keys = ("Ui", "Pm", "w", "split", "scales")
    elif planes is not None and planes.shape[1] >= 1024:
        keys = ("alpha", "wsplit", "scales")
    else:
        keys = ("Linv", "alpha", "wsplit", "scales")
    for key in keys:
        if fac.get(key) is not None:
            dist.broadcast(fac[key], src=src, group=group)
    return fac


def gather_tiles(local, M, group=None):
    """All-gather the per-rank tiles of a length-M vector (tiles as produced by tile_bounds)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    base, extra = divmod(int(M), world)
    if extra == 0:
        out = torch.empty(M, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    width = base + 1                                    # pad ragged tiles to a common width
    buf = torch.zeros(width, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world):
        lo, hi = tile_bounds(M, world, r)
        parts.append(out[r * width: r * width + (hi - lo)])
    del rank
    return torch.cat(parts)


def predict_sharded(factorize_fn, alloc_fn, predict_tile_fn, Xs, group=None):
    """
    factorize_fn() -> fac      run on rank 0 only (K assembly + Cholesky + inverse + solves)
    alloc_fn() -> fac          empty same-shaped buffers on the other ranks
    predict_tile_fn(fac, Xs_tile) -> (mean_tile, sd_tile)
    Xs: (M, d) test rows, identical on every rank (each rank only reads its tile).
    Returns the full (mean, sd) on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    M = Xs.shape[0]
    fac = factorize_fn() if rank == 0 else alloc_fn()
    if world > 1:
        broadcast_factor(fac, 0, group)
    lo, hi = tile_bounds(M, world, rank)
    mean_t, sd_t = predict_tile_fn(fac, Xs[lo:hi])
    if world == 1:
        return mean_t, sd_t
    return gather_tiles(mean_t, M, group), gather_tiles(sd_t, M, group)


def predict_model_sharded(model, Xs, src=0, group=None):
    """Tile-sharded ``model.predict_sd`` for the model object of a reconstructor / skreconstructor (ExactGPModel,
    SparseGPModel, SKExactGPModel): rank ``src``'s hyper-parameters (and inducing inputs) are broadcast first --
    training is replicas-only, so the ranks may hold different values -- then rank ``src`` factorises, the cache
    is broadcast, every rank predicts its tile of ``Xs`` and the tiles are all-gathered.  Every rank passes a model
    built on the same (X, y); returns (mean, sd) for all of ``Xs`` on every rank."""
    eng = model.engine
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    Xs = Xs.to(eng.device, model.kernel.dtype).contiguous()
    sparse = hasattr(model, "Xu")
    if world > 1:
        dist.broadcast(model._theta, src=src, group=group)
        if sparse:
            dist.broadcast(model._Xu, src=src, group=group)
        model._factor = None
    kid = model.kernel.kernel_id

    def factorize():
        fac, _ = model.factor(check=True)
        return fac

    if sparse:
        alloc = lambda: eng.alloc_sparse_factor(model._Xu.shape[0], model.kernel.dtype)
        tile = lambda fac, X: eng.sparse_predict(kid, model._theta, model._Xu, fac, X)
    else:
        alloc = lambda: eng.alloc_factor(model._X.shape[0], model.kernel.dtype, with_L=False)
        tile = lambda fac, X: eng.predict(kid, model._theta, model._X, fac, X)
    mean, sd = predict_sharded(factorize, alloc, tile, Xs, group)
    shift = getattr(model, "mean_shift", None)          # constant mean of the GPyTorch-semantics model
    return (mean + shift() if shift is not None else mean), sd
