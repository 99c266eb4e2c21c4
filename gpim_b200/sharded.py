"""
Tile-sharded dense-grid prediction and acquisition sweep over the GPUs of one box.

Test points are independent given the factor cache {Linv, alpha, theta, X}: rank ``src`` alone assembles K
and factorises it (training / Cholesky are replicas-only), the cache is broadcast, every rank predicts its
contiguous tile of X_full rows and ONE all-gather returns (mean, sd).  The reference has no multi-device path
at all; the single-device semantics being sharded are those of reconstructor.predict
(gpr.py:219-255) and boptimizer.next_point (boptim.py:278-324).

Two transports:

* **native** (CUDA tensors, exact GP): ``gpg_predict_sharded`` / ``gpg_acq_sweep_sharded`` of libgpgrid.so over
  the library's own NCCL communicator (``ensure_comm`` builds it from a ``torch.distributed`` group: the 128-byte
  unique id travels through the group, nothing else does).  The broadcast of the fp16 planes is pipelined by row
  blocks under the first tile's variance GEMM; (mean, sd) come back in one all-gather.
* **generic** (any tensors, any cache -- the inducing-point cache, and the gloo / CPU test of this plumbing):
  ``torch.distributed`` broadcasts of the cache tensors + all-gathers.
"""
import numpy as np
import torch
import torch.distributed as dist


TILE_ALIGN = 128        # rows: the engine groups test points by 128 (support ranges of K*, GEMM tiles)


def tile_bounds(M, world, rank):
    """Contiguous row tile [lo, hi) of rank `rank`.  Tile edges fall on multiples of 128 rows (the last tile takes the
    ragged rest, trailing tiles may be empty): the engine's 128-row groups of test points -- the unit of the compact
    support of K* -- are then the same groups as in an unsharded call, which is what makes sharded == unsharded hold
    bit for bit."""
    per = tile_width(M, world)
    lo = min(int(M), rank * per)
    return lo, min(int(M), lo + per)


def tile_width(M, world):
    """Common (padded) tile width: what every rank allocates so that one all-gather fits all tiles."""
    groups = -(-int(M) // TILE_ALIGN)
    return -(-groups // int(world)) * TILE_ALIGN


def cyclic_rows(M, world, rank, device=None):
    """Row indices of rank `rank` under the CYCLIC distribution of the engine's 128-row groups (group g goes to rank
    g % world).  The cost of a group depends on where its points lie (its K* support range against the triangular
    factor), so contiguous tiles are unevenly loaded; dealing the groups round-robin gives every rank the same mix."""
    M, world = int(M), int(world)
    groups = -(-M // TILE_ALIGN)
    if rank >= groups:
        return torch.zeros(0, dtype=torch.int64, device=device)
    g = torch.arange(rank, groups, world, device=device)
    idx = (g[:, None] * TILE_ALIGN + torch.arange(TILE_ALIGN, device=device)[None, :]).reshape(-1)
    return idx[idx < M]


def cyclic_merge(allp, M, world):
    """allp [world, C, tile_width(M, world)] (rank r's rows in cyclic_rows order, zero padded) -> [C, M] in row order."""
    C = allp.shape[1]
    gr = allp.shape[2] // TILE_ALIGN
    return allp.reshape(world, C, gr, TILE_ALIGN).permute(1, 2, 0, 3).reshape(C, gr * world * TILE_ALIGN)[:, :int(M)]


def _global_rank(group, group_rank):
    """torch.distributed.broadcast takes GLOBAL ranks; callers of this module speak group ranks."""
    if group is None or group is dist.group.WORLD:
        return group_rank
    return dist.get_global_rank(group, group_rank)


# ---------------------------------------------------------------------------------------------
# native transport
# ---------------------------------------------------------------------------------------------
_COMM_OF = {}          # id(engine) -> (group key, nranks, rank)


def ensure_comm(engine, group=None):
    """Give `engine` an NCCL communicator spanning `group` (default: the world).  Collective over the group; cached."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    key = (id(group) if group is not None else 0, world, rank)
    if _COMM_OF.get(id(engine)) == key and engine.comm_info() == (world, rank):
        return world, rank
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=_global_rank(group, 0), group=group)
    engine.comm_init(world, rank, box[0])
    _COMM_OF[id(engine)] = key
    return world, rank


def predict_exact_sharded(engine, kernel_id, theta, X, y, jitter, Xs, src=0, group=None, fac=None, tile_only=False,
                          factor="replicate", synced=False, layout="cyclic"):
    """Exact-GP sharded predict on the native transport.  Every rank passes same-shaped theta / X / y (rank ``src``'s
    values win) and the full (M, d) test rows, or -- with ``tile_only`` -- just its own tile.  Returns (mean, sd, info)
    for all M rows on every rank (device tensors; info = the pivot status of the factorisation).

    factor="replicate" (default): rank ``src``'s theta / X / y are broadcast (a few KB; skipped with ``synced=True``
    when the caller knows they are equal already) and EVERY rank factorises -- the same deterministic kernels on the same
    inputs give bit-identical caches, the other ranks would idle during rank ``src``'s factorisation anyway, and the
    2 N ld-byte plane broadcast (238 MB at N = 7 688) disappears from the step: the only collective left is the
    all-gather of (mean, sd).  factor="broadcast": rank ``src`` alone factorises and its cache travels in row blocks
    (gpg_predict_sharded); the choice when the other GPUs have something else to do meanwhile.

    layout="cyclic" (default): the 128-row groups of the test rows are dealt round-robin to the ranks (cyclic_rows), which
    balances the position-dependent cost of the groups; layout="tiles": contiguous tiles (tile_bounds)."""
    world, rank = ensure_comm(engine, group)
    N = X.shape[0]
    replicate = factor == "replicate"
    if fac is None:
        fac = engine.alloc_factor(N, X.dtype, with_L=(replicate or rank == src))
    theta, X = theta.contiguous(), X.contiguous()
    if replicate:
        if world > 1 and not synced:
            gsrc = _global_rank(group, src)
            for t in (theta, X, y):
                dist.broadcast(t, src=gsrc, group=group)
        engine.factorize(kernel_id, theta, X, y, jitter, out=fac)
    elif rank == src:
        engine.factorize(kernel_id, theta, X, y, jitter, out=fac)
    cyclic = False
    if tile_only:
        Xt = Xs
        counts = [None] * world
        dist.all_gather_object(counts, int(Xt.shape[0]), group=group)
        M = sum(counts)
        width = max(counts)
    else:
        M = Xs.shape[0]
        cyclic = layout == "cyclic" and world > 1
        width = tile_width(M, world)
        if cyclic:
            Xt = Xs.index_select(0, cyclic_rows(M, world, rank, Xs.device))
        else:
            lo, hi = tile_bounds(M, world, rank)
            Xt = Xs[lo:hi]
            counts = [tile_bounds(M, world, r)[1] - tile_bounds(M, world, r)[0] for r in range(world)]
    if replicate:
        local = torch.zeros(2, max(width, 1), dtype=X.dtype, device=X.device)
        if Xt.shape[0]:
            engine.predict(kernel_id, theta, X, fac, Xt, mean=local[0, :Xt.shape[0]], sd=local[1, :Xt.shape[0]])
        allp = engine.allgather_pred(local)
    else:
        _, allp = engine.predict_sharded(kernel_id, theta, X, fac, Xt, max(width, 1), root=src)
    if cyclic:
        both = cyclic_merge(allp, M, world)
        mean, sd = both[0].contiguous(), both[1].contiguous()
    elif all(c == width for c in counts):
        mean, sd = allp[:, 0, :].reshape(-1), allp[:, 1, :].reshape(-1)
    else:
        mean = torch.cat([allp[r, 0, :c] for r, c in enumerate(counts)])
        sd = torch.cat([allp[r, 1, :c] for r, c in enumerate(counts)])
    return mean, sd, fac["info"]


def acq_topk_sharded(engine, acq_id, mean_local, sd_local, idx_offset, k, group=None, **kw):
    """Global top-k of the acquisition sweep over tiles that live on the ranks (native transport)."""
    ensure_comm(engine, group)
    return engine.acq_sweep_sharded(acq_id, mean_local, sd_local, idx_offset, k, **kw)


# ---------------------------------------------------------------------------------------------
# generic transport (torch.distributed)
# ---------------------------------------------------------------------------------------------
def factor_keys(fac, engine=None):
    """Which tensors of a factor cache have to travel.  The exact-GP cache carries Linv twice (fp32 / fp64 matrix and
    its fp16 planes); gpg_predict reads one of them -- the engine's own routing rule (gpg_predict_uses_planes) says
    which.  Without an engine to ask, everything present is sent."""
    if "Ui" in fac:                  # inducing-point cache (gpg_sparse_factorize): two m x m factors and a vector
        return ("Ui", "Pm", "w", "split", "scales", "info")
    planes = fac.get("wsplit")
    if planes is not None and engine is not None and hasattr(engine, "predict_uses_planes") and \
            engine.predict_uses_planes(fac["alpha"].dtype, fac["alpha"].shape[0], True):
        return ("alpha", "wsplit", "scales", "info")
    return ("Linv", "alpha", "wsplit", "scales", "info")


def broadcast_factor(fac, src=0, group=None, engine=None):
    """Broadcast the tensors of a factor cache in place (every rank passes same-shaped buffers).  ``src`` is a rank
    of ``group``."""
    gsrc = _global_rank(group, src)
    for key in factor_keys(fac, engine):
        if fac.get(key) is not None:
            dist.broadcast(fac[key], src=gsrc, group=group)
    return fac


def gather_tiles(local, M, group=None):
    """All-gather the per-rank tiles of a length-M vector (tiles as produced by tile_bounds)."""
    world = dist.get_world_size(group)
    width = tile_width(M, world)
    if width * world == M:
        out = torch.empty(M, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    buf = torch.zeros(width, dtype=local.dtype, device=local.device)      # pad ragged / empty tiles to the common width
    buf[: local.numel()] = local
    out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[:M].clone()                              # tiles are back to back: tile r starts at r * width


def _raise_not_pd(info):
    raise torch.linalg.LinAlgError(
        f"linalg.cholesky: The factorization could not be completed because the input is not "
        f"positive-definite (the leading minor of order {info} is not positive-definite).")


def predict_sharded(factorize_fn, alloc_fn, predict_tile_fn, Xs, group=None, src=0, engine=None):
    """
    factorize_fn() -> fac      run on rank ``src`` only (K assembly + Cholesky + inverse + solves); it must NOT raise on
                               a failed factorisation (the other ranks are already waiting in the broadcast): the pivot
                               status travels in fac["info"] and every rank raises together after the collective
    alloc_fn() -> fac          empty same-shaped buffers on the other ranks
    predict_tile_fn(fac, Xs_tile) -> (mean_tile, sd_tile)
    Xs: (M, d) test rows, identical on every rank (each rank only reads its tile).
    Returns the full (mean, sd) on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    M = Xs.shape[0]
    fac = factorize_fn() if rank == src else alloc_fn()
    if world > 1:
        broadcast_factor(fac, src, group, engine)
    info = int(fac["info"].item()) if fac.get("info") is not None else 0
    if info != 0:
        _raise_not_pd(info)                             # on every rank: nobody is left behind in a collective
    lo, hi = tile_bounds(M, world, rank)
    mean_t, sd_t = predict_tile_fn(fac, Xs[lo:hi])
    if world == 1:
        return mean_t, sd_t
    return gather_tiles(mean_t, M, group), gather_tiles(sd_t, M, group)


def topk_merge_generic(vals_local, idx_local, k, group=None):
    """Merge per-rank candidate lists (values, global flat indices; already the local top-k) into the global top-k in
    the reference's order -- descending value, NaN first, ties by larger flat index (boptim.py:304-306).  Host-side:
    the generic transport's counterpart of gpg_acq_sweep_sharded."""
    world = dist.get_world_size(group)
    box = [None] * world
    dist.all_gather_object(box, (np.asarray(vals_local, dtype=np.float64), np.asarray(idx_local, dtype=np.int64)),
                           group=group)
    vals = np.concatenate([b[0] for b in box])
    idx = np.concatenate([b[1] for b in box])
    keep = idx >= 0
    vals, idx = vals[keep], idx[keep]
    nan = np.isnan(vals)
    order = np.lexsort((-idx, -np.where(nan, np.inf, vals), ~nan))     # last key first: NaN group, value, index
    order = order[:k]
    return vals[order], idx[order]


def predict_model_sharded(model, Xs, src=0, group=None):
    """Tile-sharded ``model.predict_sd`` for the model object of a reconstructor / skreconstructor (ExactGPModel,
    SparseGPModel, SKExactGPModel): rank ``src``'s hyper-parameters (and inducing inputs) win -- training is
    replicas-only, so the ranks may hold different values.  Exact GP on the native transport: the few KB of
    (theta, X, y) are broadcast and every rank factorises (predict_exact_sharded, factor="replicate"); the other models:
    rank ``src`` factorises and its cache is broadcast.  Every rank predicts its tile of ``Xs`` and the tiles are
    all-gathered.  Every rank passes a model built on the same (X, y);
    returns (mean, sd) for all of ``Xs`` on every rank.  A failed factorisation raises LinAlgError on EVERY rank."""
    eng = model.engine
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    Xs = Xs.to(eng.device, model.kernel.dtype).contiguous()
    sparse = hasattr(model, "Xu")
    kid = model.kernel.kernel_id
    shift = getattr(model, "mean_shift", None)          # constant mean of the GPyTorch-semantics model
    native = world > 1 and not sparse and Xs.is_cuda and hasattr(eng, "predict_sharded") and shift is None
    if native:
        model._factor = None
        y = model._y
        mean, sd, info = predict_exact_sharded(eng, kid, model._theta, model._X, y, model.jitter, Xs, src=src, group=group)
        info = int(info.item())
        if info != 0:
            _raise_not_pd(info)
        return mean, sd
    if world > 1:
        gsrc = _global_rank(group, src)
        dist.broadcast(model._theta, src=gsrc, group=group)
        if sparse:
            dist.broadcast(model._Xu, src=gsrc, group=group)
        model._factor = None

    def factorize():
        fac, _ = model.factor(check=False)              # never raise before the collective (see predict_sharded)
        return fac

    if sparse:
        alloc = lambda: eng.alloc_sparse_factor(model._Xu.shape[0], model.kernel.dtype)
        tile = lambda fac, X: eng.sparse_predict(kid, model._theta, model._Xu, fac, X)
    else:
        alloc = lambda: eng.alloc_factor(model._X.shape[0], model.kernel.dtype, with_L=False)
        tile = lambda fac, X: eng.predict(kid, model._theta, model._X, fac, X)
    mean, sd = predict_sharded(factorize, alloc, tile, Xs, group, src=src, engine=eng)
    return (mean + shift() if shift is not None else mean), sd
