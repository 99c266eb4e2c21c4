#!/usr/bin/env python
"""
bench.py -- the headline benchmark of the exact-GP-on-grids hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|h512|c3|c5]

A "step" is ONE pass of the hot path over one dense grid: everything the reference's
reconstructor.predict does per call (gpr.py:248): K(X,X) assembly, Cholesky of K + (noise+jitter) I,
the triangular solves, K(X*,X), the diagonal predictive variance and the mean, for all M points
of X_full.  Metric: predicted grid points per second (mean + sd), whole job.

* `value`  : inputs (X, y, theta, X_full rows) already resident in HBM, outputs left in HBM.
* `e2e`    : the same pass through the reference-facing API gpim.reconstructor (host numpy
             arrays in, numpy arrays out; host<->device copies inside the timed region).
* `roofline`: the dominant kernel (the Linv x K* product with the fused column-sum-of-squares
             epilogue), CUDA-event bracketed inside libgpgrid.so on the launching stream.
* `cpu_baseline` / `--impl reference`: oracle/gp_oracle.py (the CPU restatement of the
             reference's Pyro path -- pyro-ppl is not installable here) timed on the
             host cores on a bounded sample of the same workload.

N > 1 (torchrun, one rank per GPU): weak scaling -- the dense grid gets N times more rows (step
1/N), rank 0 factorises, one NCCL broadcast of {Linv, alpha}, every rank predicts its row tile,
one all-gather of (mean, sd).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import workloads as W  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the variance GEMM (16 384 test points), from the
# ncu --set full captures summarised in profiles/r1_05_ncu_pgemm_{c2,h512}.md
NCU_TRAFFIC_BYTES = {"c2": 2.099609e9 + 6.744320e6, "h512": 10.066302e9 + 7.414528e6}
METRIC = "predicted grid points/sec (mean+sd)"
UNIT = "points/s"


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def make_workload(name, dense=1):
    """-> dict(R, kernel, theta(list), d, label).  `dense` multiplies the number of X_full rows."""
    ft = W.FIXED_THETA
    if name in ("c2", "h512", "c5", "c1k"):
        n = {"c2": 256, "h512": 512, "c5": 1024, "c1k": 128}[name]
        R = W.spiral_scan(n)
        theta = [ft["variance"], ft["noise"], 1.0, ft["lengthscale"], ft["lengthscale"]]
        kern = "RBF"
        label = f"2D {n}x{n} sparse spiral scan, RBF, fixed theta"
    elif name == "c3":
        R = W.hyperspectral((64, 64, 16))
        theta = [ft["variance"], ft["noise"], 1.0, ft["lengthscale"], ft["lengthscale"], ft["lengthscale_z"]]
        kern = "Matern52"
        label = "3D 64x64x16 hyperspectral, Matern52, fixed theta"
    else:
        raise SystemExit(f"unknown workload {name}")
    sl = [slice(0, R.shape[0], 1.0 / dense)] + [slice(0, e, 1.0) for e in R.shape[1:]]
    Xfull = np.array(np.mgrid[tuple(sl)])                     # gprutils.get_full_grid layout (c, *dims)
    return {"name": name, "R": R, "kernel": kern, "theta": theta, "d": R.ndim, "label": label,
            "Xfull": Xfull, "jitter": ft["jitter"]}


def rows_of(Xgrid):
    return Xgrid.reshape(Xgrid.shape[0], -1).T


def train_rows(R):
    """(N, d) coordinates and (N,) values of the observed pixels (product-side layout helper)."""
    from gpim_b200 import gprutils
    X, y = gprutils.prepare_training_data(gprutils.get_sparse_grid(R), R)
    return X.numpy(), y.numpy()


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up (its start-up takes longer than a short timed
    region); stop(t0, t1) keeps the samples whose timestamp falls inside the timed window."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    @staticmethod
    def _ts(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:                                    # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                rows.append((self._ts(c[0]), float(c[2]), float(c[3]), float(c[4]),
                             [nm for nm, v in zip(names, c[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if t0 is not None and r[0] is not None and t0 - 0.02 <= r[0] <= t1 + 0.02]
        window = "timed region"
        if not inside:                                       # very short region: fall back to the loaded samples
            thr = 0.5 * max([r[3] for r in rows], default=0.0)
            inside = [r for r in rows if r[3] >= thr]
            window = "samples at >= half of peak power (timed region shorter than the polling interval)"
        if inside:
            reasons = sorted({nm for r in inside for nm in r[4]})
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=reasons, samples=len(inside), power_w_max=float(max(r[3] for r in inside)),
                       window=window)
        return out


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_predict_sample(wl, m_sample, dtype_name="f32"):
    """One reference-style predict on the host: full factorisation (K, Cholesky) + K*, TRSM and
    reductions on `m_sample` grid points; whole-grid time extrapolated linearly in M for the
    per-point stages.  Returns (points_per_s, seconds_measured, stage dict)."""
    import torch
    from oracle import gp_oracle as O
    X, y = train_rows(wl["R"])
    Xs = rows_of(wl["Xfull"])
    M = Xs.shape[0]
    sel = np.linspace(0, M - 1, m_sample).astype(np.int64)
    th = wl["theta"]
    dt = torch.float32 if dtype_name == "f32" else torch.float64
    t0 = time.perf_counter()
    _, _, st = O.predict_fixed_theta(wl["kernel"], X, y, Xs[sel], th[0], th[3:], th[1], jitter=wl["jitter"], dtype=dt,
                                     scale_mixture=th[2])
    wall = time.perf_counter() - t0
    t_fact = st["kmat"] + st["cholesky"]
    t_pts = st["kcross"] + st["trsm"] + st["reduce"]
    t_full = t_fact + t_pts * (M / m_sample)
    return M / t_full, wall, st


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload(args.workload, dense=args.gpus)
    M = rows_of(wl["Xfull"]).shape[0]
    m_sample = args.cpu_sample
    for _ in range(min(args.warmup, 1)):
        cpu_predict_sample(wl, m_sample, args.cpu_dtype)
    vals, walls = [], []
    for _ in range(args.steps):
        v, wall, st = cpu_predict_sample(wl, m_sample, args.cpu_dtype)
        vals.append(v); walls.append(wall)
    value = float(np.mean(vals))
    N = int((~np.isnan(wl["R"])).sum())
    sample = (f"full K assembly + Cholesky (N={N}) and K*/TRSM/reduce on {m_sample} of {M} grid points per step, "
              f"per-point stages scaled by M/{m_sample}; torch {torch.__version__} CPU {args.cpu_dtype}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * M / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.cpu_dtype, "data": "synthetic",
            "config": bench_config(wl, args.gpus, N, M),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample, "measured_s_per_step": float(np.mean(walls))},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure_extra(eng, name, steps, warmup, compact_support=False):
    """The same step on another workload of SURVEY 8d (1 GPU, device-resident inputs, CUDA events).
    compact_support: with GPG_OPT_COMPACT_SUPPORT (NOT the headline configuration): the variance GEMM of each
    128-row tile of test points only visits the training rows whose covariance with the tile exceeds 1e-14 x
    variance -- same outputs to fp32 rounding, far fewer MMAs when the lengthscale is short against the grid."""
    import torch
    from gpim_b200 import _lib
    from gpim_b200._lib import KERNEL_IDS
    wl = make_workload(name)
    if compact_support:
        eng.set_option(_lib.OPT_COMPACT_SUPPORT, 1)
    X, y = train_rows(wl["R"])
    Xs = rows_of(wl["Xfull"])
    N, M = X.shape[0], Xs.shape[0]
    dev, dt = eng.device, torch.float32
    kid = KERNEL_IDS[wl["kernel"]]
    th = torch.tensor(wl["theta"], dtype=dt, device=dev)
    Xd, yd, Xsd = (torch.tensor(a, dtype=dt, device=dev) for a in (X, y, Xs))
    fac = eng.alloc_factor(N, dt)
    mean_t, sd_t = torch.empty(M, dtype=dt, device=dev), torch.empty(M, dtype=dt, device=dev)

    def step():
        eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
        eng.predict(kid, th, Xd, fac, Xsd, mean=mean_t, sd=sd_t)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    for _ in range(steps):
        eng.predict(kid, th, Xd, fac, Xsd, mean=mean_t, sd=sd_t)
    e2.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ms_pred = e1.elapsed_time(e2) / steps
    assert bool(torch.isfinite(mean_t).all()) and bool(torch.isfinite(sd_t).all())
    out = {"workload": wl["label"], "N_train": N, "M_grid": M, "ms_per_step": ms, "value": M / (ms * 1e-3), "unit": UNIT,
           "factor_cached_ms": ms_pred, "factor_cached_value": M / (ms_pred * 1e-3), "steps": steps}
    if compact_support:
        eng.set_option(_lib.OPT_COMPACT_SUPPORT, 0)
        out["note"] = ("GPG_OPT_COMPACT_SUPPORT = 1 (opt-in, not the headline): variance GEMM restricted per tile to the "
                       "training rows with covariance > 1e-14 x variance; outputs equal the dense ones to fp32 rounding "
                       "(tests/test_gpu_parity.py::test_predict_compact_support_option_is_exact)")
    else:
        out["variance_gemm_tflops_algorithmic"] = float(N) * N * M / (ms_pred * 1e-3) / 1e12
    del fac, Xsd, mean_t, sd_t
    torch.cuda.empty_cache()
    return out


def measure_training(eng, name, iters):
    """Hot loop 1 of the reference (gpr.py:190-197): Adam iterations on the marginal likelihood, all on the
    device (K assembly, Cholesky, inverse, solves, K^-1, gradient reduction, Adam step), fp32."""
    import torch
    from gpim_b200._lib import KERNEL_IDS
    wl = make_workload(name)
    X, y = train_rows(wl["R"])
    d = X.shape[1]
    dev, dt = eng.device, torch.float32
    Xd, yd = torch.tensor(X, dtype=dt, device=dev), torch.tensor(y, dtype=dt, device=dev)
    bounds = [1e-4, 10.0] + [1.0] * d + [4.0] * d            # GP_sparse2Dimages.ipynb cell 11 bounds

    def run(n):
        u = torch.zeros(3 + d, dtype=dt, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        traj, theta, info = eng.fit_adam(KERNEL_IDS[wl["kernel"]], Xd, yd, wl["jitter"], u, bounds, d, n, 0.1)
        e1.record()
        torch.cuda.synchronize()
        assert int(info.item()) == 0 and bool(torch.isfinite(traj).all())
        return e0.elapsed_time(e1)

    run(2)
    ms = run(iters)
    return {"workload": wl["label"].replace("fixed theta", "Adam on (variance, lengthscales, noise)"), "N_train": int(X.shape[0]),
            "iterations": iters, "ms_per_adam_iteration": ms / iters,
            "reference_published": "7.5-9.6 ms per iteration at N ~ 5..400 on a Colab GPU "
                                   "(examples/contributed/GPIM_BEPS.ipynb:665-672); 3.6-7.0 ms on CPU at N = 5..55"}


def measure_sparse(eng, name, iters, steps):
    """reconstructor(sparse=True) at the notebook setting: len(X) // 10 inducing points picked as
    gpr.py:151 does, fp64 (the reference default) and fp32.  Reports the Adam iteration of the VFE objective
    (hyper-parameters + inducing inputs, gpg_sparse_fit_adam) and the prediction over the dense grid
    (gpg_sparse_factorize + gpg_sparse_predict, refactorising every step like the reference's forward())."""
    import torch
    from gpim_b200._lib import KERNEL_IDS
    wl = make_workload(name)
    X, y = train_rows(wl["R"])
    N, d = X.shape
    m_ind = N // 10
    Xs_host = rows_of(wl["Xfull"])
    bounds = [1e-4, 10.0] + [1.0] * d + [4.0] * d
    kid = KERNEL_IDS[wl["kernel"]]
    out = {"workload": wl["label"].replace("fixed theta", "VFE inducing points"), "N_train": int(N),
           "inducing_points": int(len(range(0, N, N // m_ind))), "M_grid": int(Xs_host.shape[0])}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        dev = eng.device
        Xd, yd = torch.tensor(X, dtype=dt, device=dev), torch.tensor(y, dtype=dt, device=dev)
        Xs = torch.tensor(Xs_host, dtype=dt, device=dev)
        theta = torch.tensor(wl["theta"], dtype=dt, device=dev)
        jitter = 1e-5 if tag == "f64" else 1e-4

        def fit(n):
            Xu = Xd[::N // m_ind].clone()
            u = torch.zeros(3 + d, dtype=dt, device=dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            traj, _, _, info = eng.sparse_fit_adam(kid, Xd, yd, Xu, jitter, u, bounds, d, n, 0.05, record_xu=False)
            e1.record()
            torch.cuda.synchronize()
            assert int(info.item()) == 0 and bool(torch.isfinite(traj).all())
            return e0.elapsed_time(e1) / n

        def predict(n):
            Xu = Xd[::N // m_ind].clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fac = eng.sparse_factorize(kid, theta, Xd, yd, Xu, jitter)
                mean, sd = eng.sparse_predict(kid, theta, Xu, fac, Xs)
            e1.record()
            torch.cuda.synchronize()
            assert int(fac["info"].item()) == 0 and bool(torch.isfinite(mean).all()) and bool(torch.isfinite(sd).all())
            return e0.elapsed_time(e1) / n

        fit(2)
        predict(1)
        ms_fit, ms_pred = fit(iters), predict(steps)
        out[tag] = {"ms_per_adam_iteration": ms_fit, "predict_ms_per_step": ms_pred,
                    "points_per_s": Xs.shape[0] / (ms_pred * 1e-3)}
    try:                                              # context only: must never take the bench line down
        out["cpu_oracle"] = cpu_sparse_sample(X, y, m_ind, wl["kernel"], wl["theta"])
    except Exception as e:                            # noqa: BLE001
        out["cpu_oracle"] = {"error": repr(e)[:200]}
    return out


def cpu_sparse_sample(X, y, m_ind, kernel, theta, iters=2):
    """The reference's CPU arithmetic for the same Adam iteration (oracle/sparse_oracle.py: the VFE objective through
    torch autograd, fp64, all host threads), timed on a bounded sample of `iters` iterations after one warm-up."""
    import torch
    from oracle.sparse_oracle import vfe_loss
    torch.set_num_threads(os.cpu_count() or 1)
    N, d = X.shape
    Xt, yt = torch.tensor(X, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    leaf = lambda v: torch.tensor(v, dtype=torch.float64, requires_grad=True)
    v, s2, al, ls = leaf(theta[0]), leaf(theta[1]), leaf(theta[2]), leaf(theta[3:3 + d])
    Xu = Xt[::N // m_ind].clone().requires_grad_(True)
    opt = torch.optim.Adam([v, s2, ls, Xu], lr=1e-3)
    times = []
    for it in range(iters + 1):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = vfe_loss(kernel, Xt, yt, Xu, v, ls, s2, al, 1e-5)
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    return {"ms_per_adam_iteration": 1e3 * sum(times[1:]) / iters, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{iters} iterations after 1 warm-up, fp64, N={N}, m={int(Xu.shape[0])}"}


def bench_config(wl, gpus, N, M):
    return {"workload": wl["label"], "name": wl["name"], "N_train": N, "M_grid": M, "kernel": wl["kernel"],
            "theta": {"variance": wl["theta"][0], "noise": wl["theta"][1], "lengthscale": wl["theta"][3:],
                      "jitter": wl["jitter"]},
            "sharding": "1 GPU" if gpus == 1 else f"X_full row tiles over {gpus} GPUs, 1 broadcast {{Linv,alpha}} + 1 all-gather",
            "l2_policy": "inputs larger than L2: every step rewrites and rereads K/L/Linv (N x N fp32 each) and the "
                         "K* tiles; no explicit flush"}


# ---------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from gpim_b200 import _lib, sharded
    from gpim_b200._lib import KERNEL_IDS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    # started early: nvidia-smi takes ~1 s to come up
    clocks = ClockSampler(local) if (rank == 0 and not args.no_clocks) else None
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the one
        # JSON line by pointing fd 1 at stderr while the process group comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    eng = _lib.get_engine(local)
    dev = eng.device

    wl = make_workload(args.workload, dense=world)
    X, y = train_rows(wl["R"])
    Xs = rows_of(wl["Xfull"])
    N, M, d = X.shape[0], Xs.shape[0], X.shape[1]
    kid = KERNEL_IDS[wl["kernel"]]
    dt = torch.float32
    th = torch.tensor(wl["theta"], dtype=dt, device=dev)
    Xd = torch.tensor(X, dtype=dt, device=dev)
    yd = torch.tensor(y, dtype=dt, device=dev)
    Xsd = torch.tensor(Xs, dtype=dt, device=dev)
    fac = eng.alloc_factor(N, dt)
    lo, hi = sharded.tile_bounds(M, world, rank)
    mean_t = torch.empty(hi - lo, dtype=dt, device=dev)
    sd_t = torch.empty(hi - lo, dtype=dt, device=dev)

    def step():
        if rank == 0:
            eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
        if world > 1:
            sharded.broadcast_factor(fac, 0)
        eng.predict(kid, th, Xd, fac, Xsd[lo:hi], mean=mean_t, sd=sd_t)
        if world > 1:
            return sharded.gather_tiles(mean_t, M), sharded.gather_tiles(sd_t, M)
        return mean_t, sd_t

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    eng.set_option(_lib.OPT_STAGE_TIMING, 1)
    eng.stage_times()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    stages = eng.stage_times()
    eng.set_option(_lib.OPT_STAGE_TIMING, 0)
    clk = clocks.stop(wall0, wall1) if clocks else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    assert bool(torch.isfinite(out[0]).all()) and bool(torch.isfinite(out[1]).all()), "non-finite prediction"
    ms_per_step = ms / args.steps
    value = M / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing API (host arrays in / out), all ranks idle but 0 at N>1
    e2e = None
    if world == 1:
        e2e = run_e2e(args, wl, eng)
    else:
        e2e = run_e2e_sharded(args, wl, eng, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:                                        # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (dense 16-bit tensor rate, kernel timed inside a long step)" \
        if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    pg_ms, pg_n = stages["pgemm"]
    m_local = hi - lo
    flops_per_step = float(N) * float(N) * float(m_local)    # N^2 FLOP per predicted point
    achieved = flops_per_step * args.steps / (pg_ms * 1e-3) / 1e12 if pg_ms > 0 else 0.0
    km_ms, km_n = stages["kmat"]
    hbm = peaks.get("hbm_gbs") or 6650.0
    kmat_gbps = (4.0 * N * N / 2 + 8.0 * d * N) * km_n / (km_ms * 1e-3) / 1e9 if km_ms > 0 else 0.0
    ch_ms, ch_n = stages["cholesky"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(wl, world, N, M),
        "roofline": {"kernel": "predict GEMM Linv x K* + colsumsq epilogue (stage pgemm)", "bound": "tensor",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": NCU_TRAFFIC_BYTES.get(wl["name"]) if world == 1 else None,
                     "traffic_note": "bytes per launch (ncu capture, profiles/r1_05_ncu_pgemm_*.md); algorithmic bytes per "
                                     "launch = 4 B x (16384 x N) K* planes + 2 N^2 B lower-triangle W planes",
                     "algorithmic_bytes_per_launch": 4.0 * 16384 * N + 2.0 * N * N,
                     "launches": pg_n, "avg_launch_ms": pg_ms / max(pg_n, 1),
                     "algorithmic_flops_per_launch": flops_per_step * args.steps / max(pg_n, 1),
                     "peak_source": peak_src,
                     "note": "fp32-faithful split-fp16 product: 3 tcgen05 MMAs per algorithmic MAC, so the tensor pipe "
                             "executes 3 x achieved",
                     "tensor_pipe_tflops": 3.0 * achieved, "tensor_pipe_frac": 3.0 * achieved / peak_tf},
        "stages_ms_per_step": {k: v[0] / args.steps for k, v in stages.items() if v[1]},
        "kmat_assembly": {"bound": "hbm", "achieved": kmat_gbps, "peak": hbm, "unit": "GB/s", "frac": kmat_gbps / hbm,
                          "note": "lower-triangle tiles only: 2 N^2 + 8 d N bytes per launch"},
        "cholesky_tflops": (N ** 3 / 3.0) * ch_n / (ch_ms * 1e-3) / 1e12 if ch_ms > 0 else None,
        "e2e": e2e, "gpu_launches": launches, "clocks": clk,
    }
    if world == 1 and not args.no_extra and args.workload == "c2":
        # the 512 x 512 reconstruction BASELINE.json's target is quoted on, same step definition
        line["extra_workloads"] = {"h512": measure_extra(eng, "h512", max(2, args.steps // 2), 2),
                                   "train_c2": measure_training(eng, "c2", 10),
                                   "c2_compact_support": measure_extra(eng, "c2", args.steps, 2, compact_support=True),
                                   "h512_compact_support": measure_extra(eng, "h512", max(2, args.steps // 2), 2,
                                                                         compact_support=True),
                                   "sparse_c2": measure_sparse(eng, "c2", 10, max(2, args.steps // 2))}
    if world == 1 and not args.no_cpu_baseline:
        import torch as _t
        _t.set_num_threads(os.cpu_count() or 1)
        v, wall, st = cpu_predict_sample(wl, args.cpu_sample, args.cpu_dtype)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
            "sample": f"oracle predict (K + Cholesky at N={N}, K*/TRSM/reduce on {args.cpu_sample} of {M} points, "
                      f"per-point stages scaled to M), {args.cpu_dtype}, {wall:.1f} s measured",
            "stages_s": st}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, wl, eng):
    """gpim.reconstructor: swap the training set in (host -> device), predict over X_full given as
    a host numpy grid, results back as numpy arrays.  Factor cache invalidated every step, as the
    reference refactorises on every predict (gpr.py:248)."""
    import torch
    import gpim_b200 as gpim
    R = wl["R"]
    Xsp = gpim.utils.get_sparse_grid(R)
    rec = gpim.reconstructor(Xsp, R, wl["Xfull"], kernel=wl["kernel"], iterations=0, verbose=0, precision="single",
                             jitter=wl["jitter"], lengthscale=[[1.0] * R.ndim, [20.0] * R.ndim])
    th = wl["theta"]
    rec.model.set_theta(th[0], th[3:], th[1], th[2])
    Xh = rec.X.cpu().pin_memory()
    yh = rec.y.cpu().pin_memory()
    Xfull = wl["Xfull"]

    def step():
        rec.model.X = Xh
        rec.model.y = yh
        return rec.predict(Xfull, verbose=0)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mean, sd = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    M = mean.size
    h2d = Xh.numel() * 4 + yh.numel() * 4 + M * R.ndim * 4
    return {"value": M / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * M * 4),
            "ms_per_step": dt * 1e3, "api": "gpim.reconstructor.predict(X_full) after model.X/.y swap (host numpy in/out)"}


def run_e2e_sharded(args, wl, eng, world, rank):
    """N > 1: host rows in, every rank copies its tile to the device, sharded predict, rank 0 reads
    the gathered result back to the host."""
    import torch
    import torch.distributed as dist
    from gpim_b200 import sharded
    from gpim_b200._lib import KERNEL_IDS
    dev = eng.device
    dt = torch.float32
    X, y = train_rows(wl["R"])
    Xs = rows_of(wl["Xfull"])
    N, M = X.shape[0], Xs.shape[0]
    kid = KERNEL_IDS[wl["kernel"]]
    Xh = torch.tensor(X, dtype=dt).pin_memory()
    yh = torch.tensor(y, dtype=dt).pin_memory()
    lo, hi = sharded.tile_bounds(M, world, rank)
    Xsh = torch.tensor(Xs[lo:hi], dtype=dt).pin_memory()
    th = torch.tensor(wl["theta"], dtype=dt, device=dev)
    fac = eng.alloc_factor(N, dt)

    def step():
        Xd = Xh.to(dev, non_blocking=True)
        yd = yh.to(dev, non_blocking=True)
        Xsd = Xsh.to(dev, non_blocking=True)
        if rank == 0:
            eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
        sharded.broadcast_factor(fac, 0)
        m, s = eng.predict(kid, th, Xd, fac, Xsd)
        m, s = sharded.gather_tiles(m, M), sharded.gather_tiles(s, M)
        if rank == 0:
            return m.cpu().numpy(), s.cpu().numpy()
        return None

    for _ in range(2):
        step()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dist.barrier(); torch.cuda.synchronize()
    sec = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    sec = float(sec.item())
    h2d = (Xh.numel() + yh.numel()) * 4 * world + M * Xs.shape[1] * 4
    return {"value": M / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * M * 4),
            "ms_per_step": sec * 1e3, "api": "sharded.predict: pinned host rows -> per-rank tiles -> gathered numpy on rank 0"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu-sample", type=int, default=1024, dest="cpu_sample")
    ap.add_argument("--cpu-dtype", default="f32", choices=["f32", "f64"], dest="cpu_dtype")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra 512 x 512 measurement")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "cuda" and world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
