#!/usr/bin/env python
"""
bench.py -- the headline benchmark of the exact-GP-on-grids hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|h512|c3|c5]

A "step" is ONE pass of the hot path over one dense grid: everything the reference's
reconstructor.predict does per call (gpr.py:248): K(X,X) assembly, Cholesky of K + (noise+jitter) I,
the triangular solves, K(X*,X), the diagonal predictive variance and the mean, for all M points
of X_full.  Metric: predicted grid points per second (mean + sd), whole job.

Headline workload (N = 1): BASELINE.json configs[1] = C2, the 256 x 256 spiral scan (RBF, fp32).  The other
single-GPU configurations -- the 512 x 512 reconstruction the north-star target is quoted on, C3 (64 x 64 x 16
Matern52) and C4 (Bayesian optimisation, EI, 128 x 128, 50 steps) -- ride in the same JSON line under `workloads`,
each with its own `value`, `e2e`, `roofline` and `cpu_baseline`.

* `value`  : inputs (X, y, theta, X_full rows) already resident in HBM, outputs left in HBM; CUDA events.
* `e2e`    : the same pass through the reference-facing API gpim.reconstructor (host numpy arrays in, numpy arrays
             out; host<->device copies inside the timed region).
* `roofline`: the dominant kernel (the Linv x K* product with the fused column-sum-of-squares epilogue), CUDA-event
             bracketed inside libgpgrid.so on the launching stream; `cholesky` and `kmat_assembly` carry the
             rooflines BASELINE.json's metric names next to it.
* `cpu_baseline`: oracle/gp_oracle.py (the CPU restatement of the reference's Pyro path -- pyro-ppl is not
             installable here) on the host cores on a bounded sample of the same workload: the FULL
             factorisation at N plus the per-point stages on a sample of grid rows.  Its outputs are kept and
             compared with the CUDA path's at the same rows (`parity_vs_oracle`).
* `--impl reference`: the same oracle on the WHOLE workload, unextrapolated, capped at 2 timed steps.
* `torch_cuda_baseline`: the oracle's torch ops on device="cuda" -- what the reference's use_gpu=True dispatches to
             on the same box (cuSOLVER potrf, cuBLAS trsm; gpr.py:136-140,248).  A baseline leg, not the product.

N > 1 (torchrun, one rank per GPU): weak scaling on the headline workload -- the dense grid gets N times more rows.
Every rank factorises (the same deterministic kernels on the same inputs: bit-identical caches; the ranks would idle
during a single rank's factorisation anyway, and the 2 N ld-byte broadcast of the cache leaves the step), predicts
its row tile, ONE all-gather of (mean, sd) through gpg_allgather_pred.  `broadcast_mode` carries the same step with
rank 0 factorising alone and the cache travelling by the pipelined NCCL broadcast of gpg_predict_sharded;
`shard_parity_max_rel` compares the gathered result with an unsharded predict on rank 0;
`strong_c5` is BASELINE.json configs[4] as configured (1024 x 1024, M fixed at 1 048 576, tiles of M / N rows).
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
from workloads import make_workload, rows_of  # noqa: E402,F401

METRIC = "predicted grid points/sec (mean+sd)"
UNIT = "points/s"
PREDICT_CHUNK = 16384            # test points per variance-GEMM launch (gpg_predict's internal tile)


def train_rows(R):
    """(N, d) coordinates and (N,) values of the observed pixels (product-side layout helper)."""
    from gpim_b200 import gprutils
    X, y = gprutils.prepare_training_data(gprutils.get_sparse_grid(R), R)
    return X.numpy(), y.numpy()


def relinf(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:                                        # noqa: BLE001
        return {}


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the variance GEMM from the newest committed ncu
    --set full capture of this workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t.get(name)
        return (float(e["bytes_per_launch"]), e.get("source")) if e else (None, None)
    except Exception:                                        # noqa: BLE001
        return None, None


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
# CPU legs: the oracle on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_predict(wl, rows, dtype_name):
    """One reference-style predict on the host (oracle): full K assembly + Cholesky at N, then K*, TRSM and the
    reductions on the grid rows `rows` (None = all of them).  -> (mean, sd, stage seconds, wall seconds)."""
    import torch
    from oracle import gp_oracle as O
    X, y = train_rows(wl["R"])
    Xs = rows_of(wl["Xfull"])
    if rows is not None:
        Xs = Xs[rows]
    th = wl["theta"]
    dt = torch.float32 if dtype_name == "f32" else torch.float64
    t0 = time.perf_counter()
    mean, sd, st = O.predict_fixed_theta(wl["kernel"], X, y, Xs, th[0], th[3:], th[1], jitter=wl["jitter"], dtype=dt,
                                         scale_mixture=th[2])
    return mean, sd, st, time.perf_counter() - t0


def cpu_baseline_block(wl, m_sample, dtype_name, cuda_out=None):
    """Bounded CPU sample of one workload: full factorisation + `m_sample` grid rows; the whole-grid time scales the
    per-point stages by M / m_sample.  cuda_out = (mean, sd) numpy arrays of the CUDA path over the whole grid:
    compared with the oracle at the sampled rows (fp64 oracle = the parity reference of north_star)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    M = rows_of(wl["Xfull"]).shape[0]
    N = int((~np.isnan(wl["R"])).sum())
    rows = W.sample_rows(M, min(m_sample, M))
    mean, sd, st, wall = cpu_predict(wl, rows, dtype_name)
    t_fact = st["kmat"] + st["cholesky"]
    t_pts = st["kcross"] + st["trsm"] + st["reduce"]
    t_full = t_fact + t_pts * (M / len(rows))
    out = {"value": M / t_full, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "dtype": dtype_name,
           "sample": f"oracle predict: K + Cholesky at the full N = {N}, K* / TRSM / reduce on {len(rows)} of {M} grid rows "
                     f"(per-point stages scaled by M / {len(rows)}); {wall:.1f} s measured, torch {torch.__version__} CPU",
           "stages_s": {k: round(v, 4) for k, v in st.items()}, "measured_s": wall}
    if cuda_out is not None:
        out["parity_vs_oracle"] = {"rows": int(len(rows)), "oracle_dtype": dtype_name,
                                   "mean_relinf": relinf(cuda_out[0][rows], mean), "sd_relinf": relinf(cuda_out[1][rows], sd),
                                   "tolerance": {"mean": 1e-4, "sd": 1e-3},
                                   "note": "max |cuda - oracle| / max |oracle| over the sampled rows (north_star)"}
    return out


def run_reference(args):
    """--impl reference: the oracle (kind "port": pyro-ppl is not installable) on the WHOLE
    workload, all host threads, nothing extrapolated.  One full C2 predict is minutes of CPU time (fp32 TRSM with
    65 536 right-hand sides runs at ~30 GFLOP/s in torch's LAPACK path), so the timed steps are capped at 2 and the
    warm-up is a small matmul that only spins the thread pool up; `steps` reports what was timed."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload(args.workload, dense=args.gpus)
    M = rows_of(wl["Xfull"]).shape[0]
    N = int((~np.isnan(wl["R"])).sum())
    steps = max(1, min(args.steps, args.cpu_full_steps))
    a = torch.randn(2048, 2048)
    (a @ a).sum().item()
    # the whole grid up to args.cpu_max_rows rows (C2: all 65 536, ~46 s per step); beyond that (--gpus N makes the grid N
    # times denser: minutes per step) the FIRST cpu_max_rows rows -- a contiguous piece of the same workload run in full
    # (factorisation included), its throughput reported as measured, nothing scaled
    rows = None if M <= args.cpu_max_rows else np.arange(args.cpu_max_rows)
    m_run = M if rows is None else len(rows)
    walls, stages = [], None
    for _ in range(steps):
        _, _, st, wall = cpu_predict(wl, rows, args.cpu_dtype)
        walls.append(wall); stages = st
    sec = float(np.mean(walls))
    value = m_run / sec
    what = (f"whole workload, nothing extrapolated: K + Cholesky at N = {N} and K* / TRSM / reduce on all {M} grid rows"
            if rows is None else
            f"bounded sample, nothing extrapolated: K + Cholesky at N = {N} and K* / TRSM / reduce on the first {m_run} of "
            f"{M} grid rows; value = {m_run} rows / measured seconds")
    sample = (f"{what}, {steps} timed step(s) (requested {args.steps}; capped: one step is {sec:.0f} s), thread-pool "
              f"warm-up only; torch {torch.__version__} CPU {args.cpu_dtype}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "steps_requested": args.steps, "warmup": 0, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 * sec, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.cpu_dtype, "data": "synthetic",
            "config": bench_config(wl, args.gpus, N, M),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample, "stages_s": {k: round(v, 3) for k, v in stages.items()}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def torch_cuda_baseline(wl, steps=3):
    """The reference's use_gpu=True dispatch on the same box: the oracle's torch ops (kernel by matmul expansion,
    torch.linalg.cholesky -> cuSOLVER potrf, solve_triangular -> cuBLAS trsm on the N x (M + 1) pack) on
    device="cuda", device-resident inputs, CUDA-synchronised wall clock.  A baseline leg: the product never runs
    through it."""
    import torch
    from oracle import gp_oracle as O
    X, y = train_rows(wl["R"])
    Xs = rows_of(wl["Xfull"])
    th = wl["theta"]
    N, M = X.shape[0], Xs.shape[0]
    dev = torch.device("cuda")
    Xd, yd, Xsd = (torch.tensor(a, dtype=torch.float32, device=dev) for a in (X, y, Xs))
    try:
        def run():
            return O.predict_fixed_theta(wl["kernel"], Xd, yd, Xsd, th[0], th[3:], th[1], jitter=wl["jitter"],
                                         dtype=torch.float32, scale_mixture=th[2], device=dev, to_numpy=False)
        run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            mean, sd, st = run()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / steps
        out = {"value": M / sec, "unit": UNIT, "ms_per_step": 1e3 * sec, "steps": steps, "dtype": "f32",
               "what": "oracle.predict_fixed_theta on device='cuda' (torch " + torch.__version__ + ": cuSOLVER potrf + cuBLAS "
                       "trsm with the whole N x M cross-kernel materialised) -- the reference's use_gpu=True path "
                       "(gpr.py:136-140,248)",
               "stages_s": {k: round(v, 5) for k, v in st.items()},
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        out["_mean"], out["_sd"] = mean.cpu().numpy(), sd.cpu().numpy()
        return out
    except Exception as e:                                   # noqa: BLE001  (e.g. out of memory at C5 sizes)
        return {"error": repr(e)[:300]}
    finally:
        del Xd, yd, Xsd
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------
# the CUDA arm: one workload, device-resident
# ---------------------------------------------------------------------------------------------
def bench_config(wl, gpus, N, M):
    return {"workload": wl["label"], "name": wl["name"], "N_train": N, "M_grid": M, "kernel": wl["kernel"],
            "theta": {"variance": wl["theta"][0], "noise": wl["theta"][1], "lengthscale": wl["theta"][3:],
                      "jitter": wl["jitter"]},
            "sharding": "1 GPU" if gpus == 1 else f"128-row groups of X_full dealt round-robin over {gpus} GPUs, every rank factorises (replicated, "
                                                  f"bit-identical caches), 1 all-gather of (mean, sd)",
            "l2_policy": "inputs larger than L2: every step rewrites and rereads K/L/Linv (N x N fp32 each) and the "
                         "K* tiles; no explicit flush"}


def measure_predict(eng, wl, steps, warmup, world=1, rank=0, keep_outputs=True, dtype_name="f32", factor_mode="replicate"):
    """`steps` timed passes (factorise + predict) with device-resident inputs, CUDA events, stage clocks.
    world > 1, factor_mode "replicate": every rank factorises (identical kernels on identical inputs: bit-identical caches,
    no factor broadcast), predicts its tile, ONE all-gather of (mean, sd) through gpg_allgather_pred;
    factor_mode "broadcast": gpg_predict_sharded (rank 0 factorises, pipelined NCCL broadcast of the cache).
    Returns a dict; on rank 0 it carries the gathered outputs."""
    import torch
    import torch.distributed as dist
    from gpim_b200 import _lib, sharded
    from gpim_b200._lib import KERNEL_IDS
    X, y = train_rows(wl["R"])
    Xs = rows_of(wl["Xfull"])
    N, M, d = X.shape[0], Xs.shape[0], X.shape[1]
    kid = KERNEL_IDS[wl["kernel"]]
    dev, dt = eng.device, (torch.float32 if dtype_name == "f32" else torch.float64)
    th = torch.tensor(wl["theta"], dtype=dt, device=dev)
    Xd = torch.tensor(X, dtype=dt, device=dev)
    yd = torch.tensor(y, dtype=dt, device=dev)
    # the engine's 128-row groups of X_full are dealt round-robin to the ranks (sharded.cyclic_rows): equal mix of
    # cheap and expensive groups on every GPU
    rows = sharded.cyclic_rows(M, world, rank).numpy() if world > 1 else np.arange(M)
    lo, hi = 0, len(rows)
    Xsd = torch.tensor(Xs[rows], dtype=dt, device=dev)
    replicate = factor_mode == "replicate"
    fac = eng.alloc_factor(N, dt, with_L=(rank == 0 or replicate))
    width = sharded.tile_width(M, world)
    pred_local = torch.zeros(2, width, dtype=dt, device=dev)
    pred_all = torch.empty(world, 2, width, dtype=dt, device=dev) if world > 1 else None

    m_loc = hi - lo

    def step():
        if rank == 0 or replicate:
            eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
        if world > 1 and not replicate:
            eng.predict_sharded(kid, th, Xd, fac, Xsd, width, root=0, pred_local=pred_local, pred_all=pred_all)
        elif world > 1:
            if m_loc:
                eng.predict(kid, th, Xd, fac, Xsd, mean=pred_local[0, :m_loc], sd=pred_local[1, :m_loc])
            eng.allgather_pred(pred_local, out=pred_all)
        else:
            eng.predict(kid, th, Xd, fac, Xsd, mean=pred_local[0], sd=pred_local[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    eng.set_option(_lib.OPT_STAGE_TIMING, 1)
    eng.stage_times()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    stages = eng.stage_times()
    macs = eng.variance_gemm_macs()                          # what the variance GEMM executed on THIS rank, all steps
    eng.set_option(_lib.OPT_STAGE_TIMING, 0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    info = int(fac["info"].item())
    assert info == 0, f"factorisation failed at pivot {info}"
    if world > 1:
        both = sharded.cyclic_merge(pred_all, M, world)
        mean, sd = both[0].contiguous(), both[1].contiguous()
    else:
        mean, sd = pred_local[0, :M], pred_local[1, :M]
    assert bool(torch.isfinite(mean).all()) and bool(torch.isfinite(sd).all()), "non-finite prediction"
    # factor-cached passes (L reused): the per-point part alone, 1 GPU only
    ms_cached = None
    if world == 1:
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(max(1, steps // 2)):
            eng.predict(kid, th, Xd, fac, Xsd, mean=pred_local[0], sd=pred_local[1])
        e3.record()
        torch.cuda.synchronize()
        ms_cached = e2.elapsed_time(e3) / max(1, steps // 2)
    out = {"N": N, "M": M, "d": d, "m_local": hi - lo, "ms_per_step": ms / steps, "value": M / (ms / steps * 1e-3),
           "launches": launches, "stages": stages, "steps": steps, "wall": (wall0, wall1), "ms_cached": ms_cached,
           "pgemm_macs": macs}
    if world > 1 and rank == 0:
        # sharded == unsharded: the same kernels on the same inputs, tile by tile -> must be bit-identical
        m1, s1 = eng.predict(kid, th, Xd, fac, torch.tensor(Xs, dtype=dt, device=dev))
        out["shard_parity_max_rel"] = max(relinf(mean.cpu().numpy(), m1.cpu().numpy()), relinf(sd.cpu().numpy(), s1.cpu().numpy()))
        del m1, s1
    if keep_outputs and rank == 0:
        out["mean"], out["sd"] = mean.cpu().numpy(), sd.cpu().numpy()
    del fac, Xsd, pred_local, pred_all
    torch.cuda.empty_cache()
    return out


def roofline_blocks(res, wl, steps, peaks, timed_region_s):
    """roofline of the dominant kernel + the two rooflines BASELINE.json's metric names (Cholesky, assembly)."""
    N, d, m_local = res["N"], res["d"], res["m_local"]
    stages = res["stages"]
    # burst peak for a sub-2-second timed region at full clocks, the sustained (power-capped) one for a long step
    burst = timed_region_s < 2.0
    key = "bf16_tflops" if burst else "bf16_tflops_sustained"
    peak_tf = peaks.get(key) or (1590.0 if burst else 1400.0)
    peak_src = (f"MEASURED_PEAKS.json {key} (measured cuBLAS bf16 rate; "
                + ("burst figure: the timed region is %.2f s at full clocks" % timed_region_s if burst else
                   "sustained figure: the timed region is %.1f s under the power cap" % timed_region_s) + ")") if peaks else \
        "fallback of B200_PROFILING.md (no MEASURED_PEAKS.json)"
    hbm = peaks.get("hbm_gbs") or 6650.0
    pg_ms, pg_n = stages["pgemm"]
    dense_flops_per_step = float(N) * float(N) * float(m_local)    # N^2 FLOP per predicted point (dense product)
    # ALGORITHMIC work of the launch = what the product has to do once K* entries below fp32 resolution are left out
    # (GPG_OPT_COMPACT_SUPPORT, the default): counted by the kernel itself at tile granularity, 2 FLOP per MAC
    flops_total = 2.0 * res["pgemm_macs"] if res.get("pgemm_macs") else dense_flops_per_step * steps
    achieved = flops_total / (pg_ms * 1e-3) / 1e12 if pg_ms > 0 else 0.0
    dense_equiv = dense_flops_per_step * steps / (pg_ms * 1e-3) / 1e12 if pg_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic(wl["name"])
    pts_per_launch = min(PREDICT_CHUNK, m_local)
    roof = {"kernel": "gemm_tc_kernel as the predict GEMM: Linv x K* + column-sum-of-squares epilogue (stage pgemm)",
            "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": 4.0 * pts_per_launch * N + 2.0 * N * N,
            "algorithmic_bytes_note": "4 B x (points per launch x N) K* planes + 2 N^2 B lower-triangle planes of Linv "
                                      "(upper bound: the compact-support ranges read less)",
            "launches": pg_n, "avg_launch_ms": pg_ms / max(pg_n, 1),
            "algorithmic_flops_per_launch": flops_total / max(pg_n, 1),
            "executed_fraction_of_dense": flops_total / (dense_flops_per_step * steps) if dense_flops_per_step else None,
            "dense_equivalent_tflops": dense_equiv,
            "dense_equivalent_note": "N^2 FLOP per point / time: what a dense product would have to sustain for the same "
                                     "points/s; not a hardware rate (the skipped K* entries are below fp32 resolution)",
            "peak_source": peak_src,
            "note": "fp32-faithful split-fp16 product: 3 tcgen05 MMAs per algorithmic MAC, so the tensor pipe "
                    "executes 3 x achieved; achieved counts the MACs the kernel executed (128 x 256 x 32 per k-block)",
            "tensor_pipe_tflops": 3.0 * achieved, "tensor_pipe_frac": 3.0 * achieved / peak_tf}
    ch_ms, ch_n = stages["cholesky"]
    ch_tf = (N ** 3 / 3.0) * ch_n / (ch_ms * 1e-3) / 1e12 if ch_ms > 0 else 0.0
    chol = {"kernel": "blocked Cholesky (stage cholesky: diagonal blocks + tcgen05 panels and SYRK updates)",
            "bound": "tensor", "achieved": ch_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ch_tf / peak_tf,
            "algorithmic_flops": N ** 3 / 3.0, "ms": ch_ms / max(ch_n, 1), "peak_source": peak_src}
    km_ms, km_n = stages["kmat"]
    kmat_gbps = (4.0 * N * N / 2 + 8.0 * d * N) * km_n / (km_ms * 1e-3) / 1e9 if km_ms > 0 else 0.0
    kmat = {"kernel": "kmat_kernel (stage kmat)", "bound": "hbm", "achieved": kmat_gbps, "peak": hbm, "unit": "GB/s",
            "frac": kmat_gbps / hbm, "note": "lower-triangle tiles only: 2 N^2 + 8 d N bytes per launch",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}
    return roof, chol, kmat


def run_e2e(wl, steps):
    """gpim.reconstructor: swap the training set in (host -> device), predict over X_full given as
    a host numpy grid, results back as numpy arrays.  Factor cache invalidated every step, as the
    reference refactorises on every predict (gpr.py:248)."""
    import torch
    import gpim_b200 as gpim
    R = wl["R"]
    Xsp = gpim.utils.get_sparse_grid(R)
    rec = gpim.reconstructor(Xsp, R, wl["Xfull"], kernel=wl["kernel"], iterations=0, verbose=0, precision="single",
                             jitter=wl["jitter"], lengthscale=[[1.0] * R.ndim, [20.0] * R.ndim], shard=False)
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
    for _ in range(steps):
        mean, sd = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    M = mean.size
    h2d = Xh.numel() * 4 + yh.numel() * 4 + M * R.ndim * 4
    del rec
    torch.cuda.empty_cache()
    return {"value": M / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * M * 4),
            "ms_per_step": dt * 1e3, "steps": steps,
            "api": "gpim.reconstructor.predict(X_full) after model.X/.y swap (host numpy in/out)"}


def run_e2e_sharded(wl, steps, world, rank):
    """N > 1 through the kept API: every rank builds the reconstructor from host arrays and calls predict(X_full)
    (a collective: rank 0 factorises, rows tiled over the ranks, numpy (mean, sd) on every rank)."""
    import torch
    import torch.distributed as dist
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

    for _ in range(2):
        step()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        mean, sd = step()
    dist.barrier(); torch.cuda.synchronize()
    sec = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=rec.model.engine.device)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    sec = float(sec.item())
    M = mean.size
    h2d = (Xh.numel() + yh.numel()) * 4 * world + M * R.ndim * 4 * world
    del rec
    torch.cuda.empty_cache()
    return {"value": M / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * M * 4 * world),
            "ms_per_step": sec * 1e3, "steps": steps,
            "api": "gpim.reconstructor.predict(X_full) on every rank (host numpy in/out; shard='auto' under NCCL)"}


def workload_block(eng, name, steps, warmup, peaks, cpu_sample, with_torch_cuda=True, with_cpu=True, e2e_steps=None):
    """One single-GPU configuration with all its blocks: value, e2e, roofline, cholesky, kmat, cpu_baseline (+ parity),
    torch_cuda_baseline."""
    wl = make_workload(name)
    res = measure_predict(eng, wl, steps, warmup)
    roof, chol, kmat = roofline_blocks(res, wl, steps, peaks, res["ms_per_step"] * steps * 1e-3)
    blk = {"config": bench_config(wl, 1, res["N"], res["M"]), "value": res["value"], "unit": UNIT,
           "ms_per_step": res["ms_per_step"], "steps": steps, "warmup": warmup, "dtype": "f32",
           "factor_cached_ms": res["ms_cached"], "factor_cached_value": res["M"] / (res["ms_cached"] * 1e-3),
           "gpu_launches": res["launches"],
           "stages_ms_per_step": {k: v[0] / steps for k, v in res["stages"].items() if v[1]},
           "roofline": roof, "cholesky": chol, "kmat_assembly": kmat,
           "e2e": run_e2e(wl, e2e_steps or steps)}
    if with_cpu:
        blk["cpu_baseline"] = cpu_baseline_block(wl, cpu_sample, "f64", cuda_out=(res["mean"], res["sd"]))
    if with_torch_cuda:
        tcb = torch_cuda_baseline(wl)
        if "_mean" in tcb:
            tcb["agrees_with_engine"] = {"mean_relinf": relinf(tcb.pop("_mean"), res["mean"]),
                                         "sd_relinf": relinf(tcb.pop("_sd"), res["sd"])}
            tcb["engine_speedup"] = res["value"] / tcb["value"]
        blk["torch_cuda_baseline"] = tcb
    return blk, res, wl


def measure_c4(eng, peaks, steps=50, gp_iterations=1000, cpu_steps=1):
    """BASELINE.json configs[3]: GP-BO with EI on a 128 x 128 grid, 100 seed pixels, `steps` exploration steps x
    `gp_iterations` Adam iterations, all defaults (fp64), through gpim.boptimizer -- host numpy in/out by construction,
    so the run IS the end-to-end number.  CPU baseline: the oracle's bo_run on `cpu_steps` exploration step(s)."""
    import torch
    import gpim_b200 as gpim
    from gpim_b200 import _lib
    n = 128
    f = W.bo_trial_func(n)
    np.random.seed(0)
    idx = np.random.randint(0, n, size=(100, 2))
    Zs = np.full((n, n), np.nan)
    for i, j in idx:
        Zs[i, j] = f((i, j))
    X_full, X_sparse = gpim.utils.get_full_grid(Zs), gpim.utils.get_sparse_grid(Zs)
    M = n * n
    out_dir = tempfile.mkdtemp()

    def run(nsteps):
        bo = gpim.boptimizer(X_sparse, Zs.copy(), X_full, f, acquisition_function="ei", exploration_steps=nsteps,
                             gp_iterations=gp_iterations, verbose=0, filename=os.path.join(out_dir, "bo"))
        torch.cuda.synchronize()
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        bo.run()
        torch.cuda.synchronize()
        return bo, time.perf_counter() - t0, eng.launch_count() - l0

    run(1)                                               # warm-up: workspace, graph capture path, caches
    eng.set_option(_lib.OPT_STAGE_TIMING, 0)
    bo, sec, launches = run(steps)
    iters = (steps + 1) * gp_iterations
    predicts = 2 * steps                                 # dense grid + measured rows (the EI incumbent) per step
    # acquisition sweep alone, against its HBM roofline (K6: 8 B read + 4 B... per point)
    mean_d, sd_d = bo.surrogate_model._last_pred_device
    eng.acq_sweep(_lib.ACQ_IDS["ei"], mean_d, sd_d, 100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.acq_sweep(_lib.ACQ_IDS["ei"], mean_d, sd_d, 100)
    e1.record()
    torch.cuda.synchronize()
    acq_ms = e0.elapsed_time(e1) / 20
    hbm = peaks.get("hbm_gbs") or 6650.0
    eb = mean_d.element_size()
    acq_gbps = (2 * eb * M) / (acq_ms * 1e-3) / 1e9
    blk = {"config": {"workload": "GP-BO, EI, 128x128 grid, 100 seed pixels, RBF, fp64 (reference defaults)", "name": "c4",
                      "exploration_steps": steps, "gp_iterations": gp_iterations, "M_grid": M, "N_train": f"100..{100 + steps}"},
           "value": M * steps / sec, "unit": UNIT, "dtype": "f64",
           "value_note": "dense-grid points predicted (mean + sd) per second over the WHOLE run: trainings, predicts, "
                         "acquisition sweeps and the host loop included",
           "seconds_whole_run": sec, "adam_iterations": iters, "ms_per_adam_iteration_incl_everything": 1e3 * sec / iters,
           "gpu_launches": launches, "dense_predicts": predicts,
           "e2e": {"value": M * steps / sec, "unit": UNIT, "h2d_bytes_per_step": int((100 + steps) * 3 * 8 + M * 2 * 8),
                   "d2h_bytes_per_step": int(2 * M * 8 + gp_iterations * 6 * 8),
                   "api": "gpim.boptimizer(...).run(): numpy in, numpy out; per exploration step the grown training set and "
                          "X_full go host -> device, (mean, sd) and the Adam trajectory come back"},
           "roofline": {"kernel": "acquisition sweep + top-k (gpg_acq_sweep) on the 128 x 128 grid", "bound": "hbm",
                        "achieved": acq_gbps, "peak": hbm, "unit": "GB/s", "frac": acq_gbps / hbm, "traffic": None,
                        "avg_launch_ms": acq_ms,
                        "note": "the run as a whole is latency-bound (N <= 150: ~35 dependent few-microsecond kernels per "
                                "Adam iteration, replayed as a CUDA graph); this entry is the one HBM-bound kernel of the "
                                "config -- 16 384 points are 256 KB, far too small to reach the HBM roofline"},
           "picks": [list(map(int, p)) for p in bo.indices_all[:5]]}
    try:
        import torch as _t
        from oracle import gp_oracle as O
        _t.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        O.bo_run(X_sparse, Zs.copy(), X_full, f, acquisition="ei", exploration_steps=cpu_steps, gp_iterations=gp_iterations)
        csec = time.perf_counter() - t0
        per_step = csec / (cpu_steps + 1)                # cpu_steps + 1 trainings dominate
        blk["cpu_baseline"] = {"value": M * steps / (per_step * (steps + 1)), "unit": UNIT, "cores": _t.get_num_threads(),
                               "kind": "port", "dtype": "f64",
                               "sample": f"oracle bo_run, {cpu_steps} exploration step(s) = {cpu_steps + 1} trainings x "
                                         f"{gp_iterations} Adam iterations + {2 * cpu_steps} dense predicts in {csec:.1f} s; "
                                         f"whole run scaled to {steps + 1} trainings",
                               "ms_per_adam_iteration": 1e3 * csec / ((cpu_steps + 1) * gp_iterations)}
    except Exception as e:                                   # noqa: BLE001
        blk["cpu_baseline"] = {"error": repr(e)[:200]}
    return blk


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
    return out


def measure_dense(eng, name, steps, warmup, peaks):
    """GPG_OPT_COMPACT_SUPPORT = 0: the variance GEMM over ALL training rows for every tile of test points -- what the
    default degenerates to when nothing of K* is negligible (long lengthscales, unordered training rows)."""
    from gpim_b200 import _lib
    eng.set_option(_lib.OPT_COMPACT_SUPPORT, 0)
    try:
        wl = make_workload(name)
        res = measure_predict(eng, wl, steps, warmup, keep_outputs=False)
    finally:
        eng.set_option(_lib.OPT_COMPACT_SUPPORT, 1)
    roof, _, _ = roofline_blocks(res, wl, steps, peaks, res["ms_per_step"] * steps * 1e-3)
    return {"ms_per_step": res["ms_per_step"], "value": res["value"], "unit": UNIT, "factor_cached_ms": res["ms_cached"],
            "roofline": {k: roof[k] for k in ("achieved", "peak", "frac", "unit", "tensor_pipe_frac", "avg_launch_ms")},
            "note": "dense variance GEMM (GPG_OPT_COMPACT_SUPPORT = 0); outputs equal the default's to fp32 rounding "
                    "(tests/test_gpu_parity.py::test_predict_compact_support_option_is_exact)"}


def measure_f64(eng, name, steps, warmup):
    """The reference's default precision (gpr.py:92, precision='double'): the same step in fp64 -- blocked SIMT Cholesky,
    variance product on DMMA tiles (mma.sync.m8n8k4.f64), compact support at fp64 resolution."""
    wl = make_workload(name)
    res = measure_predict(eng, wl, steps, warmup, keep_outputs=False, dtype_name="f64")
    N, M = res["N"], res["M"]
    st = {k: v[0] / steps for k, v in res["stages"].items() if v[1]}
    out = {"workload": wl["label"], "dtype": "f64", "ms_per_step": res["ms_per_step"],
           "value": res["value"], "unit": UNIT, "factor_cached_ms": res["ms_cached"], "stages_ms_per_step": st,
           "round_1": "307 ms per step (SIMT product, dense)"}
    if st.get("pgemm"):
        out["variance_product"] = {"kernel": "gemm_dmma_kernel (COLSUMSQ epilogue)", "ms_per_step": st["pgemm"],
                                   "dense_equivalent_tflops": N * N * M / (st["pgemm"] * 1e-3) / 1e12,
                                   "peak": 40.0, "peak_source": "B200 data sheet fp64 (tensor = vector = 40 TFLOP/s); "
                                   "no measured fp64 figure in MEASURED_PEAKS.json",
                                   "note": "dense-equivalent rate: N^2 FLOP per point / time; the compact support of K* "
                                           "at fp64 resolution skips about half of it, so the executed rate is about "
                                           "half the figure"}
    return out


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from gpim_b200 import _lib, sharded

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
        # JSON line by pointing fd 1 at stderr while the process groups come up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            eng = _lib.get_engine(local)
            sharded.ensure_comm(eng)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    eng = _lib.get_engine(local)
    peaks = load_peaks()
    warmup = max(args.warmup, 3)

    wl = make_workload(args.workload, dense=world)
    res = measure_predict(eng, wl, args.steps, warmup, world, rank)
    clk = clocks.stop(*res["wall"]) if clocks else None
    e2e = run_e2e(wl, args.steps) if world == 1 else run_e2e_sharded(wl, max(2, args.steps // 2), world, rank)

    strong = None
    bmode = None
    if world > 1 and not args.no_extra:
        rb = measure_predict(eng, wl, max(3, args.steps // 2), 3, world, rank, keep_outputs=False, factor_mode="broadcast")
        bmode = {"ms_per_step": rb["ms_per_step"], "value": rb["value"], "unit": UNIT,
                 "shard_parity_max_rel": rb.get("shard_parity_max_rel"),
                 "what": "rank 0 factorises alone; gpg_predict_sharded broadcasts its cache in row blocks under the "
                         "first K* tiles"}
        # BASELINE.json configs[4] as configured: M fixed at 1024 x 1024, N = 30 757, tiles of M / world rows
        c5 = make_workload("c5")
        r5 = measure_predict(eng, c5, max(2, args.steps // 4), 2, world, rank, keep_outputs=False)
        if rank == 0:
            st5 = {k: v[0] / r5["steps"] for k, v in r5["stages"].items() if v[1]}
            serial = sum(st5.get(k, 0.0) for k in ("kmat", "cholesky", "trtri", "solve"))
            strong = {"config": bench_config(c5, world, r5["N"], r5["M"]), "scaling": "strong", "value": r5["value"],
                      "unit": UNIT, "ms_per_step": r5["ms_per_step"], "steps": r5["steps"], "rows_per_gpu": r5["m_local"],
                      "stages_ms_per_step_rank0": st5,
                      "serial_ms_rank0": serial, "serial_fraction_of_step": serial / r5["ms_per_step"],
                      "serial_note": "K assembly + Cholesky + inverse + solves do not shard (replicas-only): every rank runs "
                                     "them on the same inputs; they are the serial fraction of the strong-scaling step",
                      "shard_parity_max_rel": r5.get("shard_parity_max_rel"), "gpu_launches": r5["launches"]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roof, chol, kmat = roofline_blocks(res, wl, args.steps, peaks, res["ms_per_step"] * args.steps * 1e-3)
    if world > 1:
        roof["traffic"] = None
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(wl, world, res["N"], res["M"]),
        "roofline": roof, "cholesky": chol, "kmat_assembly": kmat,
        "stages_ms_per_step": {k: v[0] / args.steps for k, v in res["stages"].items() if v[1]},
        "e2e": e2e, "gpu_launches": res["launches"], "clocks": clk,
    }
    if res["ms_cached"]:
        line["factor_cached"] = {"ms_per_step": res["ms_cached"], "value": res["M"] / (res["ms_cached"] * 1e-3), "unit": UNIT}
    if "shard_parity_max_rel" in res:
        line["shard_parity_max_rel"] = res["shard_parity_max_rel"]
    if strong is not None:
        line["strong_c5"] = strong
    if bmode is not None:
        line["broadcast_mode"] = bmode
    if world == 1 and not args.no_cpu_baseline:
        # the parity reference of north_star is the fp64 oracle; the f32 sample is the timing baseline of the fp32 config
        cb = cpu_baseline_block(wl, args.cpu_sample, args.cpu_dtype, cuda_out=None)
        par = cpu_baseline_block(wl, 1024, "f64", cuda_out=(res["mean"], res["sd"]))
        cb["parity_vs_oracle"] = par["parity_vs_oracle"]
        cb["f64"] = {"value": par["value"], "stages_s": par["stages_s"], "sample": par["sample"]}
        line["cpu_baseline"] = cb
        line["parity_vs_oracle"] = par["parity_vs_oracle"]
    if world == 1 and not args.no_extra:
        tcb = torch_cuda_baseline(wl)
        if "_mean" in tcb:
            tcb["agrees_with_engine"] = {"mean_relinf": relinf(tcb.pop("_mean"), res["mean"]),
                                         "sd_relinf": relinf(tcb.pop("_sd"), res["sd"])}
            tcb["engine_speedup"] = res["value"] / tcb["value"]
        line["torch_cuda_baseline"] = tcb
        wls = {}
        if args.workload == "c2":
            for name, st_, cs in (("h512", max(3, args.steps // 4), 1024), ("c3", max(3, args.steps // 4), 1024)):
                blk, _, _ = workload_block(eng, name, st_, 2, peaks, cs, with_cpu=not args.no_cpu_baseline,
                                           e2e_steps=max(2, st_ // 2))
                wls[name] = blk
            wls["c4"] = measure_c4(eng, peaks, steps=args.c4_steps)
        line["workloads"] = wls
        line["extra_workloads"] = {"train_c2": measure_training(eng, "c2", 10),
                                   "c2_f64": measure_f64(eng, "c2", 3, 1),
                                   "c2_dense": measure_dense(eng, "c2", max(3, args.steps // 2), 2, peaks),
                                   "h512_dense": measure_dense(eng, "h512", max(3, args.steps // 4), 2, peaks),
                                   "sparse_c2": measure_sparse(eng, "c2", 10, max(2, args.steps // 2))}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu-sample", type=int, default=4096, dest="cpu_sample",
                    help="grid rows of the bounded CPU sample in the CUDA arm's cpu_baseline")
    ap.add_argument("--cpu-full-steps", type=int, default=1, dest="cpu_full_steps",
                    help="--impl reference: cap on the number of full-workload steps timed (max 2)")
    ap.add_argument("--cpu-dtype", default="f32", choices=["f32", "f64"], dest="cpu_dtype")
    ap.add_argument("--cpu-max-rows", type=int, default=131072, dest="cpu_max_rows",
                    help="--impl reference: most grid rows of one CPU step (the whole grid when it has no more)")
    ap.add_argument("--c4-steps", type=int, default=50, dest="c4_steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    args = ap.parse_args()
    args.cpu_full_steps = max(1, min(args.cpu_full_steps, 2))
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
