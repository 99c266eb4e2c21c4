"""Stage clocks and accuracy of gpg_factorize under the blocked-Cholesky panel modes (development aid).
usage: python tools/chol_modes.py [workload ...]      (default: c2 h512)
Modes: 1 = per-step launches (diag block, tcgen05 panel GEMM, inner update), 3 = cooperative panel kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from gpim_b200 import _lib  # noqa: E402
from gpim_b200._lib import get_engine, KERNEL_IDS  # noqa: E402

eng = get_engine()
names = sys.argv[1:] or ["c2", "h512"]
for name in names:
    wl = W.make_workload(name)
    X, y = O.training_rows(O.sparse_grid(wl["R"]), wl["R"])
    N = len(y)
    kid = KERNEL_IDS[wl["kernel"]]
    th = torch.tensor(wl["theta"], dtype=torch.float32).cuda()
    Xd, yd = torch.tensor(X, dtype=torch.float32).cuda(), torch.tensor(y, dtype=torch.float32).cuda()
    fac = eng.alloc_factor(N, torch.float32)
    ref = None
    if N <= 8000:
        t = wl["theta"]
        K = O.kernel_matrix(wl["kernel"], torch.tensor(X), torch.tensor(X), torch.tensor(t[0]).double(),
                            torch.tensor(t[3:]).double(), torch.tensor(t[2]).double()).numpy()
        K[np.diag_indices(N)] += t[1] + wl["jitter"]
        ref = (K, np.linalg.cholesky(K), np.linalg.solve(K, y))
    for mode, ahead in ((1, 0), (3, 0), (3, 1)):
        eng.set_option(_lib.OPT_PANEL_MODE, mode)
        eng.set_option(_lib.OPT_LOOKAHEAD, ahead)
        for _ in range(2):
            eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
        torch.cuda.synchronize()
        eng.set_option(_lib.OPT_STAGE_TIMING, 1)
        eng.stage_times()
        reps = 5
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
        e1.record()
        torch.cuda.synchronize()
        st = eng.stage_times()
        eng.set_option(_lib.OPT_STAGE_TIMING, 0)
        chol = st["cholesky"][0] / reps
        line = (f"{name} N={N} panel_mode={mode} lookahead={ahead}: factorize {e0.elapsed_time(e1) / reps:8.3f} ms  cholesky {chol:8.3f} ms "
                f"({N ** 3 / 3 / (chol * 1e-3) / 1e12:6.1f} TFLOP/s)  trtri {st['trtri'][0] / reps:7.3f}  solve {st['solve'][0] / reps:6.3f}  "
                f"launches/factorize {(eng.launch_count() - l0) // reps}  info {int(fac['info'].item())}")
        if ref is not None:
            K, Lref, aref = ref
            L = torch.tril(fac["L"][:, :N]).cpu().double().numpy()
            a = fac["alpha"].cpu().double().numpy()
            line += (f"  | L relinf {np.abs(L - Lref).max() / np.abs(Lref).max():.1e}  |LL^T-K|/|K| {np.abs(L @ L.T - K).max() / np.abs(K).max():.1e}"
                     f"  alpha relinf {np.abs(a - aref).max() / np.abs(aref).max():.1e}"
                     f"  logdet err {float(fac['scalars'][1]) - np.log(np.diag(Lref)).sum():+.2e}")
        print(line, flush=True)
    eng.set_option(_lib.OPT_PANEL_MODE, 3)
    eng.set_option(_lib.OPT_LOOKAHEAD, 1)
    del fac
    torch.cuda.empty_cache()
