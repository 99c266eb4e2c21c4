"""Wall time of the reference's BO known-answer run (25x25, EI, 20 steps x 1000 Adam iterations, fp64) and of a
C4-shaped run (128x128, 100 seeds, EI) shortened to `steps` exploration steps.  usage: python tools/bo_bench.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpim_b200 as gpim  # noqa: E402
from gpim_b200._lib import get_engine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def trial_func(idx, n):
    x0, y0, fwhm = 0.4 * n, 0.6 * n, 0.18 * n
    return float(np.exp(-4 * np.log(2) * ((idx[0] - x0) ** 2 + (idx[1] - y0) ** 2) / fwhm ** 2))


for n, seeds, st in ((25, 5, 20), (128, 100, steps)):
    np.random.seed(0)
    idx = np.random.randint(0, n, size=(2, seeds))
    Z = np.full((n, n), np.nan)
    for i, j in zip(*idx):
        Z[i, j] = trial_func([i, j], n)
    X_full = gpim.utils.get_full_grid(Z)
    X_sparse = gpim.utils.get_sparse_grid(Z)
    for graph in (0, 1):
        get_engine().set_option(7, graph)
        bo = gpim.boptimizer(X_sparse, Z.copy(), X_full, lambda i, n=n: trial_func(i, n), acquisition_function="ei",
                             exploration_steps=st, use_gpu=False, verbose=0, filename="/tmp/bo_bench")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bo.run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        iters = (st + 1) * 1000
        print(f"{n}x{n} seeds={seeds} steps={st} graph={graph}: {dt:.2f} s  ({1e3 * dt / iters:.3f} ms per Adam iteration incl. "
              f"predict/acquisition), picks {bo.indices_all[-3:]}")
