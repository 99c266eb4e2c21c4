"""Timing + correctness of the acquisition sweep / top-k at large M (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpim_b200._lib import get_engine, ACQ_IDS  # noqa: E402


def reference_topk(acq, k):
    """reversed arg-sort (boptim.py:303-306): NaN first, ties by larger flat index"""
    order = np.argsort(acq, kind="stable")[::-1]
    return order[:k]


def main():
    eng = get_engine()
    g = torch.Generator(device="cuda").manual_seed(0)
    for dtype in (torch.float32, torch.float64):
        for M in (1 << 14, 1 << 20):
            mean = torch.randn(M, dtype=dtype, device="cuda", generator=g)
            sd = torch.rand(M, dtype=dtype, device="cuda", generator=g) + 0.1
            for name, k in (("ei", 100), ("ei", 1024), ("poi", 100), ("cb", 1)):
                for case in ("random", "plateau"):
                    m = mean.clone()
                    if case == "plateau":
                        m[: M // 2] = 3.0                   # half the grid ties (EI / POI / CB all equal there)
                        s_ = sd.clone(); s_[: M // 2] = 0.5
                    else:
                        s_ = sd
                    kw = dict(mu_best=0.5, xi=0.01, alpha=1.0, beta=1.0)
                    vals, idx, count, acq = eng.acq_sweep(ACQ_IDS[name], m, s_, k, want_acq=True, **kw)
                    torch.cuda.synchronize()
                    ref = reference_topk(acq.cpu().numpy(), k)
                    ok = bool(np.array_equal(ref, idx.cpu().numpy()))
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        eng.acq_sweep(ACQ_IDS[name], m, s_, k, **kw)
                    e1.record()
                    torch.cuda.synchronize()
                    print(f"{str(dtype):14s} M={M:8d} {name:3s} k={k:5d} {case:8s}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per sweep + top-k   "
                          f"order == reversed argsort: {ok}")
                    assert ok


if __name__ == "__main__":
    main()
