"""Stage timings of the hot path on one GPU (development aid; bench.py is the contract)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from gpim_b200._lib import get_engine, KERNEL_IDS  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main(n=256, path=0, dtype=torch.float32, chunk=0):
    eng = get_engine()
    eng.set_option(1, path)
    eng.set_option(2, chunk)
    if os.environ.get("GPG_COMPACT_SUPPORT"):
        eng.set_option(10, int(os.environ["GPG_COMPACT_SUPPORT"]))
    if os.environ.get("GPG_INNER_LEFT"):
        eng.set_option(11, int(os.environ["GPG_INNER_LEFT"]))
    if os.environ.get("GPG_OUTER_PANEL"):
        eng.set_option(9, int(os.environ["GPG_OUTER_PANEL"]))
    R = W.spiral_scan(n)
    X, y = O.training_rows(O.sparse_grid(R), R)
    N, M = len(y), n * n
    ft = W.FIXED_THETA
    th = torch.tensor([ft["variance"], ft["noise"], 1.0, ft["lengthscale"], ft["lengthscale"]], dtype=dtype).cuda()
    Xd, yd = torch.tensor(X, dtype=dtype).cuda(), torch.tensor(y, dtype=dtype).cuda()
    kid = KERNEL_IDS["RBF"]
    ld = (N + 63) // 64 * 64
    K = torch.empty(N, ld, dtype=dtype, device="cuda")
    t_k = timed(lambda: eng.kmat(kid, th, Xd, None, jitter=ft["jitter"], out=K))
    esz = K.element_size()
    print(f"n={n} N={N} M={M} dtype={dtype} path={path}")
    print(f"  kmat        {t_k*1e3:9.3f} ms  {esz*N*N/t_k/1e9:8.1f} GB/s")
    Kc = K.clone()
    t_c = timed(lambda: (K.copy_(Kc), eng.cholesky_(K))) - timed(lambda: K.copy_(Kc))
    print(f"  cholesky    {t_c*1e3:9.3f} ms  {N**3/3/t_c/1e12:8.2f} TFLOP/s")
    K.copy_(Kc)
    eng.cholesky_(K)
    Linv = torch.empty_like(K)
    t_t = timed(lambda: eng.trtri(K, out=Linv))
    print(f"  trtri       {t_t*1e3:9.3f} ms  {N**3/3/t_t/1e12:8.2f} TFLOP/s")
    t_f = timed(lambda: eng.factorize(kid, th, Xd, yd, ft["jitter"]))
    print(f"  factorize   {t_f*1e3:9.3f} ms")
    fac = eng.factorize(kid, th, Xd, yd, ft["jitter"])
    t_p = timed(lambda: eng.predict_grid(kid, th, Xd, fac, [n, n], [1.0, 1.0], 0, M))
    print(f"  predict     {t_p*1e3:9.3f} ms  {N*N*M/t_p/1e12:8.2f} TFLOP/s(alg)  {M/t_p:12.0f} pts/s (factor cached)")
    print(f"  end-to-end  {(t_f+t_p)*1e3:9.3f} ms  {M/(t_f+t_p):12.0f} pts/s")
    u = torch.zeros(5, dtype=dtype, device="cuda")
    t_a = timed(lambda: eng.fit_adam(kid, Xd, yd, ft["jitter"], u.clone(), [1e-4, 10, 1, 1, 4, 4], 2, 3, 0.1), reps=1) / 3
    print(f"  adam iter   {t_a*1e3:9.3f} ms")
    print("  launches", eng.launch_count())


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    path = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dt = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
    main(n, path, dt, int(sys.argv[4]) if len(sys.argv) > 4 else 0)
