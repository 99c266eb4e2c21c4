"""A short pass through every C-ABI entry on ragged sizes, for `compute-sanitizer --tool memcheck|racecheck`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpim_b200._lib import get_engine, KERNEL_IDS  # noqa: E402

eng = get_engine()
for dtype, n, d in ((torch.float32, 1111, 2), (torch.float32, 300, 3), (torch.float64, 257, 2)):
    rng = np.random.RandomState(n)
    X = torch.tensor(rng.rand(n, d) * 30.0, dtype=dtype).cuda()
    y = torch.sin(X[:, 0] / 4.0) + 0.1 * torch.tensor(rng.randn(n), dtype=dtype).cuda()
    th = torch.tensor([0.7, 0.02, 1.2] + [3.0 + k for k in range(d)], dtype=dtype).cuda()
    for kname in ("RBF", "Matern52", "RationalQuadratic"):
        kid = KERNEL_IDS[kname]
        fac = eng.factorize(kid, th, X, y, 1e-5)
        Xs = torch.tensor(rng.rand(1003, d) * 30.0, dtype=dtype).cuda()
        Xs[5, 0] = float("nan")
        m, s = eng.predict(kid, th, X, fac, Xs)
        eng.set_option(10, 1)                    # GPG_OPT_COMPACT_SUPPORT
        eng.predict(kid, th, X, fac, Xs)
        eng.set_option(10, 0)
        eng.predict_grid(kid, th, X, fac, [31] * d, [1.0] * d, 7, 500)
        eng.nll_grad(kid, th, X, y, 1e-5)
    u = torch.zeros(3 + d, dtype=dtype, device="cuda")
    eng.fit_adam(0, X, y, 1e-5, u, [1e-4, 10.0] + [1.0] * d + [10.0] * d, d, 9, 0.1)
    K = eng.kmat(0, th, X, None, jitter=1e-5)
    L, info = eng.cholesky_(K.clone())
    Li = eng.trtri(L)
    eng.solve_vec(L, Li, y)
    eng.acq_sweep(1, m.nan_to_num(), s.nan_to_num() + 0.1, 100, mu_best=0.5)
    if dtype == torch.float32:
        eng.gemm_nt(torch.randn(200, 100, device="cuda"), torch.randn(300, 100, device="cuda"))
torch.cuda.synchronize()
print("sanitize_run ok, launches", eng.launch_count())
