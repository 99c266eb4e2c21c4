"""Accuracy of gpg_factorize (f32) against numpy fp64 on random 2-D point sets: SIMT path vs tcgen05 recursion.
usage: python tools/factor_accuracy.py [noise]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gp_oracle as O  # noqa: E402
from gpim_b200._lib import get_engine, KERNEL_IDS, OPT_GEMM_PATH  # noqa: E402

eng = get_engine()
noise = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-2
if len(sys.argv) > 4:
    eng.set_option(6, int(sys.argv[4]))     # GPG_OPT_FACTOR_ALGO
if len(sys.argv) > 3:
    eng.set_option(4, int(sys.argv[2]))     # GPG_OPT_PANEL_REFINE
    eng.set_option(5, int(sys.argv[3]))     # GPG_OPT_SYRK_CHUNK
    print("panel_refine", sys.argv[2], "syrk_chunk", sys.argv[3])
v, ls, jitter = 0.5, [3.0, 4.0], 1e-5
for n in (1111, 2500, 5000):
    rng = np.random.RandomState(n)
    X = rng.rand(n, 2) * 40.0
    y = np.sin(X[:, 0] / 5.0) + 0.1 * rng.randn(n)
    K = O.kernel_matrix("RBF", torch.tensor(X), torch.tensor(X), torch.tensor(v).double(), torch.tensor(ls).double(),
                        torch.tensor(1.0).double()).numpy() + (noise + jitter) * np.eye(n)
    Lref = np.linalg.cholesky(K)
    aref = np.linalg.solve(K, y)
    th = torch.tensor([v, noise, 1.0, *ls], dtype=torch.float32).cuda()
    for path in (1, 2):
        eng.set_option(OPT_GEMM_PATH, path)
        fac = eng.factorize(KERNEL_IDS["RBF"], th, torch.tensor(X, dtype=torch.float32).cuda(),
                            torch.tensor(y, dtype=torch.float32).cuda(), jitter)
        eng.set_option(OPT_GEMM_PATH, 0)
        L = torch.tril(fac["L"][:, :n]).cpu().double().numpy()
        Li = fac["Linv"][:, :n].cpu().double().numpy()
        a = fac["alpha"].cpu().double().numpy()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.set_option(OPT_GEMM_PATH, path)
        e0.record()
        eng.factorize(KERNEL_IDS["RBF"], th, torch.tensor(X, dtype=torch.float32).cuda(), torch.tensor(y, dtype=torch.float32).cuda(), jitter, out=fac)
        e1.record()
        torch.cuda.synchronize()
        eng.set_option(OPT_GEMM_PATH, 0)
        ms = e0.elapsed_time(e1)
        print(f"n={n:5d} noise={noise:g} path={path}: N v/nz={n*v/(noise+jitter):.1e} L relinf {np.abs(L-Lref).max()/np.abs(Lref).max():.1e} "
              f"backward |LL^T-K|/|K| {np.abs(L@L.T-K).max()/np.abs(K).max():.1e} |Li L-I| {np.abs(Li@Lref-np.eye(n)).max():.1e} "
              f"alpha relinf {np.abs(a-aref).max()/np.abs(aref).max():.1e} logdet err {float(fac['scalars'][1])-np.log(np.diag(Lref)).sum():+.2e} "
              f"quad err {float(fac['scalars'][0])-0.5*y@aref:+.2e} (quad {0.5*y@aref:.1f}) {ms:.2f} ms")
