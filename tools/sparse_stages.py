"""Stage clock of the inducing-point Adam iteration at a bench workload (development aid).
usage: python tools/sparse_stages.py [c2|h512|c3] [iters] [f32|f64]
Stages: kmat = Kuu + Kuf assembly, cholesky / trtri = both m x m factorisations, pgemm = B = Ui Kuf,
pfinal = S = B B^T, solve = vectors, grad = everything after the factorisations (A'^-1, the m x m chain,
dF/dKuf = T2 B, the two fused kernel-derivative reductions, finish)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gpim_b200._lib import get_engine, KERNEL_IDS, OPT_STAGE_TIMING  # noqa: E402


def main(name="c2", iters=10, only=None):
    eng = get_engine()
    if os.environ.get("GPG_GEMM_PATH"):               # 1: all-SIMT, 0 (default): tcgen05 for the large fp32 products
        eng.set_option(1, int(os.environ["GPG_GEMM_PATH"]))
    wl = bench.make_workload(name)
    X, y = bench.train_rows(wl["R"])
    N, d = X.shape
    m_ind = N // 10
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        if only and tag != only:
            continue
        Xd, yd = torch.tensor(X, dtype=dt).cuda(), torch.tensor(y, dtype=dt).cuda()
        bounds = [1e-4, 10.0] + [1.0] * d + [4.0] * d
        for timing in (0, 1):
            eng.set_option(OPT_STAGE_TIMING, timing)
            Xu = Xd[::N // m_ind].clone()
            u = torch.zeros(3 + d, dtype=dt).cuda()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.sparse_fit_adam(KERNEL_IDS[wl["kernel"]], Xd, yd, Xu, 1e-4, u, bounds, d, iters, 0.05, record_xu=False)
            e1.record()
            torch.cuda.synchronize()
            if timing:
                st = eng.stage_times()
                print(tag, f"N={N} m={Xu.shape[0]}", f"total {e0.elapsed_time(e1) / iters:.3f} ms/iter |",
                      " ".join(f"{k} {v[0] / iters:.3f}" for k, v in st.items() if v[1]))
        eng.set_option(OPT_STAGE_TIMING, 0)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "c2", int(sys.argv[2]) if len(sys.argv) > 2 else 10,
         sys.argv[3] if len(sys.argv) > 3 else None)
