"""
Generates tests/golden/oracle_full_{c2,c3,h512}.npz: the fp64 oracle (oracle/gp_oracle.py, pinned by the reference's
own goldens, tests/test_oracle.py) on BASELINE.json's configurations at FULL training-set size -- C2 (256 x 256
spiral, N = 7 688), C3 (64 x 64 x 16 hyperspectral, Matern52, N = 19 744) and the 512 x 512 headline (N = 15 377) --
evaluated on 1 024 rows spread evenly over the dense grid.  The full-size GPU parity tests compare the tcgen05 path
against these vectors (mean 1e-4, sd 1e-3, north_star) with no oracle in the loop; C2 is also checked against the
live oracle (tests/test_gpu_parity.py::test_full_size_c2_against_live_oracle).

    python tests/golden/make_fullsize_vectors.py [c2 c3 h512]     (CPU, ~5 min for all three on 8 cores)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as W                      # noqa: E402
from oracle import gp_oracle as O          # noqa: E402

M_SAMPLE = 1024


def main(names):
    torch.set_num_threads(os.cpu_count() or 1)
    for name in names:
        wl = W.make_workload(name)
        X, y = O.training_rows(O.sparse_grid(wl["R"]), wl["R"])
        Xs = W.rows_of(wl["Xfull"])
        sel = W.sample_rows(Xs.shape[0], M_SAMPLE)
        th = wl["theta"]
        t0 = time.perf_counter()
        mean, sd, _ = O.predict_fixed_theta(wl["kernel"], X, y, Xs[sel], th[0], th[3:], th[1], jitter=wl["jitter"],
                                            dtype=torch.float64, scale_mixture=th[2])
        out = os.path.join(ROOT, "tests", "golden", f"oracle_full_{name}.npz")
        np.savez_compressed(out, sel=sel, mean=mean, sd=sd, N=X.shape[0], M=Xs.shape[0], theta=np.array(th),
                            jitter=wl["jitter"], kernel=wl["kernel"])
        print(f"{name}: N={X.shape[0]} M={Xs.shape[0]} {time.perf_counter() - t0:.1f} s -> {out}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["c2", "c3", "h512"])
