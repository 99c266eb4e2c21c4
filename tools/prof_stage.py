"""One isolated stage of the hot path at workload size, for `ncu -k regex:<kernel>` captures.
usage: python tools/prof_stage.py <kmat|chol|factor|predict|step|fit> [workload=c2] [reps=1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gpim_b200._lib import get_engine, KERNEL_IDS  # noqa: E402

stage = sys.argv[1] if len(sys.argv) > 1 else "step"
name = sys.argv[2] if len(sys.argv) > 2 else "c2"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = get_engine()
wl = bench.make_workload(name)
X, y = bench.train_rows(wl["R"])
Xs = bench.rows_of(wl["Xfull"])[:16384]
dt = torch.float32
kid = KERNEL_IDS[wl["kernel"]]
th = torch.tensor(wl["theta"], dtype=dt).cuda()
Xd, yd, Xsd = (torch.tensor(a, dtype=dt).cuda() for a in (X, y, Xs))
N = len(y)
fac = eng.alloc_factor(N, dt)
ld = fac["ld"]
if stage in ("predict", "chol"):
    eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
K = torch.empty(N, ld, dtype=dt, device="cuda")
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off captures from here
for _ in range(reps):
    if stage == "kmat":
        eng.kmat(kid, th, Xd, None, jitter=wl["jitter"], out=K)
        eng.kmat(kid, th, Xd, None, jitter=wl["jitter"], lower_only=True, out=K)
    elif stage == "chol":
        eng.kmat(kid, th, Xd, None, jitter=wl["jitter"], out=K)
        eng.cholesky_(K)
    elif stage == "factor":
        eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
    elif stage == "predict":
        eng.predict(kid, th, Xd, fac, Xsd)
    elif stage == "acq":
        g = torch.Generator(device="cuda").manual_seed(0)
        mu = torch.randn(1 << 20, device="cuda", generator=g)
        sig = torch.rand(1 << 20, device="cuda", generator=g) + 0.1
        eng.acq_sweep(1, mu, sig, 100, mu_best=1.0, xi=0.01, want_acq=True)
    elif stage == "fit":
        u = torch.zeros(3 + Xd.shape[1], dtype=dt, device="cuda")
        eng.fit_adam(kid, Xd, yd, wl["jitter"], u, [1e-4, 10.0] + [1.0] * Xd.shape[1] + [20.0] * Xd.shape[1],
                     Xd.shape[1], 2, 0.1)
    else:
        eng.factorize(kid, th, Xd, yd, wl["jitter"], out=fac)
        eng.predict(kid, th, Xd, fac, Xsd)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", stage, name, "N", N, "launches", eng.launch_count())
