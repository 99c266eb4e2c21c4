"""One factorize + predict at workload size (for ncu launch lists)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gpim_b200._lib import get_engine, KERNEL_IDS
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = get_engine()
wl = bench.make_workload(name)
X, y = bench.train_rows(wl["R"])
Xs = bench.rows_of(wl["Xfull"])[:16384]
dt = torch.float32
th = torch.tensor(wl["theta"], dtype=dt).cuda()
Xd, yd, Xsd = (torch.tensor(a, dtype=dt).cuda() for a in (X, y, Xs))
fac = eng.alloc_factor(len(y), dt)
for _ in range(reps):
    eng.factorize(KERNEL_IDS[wl["kernel"]], th, Xd, yd, wl["jitter"], out=fac)
    eng.predict(KERNEL_IDS[wl["kernel"]], th, Xd, fac, Xsd)
torch.cuda.synchronize()
print("ok", eng.launch_count())
