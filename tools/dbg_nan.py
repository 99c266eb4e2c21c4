import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W
from oracle import gp_oracle as O
from gpim_b200._lib import get_engine, KERNEL_IDS
eng = get_engine()
dtype = torch.float64
R = W.dummy_blob()
X, y = O.training_rows(O.sparse_grid(R), R)
Xs = O.to_rows(O.sparse_grid(R))
good = ~np.isnan(Xs).any(axis=1)
print("rows", Xs.shape, "good", good.sum())
th = torch.tensor([0.5, 1e-3, 1.3, 12.0, 9.0], dtype=dtype).cuda()
Xd, yd = torch.tensor(X, dtype=dtype).cuda(), torch.tensor(y, dtype=dtype).cuda()
fac = eng.factorize(0, th, Xd, yd, 1e-5)
Xsd = torch.tensor(Xs, dtype=dtype).cuda()
print("device nan rows", int(torch.isnan(Xsd).any(1).sum()))
mean = torch.full((len(Xs),), 7.0, dtype=dtype, device="cuda"); sd = torch.full((len(Xs),), 7.0, dtype=dtype, device="cuda")
eng.predict(0, th, Xd, fac, Xsd, mean=mean, sd=sd)
m = mean.cpu().numpy(); s = sd.cpu().numpy()
print("mean nan", np.isnan(m).sum(), "sd nan", np.isnan(s).sum(), "mean==7", (m == 7).sum(), "sd==7", (s==7).sum())
bad_notnan = np.nonzero(~good & ~np.isnan(m))[0]
print("bad rows with finite mean:", bad_notnan[:20], m[bad_notnan[:5]], Xs[bad_notnan[:5]])
good_nan = np.nonzero(good & np.isnan(m))[0]
print("good rows with nan mean:", good_nan[:20])
