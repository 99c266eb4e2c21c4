"""fp32 paths (tcgen05 split-fp16, SIMT) against the engine's own fp64 path at full workload size.
usage: python tools/accuracy_full.py c2|h512|c3"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gpim_b200._lib import get_engine, KERNEL_IDS, OPT_GEMM_PATH

def relinf(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())

def main(name, sub=16384, noise=None):
    eng = get_engine()
    wl = bench.make_workload(name)
    if noise is not None:
        wl["theta"][1] = noise
    X, y = bench.train_rows(wl["R"])
    Xs = bench.rows_of(wl["Xfull"])
    sel = np.linspace(0, len(Xs) - 1, min(sub, len(Xs))).astype(np.int64)
    Xs = Xs[sel]
    kid = KERNEL_IDS[wl["kernel"]]
    out = {}
    for tag, dt, path in (("f64", torch.float64, 1), ("f32_simt", torch.float32, 1), ("f32_tc", torch.float32, 2)):
        eng.set_option(OPT_GEMM_PATH, path)
        th = torch.tensor(wl["theta"], dtype=dt).cuda()
        Xd, yd, Xsd = (torch.tensor(a, dtype=dt).cuda() for a in (X, y, Xs))
        fac = eng.factorize(kid, th, Xd, yd, wl["jitter"])
        assert int(fac["info"].item()) == 0
        out[tag] = eng.predict(kid, th, Xd, fac, Xsd)
        del fac
        torch.cuda.empty_cache()
    for tag in ("f32_simt", "f32_tc"):
        print(f"{name} N={len(y)} M={len(Xs)} noise={wl['theta'][1]:g} {tag}: mean relinf {relinf(out[tag][0], out['f64'][0]):.2e} "
              f"sd relinf {relinf(out[tag][1], out['f64'][1]):.2e}")

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "c2", noise=float(sys.argv[2]) if len(sys.argv) > 2 else None)
