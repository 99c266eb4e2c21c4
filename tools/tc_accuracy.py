"""Accuracy probe of the split-fp16 tcgen05 GEMM: signed error statistics vs fp64."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpim_b200._lib import get_engine
eng = get_engine()
torch.manual_seed(0)
for K in (64, 256, 1024, 4096, 16384):
    for mode in ("pos_fp16exact", "pos", "randn"):
        M, N = 256, 256
        if mode == "randn":
            A = torch.randn(M, K, dtype=torch.float64); B = torch.randn(N, K, dtype=torch.float64)
        else:
            A = torch.rand(M, K, dtype=torch.float64) * 0.5 + 0.5; B = torch.rand(N, K, dtype=torch.float64) * 0.5 + 0.5
        if mode == "pos_fp16exact":
            A = A.half().double(); B = B.half().double()
        A32, B32 = A.float(), B.float()
        exact = A32.double() @ B32.double().T
        tcr = eng.gemm_nt(A32.cuda(), B32.cuda()).cpu().double()
        f32 = (A32.cuda() @ B32.cuda().T).cpu().double()      # cuBLAS fp32 (SIMT or TF32 off by default)
        scale = (A32.double().abs() @ B32.double().abs().T)
        e_tc = ((tcr - exact) / scale); e_f32 = ((f32 - exact) / scale)
        print(f"K={K:6d} {mode:14s} tc: mean {e_tc.mean().item():+.2e} rms {e_tc.pow(2).mean().sqrt().item():.2e} max {e_tc.abs().max().item():.2e} | "
              f"fp32: mean {e_f32.mean().item():+.2e} rms {e_f32.pow(2).mean().sqrt().item():.2e} max {e_f32.abs().max().item():.2e}")
