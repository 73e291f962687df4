"""Device time of the exact-product GEMM kernels at the DownSampleToken shapes of the seg step (B=16)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import _lib as L, ops
B, D = 16, 128
for N in (2048, 1024):
    g = torch.Generator().manual_seed(0)
    q = (torch.randn(B, N, D, generator=g) * 4).cuda(); k = torch.randn(B, N, D, generator=g).cuda(); kt = torch.randn(4, D, generator=g).cuda()
    w = torch.randn(384, D, generator=g).cuda()
    qd, kd, wd = ops.digits(q), ops.digits(k), ops.weight_digits(w)
    for _ in range(3):
        ops.ds_row_stats_exact(qd, kd, q, kt); ops.xgemm(qd, wd)
    torch.cuda.synchronize()
    L.profile(True)
    for _ in range(10):
        ops.ds_row_stats_exact(qd, kd, q, kt); ops.xgemm(qd, wd)
    rep = L.profile_report(); L.profile(False)
    items = B * (N // 128) * (N // 64)
    us = rep["xgemm_rowstat_kernel"][1] / 10 * 1e3
    print(f"N={N}: " + ", ".join(f"{k2} {v[1]/v[0]*1e3:.1f} us" for k2, v in rep.items()) +
          f"; row statistics: {us * 1e-6 * 1.965e9 * 148 / items / 80:.1f} cycles per MMA per SM (48 = shared-memory bound)")
