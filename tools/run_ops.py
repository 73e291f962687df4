"""Run the hot-path ops in isolation at the BASELINE seg shapes (ncu target / micro-bench)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import _lib as L
from samble_b200 import ops  # noqa: E402
from samble_b200.testing import synthetic_features  # noqa: E402

B, N = int(os.environ.get("B", 16)), int(os.environ.get("N", 2048))
which = sys.argv[1:] or ["xyz", "feat128", "feat64", "rowstats"]
dev = torch.device("cuda:0")
x3 = synthetic_features(B, 3, N, 1).to(dev)
x64 = synthetic_features(B, 64, N, 2).to(dev)
x128 = synthetic_features(B, 128, N, 3).to(dev)
q = torch.randn(B, N, 128, device=dev)
k = torch.randn(B, N, 128, device=dev)
ktok = torch.randn(4, 128, device=dev)


def bench(name, fn, flops=None, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    extra = f"  {flops / ms / 1e9:.1f} TFLOP/s" if flops else ""
    print(f"{name:12s} {ms * 1e3:9.1f} us{extra}   pairs/s {B * N * N / ms / 1e6:.1f} G", flush=True)


L.profile(os.environ.get("PROFILE", "0") == "1")
if "xyz" in which:
    bench("knn xyz", lambda: ops.knn_indices(x3, 32))
if "feat64" in which:
    bench("knn C=64", lambda: ops.knn_indices(x64, 32), 2.0 * B * N * N * 64)
if "feat128" in which:
    bench("knn C=128", lambda: ops.knn_indices(x128, 32), 2.0 * B * N * N * 128)
if "rowstats" in which:
    bench("ds rowstats", lambda: ops.ds_row_stats(q, k, ktok), 2.0 * B * N * N * 128)
if os.environ.get("PROFILE", "0") == "1":
    for kname, (n, ms) in sorted(L.profile_report().items(), key=lambda kv: -kv[1][1]):
        print(f"  {kname:28s} {n:5d} launches  {ms / n * 1e3:9.1f} us avg")

if "linear" in which:
    # GPU-side kernel time from the library's own event pairs (host launch overhead ~30 us would hide short kernels)
    shapes = [(B * N, 128, 384), (B * N, 128, 512), (B * N, 512, 128), (B * N, 128, 1024), (B * N, 1024, 256),
              (B * N // 2, 128, 384), (B * N, 64, 128), (B * N, 256, 128), (B * N // 4, 512, 128)]
    for (M, K, Nn) in shapes:
        xx = torch.randn(M, K, device=dev)
        ww = torch.randn(Nn, K, device=dev)
        sc, sh = torch.rand(Nn, device=dev) + 0.5, torch.randn(Nn, device=dev)
        for _ in range(3):
            ops.linear(xx, ww, scale=sc, shift=sh, lrelu=True)
        torch.cuda.synchronize()
        L.profile(True)
        for _ in range(10):
            ops.linear(xx, ww, scale=sc, shift=sh, lrelu=True)
        torch.cuda.synchronize()
        rep = L.profile_report()
        L.profile(False)
        n, ms = [v for k, v in rep.items() if k.startswith("linear")][0]
        us = ms / n * 1e3
        mma_us = 3 * 2.0 * M * K * Nn / 1151e12 * 1e6
        print(f"lin {K:4d}->{Nn:4d} M={M:6d}  {us:7.1f} us   {2.0 * M * K * Nn / us / 1e6:6.1f} TF/s(fp32-equiv)   tensor-pipe floor {mma_us:5.1f} us"
              f"   out {M * Nn * 4 / 1e6:5.0f} MB", flush=True)
