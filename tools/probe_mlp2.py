"""Fused two-layer MLP (csrc/mlp2.cu) vs the two linear_tma launches it replaces: device time per call and phase ablation
(samble_set_mlp2_debug: 1 = no conversion, 2 = no output epilogue, 4 = no MMAs; results are garbage while a bit is set)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from samble_b200 import ops, _lib as L

lib = L.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for (B, P, Hd, N2) in [(16, 2048, 512, 128), (16, 1024, 512, 128), (16, 512, 512, 128), (16, 2048, 1024, 256)]:
    x = torch.randn(B, P, 128, device="cuda")
    w1, w2 = torch.randn(Hd, 128, device="cuda") / 11, torch.randn(N2, Hd, device="cuda") / Hd ** 0.5
    # the feed-forward has no scale / shift between its layers (attention.py:187-192); the head has both (folded BatchNorm, per-cloud shift)
    s1, h1 = (torch.rand(Hd, device="cuda") + 0.5, torch.randn(B, Hd, device="cuda") * 0.1) if N2 == 256 else (None, None)
    s2, h2 = torch.rand(N2, device="cuda") + 0.5, torch.randn(N2, device="cuda") * 0.1
    res = torch.randn(B, P, N2, device="cuda") if N2 == 128 else None
    two = lambda: ops.linear(ops.linear(x, w1, scale=s1, shift=h1, lrelu=True), w2, scale=s2, shift=h2, residual=res, residual_first=True, lrelu=N2 == 256)
    one = lambda: ops.mlp2(x, w1, w2, scale1=s1, shift1=h1, scale2=s2, shift2=h2, residual=res, residual_first=True, lrelu2=N2 == 256)
    t2, t1 = timed(two), timed(one)
    parts = []
    wc = torch.zeros(148 * 5, dtype=torch.int64, device="cuda")
    for bits, name in ((0, "full"), (1, "no conversion"), (8, "no scale/shift"), (2, "no output epilogue"), (4, "no MMA"), (7, "weights stream only")):
        lib.samble_set_mlp2_debug(bits)
        t = timed(one)
        lib.samble_set_mlp2_probe(L.ptr(wc))
        one(); torch.cuda.synchronize()
        lib.samble_set_mlp2_probe(None)
        w = wc.view(148, 5)[: min(148, B * P // 128)].double()
        b_ = (w[w[:, 0].argmax()] / 1e3).tolist()
        parts.append("\n    %-22s %6.1f us | MMA thread k-cycles: total %6.1f, waits: weights %5.1f, conversion %5.1f, drain %5.1f, X %4.1f" % (name, t, *b_))
    lib.samble_set_mlp2_debug(0)
    mtiles = B * P // 128
    rounds = -(-mtiles // 148)
    mma = rounds * (Hd // 128) * (48 + 48 * N2 // 128) * 64 / 1.965e3
    print(f"M={B * P} 128->{Hd}->{N2}: two launches {t2:.1f} us, fused {t1:.1f} us, MMA-issue bound {mma:.1f}" + "".join(parts))
