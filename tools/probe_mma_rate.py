import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import _lib as L
lib = L.lib()
for nt in (64, 128, 256):
    for ctas in (1, 148):
        out = torch.zeros(ctas, dtype=torch.int64, device="cuda")
        iters = 2000
        L.check(lib.samble_selftest_mma_rate(nt, iters, ctas, L.ptr(out), L.stream()), "rate")
        torch.cuda.synchronize()
        cyc = out.double().mean().item() / (iters * 4)
        print(f"N={nt:3d} ctas={ctas:3d}: {cyc:7.1f} cycles per 128x{nt}x8 tf32 MMA  -> {128*nt*8*2/cyc*148*1.9e9/1e12:7.1f} TFLOP/s chip-wide at 1.9 GHz")

print("extended (bf16, kind::f16): N, accumulators, A tiles -> cycles per MMA (148 CTAs)")
for nt, combos in ((64, ((1, 1), (4, 1), (4, 4), (1, 4))), (128, ((1, 1), (4, 1), (4, 4), (1, 4))), (256, ((1, 1), (2, 1), (1, 4)))):
    for n_acc, n_a in combos:
        out = torch.zeros(148, dtype=torch.int64, device="cuda")
        iters = 2000
        L.check(lib.samble_selftest_mma_rate_ex(1, nt, n_acc, n_a, iters, 148, L.ptr(out), L.stream()), "rate_ex")
        torch.cuda.synchronize()
        cyc = out.double().mean().item() / (iters * 4)
        print(f"bf16 N={nt:3d} acc={n_acc} A={n_a}: {cyc:7.1f} cycles per 128x{nt}x16 MMA (compute {128*nt*16/4096:.0f}, smem {(128*32+nt*32)/128:.0f})")

print("wall clock vs SM cycles (bf16 N=64, 4 accumulators, 4 A tiles; 148 CTAs):")
for iters in (2000, 20000, 200000):
    out = torch.zeros(148, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    L.check(lib.samble_selftest_mma_rate_ex(1, 64, 4, 4, iters, 148, L.ptr(out), L.stream()), "rate_ex")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cyc = out.double().mean().item()
    print(f"  iters {iters}: {cyc / (iters * 4):.1f} cycles per MMA, kernel {ms:.3f} ms -> effective SM clock {cyc / (ms * 1e-3) / 1e6:.0f} MHz")
