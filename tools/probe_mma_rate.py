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
