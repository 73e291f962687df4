"""Which phase paces linear_tma_kernel?  Times the step's main shapes with phases switched off (samble_set_linear_debug):
1 = no operand split, 2 = no epilogue, 4 = no MMAs, 8 / 16 = no residual loads / no stores (direct epilogue), 64 = 96-wide tiles + TMA stores.  Device time from the library's per-launch event pairs (samble_profile_enable)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import ops, _lib as L
torch.manual_seed(0)
shapes = [(32768, 128, 384, False), (32768, 128, 512, True), (32768, 512, 128, False), (32768, 128, 1024, True), (32768, 1024, 256, True), (16384, 128, 384, False), (8192, 512, 128, False)]
lib = L.lib()
for M, K, Nout, act in shapes:
    x = torch.randn(16, M // 16, K, device="cuda"); w = torch.randn(Nout, K, device="cuda") / K ** 0.5
    sc, sh = torch.rand(Nout, device="cuda") + 0.5, torch.randn(Nout, device="cuda")
    res = torch.randn(16, M // 16, Nout, device="cuda") if not act else None
    row = []
    for bits in (0, 2, 64, 66):
        lib.samble_set_linear_debug(bits)
        for _ in range(3): ops.linear(x, w, scale=sc, shift=sh, lrelu=act, residual=res)
        torch.cuda.synchronize()
        L.profile(True)                      # the library's own event pair around every launch: device time, no host overhead
        for _ in range(20): ops.linear(x, w, scale=sc, shift=sh, lrelu=act, residual=res)
        torch.cuda.synchronize()
        n, ms = L.profile_report()["linear_tma_kernel"]
        L.profile(False)
        row.append(ms / n * 1e3)
    lib.samble_set_linear_debug(0)
    mma_us = 3 * 2.0 * M * K * Nout / 1151e12 * 1e6
    print(f"M={M} K={K} Nout={Nout}: default {row[0]:.1f} us, no epilogue {row[1]:.1f} | 96-wide + TMA stores for resident weights {row[2]:.1f}, no epilogue {row[3]:.1f} | MMA-issue bound {mma_us:.1f}")
