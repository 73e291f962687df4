"""samble_linear_pool: swapped orientation (weight tile as the A operand, reduction over the points down each thread's own
accumulator columns; default) against the row-per-thread butterfly epilogue (samble_set_linear_debug(128)), and without any
epilogue (bit 2).  Device time per launch from the library's own event pairs, back-to-back launches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import ops, _lib as L
torch.manual_seed(0)
lib = L.lib()


def timed(fn, name="linear_tma_kernel"):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    L.profile(True)
    for _ in range(20): fn()
    torch.cuda.synchronize()
    n, ms = L.profile_report()[name]
    L.profile(False)
    return ms / n * 1e3


for M, K, Nout in [(32768, 128, 1024), (16384, 128, 1024), (32768, 64, 1024)]:
    x = torch.randn(16, M // 16, K, device="cuda"); w = torch.randn(Nout, K, device="cuda") / K ** 0.5
    sc, sh = torch.rand(Nout, device="cuda") + 0.5, torch.randn(Nout, device="cuda")
    row = []
    for bits in (0, 128, 2):
        lib.samble_set_linear_debug(bits)
        row.append(timed(lambda: ops.linear_pool(x, w, scale=sc, shift=sh, lrelu=True)))
    lib.samble_set_linear_debug(0)
    print(f"pooled M={M} K={K} Nout={Nout}: swapped {row[0]:.1f} us | row-per-thread butterfly {row[1]:.1f} | no epilogue {row[2]:.1f} | MMA-issue bound {3 * 2.0 * M * K * Nout / 1151e12 * 1e6:.1f}")
