"""Where do the warp roles of knn_tc_kernel wait?  samble_set_knn_probe: per CTA cycle counters of the MMA-issuing thread (total, wait for
operand stages, wait for the epilogue to drain an accumulator, wait for the query tile), the epilogue (wait for accumulators) and the TMA producer
(total, wait for free stages).  The two passes of one ops.knn call write the same buffer, so each pass is probed in its own run (debug bit 32 / 64
would be nicer; here: the collect pass overwrites the threshold pass, and the threshold pass is read by launching with k such that ... no:
we simply read after each launch through the profiling hook order: threshold first, collect second -> the buffer holds the COLLECT pass)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import ops, _lib as L
from samble_b200.testing import synthetic_features
lib = L.lib()
import itertools
for (B, N, C), bits in itertools.product(((16, 2048, 128), (16, 2048, 64)), (0, 1, 2, 3)):
    x = synthetic_features(B, C, N, 1).cuda()
    for _ in range(2): ops.knn_indices(x, 32, ordered=False)
    ctas = B * (N // 128)
    buf = torch.zeros(ctas * 8, dtype=torch.int64, device="cuda")
    lib.samble_set_knn_debug(bits)
    lib.samble_set_knn_probe(L.ptr(buf))
    ops.knn_indices(x, 32, ordered=False)
    torch.cuda.synchronize()
    lib.samble_set_knn_probe(None)
    lib.samble_set_knn_debug(0)
    w = buf.view(ctas, 8).double() / 1e3
    tiles = N // 128
    print(f"B={B} N={N} C={C} [{('full', 'no MMAs', 'idle epilogue', 'no MMAs, idle epilogue')[bits]}]: collect pass, {ctas} CTAs, {tiles} candidate tiles each; k cycles, mean over CTAs (max)")
    for name, col in (("MMA thread total", 0), ("  wait operand stages", 1), ("  wait accumulator drain (epilogue)", 2), ("  wait query tile", 3), ("  descriptors + MMA issue (rest: commits, loop)", 7),
                      ("epilogue: wait for accumulators", 4), ("producer total", 5), ("  wait free stages", 6)):
        print(f"   {name:38s} {w[:, col].mean().item():7.1f} ({w[:, col].max().item():7.1f})")
    nkt = (C + 63) // 64
    print(f"   MMA floor {tiles * (3 * 4 * nkt + 1) * 66 / 1e3:.1f} k cycles, stream floor at 40 B/clk {tiles * (2 * nkt * 16384 + 4096) / 40 / 1e3:.1f} k cycles")
