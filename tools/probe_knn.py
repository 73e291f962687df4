"""Which phase paces knn_tc_kernel (threshold / collect passes)?  samble_set_knn_debug: 1 = no MMAs, 2 = idle epilogue.
Device time from the library's per-launch event pairs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import ops, _lib as L
from samble_b200.testing import synthetic_features
lib = L.lib()
for B, N, C in ((16, 2048, 128), (16, 2048, 64), (16, 1024, 128), (16, 512, 128), (8, 8192, 128)):
    x = synthetic_features(B, C, N, 1).cuda()
    out = {}
    for cs in (1, 2, 4):
        lib.samble_set_knn_debug(cs << 8)
        for _ in range(2): ops.knn_indices(x, 32, ordered=False)
        torch.cuda.synchronize(); L.profile(True)
        for _ in range(10): ops.knn_indices(x, 32, ordered=False)
        torch.cuda.synchronize(); rep = L.profile_report(); L.profile(False)
        lib.samble_set_knn_debug(0)
        print(f"   cluster of {cs}: threshold {rep['knn_tc_threshold_kernel'][1] / rep['knn_tc_threshold_kernel'][0] * 1e3:.1f} us, collect {rep['knn_tc_collect_kernel'][1] / rep['knn_tc_collect_kernel'][0] * 1e3:.1f} us")
    for bits in (0, 1, 2, 3):
        lib.samble_set_knn_debug(bits)
        try:
            for _ in range(2): ops.knn_indices(x, 32, ordered=False)
            torch.cuda.synchronize(); L.profile(True)
            for _ in range(10): ops.knn_indices(x, 32, ordered=False)
            torch.cuda.synchronize(); rep = L.profile_report(); L.profile(False)
        finally:
            lib.samble_set_knn_debug(0)
        out[bits] = {k: v[1] / v[0] * 1e3 for k, v in rep.items()}
    names = ("knn_tc_threshold_kernel", "knn_tc_collect_kernel")
    print(f"B={B} N={N} C={C}: " + " | ".join(f"{n.split('_')[2]}: full {out[0][n]:.1f} us, no MMA {out[1][n]:.1f}, no epilogue {out[2][n]:.1f}, loads only {out[3][n]:.1f}" for n in names)
          + " | others: " + ", ".join(f"{k.replace('_kernel','')} {v:.1f}" for k, v in out[0].items() if k not in names))
