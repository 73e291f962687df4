"""Which phase paces edge_mlp_tc_kernel?  samble_set_edge_debug: 1 = no gathers, 8 = no stage build, 2 = no MMAs, 4 = idle epilogue."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import ops, _lib as L
lib = L.lib()
for B, N, C1, C2 in ((16, 2048, 64, 64), (16, 2048, 64, 128)):
    g = torch.Generator().manual_seed(0)
    pr = torch.randn(B, N, 2 * C1, generator=g).cuda(); w2 = (torch.randn(C2, C1, generator=g) / 8).cuda(); b2 = torch.randn(C2, generator=g).cuda()
    idx = torch.randint(0, N, (B, N, 32), generator=g, dtype=torch.int32).cuda()
    row = {}
    for bits in (0, 1, 9, 2, 4, 6, 15):
        lib.samble_set_edge_debug(bits)
        for _ in range(2): ops.edge_mlp_max(pr, idx, w2, b2)
        torch.cuda.synchronize(); L.profile(True)
        for _ in range(10): ops.edge_mlp_max(pr, idx, w2, b2)
        torch.cuda.synchronize(); n, ms = L.profile_report()["edge_mlp_tc_kernel"]; L.profile(False)
        row[bits] = ms / n * 1e3
    lib.samble_set_edge_debug(0)
    print(f"C1={C1} C2={C2}: full {row[0]:.1f} us | no gathers {row[1]:.1f} | no gathers+build {row[9]:.1f} | no MMA {row[2]:.1f} | no epilogue {row[4]:.1f} | no MMA+epilogue {row[6]:.1f} | nothing {row[15]:.1f}")
