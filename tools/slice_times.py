"""Is the rank skew of the 8-GPU run data-dependent?  On ONE GPU: graph-replay time of the seg step for each rank's slice of
the global batch bench.py builds (clouds [16r, 16r+16) of synthetic_clouds(128, 2048, seed=2)), plus the per-kernel census."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import models, _lib as L, ops
from samble_b200.config import seg_config
from samble_b200.runtime import GraphedForward
from samble_b200.testing import fill_state_dict_, synthetic_clouds
B, N, W = 16, 2048, 8
m = models.ShapeNetModel(seg_config(M=(N // 2, N // 4)))
m.load_state_dict(fill_state_dict_(m.state_dict(), seed=1, sharpen=4.0)); m = m.eval().cuda()
xh, cath = synthetic_clouds(B * W, N, seed=2)
xc, catc = synthetic_clouds(B, N, seed=1002)
with torch.no_grad():
    m(xc.cuda(), catc.cuda())
models.freeze_boundaries(m)
g = GraphedForward(m, xh[:B].cuda(), cath[:B].cuda())
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for r in range(W):
    x, c = xh[r * B:(r + 1) * B].cuda(), cath[r * B:(r + 1) * B].cuda()
    for _ in range(3): g(x, c)
    ts = []
    for _ in range(20):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g(x, c); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ops.CONCURRENT_BRANCHES = False
    L.profile(True)
    with torch.no_grad(): m(x, c)
    torch.cuda.synchronize(); rep = L.profile_report(); L.profile(False)
    ops.CONCURRENT_BRANCHES = True
    big = {k.replace("_kernel", ""): round(v[1], 3) for k, v in rep.items() if k in ("knn_feat_repair_kernel", "knn_select_kernel", "knn_tc_collect_kernel", "knn_xyz2_kernel")}
    print(f"slice {r}: {sorted(ts)[len(ts) // 2]:.3f} ms per step   {big}")
