"""Per-kernel device time inside the DownSampleToken blocks of one seg forward (B=16, N=2048)."""
import os, sys, torch, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import models, _lib as L
from samble_b200.config import seg_config
from samble_b200.testing import fill_state_dict_, synthetic_clouds
from torch.profiler import profile, ProfilerActivity
B, N = 16, 2048
m = models.ShapeNetModel(seg_config(M=(N // 2, N // 4))); m.load_state_dict(fill_state_dict_(m.state_dict(), seed=1, sharpen=4.0)); m = m.eval().cuda()
x, cat = synthetic_clouds(B, N, 2); x, cat = x.cuda(), cat.cuda()
with torch.no_grad():
    m(x, cat); models.freeze_boundaries(m); m(x, cat)
    feats = {}
    for i, ds in enumerate(m.block.downsample_list):
        ds.register_forward_pre_hook(lambda mod, args, i=i: feats.__setitem__(i, args))
    m(x, cat)
    for i, ds in enumerate(m.block.downsample_list):
        ds._forward_pre_hooks.clear()
        args = feats[i]
        for _ in range(2): ds(*args)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3): ds(*args)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
        tot = sum(e.self_device_time_total for e in rows) / 3
        print(f"== DownSampleToken[{i}]  total {tot:.0f} us")
        for e in rows[:22]:
            print(f"   {e.self_device_time_total / 3:8.1f} us  x{e.count // 3:<3d} {e.key[:90]}")
