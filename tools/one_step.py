"""Exactly one SAMBLE seg forward (B=16, N=2048, bench.py's workload) between cudaProfilerStart/Stop, after calibration
and warm-up: the ncu target for per-step launch lists and DRAM-traffic captures.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/one_step.py
  ncu --profile-from-start off --set full --clock-control none -k regex:<ours> -o R python tools/one_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import models  # noqa: E402
from samble_b200.config import seg_config  # noqa: E402
from samble_b200.testing import fill_state_dict_, synthetic_clouds  # noqa: E402

B, N = int(os.environ.get("B", 16)), int(os.environ.get("N", 2048))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
m = models.ShapeNetModel(seg_config(M=(N // 2, N // 4)))
m.load_state_dict(fill_state_dict_(m.state_dict(), seed=1, sharpen=4.0))
m = m.eval().cuda()
x, cat = synthetic_clouds(B, N, 2)
x, cat = x.cuda(), cat.cuda()
with torch.no_grad():
    m(x, cat)
    models.freeze_boundaries(m)
    for _ in range(3):
        m(x, cat)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(x, cat)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("one step done")
