"""Where does one seg forward spend its device time?  torch.profiler table (all kernels, ours + library)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import models  # noqa: E402
from samble_b200.config import seg_config  # noqa: E402
from samble_b200.testing import fill_state_dict_, synthetic_clouds  # noqa: E402

B, N = int(os.environ.get("B", 16)), int(os.environ.get("N", 2048))
dev = torch.device("cuda:0")
cfg = seg_config(M=(N // 2, N // 4))
m = models.ShapeNetModel(cfg)
m.load_state_dict(fill_state_dict_(m.state_dict(), seed=1, sharpen=4.0))
m = m.eval().to(dev)
x, cat = synthetic_clouds(B, N, 2)
x, cat = x.to(dev), cat.to(dev)
with torch.no_grad():
    m(x, cat)
    models.freeze_boundaries(m)
    for _ in range(3):
        m(x, cat)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            m(x, cat)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
