import os, sys, torch, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import models, ops
import samble_b200.blocks as blocks, samble_b200.models as mm
from samble_b200.config import seg_config
from samble_b200.testing import fill_state_dict_, synthetic_clouds
B, N = 16, 2048
cfg = seg_config(M=(N // 2, N // 4))
m = models.ShapeNetModel(cfg); m.load_state_dict(fill_state_dict_(m.state_dict(), seed=1, sharpen=4.0)); m = m.eval().cuda()
x, cat = synthetic_clouds(B, N, 2); x, cat = x.cuda(), cat.cuda()
real = ops.linear
log = []
def timed(x_, w_, **kw):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); y = real(x_, w_, **kw); e.record(); torch.cuda.synchronize()
    log.append((tuple(x_.shape), tuple(w_.shape), kw.get("x_layout", "rows"), kw.get("out_layout", "rows"), s.elapsed_time(e)))
    return y
with torch.no_grad():
    m(x, cat); models.freeze_boundaries(m); m(x, cat)
    ops.linear = timed
    m(x, cat)
tot = sum(l[-1] for l in log)
print(f"{len(log)} linear calls, {tot:.3f} ms")
for l in sorted(log, key=lambda l: -l[-1]):
    M = l[0][0] * (l[0][2] if l[2] == "bcn" else l[0][1]) if len(l[0]) == 3 else l[0][0]
    K = l[1][1]; Nn = l[1][0]
    print(f"{l[-1]*1e3:8.1f} us  x{l[0]} w{l[1][:2]} {l[2]}->{l[3]}  {2.0*M*K*Nn/l[-1]/1e9:6.1f} TF/s")
