"""Does running the batch as S independent sub-batches on S streams (one CUDA graph, S concurrent branches) beat one pass over
the whole batch?  Every op on the path is per-cloud, so the split is exact; small-grid kernels and partial last waves of one
half would be filled by the other's work.  Prints ms per step of graph replays (L2 flushed between steps)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samble_b200 import models
from samble_b200.config import seg_config
from samble_b200.testing import fill_state_dict_, synthetic_clouds

B, N = 16, 2048
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
m = models.ShapeNetModel(seg_config(M=(N // 2, N // 4)))
m.load_state_dict(fill_state_dict_(m.state_dict(), seed=1, sharpen=4.0))
m = m.eval().cuda()
x, cat = synthetic_clouds(B, N, 2)
x, cat = x.cuda(), cat.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    m(x, cat)
    models.freeze_boundaries(m)
    ref = m(x, cat)


def build(S):
    streams = [torch.cuda.Stream() for _ in range(S)]
    xs, cs = x.chunk(S), cat.chunk(S)

    def step():
        main = torch.cuda.current_stream()
        outs = []
        for st, xi, ci in zip(streams, xs, cs):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                outs.append(m(xi, ci))
        for st in streams:
            main.wait_stream(st)
        return torch.cat(outs) if S > 1 else outs[0]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        out = step()
    return g, out


for S in (1, 2, 4):
    g, out = build(S)
    for _ in range(5):
        g.replay()
    ts = []
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    same = torch.equal(out, ref)
    print(f"{S} sub-batch(es) of {B // S} clouds: {ts[len(ts) // 2]:.3f} ms per step ({B / ts[len(ts) // 2] * 1e3:.0f} clouds/s), logits identical to the single pass: {same}")
