"""debug: upsample_list.0 in isolation on the tensors it sees inside the seg model"""
import sys, torch
sys.path.insert(0, ".")
from tests.test_gpu_backward import _prepared
from oracle import harness, samble_oracle as O
from samble_b200.testing import synthetic_clouds
from samble_b200._precision import strict_fp32
m, sd, cfg = _prepared("seg", 2, 256, (128, 64), seed=4, train=False)
x, cat = synthetic_clouds(2, 256, 6)
up = m.block.upsample_list[0]; pre = "block.upsample_list.0."
cap = {}
def prehook(mod, args): cap["args"] = args
h = up.register_forward_pre_hook(prehook)
with torch.no_grad():
    m(x.cuda(), cat.cuda())
h.remove()
pcd_up, ((sel, idx_sel, xyz_sel), _), xyz_up = cap["args"]
tens = [t.detach().clone().contiguous() for t in (pcd_up, sel, xyz_up, xyz_sel)]
def rel(g, r): return float((g.detach().cpu().double() - r.double()).abs().max() / r.double().abs().max())
probe = torch.randn(tens[0].shape, generator=torch.Generator().manual_seed(5))
log = harness._Log()
ins = [t.clone().requires_grad_(True) for t in tens]
up.zero_grad()
with harness.record_decisions(log):
    y = up(ins[0], ((ins[1], None, ins[3]), (None, None)), ins[2])
with strict_fp32():
    (y * probe.cuda()).sum().backward()
d = log[0][2]
print("3-NN distances: min", float(d.min()), "zeros", int((d == 0).sum()), "tiny (<1e-2)", int((d < 1e-2).sum()), "of", d.numel())
for dtype in (torch.float64, torch.float32):
    sdg = {k: v.to(dtype).requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items() if k.startswith(pre)}
    rin = [t.cpu().to(dtype).requires_grad_(True) for t in tens]
    klog = [(e[0], e[1]) + ((e[2].to(dtype),) if len(e) > 2 else ()) for e in log]
    with O.forcing(O.Forcing(knn_log=klog, keep_inputs=False)) as f:
        yr = O.upsample_interpolation(sdg, pre, rin[0], rin[1], rin[2], rin[3], 3)
    (yr * probe.to(dtype)).sum().backward()
    print(dtype, "fwd", rel(y, yr.detach()))
    for nm, a, b in zip(("pcd_up", "select", "xyz_up", "xyz_sel"), ins, rin):
        print("  d/d", nm, rel(a.grad, b.grad), "scale", float(b.grad.abs().max()))
    for k, v in sdg.items():
        if v.requires_grad and v.grad is not None:
            print("  ", k, rel(dict(up.named_parameters())[k[len(pre):]].grad, v.grad))
