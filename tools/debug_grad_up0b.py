"""debug: in-model vs isolated backward of upsample_list.0, native and oracle, same output gradient"""
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
def prehook(mod, args):
    cap["args"] = args
    for t in (args[0], args[1][0][0], args[1][0][2], args[2]):
        if t.requires_grad: t.retain_grad()
def posthook(mod, args, out): out.retain_grad(); cap["out"] = out
h1, h2 = up.register_forward_pre_hook(prehook), up.register_forward_hook(posthook)
log = harness._Log()
xg = x.cuda().requires_grad_(True)
with harness.record_decisions(log):
    y = m(xg, cat.cuda())
probe = torch.randn(y.shape, generator=torch.Generator().manual_seed(0))
with strict_fp32():
    (y * probe.cuda()).sum().backward()
h1.remove(); h2.remove()
pcd_up, ((sel, idx_sel, xyz_sel), _), xyz_up = cap["args"]
g_out = cap["out"].grad.clone()
in_model = [t.grad.clone() if t.grad is not None else None for t in (pcd_up, sel, xyz_up, xyz_sel)]
in_model_params = {k: p.grad.clone() for k, p in up.named_parameters()}
def rel(g, r): return float((g.detach().cpu().double() - r.detach().cpu().double()).abs().max() / r.detach().cpu().double().abs().max())
ins = [t.detach().clone().requires_grad_(True) for t in (pcd_up, sel, xyz_up, xyz_sel)]
up.zero_grad()
yy = up(ins[0], ((ins[1], None, ins[3]), (None, None)), ins[2])
with strict_fp32():
    (yy * g_out).sum().backward()
print("isolated native vs in-model native (same g_out); note pcd_up/xyz also receive gradient from other consumers in the model")
for nm, a, b in zip(("pcd_up", "select", "xyz_up", "xyz_sel"), ins, in_model):
    print("  d/d", nm, rel(a.grad, b), "strides", tuple(b.stride()), [tuple(t.stride()) for t in (pcd_up, sel, xyz_up, xyz_sel)][("pcd_up", "select", "xyz_up", "xyz_sel").index(nm)])
for k, p in up.named_parameters():
    print("  ", k, rel(p.grad, in_model_params[k]))
