"""debug (CPU only): is the ORACLE's in-model backward of upsample 0 consistent with its isolated backward?"""
import sys, torch
sys.path.insert(0, ".")
from oracle import samble_oracle as O
from samble_b200 import models
from samble_b200.config import seg_config
from samble_b200.testing import fill_state_dict_, synthetic_clouds
cfg = seg_config(M=(128, 64))
m = models.ShapeNetModel(cfg)
sd = fill_state_dict_(m.state_dict(), seed=4, sharpen=2.0)
x, cat = synthetic_clouds(2, 256, 6)
dt = torch.float64
states = [O.DSState(True) for _ in range(2)]
with torch.no_grad():
    O.seg_forward({k: v.to(dt) if v.is_floating_point() else v for k, v in sd.items()}, cfg, synthetic_clouds(2, 256, 104)[0].to(dt), cat.to(dt), states)
for s in states: s.dynamic = False
sdg = {k: (v.to(dt).requires_grad_(True) if v.is_floating_point() and "running_" not in k else (v.to(dt) if v.is_floating_point() else v)) for k, v in sd.items()}
cap = {}
real = O.upsample_interpolation
def wrapped(sd_, pre, pcd_up, selected, xyz_up, xyz_sel, K=3):
    if pre.endswith("upsample_list.0."):
        for t in (pcd_up, selected, xyz_up, xyz_sel): t.retain_grad()
        cap["in"] = (pcd_up, selected, xyz_up, xyz_sel)
    out = real(sd_, pre, pcd_up, selected, xyz_up, xyz_sel, K)
    if pre.endswith("upsample_list.0."):
        out.retain_grad(); cap["out"] = out
    return out
O.upsample_interpolation = wrapped
xr = x.to(dt).requires_grad_(True)
y = O.seg_forward(sdg, cfg, xr, cat.to(dt), states)
probe = torch.randn(y.shape, generator=torch.Generator().manual_seed(0)).to(dt)
(y * probe).sum().backward()
g_out = cap["out"].grad.clone()
pre = "block.upsample_list.0."
in_model_params = {k: v.grad.clone() for k, v in sdg.items() if k.startswith(pre) and isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
in_model = [t.grad.clone() for t in cap["in"]]
O.upsample_interpolation = real
sd2 = {k: (v.detach().clone().requires_grad_(True) if v.requires_grad else v) for k, v in sdg.items() if k.startswith(pre)}
ins = [t.detach().clone().requires_grad_(True) for t in cap["in"]]
yy = real(sd2, pre, *ins, 3)
(yy * g_out).sum().backward()
def rel(g, r): return float((g - r).abs().max() / r.abs().max())
for nm, a, b in zip(("pcd_up", "select", "xyz_up", "xyz_sel"), ins, in_model):
    print("  d/d", nm, rel(a.grad, b))
for k, v in sd2.items():
    if k in in_model_params: print("  ", k, rel(v.grad, in_model_params[k]))
