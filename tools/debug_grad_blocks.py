"""debug: block-level gradients (UpSampleInterpolation, N2P) vs the oracle in fp64 with forced neighbours"""
import sys, torch
sys.path.insert(0, ".")
from oracle import harness, samble_oracle as O
from samble_b200 import models
from samble_b200.config import seg_config
from samble_b200.testing import fill_state_dict_, synthetic_clouds, synthetic_features
from samble_b200._precision import strict_fp32
DEV = "cuda:0"
cfg = seg_config(M=(128, 64))
m = models.ShapeNetModel(cfg)
sd = fill_state_dict_(m.state_dict(), seed=4, sharpen=2.0)
m.load_state_dict(sd); m = m.eval().to(DEV)
def rel(g, r): return float((g.detach().cpu().double() - r.double()).abs().max() / r.double().abs().max())
B, N, M, C = 2, 128, 64, 128
xyz, _ = synthetic_clouds(B, N, 3)
sel = torch.stack([torch.randperm(N, generator=torch.Generator().manual_seed(b))[:M] for b in range(B)])
xyz_s = torch.gather(xyz, 2, sel.unsqueeze(1).expand(-1, 3, -1))
up_feat, sel_feat = synthetic_features(B, C, N, 1), synthetic_features(B, C, M, 2)
probe = torch.randn(B, C, N, generator=torch.Generator().manual_seed(5))
for name, pre, mod in (("up0", "block.upsample_list.0.", m.block.upsample_list[0]),):
    log = harness._Log()
    ins = [t.to(DEV).requires_grad_(True) for t in (up_feat, sel_feat, xyz, xyz_s)]
    mod.zero_grad()
    with harness.record_decisions(log):
        y = mod(ins[0], ((ins[1], None, ins[3]), (None, None)), ins[2])
    with strict_fp32():
        (y * probe.to(DEV)).sum().backward()
    sdg = {k: v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v) for k, v in sd.items() if k.startswith(pre)}
    rin = [t.double().requires_grad_(True) for t in (up_feat, sel_feat, xyz, xyz_s)]
    klog = [(e[0], e[1]) + ((e[2].double(),) if len(e) > 2 else ()) for e in log]
    with O.forcing(O.Forcing(knn_log=klog, keep_inputs=False)) as f:
        yr = O.upsample_interpolation(sdg, pre, rin[0], rin[1], rin[2], rin[3], 3)
    (yr * probe.double()).sum().backward()
    print(name, "fwd", rel(y, yr.detach()), "unforced", len(f.knn_log))
    for nm, a, b in zip(("pcd_up", "select", "xyz_up", "xyz_sel"), ins, rin):
        print("  d/d", nm, rel(a.grad, b.grad))
    for k, v in sdg.items():
        if v.requires_grad and v.grad is not None:
            print("  ", k, rel(dict(mod.named_parameters())[k[len(pre):]].grad, v.grad))
# N2P at N=128 from a contiguous input
n2p = m.block.feature_learning_layer_list[3]; pre = "block.feature_learning_layer_list.3."
x = synthetic_features(B, C, N, 9)
log = harness._Log(); xg = x.to(DEV).requires_grad_(True); n2p.zero_grad()
with harness.record_decisions(log):
    y = n2p(xg)
with strict_fp32():
    (y * probe.to(DEV)).sum().backward()
sdg = {k: v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v) for k, v in sd.items() if k.startswith(pre)}
xr = x.double().requires_grad_(True)
with O.forcing(O.Forcing(knn_log=list(log), keep_inputs=False)) as f:
    yr = O.n2p_attention(sdg, pre, xr, 32, 4)
(yr * probe.double()).sum().backward()
print("n2p fwd", rel(y, yr.detach()), "d/dx", rel(xg.grad, xr.grad))
for k, v in sdg.items():
    if v.requires_grad and v.grad is not None:
        print("  ", k, rel(dict(n2p.named_parameters())[k[len(pre):]].grad, v.grad))
