"""debug: gradients w.r.t. every block OUTPUT, native vs fp64 oracle (seg eval)"""
import sys, torch
sys.path.insert(0, ".")
from tests.test_gpu_backward import _prepared
from oracle import harness, samble_oracle as O
from samble_b200 import blocks
from samble_b200.testing import synthetic_clouds
from samble_b200._precision import strict_fp32
m, sd, cfg = _prepared("seg", 2, 256, (128, 64), seed=4, train=False)
x, cat = synthetic_clouds(2, 256, 6)
outs = []
def hook(mod, args, out):
    t = out[0][0] if isinstance(out, tuple) else out
    if t.requires_grad:
        t.retain_grad(); outs.append((type(mod).__name__, t))
hs = [mod.register_forward_hook(hook) for mod in m.modules() if isinstance(mod, (blocks.EdgeConv, blocks.Neighbor2PointAttention, blocks.DownSampleToken, blocks.UpSampleInterpolation))]
log = harness._Log()
xg = x.cuda().requires_grad_(True)
with harness.record_decisions(log):
    y = m(xg, cat.cuda())
probe = torch.randn(y.shape, generator=torch.Generator().manual_seed(0))
with strict_fp32():
    (y * probe.cuda()).sum().backward()
routs = []
def wrap(fn, name, pick=lambda o: o):
    def f(*a, **k):
        o = fn(*a, **k); t = pick(o); t.retain_grad(); routs.append((name, t)); return o
    return f
O.edgeconv = wrap(O.edgeconv, "EdgeConv"); O.n2p_attention = wrap(O.n2p_attention, "N2P")
O.downsample_token = wrap(O.downsample_token, "DS", lambda o: o["x_ds"]); O.upsample_interpolation = wrap(O.upsample_interpolation, "Up")
ds_list = list(m.block.downsample_list)
states = [O.DSState(False, [t.detach().cpu().double().clone() for t in ds.bin_boundaries]) for ds in ds_list]
sdg = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running_" not in k else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
klog = [(e[0], e[1]) + ((e[2].double(),) if len(e) > 2 else ()) for e in log]
xr = x.double().requires_grad_(True)
with O.forcing(O.Forcing(knn_log=klog, ds_idx=[ds.idx.cpu() for ds in ds_list], keep_inputs=False)):
    yr = O.seg_forward(sdg, cfg, xr, cat.double(), states)
(yr * probe.double()).sum().backward()
print(len(outs), len(routs))
for (n1, t), (n2, r) in zip(outs, routs):
    g, gr = t.grad.detach().cpu().double(), r.grad
    print(f"{n1:28s} {n2:10s} shape {tuple(r.shape)} fwd {float((t.detach().cpu().double()-r.detach()).abs().max()/r.detach().abs().max()):.1e} grad rel {float((g-gr).abs().max()/gr.abs().max()):.1e}")
# ---- focus on upsample 0
pre = "block.upsample_list.0."
up = m.block.upsample_list[0]
for k, p in up.named_parameters():
    r = sdg[pre + k].grad
    print(k, "rel", float((p.grad.cpu().double() - r).abs().max() / r.abs().max()), "native max", float(p.grad.abs().max()), "oracle max", float(r.abs().max()))
t, r = outs[7][1], routs[7][1]
g, gr = t.grad.cpu().double(), r.grad
print("g_out: max-rel", float((g - gr).abs().max() / gr.abs().max()), "L2-rel", float((g - gr).norm() / gr.norm()), "sum native", float(g.sum()), "oracle", float(gr.sum()))
print("g_out per-channel sum rel", float((g.sum((0, 2)) - gr.sum((0, 2))).abs().max() / gr.sum((0, 2)).abs().max()))
a = dict(up.named_parameters())["res_conv.1.bias"].grad.cpu().double(); b = sdg[pre + "res_conv.1.bias"].grad
d = (a - b).abs(); top = d.topk(6)[1]
print("worst channels", top.tolist(), "native", a[top].tolist(), "oracle", b[top].tolist())
print("num channels off by >1e-3 rel:", int((d > 1e-3 * b.abs().max()).sum()), "of", d.numel())
