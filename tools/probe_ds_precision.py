"""How much do the 3xTF32 kernels move the DownSample point score / sampled indices relative to the CPU oracle?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import samble_oracle as O
from samble_b200 import _lib as L, models, ops
from samble_b200.config import seg_config
from samble_b200.testing import fill_state_dict_, synthetic_features
import samble_b200.blocks as blocks

N, B = 2048, 2
for sharpen in (1.0, 4.0):
    cfg = seg_config(M=(N // 2, N // 4))
    m = models.ShapeNetModel(cfg)
    sd = fill_state_dict_(m.state_dict(), seed=9, sharpen=sharpen)
    m.load_state_dict(sd); m = m.eval().cuda()
    x = synthetic_features(B, 128, N, 73)
    ref = O.downsample_token(sd, "block.downsample_list.0.", x, N // 2, 32, 4, O.DSState(True))
    ds = m.block.downsample_list[0]
    real_linear = ops.linear
    def torch_linear(x_, w_, x_layout="rows", **kw):
        assert x_layout == "bcn"
        return torch.matmul(x_.transpose(1, 2), w_.flatten(1).t())
    for lin_name, lin in (("cublas-fp32", torch_linear), ("3xTF32", real_linear)):
        for rs_mode, rs_name in ((1, "ffma"), (0, "3xTF32")):
            blocks.ops.linear = lin
            L.lib().samble_set_ds_mode(rs_mode)
            ds.bin_boundaries = None; ds.dynamic_boundaries_enable = True
            with torch.no_grad():
                (x_ds, idx), _ = ds(x.cuda())
            rel = ((ds.attention_point_score.cpu().double() - ref["score"].double()).abs() / ref["score"].double().abs().clamp_min(1e-300))
            ov = sum(len(set(idx[b, 0].tolist()) & set(ref["idx"][b, 0].tolist())) for b in range(B)) / (B * N // 2)
            print(f"sharpen {sharpen}: qkv {lin_name:12s} rowstats {rs_name:7s}: score rel err median {rel.median():.1e} p99 {rel.flatten().kthvalue(int(0.99*rel.numel()))[0]:.1e} max {rel.max():.1e}; idx overlap {ov:.4f} exact {float((idx.cpu()==ref['idx']).float().mean()):.4f}")
    blocks.ops.linear = real_linear
    L.lib().samble_set_ds_mode(0)
