"""Model wiring around the hot-path blocks: the CALLER side of the drop-in boundary.

The reference's wiring (models/seg_model.py, models/cls_model.py) is out of the hot-path scope
and is meant to be used unmodified with samble_b200.patch.  It does not exist on the GPU box,
so these mirrors (same submodule names => same state_dict keys and shapes) give bench.py and
the parity tests a caller.  STN and the heads are stock PyTorch GEMMs; the only liberty taken
is algebraic: the seg head's conv2 over [global.repeat(N) | per-point] (seg_model.py:205-210)
is split into a per-cloud matvec plus a per-point GEMM (identical result, 18x fewer FLOPs).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import blocks, ops
from ._precision import fp32_forward

Tensor = torch.Tensor


def _cbl1d(cin, cout):
    return nn.Sequential(nn.Conv1d(cin, cout, kernel_size=1, bias=False), nn.BatchNorm1d(cout),
                         nn.LeakyReLU(negative_slope=0.2))


class STN(nn.Module):
    """models/embedding.py:42-97 (parameters and forward); plain PyTorch, not a hot-path kernel."""

    def __init__(self):
        super().__init__()

        def cbl2d(cin, cout):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False), nn.BatchNorm2d(cout),
                                 nn.LeakyReLU(negative_slope=0.2))

        def lbl(cin, cout):
            return nn.Sequential(nn.Linear(cin, cout, bias=False), nn.BatchNorm1d(cout), nn.LeakyReLU(negative_slope=0.2))

        self.conv1, self.conv2, self.conv3 = cbl2d(6, 64), cbl2d(64, 128), _cbl1d(128, 1024)
        self.linear1, self.linear2 = lbl(1024, 512), lbl(512, 256)
        self.transform = nn.Linear(256, 9)
        nn.init.constant_(self.transform.weight, 0)
        nn.init.eye_(self.transform.bias.view(3, 3))
        self.dp1, self.dp2 = nn.Dropout(p=0.5), nn.Dropout(p=0.5)
        self._fold = blocks._FoldCache()

    def _tail(self, x: Tensor) -> Tensor:
        B = x.size(0)
        x = self.conv3(x).max(dim=-1, keepdim=False)[0] if blocks.differentiable(self, x) else blocks.cbl_pool(self.conv3, x, want_mean=False)[0]
        x = self.dp2(self.linear2(self.dp1(self.linear1(x))))
        return self.transform(x).view(B, 3, 3)

    @fp32_forward
    def forward(self, x: Tensor) -> Tensor:
        """reference signature: x is the grouped (B,6,N,K) tensor of ops.group(x,32,'center_diff')."""
        return self._tail(self.conv2(self.conv1(x)).max(dim=-1, keepdim=False)[0])

    @fp32_forward
    def forward_cloud(self, x: Tensor, K: int = 32) -> Tensor:
        """same result from the raw (B,3,N) cloud: conv1 -> conv2 -> max over K is the EdgeConv pattern
        (seg_model.py:182-184 + embedding.py:81-85), so eval mode reuses the fused edge-MLP kernel."""
        if blocks.differentiable(self, x):
            return self.forward(ops.group(x, K, "center_diff")[0])
        idx = ops.fork(lambda: ops.knn_indices(x, K, ordered=False))
        params = [self.conv1[0].weight, self.conv2[0].weight, *self.conv1[1].parameters(), *self.conv1[1].buffers(),
                  *self.conv2[1].parameters(), *self.conv2[1].buffers()]
        weights = self._fold.get(params, lambda: blocks.edge_mlp_weights(self.conv1[0], self.conv1[1], self.conv2[0],
                                                                         self.conv2[1], "center_diff"))
        return self._tail(blocks.fused_edge_mlp(x, idx, weights))


class _BlockBase(nn.Module):
    def _build_common(self, cfg):
        self.embedding_list = nn.ModuleList([blocks.EdgeConv(cfg.embedding, l) for l in range(len(cfg.embedding.K))])
        if cfg.downsample.ds_which != "token":
            raise NotImplementedError("only ds_which='token' (seg.yaml:103 / cls.yaml:120) is on the hot path")
        self.downsample_list = nn.ModuleList([blocks.DownSampleToken(cfg.downsample, l) for l in range(len(cfg.downsample.M))])
        if cfg.attention.get("fl_which", "n2p") != "n2p":
            raise NotImplementedError("only fl_which='n2p' (default.yaml:234) is on the hot path")
        self.feature_learning_layer_list = nn.ModuleList(
            [blocks.Neighbor2PointAttention(cfg.attention, l) for l in range(len(cfg.attention.K))])

    def _front(self, x: Tensor):
        if blocks.differentiable(self, x):
            feats = []
            for emb in self.embedding_list:
                x = emb(x)
                feats.append(x)
            return self.feature_learning_layer_list[0](torch.cat(feats, dim=1))
        # inference: every EdgeConv writes its columns of the concatenated embedding (seg_model.py:98-102) in place, as
        # point-major rows -- no torch.cat, and the attention layer reads the rows without a transpose
        widths = [emb.conv2[0].out_channels for emb in self.embedding_list]
        B, _, N = x.shape
        both = torch.empty(B, N, sum(widths), dtype=torch.float32, device=x.device)
        c0 = 0
        for emb, w in zip(self.embedding_list, widths):
            x = emb(x, out_rows=both[..., c0:c0 + w])
            c0 += w
        return self.feature_learning_layer_list[0](both.transpose(1, 2))


class SegFeatureLearningBlock(_BlockBase):
    """models/seg_model.py:7-133."""

    def __init__(self, cfg):
        super().__init__()
        self._build_common(cfg)
        if cfg.upsample.us_which != "interpolation":
            raise NotImplementedError("only us_which='interpolation' (seg.yaml:124) is on the hot path")
        self.upsample_list = nn.ModuleList([blocks.UpSampleInterpolation(cfg.upsample, l) for l in range(len(cfg.upsample.q_in))])

    @fp32_forward
    def forward(self, x: Tensor) -> Tensor:
        x_xyz = x[:, :3, :]
        x = self._front(x)
        x_list, xyz_list, idx_list = [x], [x_xyz], []
        for i, ds in enumerate(self.downsample_list):
            (x, idx_select), _ = ds(x, x_xyz)
            x = self.feature_learning_layer_list[i + 1](x)
            x_xyz = ops.gather_by_idx(x_xyz, idx_select)
            x_list.append(x), xyz_list.append(x_xyz), idx_list.append(idx_select)
        split = int((len(self.feature_learning_layer_list) - 1) / 2)
        x = ((x_list.pop(), idx_list.pop(), xyz_list.pop()), (None, None))
        for j, up in enumerate(self.upsample_list):
            x_tmp = x_list.pop()
            xyz_tmp = xyz_list[-1 - j]
            x = up(x_tmp, x, xyz_tmp)
            x = self.feature_learning_layer_list[j + 1 + split](x)
            if j < len(self.upsample_list) - 1:
                x = ((x, idx_list.pop(), xyz_list[-1 - j]), (None, None))
        return x


class ShapeNetModel(nn.Module):
    """models/seg_model.py:136-224.  forward(x (B,3,N), category_id (B,16,1)) -> (B,50,N)."""

    def __init__(self, config):
        super().__init__()
        flb = config.feature_learning_block
        self.block = SegFeatureLearningBlock(flb)
        c = flb.attention.ff_conv2_channels_out[-1]
        self.conv = _cbl1d(c, 1024)
        self.conv1 = _cbl1d(16, 64)
        self.conv2 = _cbl1d(c + 2048 + 64, 1024)
        self.conv3 = _cbl1d(1024, 256)
        self.conv4 = nn.Conv1d(256, 50, kernel_size=1, bias=False)
        self.dp1, self.dp2 = nn.Dropout(p=0.5), nn.Dropout(p=0.5)
        self.STN_enable = flb.STN
        if self.STN_enable:
            self.STN = STN()
        self.stn_regularization_loss_factor = config.train.stn_regularization_loss_factor

    @fp32_forward
    def forward(self, x: Tensor, category_id: Tensor):
        B, C, N = x.shape
        trans = None
        if self.STN_enable:
            trans = self.STN.forward_cloud(x, 32)
            x = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1).contiguous()
        f = self.block(x)                                                     # (B,C,N)
        if blocks.differentiable(self, x, f):
            g = self.conv(f)
            g = torch.cat([g.max(dim=-1, keepdim=True)[0], g.mean(dim=-1, keepdim=True), self.conv1(category_id)], dim=1)
            w2 = self.conv2[0].weight
            ng = g.shape[1]
            y = F.conv1d(f, w2[:, ng:]) + F.conv1d(g, w2[:, :ng])
            y = self.conv2[2](self.conv2[1](y))
            y = self.dp2(self.conv3(self.dp1(y)))
            y = self.conv4(y)
        else:
            # head on the tensor cores (linear_tc.cu); BatchNorms folded; conv2 over cat([g.repeat(N), f]) ==
            # W_g g (one vector per cloud, folded into a per-cloud shift) + W_f f (per point)
            # The head stays point-major between its layers (one transpose in, the last GEMM writes (B,50,N)); the
            # 1024-channel global branch is pooled inside the GEMM epilogue and never stored.
            f_rows = ops.rows_of(f)                               # (B,N,C)
            gmax, gmean = blocks.cbl_pool(self.conv, f_rows, x_layout="rows")
            g = torch.cat([gmax, gmean, self.conv1(category_id).squeeze(-1)], dim=1)     # (B, 2048+64)
            w2 = self.conv2[0].weight
            ng = g.shape[1]
            a2, b2 = blocks.folded(self.conv2[1])
            gv = g @ w2[:, :ng, 0].t()                                        # (B,1024)
            w3 = self.conv3[0].weight
            if ops.FUSED_MLP2 and N % 128 == 0 and ops.mlp2_eligible(f_rows.shape[-1], w2.shape[0], w3.shape[0]):
                # conv2 -> conv3 in one kernel: the (B,N,1024) activation stays in tensor memory (csrc/mlp2.cu)
                a3, b3 = blocks.folded(self.conv3[1])
                y = ops.mlp2(f_rows, w2[:, ng:, 0], w3, scale1=a2, shift1=gv * a2 + b2, lrelu1=True, scale2=a3, shift2=b3, lrelu2=True)
            else:
                y = ops.linear(f_rows, w2[:, ng:, 0], scale=a2, shift=gv * a2 + b2, lrelu=True)      # (B,N,1024)
                y = blocks.cbl(self.conv3, y, x_layout="rows", out_layout="rows")
            y = ops.linear(y, self.conv4.weight, out_layout="bcn")
        return (y, trans) if self.stn_regularization_loss_factor > 0 else y


class ClsFeatureLearningBlock(_BlockBase):
    """models/cls_model.py:10-145 with res_link on (cls.yaml:98-99); the FPS branch (:117-130) is
    dead in the shipped scripts (train_modelnet.py:242) and is not built."""

    def __init__(self, cfg, fps: bool = False):
        super().__init__()
        if fps:
            raise NotImplementedError("farthest point sampling is out of the hot-path scope (SURVEY 2 #12)")
        self._build_common(cfg)
        self.res_link_enable = cfg.res_link.enable
        outs = cfg.attention.ff_conv2_channels_out
        if self.res_link_enable:
            self.conv_list = nn.ModuleList([nn.Conv1d(c, 1024, kernel_size=1, bias=False) for c in outs])
        else:
            self.conv = nn.Conv1d(outs[-1], 1024, kernel_size=1, bias=False)
        self.M_list = cfg.downsample.M

    def _pooled(self, conv: nn.Conv1d, x: Tensor) -> Tensor:
        """conv(x).max(dim=-1)[0] (cls_model.py:104,133); eval mode pools inside the GEMM epilogue."""
        if blocks.differentiable(self, x) or x.shape[-1] % 32:
            return conv(x).max(dim=-1)[0]
        return ops.linear_pool(ops.rows_of(x), conv.weight, want_mean=False)[0]

    @fp32_forward
    def forward(self, x: Tensor):
        x_xyz = x.clone()
        x = self._front(x)
        if not self.res_link_enable:
            for i, ds in enumerate(self.downsample_list):
                x = self.feature_learning_layer_list[i + 1](ds(x)[0][0])
            return self._pooled(self.conv, x)
        res = [self._pooled(self.conv_list[0], x)]
        for i, ds in enumerate(self.downsample_list):
            (x, idx_select) = ds(x, x_xyz)[0]
            x = self.feature_learning_layer_list[i + 1](x)
            x_xyz = ops.gather_by_idx(x_xyz, idx_select)
            res.append(self._pooled(self.conv_list[i + 1], x))
        self.res_link_list = res
        return torch.cat(res, dim=1), res


class ModelNetModel(nn.Module):
    """models/cls_model.py:148-205.  forward(x (B,3,N)) -> (B,40)."""

    def __init__(self, config, fps: bool = False):
        super().__init__()
        flb = config.feature_learning_block
        self.block = ClsFeatureLearningBlock(flb, fps)
        n_layers = len(flb.attention.K)
        self.res_link_enable = flb.res_link.enable

        def lbl(cin, cout):
            return nn.Sequential(nn.Linear(cin, cout), nn.BatchNorm1d(cout), nn.LeakyReLU(negative_slope=0.2), nn.Dropout(p=0.5))

        if self.res_link_enable:
            self.linear1 = lbl(1024 * n_layers, 1024)
        self.linear2 = lbl(1024, 256)
        self.linear3 = nn.Linear(256, 40)

    @fp32_forward
    def forward(self, x: Tensor) -> Tensor:
        if self.res_link_enable:
            x, _ = self.block(x)
            return self.linear3(self.linear2(self.linear1(x)))
        return self.linear3(self.linear2(self.block(x)))


def freeze_boundaries(model: nn.Module) -> None:
    """SURVEY 8c protocol step 2: after one calibration batch, stop the EMA so runs are repeatable
    and clouds/ranks decouple (inference semantics)."""
    for m in model.modules():
        if isinstance(m, blocks.DownSampleToken):
            if m.bin_boundaries is None:
                raise RuntimeError("freeze_boundaries: run one calibration batch first")
            m.dynamic_boundaries_enable = False
