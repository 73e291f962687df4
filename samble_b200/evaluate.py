"""Evaluation harness around the hot path: what test_shapenet.py:209-415 / test_modelnet.py do with a
model, minus wandb / HDF5 / DDP -- synthetic stand-ins for utils/dataloader.py's datasets, the no_grad forward loop, the
per-layer sampled indices the reference collects for its visualisations (test_shapenet.py:272-276), and predictions in the
layout utils/metrics.py expects, so `metrics.calculate_shape_IoU(pred, seg_label, category_id, mapping)` and
`metrics.calculate_accuracy(preds, labels)` of the unmodified reference can be applied to the result.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import numpy as np
import torch
from torch.utils.data import Dataset

from . import checkpoint

Tensor = torch.Tensor


class SyntheticShapeNetPart(Dataset):
    """ShapeNetPart-shaped samples (utils/dataloader.py ShapeNet datasets): pcd (3,N) fp32, seg_label (N,) int64 in [0,50),
    category one-hot (16,1).  Labels are a deterministic function of the coordinates (octant of the point, offset by the
    category), so a metric computed on them is reproducible; there are no real shapes offline."""

    def __init__(self, size: int, N: int = 2048, seed: int = 0):
        self.size, self.N, self.seed = size, N, seed

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        xyz = torch.rand(3, self.N, generator=g) * 2 - 1
        cat = i % 16
        onehot = torch.zeros(16, 1)
        onehot[cat, 0] = 1.0
        octant = (xyz[0] > 0).long() + 2 * (xyz[1] > 0).long()          # 4 "parts" per category
        seg = (cat * 3 + octant) % 50
        return xyz, seg, onehot, cat


class SyntheticModelNet(Dataset):
    """ModelNet40-shaped samples: pcd (3,N) fp32, cls_label int in [0,40)."""

    def __init__(self, size: int, N: int = 1024, seed: int = 0):
        self.size, self.N, self.seed = size, N, seed

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        return torch.rand(3, self.N, generator=g) * 2 - 1, i % 40


@torch.no_grad()
def evaluate_seg(model, batches: Iterable, device="cuda:0", metrics=None, mapping=None) -> Dict[str, object]:
    """batches yield (pcd (B,3,N), seg_label (B,N), category one-hot (B,16,1), category id (B,)).  Returns predictions /
    labels as numpy (B_total,N), per-layer sampled indices, point accuracy, and -- when the reference's utils.metrics module
    and its category mapping are passed -- shape / category mIoU computed by the reference's own functions."""
    model.eval()
    preds, labels, cats = [], [], []
    ds_idx: List[List[np.ndarray]] = [[] for _ in checkpoint.downsample_layers(model)]
    for pcd, seg, onehot, cat in batches:
        out = model(pcd.to(device), onehot.to(device))
        out = out[0] if isinstance(out, tuple) else out
        preds.append(out.argmax(dim=1).cpu().numpy())
        labels.append(np.asarray(seg))
        cats.append(np.asarray(cat))
        for l, ds in enumerate(checkpoint.downsample_layers(model)):
            ds_idx[l].append(ds.idx.cpu().numpy())
    pred, label, cat = np.concatenate(preds), np.concatenate(labels), np.concatenate(cats)
    res = dict(pred=pred, seg_label=label, category_id=cat, ds_idx=[np.concatenate(v) for v in ds_idx],
               point_accuracy=float((pred == label).mean()))
    if metrics is not None and mapping is not None:
        shape_ious = metrics.calculate_shape_IoU(pred, label, cat, mapping)
        res["shape_mIoU"] = float(np.mean(shape_ious))
        res["category_IoU"] = metrics.calculate_category_IoU(shape_ious, cat, mapping)
    return res


@torch.no_grad()
def evaluate_cls(model, batches: Iterable, device="cuda:0", metrics=None) -> Dict[str, object]:
    """batches yield (pcd (B,3,N), cls_label (B,)) -> predictions, labels, accuracy (utils/metrics.py:53-55 when given)."""
    model.eval()
    preds, labels = [], []
    ds_idx: List[List[np.ndarray]] = [[] for _ in checkpoint.downsample_layers(model)]
    for pcd, y in batches:
        preds.append(model(pcd.to(device)).argmax(dim=1).cpu().numpy())
        labels.append(np.asarray(y))
        for l, ds in enumerate(checkpoint.downsample_layers(model)):
            ds_idx[l].append(ds.idx.cpu().numpy())
    pred, label = np.concatenate(preds), np.concatenate(labels)
    acc = float(metrics.calculate_accuracy(pred, label)) if metrics is not None else float((pred == label).mean())
    return dict(pred=pred, cls_label=label, accuracy=acc, ds_idx=[np.concatenate(v) for v in ds_idx])
