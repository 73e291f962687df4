"""Install the native path into the reference's own modules (the drop-in).

    sys.path.insert(0, "/path/to/SAMBLE")
    import samble_b200.patch as patch
    patch.install()            # or:  with patch.installed(): ...
    from models.seg_model import ShapeNetModel      # unmodified reference wiring, our blocks + ops

Swaps: the hot-path functions of `utils.ops`, the three names `models.downsample`
binds at import time (downsample.py:8-12), and the four block classes.
"""
from __future__ import annotations

import contextlib
import importlib

from . import blocks, ops

_OPS = ["knn", "index_points", "select_neighbors", "select_neighbors_interpolate", "group", "neighbor_mask",
        "gather_by_idx", "update_sampling_score_bin_boundary", "bin_partition", "calculate_num_points_to_choose",
        "generating_downsampled_index"]
_BOUND_BY_NAME = ["calculate_num_points_to_choose", "bin_partition", "generating_downsampled_index"]
_BLOCKS = [("models.embedding", "EdgeConv"), ("models.attention", "Neighbor2PointAttention"),
           ("models.downsample", "DownSampleToken"), ("models.upsample", "UpSampleInterpolation")]
_saved = []


def install() -> None:
    if _saved:
        return
    ref_ops = importlib.import_module("utils.ops")
    for name in _OPS:
        _saved.append((ref_ops, name, getattr(ref_ops, name)))
        setattr(ref_ops, name, getattr(ops, name))
    ds = importlib.import_module("models.downsample")
    for name in _BOUND_BY_NAME:
        _saved.append((ds, name, getattr(ds, name)))
        setattr(ds, name, getattr(ops, name))
    for mod_name, cls_name in _BLOCKS:
        mod = importlib.import_module(mod_name)
        _saved.append((mod, cls_name, getattr(mod, cls_name)))
        setattr(mod, cls_name, getattr(blocks, cls_name))


def uninstall() -> None:
    while _saved:
        mod, name, old = _saved.pop()
        setattr(mod, name, old)


@contextlib.contextmanager
def installed():
    install()
    try:
        yield
    finally:
        uninstall()
