"""No-GPU checks of the drop-in boundary: the C-ABI library loads on a CPU-only box, exports every
symbol include/samble_b200.h declares, the ctypes prototypes cover the header exactly, bad arguments
come back as error codes, and the product refuses (loudly) to run without CUDA."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from samble_b200 import _build, _lib, ops
from samble_b200.config import cls_config, seg_config


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.lib()


def test_header_symbols_exported_and_bound(lib):
    declared = _lib.header_symbols()
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/samble_b200.h but not exported"
    assert lib.samble_abi_version() == 1


def test_invalid_arguments_return_codes_not_crashes(lib):
    z = C.c_void_p(0)
    one = C.c_void_p(16)          # never dereferenced: validation fails first
    assert lib.samble_knn(z, 0, 0, 0, z, 0, 0, 0, 1, 8, 8, 3, 4, z, 64, z, 0, z, 0, z) == -1
    assert b"null" in lib.samble_last_error()
    assert lib.samble_knn(one, 0, 0, 0, one, 0, 0, 0, 1, 8, 8, 3, 33, one, 64, z, 0, one, 1 << 20, z) == -1
    assert b"k=33" in lib.samble_last_error()
    assert lib.samble_knn(one, 0, 0, 0, one, 0, 0, 0, 1, 8, 4, 3, 5, one, 64, z, 0, one, 1 << 20, z) == -1
    assert lib.samble_knn(one, 0, 0, 0, one, 0, 0, 0, 1, 8, 8, 3, 4, one, 16, z, 0, one, 1 << 20, z) == -1
    assert lib.samble_knn(one, 0, 0, 0, one, 0, 0, 0, 1, 8, 8, 3, 4, one, 64, z, 0, one, 8, z) == -1   # workspace too small
    assert lib.samble_group(one, one, 64, 1, 3, 8, 4, 7, one, z) == -1
    assert lib.samble_ds_sample(one, one, one, 1, 8, 9, 4, one, one, one, one, one, z, z) == -1
    assert lib.samble_ds_row_stats(one, 128, one, 128, one, 1, 8, 100, 4, one, one, one, z) == -1
    assert lib.samble_knn_workspace_bytes(16, 2048, 2048, 128) > 16 * 2048 * 128 * 4
    assert lib.samble_knn_workspace_bytes(0, 1, 1, 1) == 0


def test_no_cpu_fallback():
    x = torch.randn(1, 3, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.group(x, 4, "center_diff")
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.knn(x.transpose(1, 2), x.transpose(1, 2), 4)
    with pytest.raises(ValueError):
        ops.group(x, 4, "bogus")
    with pytest.raises(ValueError):
        ops.select_neighbors(x, 4, "bogus")
    from samble_b200.models import ShapeNetModel

    m = ShapeNetModel(seg_config(M=(8, 4))).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 3, 16), torch.zeros(1, 16, 1))


def test_mlp2_host_side_contract():
    """the fused two-layer kernel's host wrapper: shape gate, and no CPU path (csrc/mlp2.cu behind ops.mlp2)."""
    assert ops.mlp2_eligible(128, 512, 128) and ops.mlp2_eligible(128, 1024, 256) and ops.mlp2_eligible(64, 256, 128)
    assert not ops.mlp2_eligible(256, 512, 128)          # input wider than one resident X tile
    assert not ops.mlp2_eligible(128, 500, 128)          # hidden width not a whole number of 128-wide chunks
    assert not ops.mlp2_eligible(128, 512, 64)           # output tile must be 128 or 256 wide
    x, w1, w2 = torch.randn(1, 128, 128), torch.randn(512, 128), torch.randn(128, 512)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mlp2(x, w1, w2)


def test_host_build_of_k_allocation_matches_oracle():
    """csrc/kalloc.h is compiled for the host too; the CUDA sampler runs the same source."""
    from oracle import samble_oracle as O
    from tests.golden import make_golden as G

    _build.build()
    host = C.CDLL(_build.HOSTLIB)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ops_small.npz"))
    rng = np.random.default_rng(0)
    for name, B, nb, M, N, seed in G.KALLOC_CASES + [("nb5", 256, 5, 300, 1024, 7), ("nb8", 256, 8, 700, 2048, 8),
                                                      ("nb3", 256, 3, 100, 512, 9), ("nb7", 64, 7, 33, 128, 10)]:
        w, cnt = G.kalloc_inputs(B, nb, M, N, seed)
        k = np.zeros((B, nb), np.int32)
        host.samble_host_num_points_to_choose(w.numpy().ctypes.data_as(C.c_void_p), cnt.numpy().ctypes.data_as(C.c_void_p),
                                              B, nb, M, k.ctypes.data_as(C.c_void_p))
        ref = O.calculate_num_points_to_choose(w, cnt, M).numpy()
        np.testing.assert_array_equal(k, ref, err_msg=name)
        if f"kalloc.{name}" in gold:
            np.testing.assert_array_equal(k, gold[f"kalloc.{name}"])
    host.samble_host_aten_row_sum.restype = C.c_float
    for n in range(1, 9):
        x = (rng.random((2000, n)) * 300).astype(np.float32)
        mine = np.array([host.samble_host_aten_row_sum(r.ctypes.data_as(C.c_void_p), n) for r in x], np.float32)
        np.testing.assert_array_equal(mine, torch.from_numpy(x).sum(dim=1).numpy())


def test_state_dict_layout_is_the_documented_one():
    """parameter names/shapes a reference checkpoint expects."""
    from samble_b200.models import ModelNetModel, ShapeNetModel

    sd = ShapeNetModel(seg_config()).state_dict()
    assert tuple(sd["block.downsample_list.0.bin_tokens"].shape) == (1, 128, 4)
    assert tuple(sd["block.downsample_list.1.q_conv.weight"].shape) == (128, 128, 1)
    assert tuple(sd["block.feature_learning_layer_list.4.k_conv.weight"].shape) == (128, 128, 1, 1)
    assert tuple(sd["block.feature_learning_layer_list.0.ff.0.weight"].shape) == (512, 128, 1)
    assert tuple(sd["block.embedding_list.1.conv1.0.weight"].shape) == (64, 128, 1, 1)
    assert tuple(sd["block.upsample_list.0.res_conv.0.weight"].shape) == (128, 256, 1)
    assert len(sd) == 188
    sdc = ModelNetModel(cls_config()).state_dict()
    assert tuple(sdc["block.downsample_list.0.bin_tokens"].shape) == (1, 128, 6) and len(sdc) == 96


def test_checkpoint_wire_format_roundtrip(tmp_path):
    """train_shapenet.py:663-675 / test_shapenet.py:173-187: DDP-prefixed state_dict, alone or with the per-layer
    [upper, lower] boundary pairs; loading installs the boundaries and freezes the EMA.  (CPU: construction and
    load_state_dict need no kernel.)"""
    import torch
    from samble_b200 import checkpoint, models
    from samble_b200.config import seg_config
    from samble_b200.testing import fill_state_dict_

    cfg = seg_config(M=(128, 64))
    src = models.ShapeNetModel(cfg)
    fill_state_dict_(src.state_dict(), seed=5)
    nb = src.block.downsample_list[0].num_bins
    for l, ds in enumerate(src.block.downsample_list):
        cuts = [1.5 - 0.5 * j - 0.1 * l for j in range(nb - 1)]
        ds.bin_boundaries = [torch.tensor([float("inf")] + cuts).reshape(1, 1, 1, nb), torch.tensor(cuts + [float("-inf")]).reshape(1, 1, 1, nb)]
    for dyn in (True, False):
        path = tmp_path / f"checkpoint_{dyn}.pt"
        checkpoint.save(src, str(path), dynamic_boundaries=dyn)
        raw = torch.load(str(path), weights_only=False)
        sd = raw["model_state_dict"] if dyn else raw
        assert all(k.startswith("module.") for k in sd) and (("bin_boundaries" in raw) == dyn)
        dst = models.ShapeNetModel(cfg)
        checkpoint.load(dst, str(path))
        for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
            assert torch.equal(a, b), k
        for a, b in zip(src.block.downsample_list, dst.block.downsample_list):
            if dyn:
                assert b.dynamic_boundaries_enable is False
                assert torch.equal(a.bin_boundaries[0], b.bin_boundaries[0]) and torch.equal(a.bin_boundaries[1], b.bin_boundaries[1])
            else:
                assert b.bin_boundaries is None or b.dynamic_boundaries_enable == a.dynamic_boundaries_enable
    with __import__("pytest").raises(RuntimeError):
        bad = checkpoint.state(src)
        bad["bin_boundaries"] = bad["bin_boundaries"][:1]
        checkpoint.load(models.ShapeNetModel(cfg), bad)
